"""Parity against the REFERENCE'S OWN CUDA PATH on the B200.

``oracle/_ref/*.cubin`` are the reference's correlation kernels (src/correlation.py:9-104), compiled from the kernel
strings where they lie through the reference's own ``cupy_kernel()`` templating (recipe: oracle/build_ref.py) and
launched with the reference's geometry (oracle/ref_cuda.py).  These tests

  * pin the CPU oracle's correlation restatements against the real kernels,
  * compare this repo's ``FunctionCorrelation`` with the real kernels on the same inputs,
  * run the whole forward the way the reference does on a GPU -- torch/cuDNN fp32 convolutions, ATen grid_sample /
    unfold / interpolate, and the reference's correlation kernels -- and compare this repo's forward with it
    (north_star tolerance: flow max <= 1e-2 px, mean <= 1e-3 px), and
  * time that reference CUDA path next to this repo's at 1024x1024 (a report, not an assertion on speed).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import lfn_oracle as O
from oracle import ref_cuda as RC
from pivlfn import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda"
SHAPES = RC.shapes()


def test_ref_manifest_lists_existing_binaries():
    """CPU check: when oracle/_ref was built, every file of the manifest exists and is an ELF cubin."""
    m = RC.manifest()
    if m is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    assert len(m["entries"]) >= 9
    for e in m["entries"]:
        for f in e["files"].values():
            with open(os.path.join(RC.REF, f), "rb") as fh:
                assert fh.read(4) == b"\x7fELF"


needs_ref = pytest.mark.skipif(not SHAPES, reason="oracle/_ref not built")


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def _report(line):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.txt"), "a") as fh:
            fh.write(line + "\n")
    print(line)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("shape", [s for s in SHAPES if s[2] * s[3] <= 128 * 128], ids=str)
def test_correlation_vs_reference_cuda_kernels(shape):
    from src.correlation import FunctionCorrelation
    B, C, H, W, s = shape
    f1, f2 = _rand((B, C, H, W), 11 * C + H), _rand((B, C, H, W), 13 * C + W)
    ref = RC.reference_correlation(f1.to(DEV), f2.to(DEV), s)
    # (1) the oracle's vectorised restatement vs the real kernels (different fp32 association order)
    orc = O.correlation(f1, f2, s)
    assert ref.shape == orc.shape
    assert (ref.cpu() - orc).abs().max().item() <= 1e-5
    # (2) the literal emulation (same 32-lane partial-sum order) vs the real kernels: the only freedom left is FMA
    # contraction inside a lane's accumulation, so a few ulp of the O(1) values
    if B * H * W <= 64 * 96:
        lit = O.correlation_literal(f1.numpy(), f2.numpy(), s)
        assert np.abs(ref.cpu().numpy() - lit).max() <= 2e-6
    # (3) this repo's drop-in operator vs the real kernels
    out = FunctionCorrelation(tensorFirst=f1.to(DEV), tensorSecond=f2.to(DEV), intStride=s)
    d = (out - ref).abs().max().item()
    _report(f"FunctionCorrelation vs reference CUDA kernels {shape}: max|diff| {d:.3e}")
    assert out.shape == ref.shape and d <= 1e-5


def _reference_cuda_forward(sd, a, b, model):
    """The reference's GPU path: stock torch ops in true fp32 (TF32 off) + its own correlation kernels."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        with torch.no_grad():
            return O.forward(sdd, a.to(DEV), b.to(DEV), model,
                             corr_fn=lambda x, y, s: RC.reference_correlation(x.contiguous(), y.contiguous(), s))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("precision", ["f16c", "tf32c", "3xtf32", "simt"])
@pytest.mark.parametrize("model,B,H,W,kind", [("piv", 1, 128, 128, "rankine"), ("piv", 2, 64, 96, "shear"),
                                             ("hui", 1, 64, 128, "uniform")])
def test_forward_vs_reference_cuda_path(model, B, H, W, kind, precision):
    from src.models import hui_liteflownet, piv_liteflownet
    sd = synth.synthetic_state_dict(model, 5)
    a, b, _ = synth.particle_batch(B, H, W, 300 + H, kind)
    ref = _reference_cuda_forward(sd, a, b, model)
    net = (piv_liteflownet if model == "piv" else hui_liteflownet)(sd, 1).to(DEV).eval()
    net.precision = precision
    with torch.no_grad():
        out = net(a.to(DEV), b.to(DEV))
    diff = (out - ref).abs()
    _report(f"forward vs reference CUDA path {model} {B}x{H}x{W} {precision}: max {diff.max().item():.3e} "
            f"mean {diff.mean().item():.3e} (|flow|max {ref.abs().max().item():.2f})")
    assert out.shape == ref.shape
    # north_star tolerance, ABSOLUTE (also for Hui, whose flows carry the x20 output scale)
    assert diff.max().item() <= 1e-2 and diff.mean().item() <= 1e-3


def _unmodified_reference_forward(sd, a, b, model, version=1):
    """The reference ITSELF on the GPU: its unmodified src/models.py + src/correlation.py (baseline/_ref/reference, installed by
    baseline/install_ref.py), its CUDA kernels compiled at run time through the cupy stand-in (baseline/cupy_stub: NVRTC +
    driver API, the reference's own launch geometry), stock torch ops in true fp32 (TF32 off)."""
    from collections import OrderedDict
    from oracle import ref_import as R
    models, _ = R.load_reference_models()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        fac = models.piv_liteflownet if model.startswith("piv") else models.hui_liteflownet
        net = fac(OrderedDict((k, v.clone()) for k, v in sd.items()), version).to(DEV).eval()
        with torch.no_grad():
            return net(a.to(DEV).clone(), b.to(DEV).clone())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _ref_installed():
    from oracle import ref_import as R
    return R.available()


needs_installed_ref = pytest.mark.skipif(not _ref_installed(), reason="baseline/_ref/reference not installed")


@pytest.mark.gpu
@needs_installed_ref
@pytest.mark.parametrize("model,version,B,H,W,kind", [("piv", 1, 1, 128, 128, "rankine"), ("piv", 1, 2, 64, 96, "shear"),
                                                     ("hui", 1, 1, 64, 128, "uniform"), ("piv2", 2, 1, 64, 64, "shear"),
                                                     ("hui2", 2, 1, 64, 64, "uniform")])
def test_forward_vs_unmodified_reference_on_gpu(model, version, B, H, W, kind):
    """This repo's drop-in (default precision) against the UNMODIFIED reference running its own CUDA path on the same B200,
    same inputs, same weights: north_star tolerance, absolute."""
    from src.models import hui_liteflownet, piv_liteflownet
    sd = synth.synthetic_state_dict(model, 5)
    a, b, _ = synth.particle_batch(B, H, W, 300 + H, kind)
    ref = _unmodified_reference_forward(sd, a, b, model, version)
    net = (piv_liteflownet if model.startswith("piv") else hui_liteflownet)(sd, version).to(DEV).eval()
    with torch.no_grad():
        out = net(a.to(DEV), b.to(DEV))
    diff = (out - ref).abs()
    _report(f"forward vs UNMODIFIED reference on the GPU {model} {B}x{H}x{W} {net.engine().precision}: max {diff.max().item():.3e} "
            f"mean {diff.mean().item():.3e} (|flow|max {ref.abs().max().item():.2f})")
    assert out.shape == ref.shape
    assert diff.max().item() <= 1e-2 and diff.mean().item() <= 1e-3


@pytest.mark.gpu
@needs_installed_ref
def test_unmodified_reference_correlation_matches_cubins_and_oracle():
    """The cupy stand-in runs the reference's kernels exactly like the ahead-of-time cubins do (same source, same geometry)."""
    from oracle import ref_import as R
    _, corr = R.load_reference_models()
    f1, f2 = _rand((2, 64, 32, 32), 5).to(DEV), _rand((2, 64, 32, 32), 6).to(DEV)
    out = corr.FunctionCorrelation(tensorFirst=f1, tensorSecond=f2, intStride=2)
    assert (out.cpu() - O.correlation(f1.cpu(), f2.cpu(), 2)).abs().max().item() <= 1e-5
    if (2, 64, 32, 32, 2) in SHAPES:
        assert torch.equal(out, RC.reference_correlation(f1, f2, 2))


@pytest.mark.gpu
@needs_installed_ref
def test_three_product_variant_vs_unmodified_reference_on_gpu(monkeypatch):
    """The fp32-equivalent variant (PIVLFN_P16=0: fp32 activations, three fp16 products per product), which bench.py reports
    separately, against the UNMODIFIED reference on the GPU: same absolute tolerance, and closer than the default variant."""
    from src.models import piv_liteflownet
    sd = synth.synthetic_state_dict("piv", 5)
    a, b, _ = synth.particle_batch(1, 128, 128, 428, "rankine")
    ref = _unmodified_reference_forward(sd, a, b, "piv", 1)
    monkeypatch.setenv("PIVLFN_P16", "0")
    net = piv_liteflownet(sd, 1).to(DEV).eval()
    assert not net.engine().p16 and net.engine().precision == "f16c"
    with torch.no_grad():
        out = net(a.to(DEV), b.to(DEV))
    diff = (out - ref).abs()
    _report(f"forward vs UNMODIFIED reference on the GPU piv 1x128x128 f16c three-product variant (PIVLFN_P16=0): max {diff.max().item():.3e} "
            f"mean {diff.mean().item():.3e}")
    assert diff.max().item() <= 1e-2 and diff.mean().item() <= 1e-3


@pytest.mark.gpu
@needs_installed_ref
@pytest.mark.parametrize("shape", [(2, 16, 12, 16, 2), (1, 8, 9, 11, 1), (1, 64, 32, 32, 2)], ids=str)
def test_correlation_backward_vs_unmodified_reference(shape):
    """gradFirst / gradSecond of the drop-in operator against the reference's own backward (its updateGradFirst / updateGradSecond
    CUDA kernels, compiled at run time through the cupy stand-in)."""
    from oracle import ref_import as R
    from src.correlation import FunctionCorrelation
    _, corr = R.load_reference_models()
    B, C, H, W, s = shape
    f1, f2 = _rand((B, C, H, W), 31).to(DEV), _rand((B, C, H, W), 32).to(DEV)
    go = _rand((B, 49, -(-H // s), -(-W // s)), 33).to(DEV)
    a, b = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    corr.FunctionCorrelation(tensorFirst=a, tensorSecond=b, intStride=s).backward(go)
    x, y = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    FunctionCorrelation(tensorFirst=x, tensorSecond=y, intStride=s).backward(go)
    d1, d2 = (x.grad - a.grad).abs().max().item(), (y.grad - b.grad).abs().max().item()
    _report(f"correlation backward vs UNMODIFIED reference {shape}: max|diff| gradFirst {d1:.3e} gradSecond {d2:.3e}")
    assert d1 <= 1e-5 and d2 <= 1e-5


@pytest.mark.gpu
@needs_installed_ref
def test_cfg2_bench_batch_vs_unmodified_reference():
    """BASELINE configs[1] at its own shape: the 64-pair 256x256 batch bench.py times goes through piv_liteflownet in ONE
    forward (default precision); 8 of its pairs (every 9th: all 8 pool images, several rolls) are checked against the
    UNMODIFIED reference on the GPU (its own correlation kernels, fp32 convolutions with TF32 off).  Absolute north_star
    tolerance: flow max <= 1e-2 px, mean <= 1e-3 px."""
    import bench
    from src.models import piv_liteflownet
    sd = synth.synthetic_state_dict("piv", 0)
    a, b = bench.synthetic_batch(bench.BATCH, 10_000)
    assert a.shape == (64, 3, 256, 256)
    net = piv_liteflownet(sd, 1).to(DEV).eval()
    with torch.no_grad():
        out = net(a.to(DEV), b.to(DEV))
    idx = list(range(0, 64, 9))                             # 0, 9, ..., 63
    ref = _unmodified_reference_forward(sd, a[idx], b[idx], "piv")
    diff = (out[idx] - ref).abs()
    _report(f"cfg2 (batch 64 of 256x256, {net.engine().precision}) vs UNMODIFIED reference on the GPU, pairs {idx}: "
            f"max {diff.max().item():.3e} mean {diff.mean().item():.3e} (|flow|max {ref.abs().max().item():.2f})")
    assert ref.abs().max().item() > 1.0
    assert diff.max().item() <= 1e-2 and diff.mean().item() <= 1e-3


@pytest.mark.gpu
@needs_ref
def test_reference_cuda_path_rate_report_1024():
    """Times the reference's GPU path (eager torch/cuDNN + its correlation kernels, one 1024x1024 pair per step as
    run.py does) in fp32 and with torch's default TF32 convolutions, next to this repo's forward; also checks they
    agree.  Written to gpurun_out/ref_cuda_rate.json."""
    from src.models import piv_liteflownet
    sd = synth.synthetic_state_dict("piv", 0)
    a, b, _ = synth.particle_batch(1, 1024, 1024, 77, "rankine")
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    ad, bd = a.to(DEV), b.to(DEV)
    corr = lambda x, y, s: RC.reference_correlation(x.contiguous(), y.contiguous(), s)
    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("fp32", False), ("tf32_default", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    ref = O.forward(sdd, ad, bd, "piv", corr_fn=corr)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    r = O.forward(sdd, ad, bd, "piv", corr_fn=corr)
                e1.record()
                torch.cuda.synchronize()
            res[tag] = {"ms_per_pair": e0.elapsed_time(e1) / 3, "pairs_per_s": 3e3 / e0.elapsed_time(e1)}
            if not tf32:
                ref32 = ref
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    net = piv_liteflownet(sd, 1).to(DEV).eval()
    with torch.no_grad():
        for _ in range(3):
            out = net(ad.clone(), bd.clone())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = net(ad.clone(), bd.clone())
        e1.record()
        torch.cuda.synchronize()
    res["pivlfn_" + net.engine().precision] = {"ms_per_pair": e0.elapsed_time(e1) / 5, "pairs_per_s": 5e3 / e0.elapsed_time(e1)}
    diff = (out - ref32).abs()
    res["flow_max_abs_diff_vs_fp32_reference"] = diff.max().item()
    res["flow_mean_abs_diff_vs_fp32_reference"] = diff.mean().item()
    res["note"] = ("reference CUDA path = stock torch 2.11/cuDNN ops driven by the oracle restatement of src/models.py + the "
                   "reference's correlation kernels (cubins), batch 1 of 1024x1024, eager, CUDA-event timed")
    _report("reference CUDA path @1024x1024: " + json.dumps(res))
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        json.dump(res, open(os.path.join(d, "ref_cuda_rate.json"), "w"), indent=1)
    assert diff.max().item() <= 1e-2 and diff.mean().item() <= 1e-3


@pytest.mark.gpu
@needs_ref
def test_correlation_operator_rate_report():
    """The drop-in FunctionCorrelation operator next to the reference's own three kernels (+ its three zero-fills) on the
    level-1 shape of a 1024x1024 frame and on a stride-1 level: a report (gpurun_out/parity_report.txt), plus agreement."""
    from src.correlation import FunctionCorrelation
    res = {}
    for shape in [(1, 64, 1024, 1024, 2), (1, 96, 128, 128, 1)]:
        B, C, H, W, s = shape
        f1, f2 = _rand((B, C, H, W), 3).to(DEV), _rand((B, C, H, W), 4).to(DEV)
        ref = RC.reference_correlation(f1, f2, s)
        out = FunctionCorrelation(tensorFirst=f1, tensorSecond=f2, intStride=s)
        assert (out - ref).abs().max().item() <= 1e-5
        t = {}
        for name, fn in (("reference", lambda: RC.reference_correlation(f1, f2, s)),
                         ("pivlfn", lambda: FunctionCorrelation(tensorFirst=f1, tensorSecond=f2, intStride=s))):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t[name] = e0.elapsed_time(e1) / 5
        res[str(shape)] = t
    _report("FunctionCorrelation operator, ms per call (reference CUDA kernels vs pivlfn_corr_nchw): " + json.dumps(res))


@pytest.mark.gpu
@needs_ref
def test_reference_cuda_path_rate_report_hui_cfg3():
    """BASELINE configs[2]: original LiteFlowNet ('hui') on batch 16 of 1024x436 pairs through estimate() (resized to
    1024x448, output at half resolution resized back).  Reference GPU path (eager torch/cuDNN + its correlation kernels,
    fp32 and torch's TF32 default) next to this repo; also checks that they agree.  A report, not an assertion on speed."""
    from inference import estimate
    from src.models import hui_liteflownet
    sd = synth.synthetic_state_dict("hui", 0)
    a, b, _ = synth.particle_batch(2, 436, 1024, 11, "shear")
    a, b = a.repeat(8, 1, 1, 1).contiguous(), b.repeat(8, 1, 1, 1).contiguous()          # batch 16
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    ad, bd = a.to(DEV), b.to(DEV)
    corr = lambda x, y, s: RC.reference_correlation(x.contiguous(), y.contiguous(), s)

    def ref_estimate():
        # inference.py:39-61 around the reference forward
        import torch.nn.functional as F
        x1 = F.interpolate(ad, size=(448, 1024), mode="bilinear", align_corners=False)
        x2 = F.interpolate(bd, size=(448, 1024), mode="bilinear", align_corners=False)
        raw = O.forward(sdd, x1, x2, "hui", corr_fn=corr)
        flow = F.interpolate(raw, size=(436, 1024), mode="bilinear", align_corners=False)
        flow[:, 1] *= 436.0 / 448.0
        return flow

    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("fp32", False), ("tf32_default", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                ref = ref_estimate()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    r = ref_estimate()
                e1.record()
                torch.cuda.synchronize()
            res[tag] = {"ms_per_batch": e0.elapsed_time(e1) / 2, "pairs_per_s": 16 * 2e3 / e0.elapsed_time(e1)}
            if not tf32:
                ref32 = ref
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    net = hui_liteflownet(sd, 1).to(DEV).eval()
    with torch.no_grad():
        for _ in range(2):
            out = estimate(net, ad, bd, tensor=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = estimate(net, ad, bd, tensor=True)
        e1.record()
        torch.cuda.synchronize()
    res["pivlfn_" + net.engine().precision] = {"ms_per_batch": e0.elapsed_time(e1) / 5, "pairs_per_s": 16 * 5e3 / e0.elapsed_time(e1)}
    diff = (out - ref32).abs()
    res["flow_max_abs_diff_vs_fp32_reference"] = diff.max().item()
    res["flow_mean_abs_diff_vs_fp32_reference"] = diff.mean().item()
    res["flow_abs_max"] = ref32.abs().max().item()
    _report("reference CUDA path, Hui LiteFlowNet batch 16 of 1024x436 through estimate(): " + json.dumps(res))
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        json.dump(res, open(os.path.join(d, "ref_cuda_rate_hui_cfg3.json"), "w"), indent=1)
    # north_star tolerance, absolute
    assert diff.max().item() <= 1e-2 and diff.mean().item() <= 1e-3
