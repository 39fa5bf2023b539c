"""Row-slab tiling of one frame across ranks (BASELINE configs[4]): P ranks emulated in lock-step on ONE GPU
(pivlfn.tiled.LoopbackGroup: the communication steps -- halo exchange, flow-mean all-reduce, coarse-level all-gather --
are tensor copies) must reproduce the single-GPU forward.  The same step lists run under torch.distributed with NCCL
(pivlfn.tiled.DistGroup, tools/tiled_demo.py)."""
import pytest
import torch

from pivlfn import synth
from pivlfn.arch import CFGS
from pivlfn.model import Engine
from pivlfn.tiled import LoopbackGroup, TiledPlan, make_tiled_plan

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("model,H,W,P,precision", [("piv", 256, 128, 2, "simt"), ("piv", 384, 96, 3, "3xtf32"),
                                                    ("hui", 256, 64, 2, "tf32c"), ("piv", 512, 64, 4, "f16c"),
                                                    ("hui", 256, 128, 2, "f16c"), ("piv", 384, 64, 3, "f16c")])
def test_tiled_equals_single_gpu(model, H, W, P, precision):
    sd = {k: v.to(DEV) for k, v in synth.synthetic_state_dict(model, 0).items()}
    eng = Engine(CFGS[model], sd, torch.device(DEV), precision, use_graph=False)
    i1, i2, _ = synth.particle_pair(H, W, 31, "shear")
    a = synth.to_rgb_tensor(i1)[None].to(DEV)
    b = synth.to_rgb_tensor(i2)[None].to(DEV)
    ref = eng.forward(a.clone(), b.clone())
    plans = [make_tiled_plan(eng, H, W, r, P, halo=24, warp_reach=16) for r in range(P)]
    assert 1 <= plans[0].Lt <= 6
    for p in plans:
        p.load_inputs(a, b)
    LoopbackGroup(plans).run()          # includes the displacement-bound / fp16-range check on every rank
    torch.cuda.synchronize()
    out = torch.cat([p.owned_output() for p in plans], dim=2)
    assert out.shape == ref.shape
    diff = (out - ref).abs()
    n_ex = sum(1 for s in plans[0].steps if s.kind == "exchange")
    n_op = sum(1 for s in plans[0].steps if s.kind == "op")
    print(f"tiled {model} {H}x{W} P={P} {precision}: Lt={plans[0].Lt} exchanges={n_ex} ops={n_op} max|diff|={diff.max().item():.3e}")
    # identical kernels on identical data; only the flow-mean summation order differs -> fp32 round-off
    # (the tiled plan sums the flow means in another order: an fp32 ulp can flip an e5m2 rounding of a P16 store, ~1e-4 px)
    assert diff.max().item() <= 1e-3


def test_tiled_raises_when_displacement_exceeds_warp_reach():
    """The backwarp reach is data dependent (flowU for the cost volume, flowM for the Subpixel warp, flowS for the brightness
    error): a run whose vertical displacements exceed the provisioned reach must raise on every rank instead of
    returning seams computed from stale halo rows."""
    from pivlfn.tiled import TiledBoundsError
    sd = {k: v.to(DEV) for k, v in synth.synthetic_state_dict("piv", 0).items()}
    eng = Engine(CFGS["piv"], sd, torch.device(DEV), "f16c", use_graph=False)
    i1, i2, _ = synth.particle_pair(256, 64, 31, "shear")
    a = synth.to_rgb_tensor(i1)[None].to(DEV)
    b = synth.to_rgb_tensor(i2)[None].to(DEV)
    ref = eng.forward(a.clone(), b.clone())
    assert ref[:, 1].abs().max().item() > 1.5          # the synthetic weights produce multi-pixel vertical flows
    plans = [make_tiled_plan(eng, 256, 64, r, 2, halo=8, warp_reach=1) for r in range(2)]
    assert len(plans[0].warp_flows) >= 3
    for p in plans:
        p.load_inputs(a, b)
    with pytest.raises(TiledBoundsError):
        LoopbackGroup(plans).run()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_tiled_dist_group_real_nccl_two_ranks():
    """The real one-process-per-GPU path: pivlfn.tiled.DistGroup under torch.distributed (NCCL send/recv of halo rows, flow-mean
    all-reduce, coarse-level all-gather, the all-rank bounds check) on 2 GPUs reproduces every rank's rows of the single-GPU
    forward."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "tools", "tiled_demo.py"), "512", "256", "--check"],
                       capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    print(line)
    assert d["n_gpus"] == 2 and d["exchanges_per_forward"] > 0
    assert d["max_abs_diff_vs_single_gpu"] <= 1e-3
