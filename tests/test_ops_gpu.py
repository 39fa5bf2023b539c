"""Operator-level parity on the B200: every C-ABI kernel against the CPU oracle (oracle/lfn_oracle.py) or the
torch fp32 op it replaces, on the same seeded inputs.  Everything goes through the C ABI (pivlfn.ops -> ctypes)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import lfn_oracle as O
from pivlfn import ops
from pivlfn.model import pack_conv

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _nhwc(x):       # NCHW cpu -> NHWC cuda (channel count padded to a multiple of 4 with zeros)
    B, C, H, W = x.shape
    cp = (C + 3) & ~3
    t = torch.zeros(B, H, W, cp)
    t[..., :C] = x.permute(0, 2, 3, 1)
    return t.to(DEV)


def _nchw(t, C):
    return t[..., :C].permute(0, 3, 1, 2).cpu()


def _rand(*shape, seed=0, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ---- FunctionCorrelation (public NCHW operator) -------------------------------------------------------------
@pytest.mark.parametrize("C", [64, 96, 128, 192, 5])
@pytest.mark.parametrize("s", [1, 2])
@pytest.mark.parametrize("hw", [(16, 24), (17, 31), (40, 33), (1, 1), (3, 70)])
def test_function_correlation_vs_oracle(C, s, hw):
    from src.correlation import FunctionCorrelation, ModuleCorrelation
    H, W = hw
    f1, f2 = _rand(2, C, H, W, seed=C + s), _rand(2, C, H, W, seed=C + s + 1)
    ref = O.correlation(f1, f2, s)
    out = FunctionCorrelation(tensorFirst=f1.to(DEV), tensorSecond=f2.to(DEV), intStride=s)
    assert out.shape == ref.shape and out.is_cuda and out.dtype == torch.float32
    # fp32 sums of C products in a different association order: tolerance 1e-5 relative to the value scale (~1)
    assert (out.cpu() - ref).abs().max().item() <= 1e-5
    out2 = ModuleCorrelation()(f1.to(DEV), f2.to(DEV), s)
    assert torch.equal(out, out2)


def test_function_correlation_literal_reference_order():
    f1, f2 = _rand(1, 64, 6, 7, seed=3), _rand(1, 64, 6, 7, seed=4)
    lit = O.correlation_literal(f1.numpy(), f2.numpy(), 2)
    out = ops.corr_nchw(f1.to(DEV), f2.to(DEV), 2).cpu().numpy()
    assert np.abs(out - lit).max() <= 1e-5


def test_function_correlation_errors():
    from src.correlation import FunctionCorrelation
    a = torch.zeros(1, 8, 8, 8, device=DEV)
    with pytest.raises(AssertionError):
        FunctionCorrelation(a.permute(0, 1, 3, 2), a, 1)
    with pytest.raises(NotImplementedError):
        FunctionCorrelation(a.cpu(), a.cpu(), 1)
    with pytest.raises(AssertionError):
        FunctionCorrelation(a, a, 3)


@pytest.mark.parametrize("C,s,H,W", [(8, 1, 9, 11), (16, 2, 12, 16), (5, 2, 7, 9)])
def test_function_correlation_backward_vs_autograd_of_the_oracle(C, s, H, W):
    """The operator is differentiable like the reference's (src/correlation.py:348-405): gradients against torch autograd through
    the oracle's pure-torch restatement, fp64."""
    from src.correlation import FunctionCorrelation
    f1, f2 = _rand(2, C, H, W, seed=21), _rand(2, C, H, W, seed=22)
    go = _rand(2, 49, -(-H // s), -(-W // s), seed=23)
    a, b = f1.double().requires_grad_(True), f2.double().requires_grad_(True)
    O.correlation(a, b, s).backward(go.double())
    x, y = f1.to(DEV).requires_grad_(True), f2.to(DEV).requires_grad_(True)
    FunctionCorrelation(x, y, s).backward(go.to(DEV))
    assert (x.grad.cpu().double() - a.grad).abs().max().item() <= 1e-5
    assert (y.grad.cpu().double() - b.grad).abs().max().item() <= 1e-5
    # only one input needs a gradient
    x2 = f1.to(DEV).requires_grad_(True)
    FunctionCorrelation(x2, f2.to(DEV), s).backward(go.to(DEV))
    assert torch.equal(x2.grad, x.grad)
    # like the reference (src/correlation.py:352), a non-contiguous gradient (here: the stride-0 expansion of a scalar) is refused
    with pytest.raises(AssertionError):
        FunctionCorrelation(f1.to(DEV).requires_grad_(True), f2.to(DEV), s).sum().backward()


# ---- model-internal correlation: backwarp fused into the tile load + LeakyReLU -----------------------------
@pytest.mark.parametrize("C,s,H,W", [(64, 2, 32, 48), (96, 1, 16, 24), (128, 1, 8, 8), (192, 1, 2, 3), (64, 2, 64, 64)])
@pytest.mark.parametrize("with_flow", [False, True])
def test_corr_nhwc_fused_warp(C, s, H, W, with_flow):
    f1, f2 = _rand(2, C, H, W, seed=1), _rand(2, C, H, W, seed=2)
    flow = _rand(2, 2, H, W, seed=3, scale=2.0) if with_flow else None
    scale = 1.25
    f2w = O.backwarp(f2, flow * scale) if with_flow else f2
    ref = O.lrelu(O.correlation(f1, f2w, s))
    Ho, Wo = -(-H // s), -(-W // s)
    out = torch.zeros(2, Ho, Wo, 52, device=DEV)
    fl = flow.permute(0, 2, 3, 1).contiguous().to(DEV) if with_flow else None
    ops.corr_nhwc(ops.view(_nhwc(f1)), ops.view(_nhwc(f2)), fl, scale, ops.view(out, 0, 49), 2, H, W, s, True)
    assert (_nchw(out, 49) - ref).abs().max().item() <= 2e-5


def test_warp_vs_reference_vector(golden_dir):
    d = np.load(os.path.join(golden_dir, "backwarp.npz"))
    inp, flow = torch.from_numpy(d["inp"]), torch.from_numpy(d["flow"])
    from src.models import backwarp
    out = backwarp(inp.to(DEV), flow.to(DEV)).cpu().numpy()
    assert np.abs(out - d["out"]).max() <= 2e-5


def test_warp_large_and_nonfinite_flow():
    x = _rand(1, 8, 16, 16, seed=5)
    flow = torch.zeros(1, 2, 16, 16)
    flow[0, 0, 0, 0] = 1e30
    flow[0, 1, 1, 1] = -1e30
    flow[0, 0, 2, 2] = float("inf")
    from src.models import backwarp
    out = backwarp(x.to(DEV), flow.to(DEV)).cpu()
    assert out[0, :, 0, 0].abs().max() == 0 and out[0, :, 1, 1].abs().max() == 0
    assert torch.isfinite(out[0, :, 3:, 3:]).all()
    assert (out[0, :, 5, 5] - x[0, :, 5, 5]).abs().max() <= 1e-6


# ---- convolutions ---------------------------------------------------------------------------------------------
CONV_CASES = [  # cin, cout, kh, kw, stride, lrelu, H, W
    (3, 32, 7, 7, 1, True, 20, 24), (32, 32, 3, 3, 2, True, 16, 24), (32, 64, 1, 1, 1, True, 9, 11),
    (49, 128, 3, 3, 1, True, 16, 16), (130, 128, 3, 3, 1, True, 8, 12), (32, 2, 7, 7, 1, False, 12, 12),
    (32, 49, 7, 1, 1, False, 10, 14), (49, 49, 1, 7, 1, False, 10, 14), (32, 9, 3, 3, 1, False, 4, 4),
    (128, 192, 3, 3, 2, True, 8, 8), (64, 32, 3, 3, 1, True, 33, 17), (96, 96, 3, 3, 1, True, 5, 5),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_vs_torch(case):
    cin, cout, kh, kw, st, act, H, W = case
    w, b = _rand(cout, cin, kh, kw, seed=1, scale=1.0 / math.sqrt(cin * kh * kw)), _rand(cout, seed=2)
    x = _rand(2, cin, H, W, seed=3)
    ref = F.conv2d(x, w, b, stride=st, padding=(kh // 2, kw // 2))
    ref = O.lrelu(ref) if act else ref
    cw = pack_conv(w.to(DEV), b.to(DEV), st)
    xin = _nhwc(x)
    Ho, Wo = ref.shape[2], ref.shape[3]
    y = torch.zeros(2, Ho, Wo, (cout + 3) & ~3, device=DEV)
    res = _rand(2, cout, Ho, Wo, seed=4) if cout == 2 else None
    ops.conv_simt(ops.view(xin, 0, cin), 2, H, W, cw.w_simt, cw.bias, ops.view(y, 0, cout), kh, kw, st, act,
                  ops.view(_nhwc(res), 0, cout) if res is not None else None)
    if res is not None:
        ref = ref + res
    assert (_nchw(y, cout) - ref).abs().max().item() <= 2e-5


TC_CASES = [  # cin, cout, kh, kw, H, W, lrelu
    (32, 32, 3, 3, 16, 16, True), (64, 64, 3, 3, 8, 24, True), (49, 128, 3, 3, 16, 16, True),
    (130, 128, 3, 3, 8, 16, True), (131, 128, 3, 3, 16, 8, True), (128, 64, 3, 3, 32, 32, True),
    (64, 32, 3, 3, 33, 17, True), (96, 96, 3, 3, 5, 5, True), (128, 128, 3, 3, 64, 64, True),
    (195, 128, 3, 3, 2, 3, True), (128, 96, 3, 3, 4, 4, True),
    (32, 2, 7, 7, 24, 40, False), (32, 2, 5, 5, 9, 9, False), (32, 2, 3, 3, 2, 2, False),   # flow heads (+ residual)
    (32, 49, 7, 1, 20, 12, False), (49, 49, 1, 7, 20, 12, False), (32, 25, 5, 1, 8, 8, False),
    (25, 25, 1, 5, 8, 8, False), (32, 9, 3, 3, 4, 4, False),                                 # conv_dist_R
    (32, 64, 1, 1, 16, 16, True), (32, 128, 1, 1, 16, 24, True), (96, 128, 1, 1, 7, 9, True),  # NetC_ext / moduleFeat
]


@pytest.mark.parametrize("passes,tol", [(3, 1e-4), (2, 1e-4), (4, 1e-4), (5, 1e-4), (1, 4e-3)])
@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_vs_torch(case, passes, tol):
    """tcgen05 implicit-GEMM convolution.  3 passes (3xTF32) is the fp32-equivalent mode: tolerance 1e-4
    absolute on O(1) outputs (the tensor core truncates when it aligns addends into its fp32 accumulator);
    2 passes = TF32 main product + bf16 low-order products (same tolerance); 1 pass is plain TF32: 4e-3."""
    cin, cout, kh, kw, H, W, act = case
    w, b = _rand(cout, cin, kh, kw, seed=1, scale=1.0 / math.sqrt(cin * kh * kw)), _rand(cout, seed=2)
    x = _rand(2, cin, H, W, seed=3)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=(kh // 2, kw // 2)).float()
    ref = O.lrelu(ref) if act else ref
    res = _rand(2, cout, H, W, seed=4) if cout == 2 else None
    if res is not None:
        ref = ref + res
    cw = pack_conv(w.to(DEV), b.to(DEV), 1)
    assert cw.w_hi is not None
    if passes == 5 and cw.w_f16s is None:
        pytest.skip("the single-accumulator fp16 variant is packed for Cout > 64 only")
    xin = _nhwc(x)
    pad = 4 if cout % 4 == 0 else 3
    y = torch.zeros(2, H, W, cout + pad, device=DEV)       # written through a strided view, like Sbuf/Rbuf slices
    ops.conv_tc(ops.view(xin, 0, cin), 2, H, W, cw.w_hi, cw.w_lo, cw.bias, ops.view(y, 0, cout), kh, kw, act, passes,
                ops.view(_nhwc(res), 0, cout) if res is not None else None, cw.pack16(passes))
    err = (_nchw(y, cout) - ref).abs().max().item()
    print(f"conv_tc {case} passes={passes}: max err {err:.2e}")
    assert err <= tol
    assert y[..., cout:].abs().max().item() == 0          # never writes outside its channel slice


@pytest.mark.parametrize("passes,tol", [(3, 1e-4), (2, 1e-4), (4, 1e-4), (1, 4e-3)])
@pytest.mark.parametrize("case", [(32, 32, 16, 24), (32, 64, 64, 64), (64, 96, 17, 31), (96, 128, 8, 8), (32, 32, 256, 128)])
def test_conv_tc_stride2_vs_torch(case, passes, tol):
    """3x3 stride-2 convolutions of NetC (src/models.py:77-106): every tap's box is fetched with TMA element stride 2."""
    cin, cout, H, W = case
    w, b = _rand(cout, cin, 3, 3, seed=1, scale=1.0 / math.sqrt(cin * 9)), _rand(cout, seed=2)
    x = _rand(2, cin, H, W, seed=3)
    ref = O.lrelu(F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1).float())
    cw = pack_conv(w.to(DEV), b.to(DEV), 2)
    assert cw.w_hi is not None
    Ho, Wo = ref.shape[2], ref.shape[3]
    y = torch.zeros(2, Ho, Wo, cout, device=DEV)
    ops.conv_tc(ops.view(_nhwc(x), 0, cin), 2, H, W, cw.w_hi, cw.w_lo, cw.bias, ops.view(y), 3, 3, True, passes, None,
                cw.pack16(passes), stride=2)
    assert (_nchw(y, cout) - ref).abs().max().item() <= tol


@pytest.mark.parametrize("passes,tol", [(3, 1e-4), (2, 1e-4), (4, 1e-4), (1, 2e-2)])
@pytest.mark.parametrize("hw", [(16, 16), (40, 56), (8, 8)])
def test_conv_stem_tc_vs_torch(hw, passes, tol):
    """NetC.conv1 (7x7, 3 -> 32) through the overlapping-window tensor map on the zero-bordered image."""
    from pivlfn.model import pack_stem
    H, W = hw
    w, b = _rand(32, 3, 7, 7, seed=1, scale=6.0 / math.sqrt(147)), _rand(32, seed=2)
    a, c = torch.rand(2, 3, H, W, generator=torch.Generator().manual_seed(5)), torch.rand(2, 3, H, W, generator=torch.Generator().manual_seed(6))
    mean = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6)
    xs = torch.cat([a - torch.tensor(mean[:3]).view(1, 3, 1, 1), c - torch.tensor(mean[3:]).view(1, 3, 1, 1)])
    ref = O.lrelu(F.conv2d(xs.double(), w.double(), b.double(), padding=3).float())
    img = torch.empty(4, H, W, 4, device=DEV)
    img_pad = torch.zeros(4, H, W + 8, 4, device=DEV)
    ops.prep_images(a.to(DEV), c.to(DEV), img, mean, img_pad)
    assert torch.equal(img_pad[:, :, 4:W + 4], img) and img_pad[:, :, :4].abs().max() == 0 and img_pad[:, :, W + 4:].abs().max() == 0
    cw = pack_stem(w.to(DEV), b.to(DEV))
    y = torch.zeros(4, H, W, 32, device=DEV)
    ops.conv_stem_tc(img_pad, 4, H, W, cw.w_hi, cw.w_lo, cw.bias, ops.view(y), True, passes, cw.pack16(passes))
    assert (_nchw(y, 32) - ref).abs().max().item() <= tol


# ---- glue ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [2, 49])
def test_deconv_vs_torch(C):
    x, w = _rand(2, C, 7, 9, seed=1), _rand(C, 1, 4, 4, seed=2)
    ref = F.conv_transpose2d(x, w, None, stride=2, padding=1, groups=C)
    y = torch.zeros(2, 14, 18, (C + 3) & ~3, device=DEV)
    ops.deconv4x4s2_dw(ops.view(_nhwc(x), 0, C), 2, 7, 9, w.reshape(C, 16).contiguous().to(DEV), ops.view(y, 0, C))
    assert (_nchw(y, C) - ref).abs().max().item() <= 1e-5


def test_prep_and_pyramid_vs_torch():
    a, b = torch.rand(2, 3, 32, 64, generator=torch.Generator().manual_seed(1)), torch.rand(2, 3, 32, 64, generator=torch.Generator().manual_seed(2))
    mean = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6)
    ad, bd = a.to(DEV), b.to(DEV)
    out = torch.empty(4, 32, 64, 4, device=DEV)
    ops.prep_images(ad, bd, out, mean)
    ra = a - torch.tensor(mean[:3]).view(1, 3, 1, 1)
    rb = b - torch.tensor(mean[3:]).view(1, 3, 1, 1)
    assert (ad.cpu() - ra).abs().max() <= 1e-7 and (bd.cpu() - rb).abs().max() <= 1e-7      # in-place mutation
    assert (_nchw(out[:2], 3) - ra).abs().max() <= 1e-7 and (_nchw(out[2:], 3) - rb).abs().max() <= 1e-7
    assert out[..., 3].abs().max() == 0
    half = torch.empty(4, 16, 32, 4, device=DEV)
    ops.avgpool2(out, half)
    ref = F.interpolate(ra, size=(16, 32), mode="bilinear", align_corners=False)
    assert (_nchw(half[:2], 3) - ref).abs().max() <= 1e-6


@pytest.mark.parametrize("K,H,W", [(7, 16, 20), (5, 9, 9), (3, 4, 6), (3, 1, 1)])
def test_reg_tail_vs_oracle(K, H, W):
    KK = K * K
    dist = _rand(2, KK, H, W, seed=1, scale=1.5)
    flow = _rand(2, 2, H, W, seed=2, scale=3.0)
    wx, bx, wy, by = _rand(1, KK, 1, 1, seed=3), _rand(1, seed=4), _rand(1, KK, 1, 1, seed=5), _rand(1, seed=6)
    negsq = dist.pow(2.0).neg()
    d = (negsq - negsq.max(1, True)[0]).exp()
    div = d.sum(1, True).reciprocal()
    ux = F.unfold(flow[:, 0:1], kernel_size=K, padding=K // 2).view_as(d)
    uy = F.unfold(flow[:, 1:2], kernel_size=K, padding=K // 2).view_as(d)
    ref = torch.cat([F.conv2d(d * ux, wx, bx) * div, F.conv2d(d * uy, wy, by) * div], 1)
    out = torch.empty(2, H, W, 2, device=DEV)
    out_nchw = torch.empty(2, 2, H, W, device=DEV)
    ops.reg_tail(ops.view(_nhwc(dist), 0, KK), flow.permute(0, 2, 3, 1).contiguous().to(DEV), wx.reshape(-1).to(DEV),
                 bx.to(DEV), wy.reshape(-1).to(DEV), by.to(DEV), out, out_nchw, 5.0, K)
    assert (_nchw(out, 2) - ref).abs().max().item() <= 2e-5
    assert (out_nchw.cpu() - 5.0 * ref).abs().max().item() <= 1e-4


def test_reg_input_vs_oracle():
    H, W = 12, 20
    i1, i2 = torch.rand(2, 3, H, W, generator=torch.Generator().manual_seed(1)), torch.rand(2, 3, H, W, generator=torch.Generator().manual_seed(2))
    flow = _rand(2, 2, H, W, seed=3, scale=2.0)
    scale = 2.5
    rm = flow - flow.view(2, 2, -1).mean(2, True).view(2, 2, 1, 1)
    err = (i1 - O.backwarp(i2, flow * scale)).pow(2.0).sum(1, True).sqrt()
    fl = flow.permute(0, 2, 3, 1).contiguous().to(DEV)
    part = torch.empty(2, ops.flow_mean_parts(), 2, device=DEV)
    ops.flow_mean(fl, part)
    out = torch.zeros(2, H, W, 8, device=DEV)
    ops.reg_input(_nhwc(i1), _nhwc(i2), fl, scale, part, ops.view(out, 4, 3))
    got = out.cpu()
    assert (got[..., 4] - err[:, 0]).abs().max() <= 2e-6
    assert (got[..., 5:7].permute(0, 3, 1, 2) - rm).abs().max() <= 2e-6
    assert got[..., :4].abs().max() == 0 and got[..., 7].abs().max() == 0


@pytest.mark.parametrize("size", [((50, 70), (64, 96)), ((64, 96), (50, 70)), ((436, 1024), (448, 1024)), ((32, 32), (32, 32))])
def test_resize_vs_torch(size):
    (H, W), (Ho, Wo) = size
    x = _rand(2, 2, H, W, seed=1)
    ref = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=False)
    ref[:, 0] *= 1.5
    ref[:, 1] *= 0.25
    out = ops.resize_bilinear(x.to(DEV), Ho, Wo, 1.5, 0.25).cpu()
    assert (out - ref).abs().max().item() <= 1e-5


def test_copy_nhwc_slices():
    src = _rand(1, 3, 5, 8, seed=1).to(DEV).contiguous()
    dst = torch.zeros(1, 3, 5, 12, device=DEV)
    ops.copy(ops.view(src, 0, 8), ops.view(dst, 4, 8), 15)
    assert torch.equal(dst[..., 4:], src) and dst[..., :4].abs().max() == 0
    dst2 = torch.zeros(1, 3, 5, 12, device=DEV)
    ops.copy(ops.view(src, 2, 2), ops.view(dst2, 6, 2), 15)
    assert torch.equal(dst2[..., 6:8], src[..., 2:4])


@pytest.mark.parametrize("passes,tol", [(3, 1e-4), (2, 1e-4), (4, 1e-4), (1, 4e-3)])
@pytest.mark.parametrize("K,cin,H,W", [(7, 32, 24, 40), (5, 32, 16, 16), (7, 32, 64, 8)])
def test_flow_head_pairs_plus_gather_vs_torch(K, cin, H, W, passes, tol):
    """Flow head restated as a 1x1 convolution to 2*K*K channels (tap planes) + gather-sum: equals the KxK convolution."""
    w, b = _rand(2, cin, K, K, seed=1, scale=1.0 / math.sqrt(cin * K * K)), _rand(2, seed=2)
    x = _rand(2, cin, H, W, seed=3)
    res = _rand(2, 2, H, W, seed=4)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=K // 2).float() + res
    w2 = w.permute(2, 3, 0, 1).reshape(-1, cin, 1, 1)
    pk = pack_conv(w2.to(DEV), None, 1)
    planes = torch.full((K * K, 2 * H * W, 2), float("nan"), device=DEV)
    ops.conv1x1_pairs_tc(ops.view(_nhwc(x), 0, cin), 2, H, W, pk.w_hi, pk.w_lo, pk.pack16(passes), planes,
                         K * K, passes)
    out = torch.zeros(2, H, W, 2, device=DEV)
    ops.flow_head_sum(planes, K, b.to(DEV), ops.view(_nhwc(res), 0, 2), ops.view(out), 2, H, W)
    assert (_nchw(out, 2) - ref).abs().max().item() <= tol


@pytest.mark.parametrize("K,H,W", [(7, 24, 40), (7, 8, 8), (5, 16, 16), (5, 37, 70), (3, 4, 4), (3, 9, 33), (7, 64, 96)])
def test_flow_head_simt_vs_torch(K, H, W):
    """Exact-fp32 CUDA-core flow head (KxK, 32 -> 2, + bias + residual flow) against torch in float64: fp32 round-off only."""
    w, b = _rand(2, 32, K, K, seed=1, scale=1.0 / math.sqrt(32 * K * K)), _rand(2, seed=2)
    x = _rand(2, 32, H, W, seed=3)
    res = _rand(2, 2, H, W, seed=4)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=K // 2).float() + res
    wh = w.permute(2, 3, 1, 0).reshape(K * K, 32, 2).contiguous().to(DEV)
    xin = torch.zeros(2, H, W, 36, device=DEV)                 # read through a strided view
    xin[..., :32] = _nhwc(x)
    out = torch.full((2, H, W, 4), float("nan"), device=DEV)   # written through a strided view
    ops.flow_head(ops.view(xin, 0, 32), 2, H, W, wh, b.to(DEV), ops.view(_nhwc(res), 0, 2), ops.view(out, 0, 2), K)
    assert (_nchw(out, 2) - ref).abs().max().item() <= 2e-5
    assert torch.isnan(out[..., 2:]).all()
    out2 = torch.zeros(2, H, W, 2, device=DEV)
    ops.flow_head(ops.view(xin, 0, 32), 2, H, W, wh, None, None, ops.view(out2), K)
    ref2 = F.conv2d(x.double(), w.double(), None, padding=K // 2).float()
    assert (_nchw(out2, 2) - ref2).abs().max().item() <= 2e-5


@pytest.mark.parametrize("case", [(32, 32, 16, 24), (32, 64, 64, 64), (64, 96, 34, 62), (96, 128, 16, 16), (32, 32, 256, 128)])
def test_conv_s2_halo_vs_torch(case):
    """3x3 stride-2 convolutions of NetC restated over the four input parities (TMA element strides do the space-to-depth),
    fp16 split-operand modes 4 (Cout <= 64) and 5 (Cout > 64): fp32-equivalent tolerance."""
    cin, cout, H, W = case
    w, b = _rand(cout, cin, 3, 3, seed=1, scale=1.0 / math.sqrt(cin * 9)), _rand(cout, seed=2)
    x = _rand(2, cin, H, W, seed=3)
    ref = O.lrelu(F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1).float())
    cw = pack_conv(w.to(DEV), b.to(DEV), 2)
    assert cw.w_s2 is not None and cw.s2_passes == (5 if cout > 64 else 4)
    Ho, Wo = ref.shape[2], ref.shape[3]
    y = torch.zeros(2, Ho, Wo, cout, device=DEV)
    ops.conv_s2_tc(ops.view(_nhwc(x), 0, cin), 2, H, W, cw.w_s2, cw.bias, ops.view(y), True, cw.s2_passes)
    err = (_nchw(y, cout) - ref).abs().max().item()
    print(f"conv_s2 {case}: max err {err:.2e}")
    assert err <= 1e-4
