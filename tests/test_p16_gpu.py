"""Operator-level parity of the P16 pipeline on the B200 (csrc/p16.cuh, conv_p16.cu, p16_ops.cu, the P16 variants of
corr.cu and of the stem): every kernel against the torch op it replaces, on the same seeded inputs, through the C ABI.
P16 stores x as (hi = f16(x), lo8 = e5m2((x - hi) * 2^11), hi8 = e5m2(x)): x = hi + 2^-11 lo8 up to 2^-14 |x|.  The convolution
computes a_hi * W_hi (f16) + [lo8 | hi8] * [W 2^-11 ; W - W_hi] (e5m2): the tests pin it to an fp64 emulation of exactly these three
products at fp32-accumulation tolerance (a layout or descriptor bug cannot hide inside a loose bound), and bound the emulation
against the exact convolution separately."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import lfn_oracle as O
from pivlfn import ops
from pivlfn.model import P16_MODE, pack_conv, pack_stem

pytestmark = pytest.mark.gpu
DEV = "cuda"
from p16_host import (E5, STORE_REL, conv_emul, e5m2, p16_parts, p16_ref_decode, p16_ref_encode, p16_ref_fields, p16_round,  # noqa: F401
                      scheme_bound)


def _rand(*shape, seed=0, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _r16(c):
    return (c + 15) & ~15


def to_p16(x_nchw, ld=None, off=0):
    """NCHW cpu fp32 -> P16 device buffer [N,H,W,ld] (zeros elsewhere) holding the tensor at channel word offset off"""
    B, C, H, W = x_nchw.shape
    ld = _r16(C) + off if ld is None else ld
    buf = torch.zeros(B, H, W, ld, device=DEV)
    buf[..., off:off + _r16(C)] = p16_ref_encode(x_nchw.permute(0, 2, 3, 1).contiguous()).to(DEV)
    return buf


def from_p16(buf, C, off=0):
    return p16_ref_decode(buf[..., off:off + _r16(C)].cpu(), C).permute(0, 3, 1, 2).contiguous()


def test_encode_decode_kernels_match_host_restatement():
    x = _rand(2, 5, 7, 49, seed=1, scale=3.0)
    x[0, 0, 0, :4] = torch.tensor([0.0, 1e-7, -65504.0, 6.1e-5])
    xd = x.to(DEV)
    y = torch.zeros(2, 5, 7, 80, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.p16_encode(ops.view(xd), ops.view(y, 16, 64), 2 * 5 * 7, flag)
    ref = p16_ref_encode(x)
    assert torch.equal(y[..., 16:80].cpu().view(torch.int32), ref.view(torch.int32))
    assert y[..., :16].abs().max().item() == 0 and int(flag.item()) == 0
    back = torch.zeros(2, 5, 7, 52, device=DEV)
    ops.p16_decode(ops.view(y, 16, 64), 49, ops.view(back, 0, 49), 2 * 5 * 7)
    assert torch.equal(back[..., :49].cpu(), p16_ref_decode(ref, 49))
    # 14 significant bits down to |x| ~ 6e-5 (fp16 normals): |x - (hi + lo8 / 2048)| <= 2^-14 |x|
    excess = ((back[..., :49].cpu() - x).abs() - 2.0 ** -14 * x.abs()).max().item()
    assert excess <= 2.0 ** -28
    hi, lo8, hi8 = p16_ref_fields(y[..., 16:80].cpu())
    assert torch.equal(hi8[..., :49], e5m2(x).float())
    # out of range -> flag
    xd[1, 2, 3, 4] = 7.0e4
    ops.p16_encode(ops.view(xd), ops.view(y, 16, 64), 2 * 5 * 7, flag)
    assert int(flag.item()) == 1


CONV_CASES = [  # cin, cout, kh, kw, stride, lrelu, H, W, out_fmt
    (32, 64, 1, 1, 1, True, 16, 24, 0), (32, 128, 1, 1, 1, True, 9, 12, 0), (49, 128, 3, 3, 1, True, 16, 16, 0),
    (130, 128, 3, 3, 1, True, 40, 12, 0), (132, 128, 3, 3, 1, True, 8, 8, 0), (128, 128, 3, 3, 1, True, 34, 16, 0),
    (128, 64, 3, 3, 1, True, 16, 16, 0), (64, 64, 3, 3, 1, True, 33, 20, 0), (64, 32, 3, 3, 1, True, 70, 8, 0),
    (32, 32, 3, 3, 1, True, 16, 16, 0), (32, 49, 7, 1, 1, False, 20, 16, 0), (49, 49, 1, 7, 1, False, 20, 16, 1),
    (32, 25, 5, 1, 1, False, 12, 12, 0), (25, 25, 1, 5, 1, False, 12, 12, 1), (32, 9, 3, 3, 1, False, 8, 8, 1),
    (96, 96, 3, 3, 1, True, 12, 12, 0), (128, 96, 3, 3, 1, True, 12, 12, 0), (96, 64, 3, 3, 1, True, 12, 12, 0),
    (386, 128, 3, 3, 1, True, 8, 8, 0), (258, 128, 3, 3, 1, True, 4, 4, 0), (64, 32, 3, 3, 1, True, 2, 2, 0),
    (32, 32, 3, 3, 2, True, 32, 48, 0), (32, 64, 3, 3, 2, True, 16, 16, 0), (64, 96, 3, 3, 2, True, 16, 24, 0),
    (96, 128, 3, 3, 2, True, 8, 8, 0), (128, 96, 3, 3, 2, True, 8, 8, 0), (128, 96, 3, 3, 2, True, 4, 4, 0),
    (32, 64, 1, 1, 1, True, 16, 24, 1), (64, 64, 3, 3, 1, True, 6, 6, 0), (32, 14, 1, 7, 1, False, 24, 16, 2),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=str)
def test_conv_p16_vs_torch(case):
    cin, cout, kh, kw, st, act, H, W, fmt = case
    w, b = _rand(cout, cin, kh, kw, seed=1, scale=1.0 / math.sqrt(cin * kh * kw)), _rand(cout, seed=2)
    x = _rand(2, cin, H, W, seed=3)
    pad = (kh // 2, kw // 2)
    ref = conv_emul(x, w, b, st, pad)
    exact = F.conv2d(p16_round(x).double(), w.double(), b.double(), stride=st, padding=pad)
    assert ((ref - exact).abs() <= scheme_bound(x, w, st, pad) + 1e-12).all()
    assert (ref - exact).abs().max().item() <= 2.0 ** -12 * exact.abs().max().item()
    ref = torch.where(ref >= 0, ref, 0.1 * ref) if act else ref
    cw = pack_conv(w, b, st).to_(DEV)
    w_img, mode = cw.w_f8, P16_MODE
    xin = to_p16(x, ld=_r16(cin) + 16, off=16)                      # a slice of a wider buffer
    Ho, Wo = ref.shape[2], ref.shape[3]
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    if fmt == 0:
        y = torch.full((2, Ho, Wo, _r16(cout) + 32), 7.0, device=DEV)
        ops.conv_p16(ops.view(xin, 16, _r16(cin)), 2, H, W, cin, w_img, mode, cw.bias, ops.view(y, 16, _r16(cout)), cout, kh, kw, st,
                     act, ops.OUT_P16, 0, flag)
        out = from_p16(y, cout, off=16)
        # neighbours untouched, pad channels of the last group exactly zero
        assert (y[..., :16] == 7.0).all() and (y[..., 16 + _r16(cout):] == 7.0).all()
        if cout % 16:
            pad = p16_ref_decode(y[..., 16:16 + _r16(cout)].cpu(), _r16(cout))[..., cout:]
            assert pad.abs().max().item() == 0
    elif fmt == 1:
        ld = (cout + 3) & ~3
        y = torch.zeros(2, Ho, Wo, ld, device=DEV)
        ops.conv_p16(ops.view(xin, 16, _r16(cin)), 2, H, W, cin, w_img, mode, cw.bias, ops.view(y), cout, kh, kw, st, act,
                     ops.OUT_F32, 0, flag)
        out = y[..., :cout].permute(0, 3, 1, 2).cpu()
    else:
        npl = cout // 2
        planes = torch.zeros(npl, 2 * Ho * Wo, 2, device=DEV)
        ops.conv_p16(ops.view(xin, 16, _r16(cin)), 2, H, W, cin, w_img, mode, cw.bias, ops.view(planes.view(1, npl, 2 * Ho * Wo, 2)),
                     cout, kh, kw, st, act, ops.OUT_PLANES, 2 * 2 * Ho * Wo, flag)
        out = planes.view(npl, 2, Ho, Wo, 2).permute(1, 0, 4, 2, 3).reshape(2, cout, Ho, Wo).cpu()
    assert int(flag.item()) == 0
    err = (out.double() - ref).abs().max().item()
    # fp32 accumulation of K = cin*kh*kw products of O(1/sqrt(K)) on the tensor cores (truncating adder): grows like sqrt(K)
    tol = 1e-5 + 1.2e-6 * math.sqrt(cin * kh * kw) + (STORE_REL * ref.abs().max().item() if fmt == 0 else 0.0)
    assert err <= tol, err


def test_conv_p16_range_flag():
    w, b = _rand(32, 32, 3, 3, seed=1), _rand(32, seed=2)
    x = _rand(1, 32, 16, 16, seed=3, scale=3000.0)                   # outputs ~ 3000 * sqrt(288) ~ 5e4 .. 2e5
    cw = pack_conv(w, b, 1).to_(DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    y = torch.zeros(1, 16, 16, 32, device=DEV)
    ops.conv_p16(ops.view(to_p16(x)), 1, 16, 16, 32, cw.w_f8, P16_MODE, cw.bias, ops.view(y), 32, 3, 3, 1, True, ops.OUT_P16, 0, flag)
    assert int(flag.item()) == 1


def test_conv_stem_p16_vs_torch():
    w, b = _rand(32, 3, 7, 7, seed=1, scale=1.0 / math.sqrt(147)), _rand(32, seed=2)
    x = _rand(2, 3, 24, 40, seed=3)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=3)
    ref = torch.where(ref >= 0, ref, 0.1 * ref)
    cw = pack_stem(w, b).to_(DEV)
    img_pad = torch.zeros(2, 24, 40 + 8, 4, device=DEV)
    img_pad[:, :, 4:44, :3] = x.permute(0, 2, 3, 1).to(DEV)
    y = torch.zeros(2, 24, 40, 32, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.conv_stem_p16(img_pad, 2, 24, 40, cw.w_f16, cw.bias, ops.view(y), True, flag)
    out = from_p16(y, 32)
    assert (out.double() - ref).abs().max().item() <= 3e-5 + STORE_REL * ref.abs().max().item() and int(flag.item()) == 0


@pytest.mark.parametrize("in_p16", [False, True])
@pytest.mark.parametrize("C,H,W", [(64, 16, 24), (96, 9, 7), (192, 2, 3)])
def test_warp_p16_vs_oracle(in_p16, C, H, W):
    x = _rand(2, C, H, W, seed=5)
    flow = _rand(2, 2, H, W, seed=6, scale=2.5)
    scale = 1.25
    src = p16_round(x) if in_p16 else x
    ref = O.backwarp(src, flow * scale)
    xin = to_p16(x) if in_p16 else x.permute(0, 2, 3, 1).contiguous().to(DEV)
    y = torch.zeros(2, H, W, 2 * C + 16, device=DEV)
    fl = flow.permute(0, 2, 3, 1).contiguous().to(DEV)
    ops.warp_p16(ops.view(xin), in_p16, fl, scale, ops.view(y, C, C), 2, H, W, C)
    out = from_p16(y, C, off=C)
    assert (out - ref).abs().max().item() <= 1e-5 + STORE_REL * ref.abs().max().item()
    assert y[..., :C].abs().max().item() == 0 and y[..., 2 * C:].abs().max().item() == 0


@pytest.mark.parametrize("H,W", [(8, 12), (5, 3)])
def test_deconv_p16_vs_torch(H, W):
    x = _rand(2, 49, H, W, seed=7)
    w = _rand(49, 1, 4, 4, seed=8)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, groups=49)
    xin = torch.zeros(2, H, W, 52, device=DEV)
    xin[..., :49] = x.permute(0, 2, 3, 1).to(DEV)
    y = torch.full((2, 2 * H, 2 * W, 64), 3.0, device=DEV)
    ops.deconv4x4s2_dw_p16(ops.view(xin, 0, 49), 2, H, W, 49, w.reshape(49, 16).contiguous().to(DEV), ops.view(y))
    out = from_p16(y, 64)
    assert (out[:, :49] - ref).abs().max().item() <= 1e-5 + STORE_REL * ref.abs().max().item() and out[:, 49:].abs().max().item() == 0


def test_reg_input_p16_matches_fp32_kernel():
    B, H, W = 2, 12, 20
    img = torch.zeros(2 * B, H, W, 4, device=DEV)
    img[..., :3] = _rand(2 * B, H, W, 3, seed=9).to(DEV)
    flow = _rand(B, H, W, 2, seed=10, scale=2.0).to(DEV)
    partial = torch.zeros(B, ops.flow_mean_parts(), 2, device=DEV)
    ops.flow_mean(flow, partial)
    ref = torch.zeros(B, H, W, 4, device=DEV)
    ops.reg_input(img[:B], img[B:], flow, 2.5, partial, ops.view(ref, 0, 3))
    y = torch.zeros(B, H, W, 144, device=DEV)
    ops.reg_input_p16(img[:B], img[B:], flow, 2.5, partial, ops.view(y, 128, 16))
    out = p16_ref_decode(y[..., 128:144].cpu(), 16)
    r = ref[..., :3].cpu()
    assert (out[..., :3] - r).abs().max().item() <= STORE_REL * r.abs().max().item() + 1e-9
    assert out[..., 3:].abs().max().item() == 0 and y[..., :128].abs().max().item() == 0


@pytest.mark.parametrize("f2_p16,out_p16", [(False, False), (True, False), (True, True), (False, True)])
@pytest.mark.parametrize("C,s,H,W", [(64, 2, 32, 48), (96, 1, 16, 24), (192, 1, 2, 3)])
def test_corr_p16_variants(C, s, H, W, f2_p16, out_p16):
    f1, f2 = _rand(2, C, H, W, seed=1), _rand(2, C, H, W, seed=2)
    flow = _rand(2, 2, H, W, seed=3, scale=2.0)
    scale = 1.25
    f2src = p16_round(f2) if f2_p16 else f2
    ref = O.lrelu(O.correlation(p16_round(f1), O.backwarp(f2src, flow * scale), s))
    Ho, Wo = -(-H // s), -(-W // s)
    a = to_p16(f1, ld=2 * C + 16)                              # the f1 slice of a Subpixel concat buffer
    b = to_p16(f2) if f2_p16 else f2.permute(0, 2, 3, 1).contiguous().to(DEV)
    fl = flow.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.zeros(2, Ho, Wo, 64 if out_p16 else 52, device=DEV)
    ops.corr_p16(ops.view(a, 0, C), True, ops.view(b), f2_p16, fl, scale, ops.view(out) if out_p16 else ops.view(out, 0, 49),
                 out_p16, 2, H, W, C, s, True)
    got = from_p16(out, 64) if out_p16 else out[..., :49].permute(0, 3, 1, 2).cpu()
    assert (got[:, :49] - ref).abs().max().item() <= 3e-5 + (STORE_REL * ref.abs().max().item() if out_p16 else 0.0)
    if out_p16:
        assert got[:, 49:].abs().max().item() == 0


@pytest.mark.parametrize("cols", [False, True])
@pytest.mark.parametrize("K,H,W", [(7, 24, 16), (5, 9, 12), (3, 4, 4), (7, 2, 2), (7, 70, 9)])
def test_flow_head_rows_vs_torch(K, H, W, cols):
    """The tensor-core flow head: 1xK convolution to 2K row planes + K-row gather-sum (or the transposed Kx1 / column form)
    == the KxK 32 -> 2 convolution."""
    w, b = _rand(2, 32, K, K, seed=1, scale=1.0 / math.sqrt(32 * K * K)), _rand(2, seed=2)
    x = _rand(2, 32, H, W, seed=3)
    res = _rand(2, H, W, 2, seed=4)
    ref = conv_emul(x, w, b, 1, K // 2) + res.permute(0, 3, 1, 2).double()
    if cols:
        rw = pack_conv(w.permute(3, 0, 1, 2).reshape(2 * K, 32, K, 1), None, 1).to_(DEV)
    else:
        rw = pack_conv(w.permute(2, 0, 1, 3).reshape(2 * K, 32, 1, K), None, 1).to_(DEV)
    planes = torch.zeros(K, 2 * H * W, 2, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.conv_p16(ops.view(to_p16(x)), 2, H, W, 32, rw.w_f8, P16_MODE, None, ops.view(planes.view(1, K, 2 * H * W, 2)), 2 * K,
                 K if cols else 1, 1 if cols else K, 1, False, ops.OUT_PLANES, 2 * 2 * H * W, flag)
    out = torch.zeros(2, H, W, 2, device=DEV)
    sb = torch.zeros(2, H, W, 144, device=DEV)
    (ops.head_cols_sum if cols else ops.head_rows_sum)(planes, K, b.to(DEV), res.to(DEV), out, ops.view(sb, 128, 16), 2, H, W, flag)
    got = out.permute(0, 3, 1, 2).cpu().double()
    assert (got - ref).abs().max().item() <= 2e-5 and int(flag.item()) == 0
    slot = p16_ref_decode(sb[..., 128:144].cpu(), 16)
    assert (slot[..., :2] - out.cpu()).abs().max().item() <= 2.0 ** -14 * out.abs().max().item()
    assert slot[..., 2:].abs().max().item() == 0


@pytest.mark.parametrize("src_p16", [False, True])
@pytest.mark.parametrize("cm,H,W", [(64, 40, 24), (96, 16, 16), (192, 4, 4)])
def test_conv_p16_with_fused_backwarp(cm, H, W, src_p16):
    """conv_S.0 (src/models.py:209-217): cat[f1, backwarp(f2, scale * flow), flow] -> 3x3 conv, with the middle cm channels
    gathered inside the kernel instead of being read from memory."""
    cin = 2 * cm + 2
    w, b = _rand(128, cin, 3, 3, seed=1, scale=1.0 / math.sqrt(cin * 9)), _rand(128, seed=2)
    f1, f2 = _rand(2, cm, H, W, seed=3), _rand(2, cm, H, W, seed=4)
    flow = _rand(2, 2, H, W, seed=5, scale=2.0)
    scale = 1.25
    f2src = p16_round(f2) if src_p16 else f2
    f2w = O.backwarp(f2src, flow * scale)
    # (the in-kernel warp encodes its result to P16 exactly like conv_emul encodes its input; f1 and flow come from the buffer)
    ref = conv_emul(torch.cat([f1, f2w, flow], 1), w, b, 1, 1)
    ref = torch.where(ref >= 0, ref, 0.1 * ref)
    cw = pack_conv(w, b, 1).to_(DEV)
    sbuf = torch.zeros(2, H, W, cm + 16, device=DEV)
    sbuf[..., :cm] = p16_ref_encode(f1.permute(0, 2, 3, 1).contiguous()).to(DEV)
    sbuf[..., cm:] = p16_ref_encode(flow.permute(0, 2, 3, 1).contiguous()).to(DEV)
    src = to_p16(f2) if src_p16 else f2.permute(0, 2, 3, 1).contiguous().to(DEV)
    y = torch.zeros(2, H, W, 128, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.conv_p16_warp(ops.view(sbuf), 2, H, W, cin, cw.w_f8, P16_MODE, cw.bias, ops.view(y), 128, 3, 3, True, ops.view(src), src_p16,
                      flow.permute(0, 2, 3, 1).contiguous().to(DEV), scale, cm, cm, flag)
    out = from_p16(y, 128)
    err = (out.double() - ref).abs().max().item()
    # a warped value that differs by one fp32 ulp from the CPU's can round to the neighbouring lo8 / hi8 code: 2^-14 of one product
    assert err <= 1e-5 + 1.2e-6 * math.sqrt(cin * 9) + STORE_REL * ref.abs().max().item() + 1e-4, err
    assert int(flag.item()) == 0


@pytest.mark.parametrize("K,cin,kh,kw,B,H,W", [(7, 49, 1, 7, 2, 40, 24), (7, 49, 1, 7, 1, 13, 9), (5, 25, 1, 5, 3, 9, 12), (3, 32, 3, 3, 2, 4, 4),
                                               (7, 49, 1, 7, 5, 64, 64)])
def test_conv_p16_tail_is_conv_plus_reg_tail(K, cin, kh, kw, B, H, W):
    """The Regularization tail fused into the conv_dist epilogue (OUT_TAIL) against the two launches it replaces
    (conv_p16 with fp32 output + reg_tail): same accumulators, same arithmetic in the same order -> equal to rounding; and against
    the fp64 restatement of src/models.py:279-300 on the same distances."""
    KK = K * K
    x = to_p16(_rand(B, cin, H, W, seed=1))
    w = _rand(KK, cin, kh, kw, seed=2) / math.sqrt(cin * kh * kw)
    cw = pack_conv(w, _rand(KK, seed=3), 1).to_(DEV)
    flow = _rand(B, H, W, 2, seed=4, scale=2.0).to(DEV)
    wx, wy = _rand(KK, seed=5).to(DEV), _rand(KK, seed=6).to(DEV)
    bx, by = _rand(1, seed=7).to(DEV), _rand(1, seed=8).to(DEV)
    dist = torch.zeros(B, H, W, (KK + 3) & ~3, device=DEV)
    ops.conv_p16(ops.view(x), B, H, W, cin, cw.w_f8, P16_MODE, cw.bias, ops.view(dist), KK, kh, kw, 1, False, ops.OUT_F32, 0, None)
    ref_flow, ref_out = torch.empty_like(flow), torch.empty(B, 2, H, W, device=DEV)
    ops.reg_tail(ops.view(dist, 0, KK), flow, wx, bx, wy, by, ref_flow, ref_out, 2.5, K)
    got_flow, got_out = torch.full_like(flow, float("nan")), torch.full((B, 2, H, W), float("nan"), device=DEV)
    ops.conv_p16_tail(ops.view(x), B, H, W, cin, cw.w_f8, cw.bias, kh, kw, K, flow, wx, bx, wy, by, got_flow, got_out, 2.5)
    torch.cuda.synchronize()
    assert not torch.isnan(got_flow).any() and not torch.isnan(got_out).any()          # every pixel was written
    tol = 4e-6 * max(1.0, ref_flow.abs().max().item())
    assert (got_flow - ref_flow).abs().max().item() <= tol
    assert (got_out - ref_out).abs().max().item() <= 2.5 * tol
    # fp64 restatement on the kernel's own distances
    d = dist[..., :KK].double().permute(0, 3, 1, 2)
    neg = -d * d
    e = (neg - neg.max(1, True)[0]).exp()                   # the bias of ScaleX / ScaleY is divided by the sum as well (:288-300)
    fl = flow.double().permute(0, 3, 1, 2)
    un = [F.unfold(fl[:, c:c + 1], K, padding=K // 2).view(B, KK, H, W) for c in range(2)]
    u = (((e * un[0]) * wx.double().view(1, KK, 1, 1)).sum(1) + bx.double()) / e.sum(1)
    v = (((e * un[1]) * wy.double().view(1, KK, 1, 1)).sum(1) + by.double()) / e.sum(1)
    want = torch.stack([u, v], dim=-1)
    assert (got_flow.double() - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())
    got_only = torch.empty_like(flow)
    ops.conv_p16_tail(ops.view(x), B, H, W, cin, cw.w_f8, cw.bias, kh, kw, K, flow, wx, bx, wy, by, got_only, None, 1.0)
    assert torch.equal(got_only, got_flow)
