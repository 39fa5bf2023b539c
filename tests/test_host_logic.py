"""Host-side logic on the CPU: architecture tables vs the reference's state_dict contract, weight packing,
drop-in API surface and error behaviour, sharding arithmetic, synthetic input generator."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pivlfn import arch, shard, synth
from pivlfn.model import pack_conv, tf32_round


@pytest.mark.parametrize("name", ["piv", "hui", "piv2", "hui2"])
def test_param_specs_match_reference_state_dict(golden_dir, name):
    ref = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))[name]
    mine = [[k, list(v)] for k, v in arch.param_specs(arch.CFGS[name]).items()]
    assert mine == ref          # same keys, same shapes, same ORDER (convert.py relies on the order)


def test_conv_flops_constants():
    # SURVEY.md section 8(a) / BASELINE.md section 2
    assert arch.conv_flops_per_pixel(arch.CFGS["piv"]) == pytest.approx(2390317.64, rel=1e-6)
    assert arch.conv_flops_per_pixel(arch.CFGS["hui"]) == pytest.approx(650947.64, rel=1e-6)


def test_dropin_modules_have_reference_keys_and_load_strictly(golden_dir):
    from src.models import LiteFlowNet, LiteFlowNet2, hui_liteflownet, piv_liteflownet
    ref = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    for name, fn in (("piv", lambda p: piv_liteflownet(p, 1)), ("hui", lambda p: hui_liteflownet(p, 1)),
                     ("piv2", lambda p: piv_liteflownet(p, 2)), ("hui2", lambda p: hui_liteflownet(p, 2))):
        net = fn(None)
        assert [[k, list(v.shape)] for k, v in net.state_dict().items()] == ref[name]
        sd = synth.synthetic_state_dict(name, 0)
        net = fn(sd)
        assert torch.equal(net.state_dict()["NetC.conv1.0.weight"], sd["NetC.conv1.0.weight"])
        bad = dict(sd)
        bad.pop("NetC.conv1.0.bias")
        with pytest.raises(RuntimeError):
            fn(bad)
    with pytest.raises(ValueError):
        piv_liteflownet(None, 3)
    with pytest.raises(ValueError):
        hui_liteflownet(None, 0)
    assert LiteFlowNet().lowest_level == 2 and LiteFlowNet2().lowest_level == 3
    assert LiteFlowNet().SCALEFACTOR[1] == 20.0 and piv_liteflownet().SCALEFACTOR[1] == 5.0


def test_no_cpu_fallback():
    """CPU tensors are refused like the reference's correlation op (src/correlation.py:339-340)."""
    from src.correlation import FunctionCorrelation, ModuleCorrelation
    from src.models import piv_liteflownet
    a = torch.zeros(1, 4, 8, 8)
    with pytest.raises(NotImplementedError):
        FunctionCorrelation(tensorFirst=a, tensorSecond=a, intStride=1)
    with pytest.raises(NotImplementedError):
        ModuleCorrelation()(a, a, 2)
    with pytest.raises(AssertionError):
        FunctionCorrelation(a.permute(0, 1, 3, 2), a, 1)      # non-contiguous (src/correlation.py:297-298)
    with pytest.raises(NotImplementedError):
        piv_liteflownet()(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32))


def test_tf32_round():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -3.14159265, 1e-30, 0.0])
    r = tf32_round(x)
    assert torch.equal(r.view(torch.int32) & 0x1FFF, torch.zeros(6, dtype=torch.int32))
    assert r[1].item() == 1.0 + 2 ** -10        # tie rounds away from zero (rna)
    assert ((r - x).abs() <= x.abs() * 2 ** -11 + 1e-45).all()
    lo = tf32_round(x - r)
    assert ((x - r - lo).abs() <= x.abs() * 2 ** -21 + 1e-45).all()


def _conv_from_simt_pack(x_nhwc, cw):
    """Evaluate the packed [KH*KW*Cin, CoutP] layout the way the kernel indexes it."""
    N, H, W, C = x_nhwc.shape
    ph, pw = cw.kh // 2, cw.kw // 2
    xp = F.pad(x_nhwc.permute(0, 3, 1, 2), (pw, pw, ph, ph)).permute(0, 2, 3, 1)
    Ho, Wo = (H + 2 * ph - cw.kh) // cw.stride + 1, (W + 2 * pw - cw.kw) // cw.stride + 1
    y = torch.zeros(N, Ho, Wo, cw.w_simt.shape[1])
    for ky in range(cw.kh):
        for kx in range(cw.kw):
            tap = ky * cw.kw + kx
            win = xp[:, ky:ky + (Ho - 1) * cw.stride + 1:cw.stride, kx:kx + (Wo - 1) * cw.stride + 1:cw.stride]
            y += win @ cw.w_simt[tap * C:(tap + 1) * C]
    return y[..., :cw.cout] + (cw.bias if cw.bias is not None else 0.0)


@pytest.mark.parametrize("cin,cout,kh,kw,stride", [(3, 32, 7, 7, 1), (32, 32, 3, 3, 2), (32, 49, 7, 1, 1),
                                                    (49, 49, 1, 7, 1), (32, 2, 5, 5, 1), (131, 128, 3, 3, 1)])
def test_pack_conv_layouts(cin, cout, kh, kw, stride):
    g = torch.Generator().manual_seed(cin + cout)
    w = torch.randn(cout, cin, kh, kw, generator=g)
    b = torch.randn(cout, generator=g)
    x = torch.randn(1, cin, 9, 10, generator=g)
    ref = F.conv2d(x, w, b, stride=stride, padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
    cw = pack_conv(w, b, stride)
    assert (_conv_from_simt_pack(x.permute(0, 2, 3, 1), cw) - ref).abs().max() < 1e-3
    if cw.w_hi is not None:
        coutp, cinp = cw.w_hi.shape[0], cw.w_hi.shape[2]
        assert cinp % 32 == 0 and coutp % 16 == 0 and cw.w_hi.shape == (coutp, kh * kw, cinp)
        wsum = (cw.w_hi + cw.w_lo)[:cout, :, :cin].reshape(cout, kh, kw, cin).permute(0, 3, 1, 2)
        assert (wsum - w).abs().max() <= w.abs().max() * 2 ** -20
        assert cw.w_hi[:, :, cin:].abs().max() == 0 if cinp > cin else True
        assert cw.w_hi[cout:].abs().max() == 0 if coutp > cout else True
    # input-channel permutation + padding (how Rbuf's channel order is absorbed into the weights)
    perm = torch.cat([torch.arange(3, cin), torch.arange(0, 3)]) if cin > 3 else torch.arange(cin)
    cw2 = pack_conv(w, b, stride, cin_pad=1, in_perm=perm)
    xp = torch.cat([x[:, perm], torch.full((1, 1, 9, 10), 7.0)], 1)     # garbage in the pad channel is ignored
    assert (_conv_from_simt_pack(xp.permute(0, 2, 3, 1), cw2) - ref).abs().max() < 1e-3


def test_pack_stem_layout():
    from pivlfn.model import pack_stem
    g = torch.Generator().manual_seed(3)
    w, b = torch.randn(32, 3, 7, 7, generator=g), torch.randn(32, generator=g)
    cw = pack_stem(w, b)
    assert cw.stem and cw.w_hi.shape == (32, 7, 32) and cw.cin == 4
    full = (cw.w_hi + cw.w_lo).reshape(32, 7, 8, 4)
    assert (full[:, :, :7, :3].permute(0, 3, 1, 2) - w).abs().max() <= w.abs().max() * 2 ** -20
    assert full[:, :, 7].abs().max() == 0 and full[..., 3].abs().max() == 0


def test_pair_and_frame_ranges():
    for n in (0, 1, 7, 64, 999):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.pair_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    # run.py -n 1000 => 999 pairs; neighbouring ranks share exactly one boundary frame
    fr = [shard.frame_range(1000, r, 8) for r in range(8)]
    assert fr[0][0] == 0 and fr[-1][1] == 999
    assert all(fr[i][1] == fr[i + 1][0] for i in range(7))
    assert sum(hi - lo for lo, hi in fr) == 999
    with pytest.raises(ValueError):
        shard.pair_range(4, 2, 2)


def test_row_slabs():
    slabs = shard.row_slabs(4096, 8)
    assert slabs[0] == (0, 512) and slabs[-1] == (3584, 4096)
    assert all(a % 32 == 0 and b % 32 == 0 for a, b in slabs)
    assert [b - a for a, b in shard.row_slabs(96, 2)] == [64, 32]
    with pytest.raises(ValueError):
        shard.row_slabs(100, 2)


def test_particle_generator_statistics_and_determinism():
    i1, i2, gt = synth.particle_pair(128, 128, 5, "uniform")
    j1, j2, _ = synth.particle_pair(128, 128, 5, "uniform")
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2)
    assert i1.dtype == np.uint8 and i1.shape == (128, 128) and gt.shape == (128, 128, 2)
    assert np.allclose(gt[..., 0], 2.5) and np.allclose(gt[..., 1], -1.5)
    assert 10 < i1.mean() < 80 and i1.max() > 150            # same ball park as the demo DNS image (mean 43)
    # frame 2 is frame 1 displaced by (2.5, -1.5): integer part visible as a shift of the correlation peak
    a = i1.astype(np.float64) - i1.mean()
    b = i2.astype(np.float64) - i2.mean()
    cc = np.fft.ifft2(np.fft.fft2(b) * np.conj(np.fft.fft2(a))).real
    dy, dx = np.unravel_index(np.argmax(cc), cc.shape)
    dy = dy - 128 if dy > 64 else dy
    dx = dx - 128 if dx > 64 else dx
    assert dx in (2, 3) and dy in (-1, -2)
    t = synth.to_rgb_tensor(i1)
    assert t.shape == (3, 128, 128) and float(t.max()) <= 1.0 and torch.equal(t[0], t[2])


def test_fp16_operand_packs_reconstruct_the_weights():
    """f16c packs (host logic): [f16(w) | f16((w - f16(w)) * 2^11)] and the single-accumulator pack of W = 256 w,
    [f16(W) | f16(W - f16(W)) | f16(f16(W) / 2^11)], reconstruct the weights to ~2^-22 relative; a weight outside the fp16
    range is reported (None) instead of being saturated."""
    from pivlfn.model import _pack_f16, _pack_f16_single
    g = torch.Generator().manual_seed(9)
    w = torch.randn(32, 9, 64, generator=g) * torch.logspace(-4, 1, 64)[None, None, :]       # 5 decades of magnitudes
    p = _pack_f16(w)
    assert p.dtype == torch.float16 and p.shape == (2, 32, 9, 64)
    rec = p[0].double() + p[1].double() / 2048.0
    # absolute floor: the scaled residual bottoms out at the fp16 subnormal spacing, 2^-24 / 2^11
    assert ((rec - w.double()).abs() <= w.abs().double() * 2.0 ** -21 + 2e-11).all()
    s = _pack_f16_single(w)
    assert s.shape == (3, 32, 9, 64)
    rec = (s[0].double() + s[1].double()) / 256.0
    assert ((rec - w.double()).abs() <= w.abs().double() * 2.0 ** -21 + 2.5e-10).all()       # floor 2^-24 / 2^8
    assert torch.equal(s[2], (s[0].float() / 2048.0).half())
    big = w.clone()
    big[0, 0, 0] = 1.0e5
    assert _pack_f16(big) is None and _pack_f16_single(big) is None


def test_stride2_restatement_equals_the_strided_convolution():
    """3x3 stride-2 convolution == 2x2-tap stride-1 convolution over the four input parities with the restated weights
    (host logic of pivlfn_conv_s2_tc; block taps at offsets -1 and 0, zero padding)."""
    from pivlfn.model import _restate_s2
    g = torch.Generator().manual_seed(4)
    w = torch.randn(8, 5, 3, 3, generator=g)
    x = torch.randn(2, 5, 12, 10, generator=g)
    ref = F.conv2d(x, w, None, stride=2, padding=1)
    w2 = _restate_s2(w)                                                   # [8, 4 taps, 4*5]
    # space-to-depth: channel (py*2 + px)*Cin + c holds x[c, 2y + py, 2x + px]
    s2d = torch.cat([x[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], 1)
    wk = w2.reshape(8, 2, 2, 20).permute(0, 3, 1, 2)                      # [cout, 4*cin, by, bx]
    out = F.conv2d(F.pad(s2d, (1, 0, 1, 0)), wk, None)                    # taps at block offsets -1, 0
    assert out.shape == ref.shape and (out - ref).abs().max() <= 1e-5
    assert (w2 == 0).float().mean() > 0.4                                 # 7 of 16 (tap, parity) products are zero


def test_weight_stage_image_is_the_swizzled_shared_memory_layout():
    """pivlfn.model.stage_image: [parts, CoutP, taps, CinP] -> [chunk][tap][part][row][64 B] with the 64B swizzle; checked
    against an independent byte-offset statement of the layout include/pivlfn.h documents, and for invertibility."""
    from pivlfn.model import stage_image
    parts, coutp, ntaps, cinp = 3, 48, 9, 96
    g = torch.Generator().manual_seed(5)
    pack = torch.randn(parts, coutp, ntaps, cinp, generator=g).to(torch.float16)
    img = stage_image(pack)
    assert img.shape == (cinp // 32, ntaps, parts, coutp, 4, 8) and img.is_contiguous()
    flat = img.reshape(-1)
    rng = np.random.default_rng(0)
    for _ in range(500):
        p, r, t, k = (int(rng.integers(n)) for n in (parts, coutp, ntaps, cinp))
        c, kk = divmod(k, 32)
        byte = ((((c * ntaps + t) * parts + p) * coutp + r) * 64) + (((kk // 8) ^ ((r >> 1) & 3)) * 16) + (kk % 8) * 2
        assert flat[byte // 2] == pack[p, r, t, k]
    # a stage of tps taps is a contiguous range of tps * parts * CoutP * 64 bytes
    assert img[1, 3:6].reshape(-1).data_ptr() - img[1, 3].data_ptr() == 0
    assert stage_image(None) is None


def test_f8_weight_image_layout_scale_and_reconstruction():
    """pivlfn.model._pack_f8 (the weight image of pivlfn_conv_p16): per-layer power-of-two scale S with max|w| S in [8192, 16384),
    tile 0 = f16(S w), tile 1 = per 16 channels [e4m3(S w / 2048) x 16 | e4m3(S w - f16(S w)) x 16], both at the byte offsets of
    the swizzled ring image, and a 16-byte trailer [1 / S, S, 0, 0]; W_hi + W_lo reconstructs the weight to 2^-14 relative (the
    correction tile has 4 significant bits), whatever the magnitude of the layer (no fp16 overflow: 1e5 packs fine)."""
    from pivlfn.model import W8, _pack_f8, f8_scale
    g = torch.Generator().manual_seed(11)
    for mag in (1e-3, 1.0, 1e5):
        coutp, ntaps, cinp = 48, 9, 64
        w = torch.randn(coutp, ntaps, cinp, generator=g) * mag * torch.logspace(-2, 0, cinp)[None, None, :]
        img = _pack_f8(w)
        S = f8_scale(w)
        assert math.log2(S) == round(math.log2(S)) and 8192.0 <= float(w.abs().max()) * S < 16384.0
        body = (cinp // 32) * ntaps * 2 * coutp * 64
        assert img.dtype == torch.uint8 and img.numel() == body + 16
        tr = img[body:].view(torch.float32)
        assert float(tr[0]) == 1.0 / S and float(tr[1]) == S and float(tr[2]) == 0.0
        W = w * S
        hi = W.to(torch.float16)
        rng = np.random.default_rng(1)
        for _ in range(300):
            r, t, k = (int(rng.integers(n)) for n in (coutp, ntaps, cinp))
            c, kk = divmod(k, 32)
            row0 = (((c * ntaps + t) * 2 + 0) * coutp + r) * 64            # tile 0: 32 fp16 channels per 64-byte row
            row1 = (((c * ntaps + t) * 2 + 1) * coutp + r) * 64            # tile 1: [lo x 16 | hi x 16] per 16 channels
            sw = (r >> 1) & 3
            b0 = row0 + (((kk // 8) ^ sw) * 16) + (kk % 8) * 2
            assert img[b0:b0 + 2].view(torch.float16)[0] == hi[r, t, k]
            grp, j = divmod(kk, 16)                                         # 32 bytes per group: 16-byte unit 2*grp (+1 for W_lo)
            b_lo = row1 + (((2 * grp) ^ sw) * 16) + j
            b_hi = row1 + (((2 * grp + 1) ^ sw) * 16) + j
            assert img[b_lo:b_lo + 1].view(W8)[0].float() == (W[r, t, k] / 2048.0).to(W8).float()
            assert img[b_hi:b_hi + 1].view(W8)[0].float() == (W[r, t, k] - hi[r, t, k].float()).to(W8).float()
        rec = (hi.double() + (W - hi.float()).to(W8).double()) / S
        # e4m3 residual: 4 significant bits of a term <= 2^-11 |w|; absolute floor = its subnormal step, 2^-23 of the layer's largest weight
        assert ((rec - w.double()).abs() <= w.abs().double() * 2.0 ** -14 + float(w.abs().max()) * 2.0 ** -22).all()
    bad = torch.randn(16, 1, 32, generator=g)
    bad[0, 0, 0] = float("inf")
    assert _pack_f8(bad) is None
