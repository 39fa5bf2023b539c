"""CPU checks of the P16 format and of the product scheme it serves (csrc/p16.cuh, pivlfn.model._pack_f8), on the host
restatement the GPU tests pin the kernels to (tests/p16_host.py): storage error bound, byte layout of a group, and the distance
between the three-product arithmetic (f16 main + two e5m2 corrections) and the exact convolution."""
import math

import pytest
import torch
import torch.nn.functional as F

from p16_host import conv_emul, e5m2, p16_ref_decode, p16_ref_encode, p16_ref_fields, p16_round, scheme_bound


def _rand(*shape, seed=0, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_p16_storage_error_and_group_layout():
    x = _rand(3, 5, 49, seed=1) * torch.logspace(-3, 3, 49)
    x[0, 0, :3] = torch.tensor([0.0, -65504.0, 6.1e-5])
    w = p16_ref_encode(x)
    assert w.shape == (3, 5, 64) and w.dtype == torch.float32                      # 49 channels -> 4 groups of 16 words
    back = p16_ref_decode(w, 49)
    assert ((back - x).abs() <= 2.0 ** -14 * x.abs() + 2.0 ** -28).all()
    # byte layout of one group: 16 x f16 | 16 x e5m2 (scaled residual) | 16 x e5m2 (value); pad channels are zero bytes
    by = w.view(torch.uint8).reshape(3, 5, 4, 64)
    g0 = by[1, 2, 0]
    assert torch.equal(g0[:32].view(torch.float16), x[1, 2, :16].half())
    assert torch.equal(g0[32:48].view(torch.float8_e5m2).float(), e5m2((x[1, 2, :16] - x[1, 2, :16].half().float()) * 2048.0).float())
    assert torch.equal(g0[48:].view(torch.float8_e5m2).float(), e5m2(x[1, 2, :16]).float())
    assert by[:, :, 3, 2:32].abs().sum() == 0 and by[:, :, 3, 33:48].abs().sum() == 0 and by[:, :, 3, 49:].abs().sum() == 0
    hi, lo8, hi8 = p16_ref_fields(w)
    assert float(hi8[0, 0, 1]) == -57344.0                                         # saturating, like cvt.rn.satfinite


@pytest.mark.parametrize("cin,cout,k,stride", [(128, 128, 3, 1), (64, 32, 3, 1), (32, 64, 1, 1), (32, 32, 3, 2), (49, 49, 7, 1)])
def test_product_scheme_error_against_the_exact_convolution(cin, cout, k, stride):
    kh, kw = (1, k) if k == 7 else (k, k)
    x = _rand(1, cin, 12, 12, seed=2, scale=2.0)
    w = _rand(cout, cin, kh, kw, seed=3, scale=1.0 / math.sqrt(cin * kh * kw))
    b = _rand(cout, seed=4)
    pad = (kh // 2, kw // 2)
    emu = conv_emul(x, w, b, stride, pad)
    exact = F.conv2d(p16_round(x).double(), w.double(), b.double(), stride=stride, padding=pad)
    err = (emu - exact).abs()
    assert (err <= scheme_bound(x, w, stride, pad) + 1e-12).all()                  # worst case 2^-12 per product
    assert err.max().item() <= 2.0 ** -13 * exact.abs().max().item()               # observed ~2^-16: random signs
    # against single-rounded fp16 operands (what one f16 MMA alone would give): at least 8x closer
    one = F.conv2d(x.half().double(), w.half().double(), b.double(), stride=stride, padding=pad)
    assert err.mean().item() * 8 <= (one - F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)).abs().mean().item()
