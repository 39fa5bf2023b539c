"""Host restatement of the P16 activation format (csrc/p16.cuh) and of the arithmetic pivlfn_conv_p16 performs on it
(test infrastructure shared by tests/test_p16_gpu.py and tests/test_p16_host.py)."""
import torch
import torch.nn.functional as F

from pivlfn.model import W8, f8_scale

E5 = torch.float8_e5m2
STORE_REL = 2.0 ** -13          # |x - decode(encode(x))| <= 2^-14 |x|; one lo8 step of slack for a 1-ulp different fp32 input


def e5m2(x):
    """fp32 -> e5m2 (round to nearest even, saturating like cvt.rn.satfinite.e5m2x2.f32)"""
    return x.to(torch.float32).clamp(-57344.0, 57344.0).to(E5)


def p16_parts(x):
    """[..., C] fp32 -> (hi, lo8, hi8) as fp32 tensors holding the stored values: host restatement of csrc/p16.cuh."""
    x = x.to(torch.float32)
    hi = x.to(torch.float16)
    lo8 = e5m2((x - hi.float()) * 2048.0)
    return hi, lo8, e5m2(x)


def p16_ref_encode(x):
    """[..., C] fp32 -> [..., 16 * G] words (viewed as float32): per 16-channel group 32 bytes hi | 16 bytes lo8 | 16 bytes hi8."""
    C = x.shape[-1]
    G = (C + 15) // 16
    xp = F.pad(x, (0, 16 * G - C))
    hi, lo8, hi8 = p16_parts(xp)
    lead = x.shape[:-1]
    g = torch.cat([hi.reshape(*lead, G, 16).contiguous().view(torch.uint8), lo8.reshape(*lead, G, 16).view(torch.uint8),
                   hi8.reshape(*lead, G, 16).view(torch.uint8)], dim=-1)                        # [..., G, 64] bytes
    return g.reshape(*lead, G * 64).contiguous().view(torch.float32)


def p16_ref_fields(wds):
    """words [..., 16 * G] -> (hi, lo8, hi8) fp32 [..., 16 * G]"""
    by = wds.contiguous().view(torch.uint8)
    G = by.shape[-1] // 64
    g = by.reshape(*by.shape[:-1], G, 64)
    hi = g[..., :32].contiguous().view(torch.float16).float()
    lo8 = g[..., 32:48].contiguous().view(E5).float()
    hi8 = g[..., 48:64].contiguous().view(E5).float()
    lead = by.shape[:-1]
    return hi.reshape(*lead, G * 16), lo8.reshape(*lead, G * 16), hi8.reshape(*lead, G * 16)


def p16_ref_decode(wds, C):
    hi, lo8, _ = p16_ref_fields(wds)
    return (hi + lo8 / 2048.0)[..., :C]


def p16_round(x_nchw):
    """what a P16 tensor holds for these fp32 values (NCHW in / out)"""
    x = x_nchw.permute(0, 2, 3, 1).contiguous()
    return p16_ref_decode(p16_ref_encode(x), x.shape[-1]).permute(0, 3, 1, 2).contiguous()


def conv_emul(x_nchw, w, b, stride=1, padding=0):
    """fp64 value of what pivlfn_conv_p16 computes from the P16 encoding of x and the _pack_f8 tiles of w (before the activation)."""
    hi, lo8, hi8 = (t.permute(0, 3, 1, 2).double() for t in p16_parts(x_nchw.permute(0, 2, 3, 1).contiguous()))
    S = f8_scale(w)
    W = w.float() * S
    Wh = W.to(torch.float16).float()
    c_lo, c_hi = (W / 2048.0).to(W8).double(), (W - Wh).to(W8).double()
    kw = dict(stride=stride, padding=padding)
    y = (F.conv2d(hi, Wh.double(), None, **kw) + F.conv2d(lo8, c_lo, None, **kw) + F.conv2d(hi8, c_hi, None, **kw)) / S
    return y if b is None else y + b.double().view(1, -1, 1, 1)


def scheme_bound(x_nchw, w, stride=1, padding=0):
    """bound on |conv_emul - exact conv of the P16-rounded input|: each product carries at most 2^-12 relative error (lo8 and
    W 2^-11 rounded to 3 bits: 2^-14 each, hi8 * W_lo: 2^-13; random signs: the observed error is ~16x smaller);
    sum_k |a_k w_k| per output."""
    return 2.0 ** -12 * F.conv2d(p16_round(x_nchw).abs().double(), w.abs().double(), None, stride=stride, padding=padding)


