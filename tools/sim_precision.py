"""CPU emulation of candidate operand formats for the convolutions: what flow error does a product scheme cost?

Runs the oracle network (oracle/lfn_oracle.py, fp64) with every convolution replaced by an emulation of a split-operand
scheme, on synthetic particle pairs with the synthetic weights, and prints max / mean |flow - fp64 flow| in pixels.
    python tools/sim_precision.py [H] [B]
schemes:
  f16x3   a = hi + lo*2^-11 (both fp16); D = a_hi*w_hi + 2^-11 (a_lo*w_hi + a_hi*w_lo)        (what conv_p16 computes)
  tf32    both operands rounded to 11 significant bits, one product
  f16+e5  a = hi(fp16) + lo8*2^-11 with lo8 in e5m2; corrections a_lo8*w_hi8 + a_hi8*w_lo8 in e5m2 (an fp8 MMA at twice the rate)
  f16+e4  the same with e4m3 corrections
This is analysis tooling (it imports oracle/): not part of the product."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from oracle import lfn_oracle as O  # noqa: E402
from pivlfn import synth  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.set_num_threads(16)


def f16(x):
    return x.to(torch.float32).to(torch.float16).to(torch.float64)


def q8(x, dt):
    return x.to(torch.float32).to(dt).to(torch.float64)


def tf32(x):
    x32 = x.to(torch.float32)
    i = x32.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF           # round to 10 explicit mantissa bits
    return i.view(torch.float32).to(torch.float64)


def make_conv(scheme):
    def conv(x, sd, prefix, stride=1, padding=0, act=True):
        w, b = sd[prefix + ".weight"], sd.get(prefix + ".bias")
        kw = dict(stride=stride, padding=padding)
        if scheme == "fp64":
            y = F.conv2d(x, w, None, **kw)
        elif scheme == "tf32":
            y = F.conv2d(tf32(x), tf32(w), None, **kw)
        else:
            xh, wh = f16(x), f16(w)
            xl, wl = (x - xh) * 2048.0, (w - wh) * 2048.0
            if scheme == "f16x3":
                xl, wl = f16(xl), f16(wl)
                corr = F.conv2d(xl, wh, None, **kw) + F.conv2d(xh, wl, None, **kw)
            elif scheme == "f16+mix":
                sw = 2.0 ** math.floor(math.log2(16384.0 / float(w.abs().max())))          # per-layer weight scale (f8_scale)
                e5, e4 = torch.float8_e5m2, torch.float8_e4m3fn
                corr = (F.conv2d(q8(xl, e5), q8(wh * sw / 2048.0, e4), None, **kw) * 2048.0
                        + F.conv2d(q8(xh, e5), q8((w * sw - f16(w * sw)), e4), None, **kw) * 2048.0) / sw
            else:
                dt = torch.float8_e5m2 if scheme == "f16+e5" else torch.float8_e4m3fn
                s = 1.0 if scheme == "f16+e5" else 64.0          # e4m3: keep small weights out of the subnormals
                corr = (F.conv2d(q8(xl, dt), q8(wh * s, dt), None, **kw) + F.conv2d(q8(xh, dt), q8(wl * s, dt), None, **kw)) / s
            y = F.conv2d(xh, wh, None, **kw) + corr / 2048.0
        if b is not None:
            y = y + b.view(1, -1, 1, 1)
        y = O.lrelu(y) if act else y
        if scheme in ("f16+e5", "f16+e4", "f16+mix") and act:
            # the stored activation is hi + lo8 * 2^-11
            dt = torch.float8_e4m3fn if scheme == "f16+e4" else torch.float8_e5m2
            yh = f16(y)
            y = yh + q8((y - yh) * 2048.0, dt) / 2048.0
        elif scheme == "f16x3" and act:
            yh = f16(y)
            y = yh + f16((y - yh) * 2048.0) / 2048.0
        return y
    return conv


def main():
    sd = {k: v.double() for k, v in synth.synthetic_state_dict("piv", 0).items()}
    a, b, _ = synth.particle_batch(B, H, H, seed0=0)
    a, b = a.double(), b.double()
    orig = O._conv
    outs = {}
    for scheme in ("fp64", "f16x3", "f16+e5", "f16+mix", "f16+e4", "tf32"):
        O._conv = make_conv(scheme)
        with torch.no_grad():
            outs[scheme] = O.forward(sd, a.clone(), b.clone(), "piv")
        O._conv = orig
        if scheme != "fp64":
            d = (outs[scheme] - outs["fp64"]).abs()
            print(f"{scheme:8s} max |dflow| {d.max().item():.3e} px   mean {d.mean().item():.3e} px   (flow max {outs['fp64'].abs().max().item():.2f})", flush=True)


if __name__ == "__main__":
    main()
