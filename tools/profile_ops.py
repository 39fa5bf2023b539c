"""Level-1 glue operators of the P16 pipeline alone (cost volume with fused backwarp, standalone backwarp, regularisation tail,
depthwise up-convolution, flow-head row sum, conv_S.0 with the fused backwarp) for ncu and for quick timing.
    python tools/profile_ops.py [B] [H]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
from pivlfn import ops  # noqa: E402
from pivlfn.model import pack_conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
cm = 64
flag = torch.zeros(1, dtype=torch.int32, device=dev)
f1 = torch.randn(B, H, H, cm, generator=g).to(dev)
Sbuf = torch.zeros(B, H, H, cm + 16, device=dev)                     # [f1 | flow group], P16
ops.p16_encode(ops.view(f1), ops.view(Sbuf, 0, cm), B * H * H, flag)
f2 = torch.randn(B, H, H, cm, generator=g).to(dev)                   # fp32 NHWC (NetC_ext of the second image)
flow = (2.0 * torch.randn(B, H, H, 2, generator=g)).to(dev)
corr = torch.zeros(B, H // 2, H // 2, 52, device=dev)
dist = torch.randn(B, H, H, 52, generator=g).to(dev)
flow_out = torch.empty_like(flow)
wx, wy = torch.randn(49, generator=g).to(dev), torch.randn(49, generator=g).to(dev)
bx, by = torch.zeros(1, device=dev), torch.zeros(1, device=dev)


def timeit(fn, name, byts, unit="GB/s algorithmic"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.4f} ms, {byts / ms / 1e6:.0f} {unit}")


timeit(lambda: ops.corr_p16(ops.view(Sbuf, 0, cm), True, ops.view(f2), False, flow, 5.0, ops.view(corr, 0, 49), False, B, H, H, cm, 2,
                            True, flag),
       "corr (f1 P16, f2 fp32) s=2 C=64 + backwarp", 4.0 * B * (2 * cm * H * H + 2 * H * H + 49 * (H // 2) ** 2))
scratch = torch.zeros(B, H, H, cm, device=dev)
timeit(lambda: ops.warp_p16(ops.view(f2), False, flow, 5.0, ops.view(scratch), B, H, H, cm, flag), "warp_p16 C=64",
       4.0 * B * H * H * (2 * cm + 2))
timeit(lambda: ops.reg_tail(ops.view(dist, 0, 49), flow, wx, bx, wy, by, flow_out, None, 5.0, 7), "reg_tail K=7",
       4.0 * B * H * H * (49 + 4))
corr_half = torch.randn(B, H // 2, H // 2, 52, generator=g).to(dev)
corr_up = torch.zeros(B, H, H, 64, device=dev)
wdw = torch.randn(49, 16, generator=g).to(dev)
timeit(lambda: ops.deconv4x4s2_dw_p16(ops.view(corr_half, 0, 49), B, H // 2, H // 2, 49, wdw, ops.view(corr_up), flag),
       "deconv4x4s2_dw_p16 C=49 (H/2 -> H)", 4.0 * B * ((H // 2) ** 2 * 52 + H * H * 64))
planes = torch.randn(7, B * H * H, 2, generator=g).to(dev)
timeit(lambda: ops.head_rows_sum(planes, 7, bx.repeat(2), flow, flow_out, ops.view(Sbuf, cm, 16), B, H, H, flag), "head_rows_sum K=7",
       4.0 * B * H * H * (14 + 4 + 16))
# conv_S.0 with the Subpixel backwarp fused in (gather warps): 130 -> 128, 3x3
cin = 2 * cm + 2
w = torch.randn(128, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
cw = pack_conv(w, torch.randn(128, generator=g), 1).to_(dev)
y = torch.empty(B, H, H, 128, device=dev)
timeit(lambda: ops.conv_p16_warp(ops.view(Sbuf), B, H, H, cin, cw.w_f8, 6, cw.bias, ops.view(y), 128, 3, 3, True, ops.view(f2), False, flow,
                                 5.0, cm, cm, flag),
       "conv_S.0 130->128 3x3 + fused backwarp (TFLOP/s in the rate column / 1000)", 2.0 * B * H * H * cin * 128 * 9, unit="GFLOP/s")
# conv_dist_R.1 (1x7, 49 -> 49) + Regularization tail: two launches against the fused epilogue
d0 = torch.zeros(B, H, H, 64, device=dev)
ops.p16_encode(ops.view(torch.randn(B, H, H, 49, generator=g).to(dev)), ops.view(d0), B * H * H, flag)
wd = torch.randn(49, 49, 1, 7, generator=g) / math.sqrt(49 * 7)
cd = pack_conv(wd, torch.randn(49, generator=g), 1).to_(dev)
dist2 = torch.zeros(B, H, H, 52, device=dev)


def two():
    ops.conv_p16(ops.view(d0), B, H, H, 49, cd.w_f8, 6, cd.bias, ops.view(dist2), 49, 1, 7, 1, False, ops.OUT_F32, 0, None)
    ops.reg_tail(ops.view(dist2, 0, 49), flow, wx, bx, wy, by, flow_out, None, 5.0, 7)


timeit(two, "conv_dist_R.1 + reg_tail (2 launches)", 2.0 * B * H * H * 49 * 49 * 7, unit="GFLOP/s")
timeit(lambda: ops.conv_p16_tail(ops.view(d0), B, H, H, 49, cd.w_f8, cd.bias, 1, 7, 7, flow, wx, bx, wy, by, flow_out, None, 5.0),
       "conv_dist_R.1 with the tail in its epilogue", 2.0 * B * H * H * 49 * 49 * 7, unit="GFLOP/s")
