"""Level-1 glue operators alone (cost volume with fused backwarp, backwarp, regularisation tail, depthwise up-convolution,
flow head) for ncu and for quick timing.   python tools/profile_ops.py [B] [H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
from pivlfn import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
cm = 64
Sbuf = torch.randn(B, H, H, 2 * cm + 8, generator=g).to(dev)
f2 = torch.randn(B, H, H, cm, generator=g).to(dev)
flow = (2.0 * torch.randn(B, H, H, 2, generator=g)).to(dev)
corr = torch.zeros(B, H // 2, H // 2, 52, device=dev)
dist = torch.randn(B, H, H, 52, generator=g).to(dev)
flow_out = torch.empty_like(flow)
wx, wy = torch.randn(49, generator=g).to(dev), torch.randn(49, generator=g).to(dev)
bx, by = torch.zeros(1, device=dev), torch.zeros(1, device=dev)


def timeit(fn, name, byts):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); fn()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.4f} ms, {byts / ms / 1e6:.0f} GB/s algorithmic")


timeit(lambda: ops.corr_nhwc(ops.view(Sbuf, 0, cm), ops.view(f2), flow, 5.0, ops.view(corr, 0, 49), B, H, H, 2, True),
       "corr_nhwc s=2 C=64 + backwarp", 4.0 * B * (2 * cm * H * H + 2 * H * H + 49 * (H // 2) ** 2))
timeit(lambda: ops.warp(ops.view(f2), flow, 5.0, ops.view(Sbuf, cm, cm), B, H, H), "warp_nhwc C=64", 4.0 * B * H * H * (2 * cm + 2))
timeit(lambda: ops.reg_tail(ops.view(dist, 0, 49), flow, wx, bx, wy, by, flow_out, None, 5.0, 7), "reg_tail K=7",
       4.0 * B * H * H * (49 + 4))

import math
corr_half = torch.randn(B, H // 2, H // 2, 52, generator=g).to(dev)
corr_up = torch.zeros(B, H, H, 52, device=dev)
wdw = torch.randn(49, 16, generator=g).to(dev)
timeit(lambda: ops.deconv4x4s2_dw(ops.view(corr_half, 0, 49), B, H // 2, H // 2, wdw, ops.view(corr_up, 0, 49)),
       "deconv4x4s2_dw C=49 (H/2 -> H)", 4.0 * B * (H // 2) ** 2 * 52 * 5)
x32 = torch.randn(B, H, H, 32, generator=g).to(dev)
wh = (torch.randn(49, 32, 2, generator=g) / math.sqrt(32 * 49)).to(dev)
bh = torch.zeros(2, device=dev)
timeit(lambda: ops.flow_head(ops.view(x32), B, H, H, wh, bh, ops.view(flow), ops.view(flow_out), 7),
       "flow_head K=7 (fp32 FMA: GFLOP/s in the GB/s column)", 2.0 * B * H * H * 32 * 2 * 49)
