"""Times the level-1 convolution shapes of PIV-LiteFlowNet-en on the P16 kernel (pivlfn_conv_p16), one layer at a time, with
CUDA events (inputs >> L2 at batch 64).  Experiment switches are read by the library from the environment
(PIVLFN_P16_ISSUERS, PIVLFN_P16_COLLECT, PIVLFN_P16_NT).
    python tools/profile_conv16.py [B] [H] [layer-filter]        e.g.  python tools/profile_conv16.py 64 256 128->128"""
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
filt = sys.argv[3] if len(sys.argv) > 3 else ""
dev = torch.device("cuda", 0)
LAYERS = [  # cin, cout, kh, kw, stride, out_fmt, name
    (128, 128, 3, 3, 1, 0, "conv_R.2"), (132, 128, 3, 3, 1, 0, "conv_R.0"), (130, 128, 3, 3, 1, 0, "conv_S.0"),
    (49, 128, 3, 3, 1, 0, "conv_M.0"), (128, 64, 3, 3, 1, 0, "conv_*.2/4"), (64, 64, 3, 3, 1, 0, "conv_R.6"),
    (64, 32, 3, 3, 1, 0, "conv_*.4/8"), (32, 32, 3, 3, 1, 0, "conv_R.10"), (32, 49, 7, 1, 1, 0, "conv_dist_R.0"),
    (49, 49, 1, 7, 1, 1, "conv_dist_R.1"), (32, 64, 1, 1, 1, 0, "NetC_ext"), (32, 64, 1, 1, 1, 1, "NetC_ext (fp32 out)"),
    (32, 128, 1, 1, 1, 0, "moduleFeat"), (32, 14, 1, 7, 1, 2, "flow head rows 1x7"), (32, 14, 7, 1, 1, 2, "flow head cols 7x1"),
    (32, 32, 3, 3, 2, 0, "NetC.conv2.0 s2"),
]
g = torch.Generator(device="cpu").manual_seed(0)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
print(f"# conv_p16 B={B} {H}x{H}  env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("PIVLFN_")))
for cin, cout, kh, kw, st, fmt, name in LAYERS:
    tag = f"{cin}->{cout} {kh}x{kw} s{st}"
    if filt and filt not in tag and filt not in name:
        continue
    w = torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin * kh * kw)
    b = torch.randn(cout, generator=g)
    cw = pack_conv(w, b, st).to_(dev)
    w_img, mode = cw.w_f8, 6
    cw16 = (cin + 15) & ~15
    x = torch.zeros(B, H, H, cw16, device=dev)
    ops.p16_encode(ops.view(torch.randn(B, H, H, cin, generator=g).to(dev)), ops.view(x), B * H * H, flag)
    Ho = H // st
    if fmt == 0:
        y = ops.view(torch.empty(B, Ho, Ho, (cout + 15) & ~15, device=dev))
        ps = 0
    elif fmt == 1:
        y = ops.view(torch.empty(B, Ho, Ho, (cout + 3) & ~3, device=dev))
        ps = 0
    else:
        y = ops.view(torch.empty(1, cout // 2, B * Ho * Ho, 2, device=dev))
        ps = 2 * B * Ho * Ho
    fn = lambda: ops.conv_p16(ops.view(x), B, H, H, cin, w_img, mode, cw.bias, y, cout, kh, kw, st, True, fmt, ps, flag)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 3)
    fl = 2.0 * B * Ho * Ho * cin * cout * kh * kw
    print(f"{best:8.4f} ms  {fl / best / 1e9:7.1f} TFLOP/s  mode{mode} out{fmt}  {tag:18s} {name}")
    del x, y
