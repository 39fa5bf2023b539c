"""Per-layer comparison with cuDNN (VERDICT r1 item 6 / SURVEY.md section 2.2: "the kernel to beat"): every level-1 convolution of
PIV-LiteFlowNet-en at batch 64 of 256 x 256, torch.nn.functional.conv2d + bias + LeakyReLU(0.1) as the reference runs it
(NCHW, and channels_last for cuDNN's preferred tensor-core layout) in true fp32 (TF32 off) and with torch's TF32-default
convolutions, next to this repo's pivlfn_conv_p16 (split operands: one fp16 product + one fp8 product of corrections, flow error ~4e-4 px).  CUDA events, best of 3
x 3 launches, inputs >> L2.
    python tools/cudnn_layers.py [B] [H] > profiles/r2_cudnn_layers_b64_256.txt"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from pivlfn import ops  # noqa: E402
from pivlfn.model import pack_conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
LAYERS = [  # cin, cout, kh, kw, stride, lrelu, name
    (32, 32, 3, 3, 2, True, "NetC.conv2.0 (s2)"), (32, 64, 1, 1, 1, True, "NetC_ext.1"), (32, 128, 1, 1, 1, True, "NetE_R.0.moduleFeat"),
    (49, 128, 3, 3, 1, True, "conv_M.0"), (128, 64, 3, 3, 1, True, "conv_M.2 / conv_S.2 / conv_R.4"), (64, 32, 3, 3, 1, True, "conv_M.4 / conv_S.4 / conv_R.8"),
    (130, 128, 3, 3, 1, True, "conv_S.0"), (131, 128, 3, 3, 1, True, "conv_R.0"), (128, 128, 3, 3, 1, True, "conv_R.2"),
    (64, 64, 3, 3, 1, True, "conv_R.6"), (32, 32, 3, 3, 1, True, "conv_R.10"), (32, 49, 7, 1, 1, False, "conv_dist_R.0"),
    (49, 49, 1, 7, 1, False, "conv_dist_R.1"), (32, 2, 7, 7, 1, False, "flow head 7x7 (ours: 7x1 columns + column sum)"),
]


def best_ms(fn, reps=3, inner=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


g = torch.Generator(device="cpu").manual_seed(0)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
print(f"# level-1 layers, batch {B} of {H}x{H}, ms per launch (TFLOP/s of useful 2*MAC); torch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}")
print(f"# {'layer':44s} {'cuDNN fp32 NCHW':>18s} {'cuDNN fp32 NHWC':>18s} {'cuDNN TF32 NCHW':>18s} {'cuDNN TF32 NHWC':>18s} {'pivlfn (f16c)':>18s}  speed-up vs fp32 / TF32 (best layout)")
tot = {"fp32": 0.0, "tf32": 0.0, "ours": 0.0}
for cin, cout, kh, kw, st, act, name in LAYERS:
    w = (torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin * kh * kw)).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    x = torch.randn(B, cin, H, H, generator=g).to(dev)
    xcl = x.contiguous(memory_format=torch.channels_last)
    wcl = w.contiguous(memory_format=torch.channels_last)
    fl = 2.0 * B * (H // st) ** 2 * cin * cout * kh * kw
    res = {}
    with torch.no_grad():
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for lay, (xi, wi) in (("NCHW", (x, w)), ("NHWC", (xcl, wcl))):
                def ref():
                    y = F.conv2d(xi, wi, b, stride=st, padding=(kh // 2, kw // 2))
                    return F.leaky_relu(y, 0.1, inplace=True) if act else y
                res[(tf32, lay)] = best_ms(ref)
    torch.backends.cudnn.allow_tf32 = True
    # ours
    if cout == 2:
        K = kh
        rw = pack_conv(w.cpu().permute(3, 0, 1, 2).reshape(2 * K, cin, K, 1), None, 1).to_(dev)
        xin = torch.zeros(B, H, H, (cin + 15) & ~15, device=dev)
        ops.p16_encode(ops.view(x.permute(0, 2, 3, 1).contiguous()), ops.view(xin), B * H * H, flag)
        planes = torch.empty(K, B * H * H, 2, device=dev)
        out = torch.empty(B, H, H, 2, device=dev)

        def ours():
            ops.conv_p16(ops.view(xin), B, H, H, cin, rw.w_f8, 6, None, ops.view(planes.view(1, K, B * H * H, 2)), 2 * K, K, 1, 1, False,
                         ops.OUT_PLANES, 2 * B * H * H, flag)
            ops.head_cols_sum(planes, K, b, None, out, None, B, H, H, flag)
    else:
        cw = pack_conv(w.cpu(), b.cpu(), st).to_(dev)
        w_img, mode = cw.w_f8, 6
        xin = torch.zeros(B, H, H, (cin + 15) & ~15, device=dev)
        ops.p16_encode(ops.view(x.permute(0, 2, 3, 1).contiguous()), ops.view(xin), B * H * H, flag)
        y = torch.empty(B, H // st, H // st, (cout + 15) & ~15, device=dev)

        def ours():
            ops.conv_p16(ops.view(xin), B, H, H, cin, w_img, mode, cw.bias, ops.view(y), cout, kh, kw, st, act, ops.OUT_P16, 0, flag)
    mo = best_ms(ours)
    f32 = min(res[(False, "NCHW")], res[(False, "NHWC")])
    t32 = min(res[(True, "NCHW")], res[(True, "NHWC")])
    tot["fp32"] += f32; tot["tf32"] += t32; tot["ours"] += mo
    cell = lambda ms: f"{ms:8.3f} ({fl / ms / 1e9:6.1f})"
    print(f"  {name + f'  {cin}->{cout} {kh}x{kw} s{st}':44s} {cell(res[(False, 'NCHW')]):>18s} {cell(res[(False, 'NHWC')]):>18s} "
          f"{cell(res[(True, 'NCHW')]):>18s} {cell(res[(True, 'NHWC')]):>18s} {cell(mo):>18s}  {f32 / mo:5.2f}x / {t32 / mo:5.2f}x")
    del x, xcl, xin
print(f"# sum over these layers (best layout each): cuDNN fp32 {tot['fp32']:.2f} ms, cuDNN TF32 {tot['tf32']:.2f} ms, pivlfn {tot['ours']:.2f} ms "
      f"-> {tot['fp32'] / tot['ours']:.2f}x / {tot['tf32'] / tot['ours']:.2f}x")
print("# pivlfn: flow max |diff| ~5e-4 px vs the fp32 reference; cuDNN TF32 rounds both operands to 10 mantissa bits")
print("# (flow max |diff| 2.4e-2 px for the single-pass tf32 mode of this repo, outside the 1e-2 px north_star tolerance)")
