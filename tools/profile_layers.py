"""Per-launch timing of one eager PIV-LiteFlowNet-en forward, labelled with the layer each launch belongs to:
CUDA events around every operator call (no profiler), convolution FLOPs / memory-op bytes next to the time.
    python tools/profile_layers.py [B] [H] [precision] [reps]
Output: one line per launch in launch order, then totals per operator family."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
from pivlfn import model as M, ops, synth  # noqa: E402
from pivlfn.arch import CFGS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
prec = sys.argv[3] if len(sys.argv) > 3 else "f16c"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
sd = {k: v.to(dev) for k, v in synth.synthetic_state_dict("piv", 0).items()}
eng = M.Engine(CFGS["piv"], sd, dev, prec, use_graph=False)
i1, i2, _ = synth.particle_pair(H, H, 3, "rankine")
a = synth.to_rgb_tensor(i1)[None].repeat(B, 1, 1, 1).to(dev)
b = synth.to_rgb_tensor(i2)[None].repeat(B, 1, 1, 1).to(dev)

records = []          # (label, op, e0, e1, work, unit)
cur = {"label": ""}

_conv = M.Plan._conv


def conv_hook(self, key, x, n, h, w, y, lrelu=True, res=None):
    cur["label"] = key
    _conv(self, key, x, n, h, w, y, lrelu, res)
    cur["label"] = ""


M.Plan._conv = conv_hook

from pivlfn import plan16 as P16  # noqa: E402

_conv16 = P16.Plan16._conv


def conv16_hook(self, key, x, n, h, w, y, lrelu=True, out_fmt=0, cin=None):
    cur["label"] = key
    _conv16(self, key, x, n, h, w, y, lrelu, out_fmt, cin)
    cur["label"] = ""


P16.Plan16._conv = conv16_hook


def wrap(name, work):
    fn = getattr(ops, name)

    def f(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*args, **kw)
        e1.record()
        wk, unit, lab = work(*args, **kw)
        records.append(((cur["label"] + " " + lab).strip(), name, e0, e1, wk, unit))
    setattr(ops, name, f)


def w_conv_tc(x, N, Hh, Ww, w_hi, w_lo, bias, y, KH, KW, lrelu, passes, res=None, w_c16=None, stride=1):
    return 2.0 * N * (Hh // stride) * (Ww // stride) * x.C * y.C * KH * KW, "F", f"{x.C}->{y.C} {KH}x{KW} s{stride} @{Hh}x{Ww}"


def w_conv_simt(x, N, Hh, Ww, w, bias, y, KH, KW, stride, lrelu, res=None):
    return 2.0 * N * (Hh // stride) * (Ww // stride) * x.C * y.C * KH * KW, "F", f"{x.C}->{y.C} {KH}x{KW} s{stride} @{Hh}x{Ww}"


def w_s2(x, N, Hh, Ww, w16, bias, y, lrelu, passes):
    return 2.0 * N * (Hh // 2) * (Ww // 2) * x.C * y.C * 9, "F", f"{x.C}->{y.C} 3x3 s2 (parity phases) @{Hh}x{Ww}"


def w_stem(img_pad, N, Hh, Ww, w_hi, w_lo, bias, y, lrelu, passes, w_c16=None):
    return 2.0 * N * Hh * Ww * 3 * 32 * 49, "F", f"stem 3->32 7x7 @{Hh}x{Ww}"


def w_pairs(x, N, Hh, Ww, w_hi, w_lo, w_c16, planes, npair, passes):
    return 2.0 * N * Hh * Ww * x.C * 2 * npair, "F", f"flow head as 1x1 {x.C}->{2 * npair} @{Hh}x{Ww}"


def w_head(x, N, Hh, Ww, w, bias, res, out, K, out2=None):
    return 2.0 * N * Hh * Ww * 32 * 2 * K * K, "F", f"flow head (fp32 CUDA cores) 32->2 {K}x{K} @{Hh}x{Ww}"


def w_headsum(planes, K, bias, res, out, N, Hh, Ww):
    return 4.0 * N * Hh * Ww * (2 * K * K + 4), "B", f"flow_head_sum K={K} @{Hh}x{Ww}"


def w_deconv(x, N, Hh, Ww, w, y):
    return 4.0 * N * Hh * Ww * x.C * 5, "B", f"deconv C={x.C} @{Hh}x{Ww}"


def w_warp(x, flow, scale, y, N, Hh, Ww):
    return 4.0 * N * Hh * Ww * (2 * x.C + 2), "B", f"warp C={x.C} @{Hh}x{Ww}"


def w_corr(f1, f2, flow, scale, out, N, Hh, Ww, stride, lrelu=True):
    ho, wo = (Hh + stride - 1) // stride, (Ww + stride - 1) // stride
    return 4.0 * N * (2 * f1.C * Hh * Ww + (2 * Hh * Ww if flow is not None else 0) + 49 * ho * wo), "B", \
        f"corr C={f1.C} s{stride} @{Hh}x{Ww}"


def w_regtail(dist, flow_in, wx, bx, wy, by, flow_out, out_nchw, final_scale, K):
    N, Hh, Ww, _ = flow_in.shape
    return 4.0 * N * Hh * Ww * (K * K + 4), "B", f"reg_tail K={K} @{Hh}x{Ww}"


def w_reginput(img1, img2, flow, scale, partial, out):
    N, Hh, Ww, _ = flow.shape
    return 4.0 * N * Hh * Ww * 13, "B", f"reg_input @{Hh}x{Ww}"


def w_copy(src, dst, npix):
    return 8.0 * npix * src.C, "B", f"copy C={src.C}"


def w_conv_p16(x, N, Hh, Ww, cin, w_img, mode, bias, y, cout, KH, KW, stride=1, lrelu=True, out_fmt=0, plane_stride=0, flag=None):
    return 2.0 * N * (Hh // stride) * (Ww // stride) * cin * cout * KH * KW, "F", \
        f"{cin}->{cout} {KH}x{KW} s{stride} mode{mode} out{out_fmt} @{Hh}x{Ww}"


def w_conv_p16_warp(x, N, Hh, Ww, cin, w_img, mode, bias, y, cout, KH, KW, lrelu, wsrc, wsrc_p16, wflow, wscale, wc0, wn, flag=None):
    return 2.0 * N * Hh * Ww * cin * cout * KH * KW, "F", f"{cin}->{cout} {KH}x{KW} s1 mode{mode} + fused backwarp of {wn} ch @{Hh}x{Ww}"


def w_conv_p16_tail(x, N, Hh, Ww, cin, w_img, bias, KH, KW, K, *rest):
    return 2.0 * N * Hh * Ww * cin * K * K * KH * KW, "F", f"{cin}->{K * K} {KH}x{KW} s1 + Regularization tail in the epilogue @{Hh}x{Ww}"


def w_stem16(img_pad, N, Hh, Ww, w_img, bias, y, lrelu=True, flag=None):
    return 2.0 * N * Hh * Ww * 3 * 32 * 49, "F", f"stem 3->32 7x7 @{Hh}x{Ww}"


def w_corr16(f1, f1_p16, f2, f2_p16, flow, scale, out, out_p16, N, Hh, Ww, C, stride, lrelu=True, flag=None):
    ho, wo = (Hh + stride - 1) // stride, (Ww + stride - 1) // stride
    return 4.0 * N * (2 * C * Hh * Ww + (2 * Hh * Ww if flow is not None else 0) + 49 * ho * wo), "B", \
        f"corr C={C} s{stride} @{Hh}x{Ww}"


def w_warp16(x, in_p16, flow, scale, y, N, Hh, Ww, C, flag=None):
    return 4.0 * N * Hh * Ww * (2 * C + 2), "B", f"warp C={C} @{Hh}x{Ww}"


def w_deconv16(x, N, Hh, Ww, C, w, y, flag=None):
    return 4.0 * N * Hh * Ww * (52 + 4 * 64), "B", f"deconv C={C} @{Hh}x{Ww} (fp32 -> P16)"


def w_reginput16(img1, img2, flow, scale, partial, out, flag=None):
    N, Hh, Ww, _ = flow.shape
    return 4.0 * N * Hh * Ww * 13, "B", f"reg_input @{Hh}x{Ww}"


def w_rows(planes, K, bias, res, out, out_p16, N, Hh, Ww, flag=None):
    return 4.0 * N * Hh * Ww * (2 * K + 4 + (8 if out_p16 is not None else 0)), "B", f"head_rows/cols_sum K={K} @{Hh}x{Ww}"


def w_small(*a, **k):
    return 0.0, "B", ""


for nm, wk in (("conv_tc", w_conv_tc), ("conv_simt", w_conv_simt), ("conv_stem_tc", w_stem), ("conv_s2_tc", w_s2), ("conv1x1_pairs_tc", w_pairs),
               ("flow_head_sum", w_headsum), ("flow_head", w_head), ("deconv4x4s2_dw", w_deconv), ("warp", w_warp), ("corr_nhwc", w_corr),
               ("reg_tail", w_regtail), ("reg_input", w_reginput), ("copy", w_copy), ("flow_mean", w_small),
               ("prep_images", w_small), ("avgpool2", w_small), ("conv_p16", w_conv_p16), ("conv_p16_warp", w_conv_p16_warp), ("conv_p16_tail", w_conv_p16_tail), ("conv_stem_p16", w_stem16),
               ("corr_p16", w_corr16), ("warp_p16", w_warp16), ("deconv4x4s2_dw_p16", w_deconv16),
               ("reg_input_p16", w_reginput16), ("head_rows_sum", w_rows), ("head_cols_sum", w_rows)):
    wrap(nm, wk)

acc = None
for r in range(reps + 1):
    records.clear()
    eng.forward(a.clone(), b.clone())
    torch.cuda.synchronize()
    ms = [rec[2].elapsed_time(rec[3]) for rec in records]
    if r == 0:
        continue                      # warm-up
    acc = ms if acc is None else [min(x, y) for x, y in zip(acc, ms)]

tot = sum(acc)
print(f"# PIV-en {prec} B={B} {H}x{H}: {len(acc)} launches, sum of per-launch best times {tot:.3f} ms "
      f"({B / tot * 1e3:.0f} pairs/s if back to back)")
fam = collections.defaultdict(lambda: [0.0, 0.0])
for (label, op, _, _, wk, unit), t in zip(records, acc):
    rate = (f"{wk / t / 1e9:8.1f} TFLOP/s" if unit == "F" else f"{wk / t / 1e6:8.0f} GB/s") if wk else ""
    print(f"{t:8.4f} ms {100 * t / tot:5.2f}%  {op:18s} {label:72s} {rate}")
    fam[op][0] += t
    fam[op][1] += wk
print("# per operator family")
for op, (t, wk) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"# {t:8.3f} ms {100 * t / tot:5.1f}%  {op:18s} " + (f"{wk / t / 1e9:8.1f} T(FLOP|B/1000)/s" if wk else ""))
