"""A few launches of one tensor-core convolution layer for `ncu --set full`.
    python tools/profile_conv.py [precision] [B] [H] [cin] [cout] [k]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
from pivlfn import ops  # noqa: E402
from pivlfn.model import pack_conv  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
cin = int(sys.argv[4]) if len(sys.argv) > 4 else 128
cout = int(sys.argv[5]) if len(sys.argv) > 5 else 128
kspec = sys.argv[6] if len(sys.argv) > 6 and not sys.argv[6].startswith("--") else "3"
kh, kw = (int(v) for v in kspec.split("x")) if "x" in kspec else (int(kspec), int(kspec))
k = kh
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
w = (torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5).to(dev)
b = torch.randn(cout, generator=g).to(dev)
cw = pack_conv(w, b, 1)
x = torch.randn(B, H, H, (cin + 3) & ~3, generator=g).to(dev)
y = torch.empty(B, H, H, (cout + 3) & ~3, device=dev)
passes = cw.passes_for({"3xtf32": 3, "tf32c": 2, "f16c": 4}.get(prec, 1))
trace = "--trace" in sys.argv
if trace:
    import ctypes
    from pivlfn import _lib
    dbg = torch.zeros(16, dtype=torch.int64, device=dev)
    fn = _lib.load().pivlfn_debug_set_conv_trace
    fn.argtypes = [ctypes.c_void_p]
    fn.restype = None
    fn(dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    if i == 1:
        e0.record()
    ops.conv_tc(ops.view(x, 0, cin), B, H, H, cw.w_hi, cw.w_lo, cw.bias, ops.view(y, 0, cout), kh, kw, True, passes, None,
                cw.pack16(passes))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{prec} conv {cin}->{cout} {kh}x{kw} @ {B}x{H}x{H}: {ms:.3f} ms, {2.0 * B * H * H * cin * cout * kh * kw / ms / 1e9:.1f} TFLOP/s")
if trace:
    torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    names = ["mma_total", "mma_wait_acc_empty", "mma_wait_A", "mma_wait_B", "items", "prod_wait_b_empty", "prod_wait_slot",
             "epi_wait_acc_full", "epi_busy", "split_wait_prev_chunk", "split_wait_raw", "split_busy", "epi_tmem_ld", "prod_total", "a_load_issue_to_seen"]
    print("  trace (CTA 0, cycles): " + ", ".join(f"{n}={v}" for n, v in zip(names, t)))
