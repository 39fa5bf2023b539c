"""One eager (no CUDA graph) PIV-LiteFlowNet-en forward for profilers: every kernel is a separate launch.
    python tools/profile_forward.py [B] [H] [precision] [n_forwards]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import torch  # noqa: E402
from pivlfn import synth  # noqa: E402
from pivlfn.arch import CFGS  # noqa: E402
from pivlfn.model import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
prec = sys.argv[3] if len(sys.argv) > 3 else "3xtf32"
nfw = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
sd = {k: v.to(dev) for k, v in synth.synthetic_state_dict("piv", 0).items()}
eng = Engine(CFGS["piv"], sd, dev, prec, use_graph=False)
i1, i2, _ = synth.particle_pair(H, H, 3, "rankine")
a = synth.to_rgb_tensor(i1)[None].repeat(B, 1, 1, 1).to(dev)
b = synth.to_rgb_tensor(i2)[None].repeat(B, 1, 1, 1).to(dev)
for _ in range(nfw):
    out = eng.forward(a.clone(), b.clone())
torch.cuda.synchronize()
print("launches per forward:", eng.launches // nfw, "flow absmax", float(out.abs().max()))
