"""One large PIV frame tiled by rows across the GPUs of one box (BASELINE configs[4] shape), NCCL halo exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/tiled_demo.py [H] [W] [--check]

Every rank builds the same synthetic frame pair and weights, runs its slab through pivlfn.tiled (DistGroup) and, with
--check, also the single-GPU forward of the whole frame to compare its own rows.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from pivlfn import synth  # noqa: E402
from pivlfn.arch import CFGS  # noqa: E402
from pivlfn.model import Engine  # noqa: E402
from pivlfn.tiled import DistGroup, make_tiled_plan  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
H = int(args[0]) if len(args) > 0 else 4096
W = int(args[1]) if len(args) > 1 else H
check = "--check" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")  # (torchrun sets it)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

sd = {k: v.to(dev) for k, v in synth.synthetic_state_dict("piv", 0).items()}
eng = Engine(CFGS["piv"], sd, dev, os.environ.get("PIVLFN_PRECISION"), use_graph=False)
# a 512x512 synthetic particle pair tiled to the frame size (host-side generation is O(particles) python)
i1, i2, _ = synth.particle_pair(512, 512, 7, "shear")
reps = ((H + 511) // 512, (W + 511) // 512)
a = synth.to_rgb_tensor(np.tile(i1, reps)[:H, :W])[None]
b = synth.to_rgb_tensor(np.tile(i2, reps)[:H, :W])[None]

plan = make_tiled_plan(eng, H, W, rank, world, halo=24, warp_reach=16)
group = DistGroup(plan)
plan.load_inputs(a, b)
group.run()                      # warm-up
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 3
e0.record()
for _ in range(iters):
    group.run()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
out = plan.owned_output()
res = {"workload": f"PIV-LiteFlowNet-en one {H}x{W} pair tiled by rows", "n_gpus": world, "ms_per_frame": float(ms.item()),
       "frames_per_s": 1e3 / float(ms.item()), "tiled_levels": plan.Lt, "halo_rows": plan.E,
       "exchanges_per_forward": sum(1 for s in plan.steps if s.kind == "exchange"),
       "ops_per_forward": sum(1 for s in plan.steps if s.kind == "op")}
if check:
    ref = eng.forward(a.to(dev), b.to(dev))
    own = H // world
    d = (out - ref[:, :, rank * own:(rank + 1) * own]).abs().max()
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    res["max_abs_diff_vs_single_gpu"] = float(d.item())
    # single-GPU time for the same frame
    torch.cuda.synchronize()
    e0.record()
    eng.forward(a.to(dev), b.to(dev))
    e1.record()
    torch.cuda.synchronize()
    res["single_gpu_ms_per_frame"] = e0.elapsed_time(e1)
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
