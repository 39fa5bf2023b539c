"""Row-slab tiling of one frame across ranks (BASELINE configs[4]): P ranks emulated in lock-step on ONE GPU
(pivlfn.tiled.LoopbackGroup: the communication steps -- halo exchange, flow-mean all-reduce, coarse-level all-gather --
are tensor copies) must reproduce the single-GPU forward.  The same step lists run under torch.distributed with NCCL
(pivlfn.tiled.DistGroup, tools/tiled_demo.py)."""
import pytest
import torch

from pivlfn import synth
from pivlfn.arch import CFGS
from pivlfn.model import Engine
from pivlfn.tiled import LoopbackGroup, TiledPlan

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("model,H,W,P,precision", [("piv", 256, 128, 2, "simt"), ("piv", 384, 96, 3, "3xtf32"),
                                                    ("hui", 256, 64, 2, "tf32c"), ("piv", 512, 64, 4, "f16c")])
def test_tiled_equals_single_gpu(model, H, W, P, precision):
    sd = {k: v.to(DEV) for k, v in synth.synthetic_state_dict(model, 0).items()}
    eng = Engine(CFGS[model], sd, torch.device(DEV), precision, use_graph=False)
    i1, i2, _ = synth.particle_pair(H, W, 31, "shear")
    a = synth.to_rgb_tensor(i1)[None].to(DEV)
    b = synth.to_rgb_tensor(i2)[None].to(DEV)
    ref = eng.forward(a.clone(), b.clone())
    plans = [TiledPlan(eng, H, W, r, P, halo=24, warp_reach=16) for r in range(P)]
    assert 1 <= plans[0].Lt <= 6
    for p in plans:
        p.load_inputs(a, b)
    LoopbackGroup(plans).run()
    torch.cuda.synchronize()
    for p in plans:
        p.check_warp_reach()
    out = torch.cat([p.owned_output() for p in plans], dim=2)
    assert out.shape == ref.shape
    diff = (out - ref).abs()
    n_ex = sum(1 for s in plans[0].steps if s.kind == "exchange")
    n_op = sum(1 for s in plans[0].steps if s.kind == "op")
    print(f"tiled {model} {H}x{W} P={P} {precision}: Lt={plans[0].Lt} exchanges={n_ex} ops={n_op} max|diff|={diff.max().item():.3e}")
    # identical kernels on identical data; only the flow-mean summation order differs -> fp32 round-off
    assert diff.max().item() <= 2e-4
