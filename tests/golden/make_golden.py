"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/src/models.py) on the CPU of the build container.

    python tests/golden/make_golden.py

Shims (see oracle/ref_import.py): a `cupy` stub + pure-torch correlation patched over
`src.models.FunctionCorrelation`, and `Tensor.cuda` as a no-op.  Weights are the deterministic
synthetic weights of `pivlfn.synth.synthetic_state_dict` (the reference's pretrained blobs are
not in the mount); inputs are synthetic particle pairs from `pivlfn.synth.particle_pair`.
The files written here are committed; nothing at test/bench time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))

from oracle import lfn_oracle as O  # noqa: E402
from oracle import ref_import as R  # noqa: E402
from pivlfn import synth  # noqa: E402

CASES = [
    # name, model, version, B, H, W, flow, weight seed, image seed
    ("piv_b2_64x96", "piv", 2, 64, 96, "rankine", 0, 100),
    ("piv_b1_128x128", "piv", 1, 128, 128, "shear", 1, 200),
    ("hui_b1_64x128", "hui", 1, 64, 128, "uniform", 2, 300),
    ("piv2_b1_64x64", "piv2", 1, 64, 64, "rankine", 3, 400),
    ("hui2_b1_64x64", "hui2", 1, 64, 64, "shear", 4, 500),
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    models, _ = R.load_reference_models(O.correlation)
    fac = {"piv": lambda sd: models.piv_liteflownet(sd, 1), "hui": lambda sd: models.hui_liteflownet(sd, 1),
           "piv2": lambda sd: models.piv_liteflownet(sd, 2), "hui2": lambda sd: models.hui_liteflownet(sd, 2)}
    for name, model, B, H, W, flow, wseed, iseed in CASES:
        sd = synth.synthetic_state_dict(model, wseed)
        net = fac[model](sd)
        ims = [synth.particle_pair(H, W, iseed + i, flow) for i in range(B)]
        a = torch.stack([synth.to_rgb_tensor(p[0]) for p in ims])
        b = torch.stack([synth.to_rgb_tensor(p[1]) for p in ims])
        out = {"img1_u8": np.stack([p[0] for p in ims]), "img2_u8": np.stack([p[1] for p in ims]),
               "wseed": np.int64(wseed), "model": np.array(model)}
        with torch.no_grad(), R.cpu_cuda_noop():
            net.eval()
            a_in, b_in = a.clone(), b.clone()
            out["flow"] = net(a_in, b_in).numpy()
            # the reference mutates its inputs in place (src/models.py:321-323): record it
            out["img1_after"] = (a_in - a)[0, :, 0, 0].numpy()
            net.train()   # training mode returns every level's [M, S, R] flows (src/models.py:365-367)
            tr = net(a.clone(), b.clone())
            net.eval()
        for k, lv in enumerate(tr):
            if len(lv) == 3:
                for tag, t in zip("MSR", lv):
                    out[f"lvl{k}_{tag}"] = t.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "flow absmax %.3f" % np.abs(out["flow"]).max(), {k: v.shape for k, v in out.items() if k.startswith("flow")})

    # state_dict key order and shapes of the four reference models (the drop-in contract)
    import json
    keys = {}
    for name, fn in (("piv", lambda: models.piv_liteflownet(None, 1)), ("hui", lambda: models.hui_liteflownet(None, 1)),
                     ("piv2", lambda: models.piv_liteflownet(None, 2)), ("hui2", lambda: models.hui_liteflownet(None, 2))):
        keys[name] = [[k, list(v.shape)] for k, v in fn().state_dict().items()]
    json.dump(keys, open(os.path.join(HERE, "state_dict_keys.json"), "w"))

    # operator-level vectors from the reference's own backwarp (src/models.py:20-35)
    g = torch.Generator().manual_seed(7)
    inp = torch.randn(2, 5, 9, 13, generator=g)
    flw = 3.0 * torch.randn(2, 2, 9, 13, generator=g)
    with torch.no_grad(), R.cpu_cuda_noop():
        models.backwarp_tensorGrid.clear()
        w = models.backwarp(inp, flw)
    np.savez_compressed(os.path.join(HERE, "backwarp.npz"), inp=inp.numpy(), flow=flw.numpy(), out=w.numpy())
    print("backwarp", tuple(w.shape))


if __name__ == "__main__":
    main()
