"""Golden vectors for the evaluation helpers, made by importing the UNMODIFIED reference functions in the
build container:  python tests/golden/make_eval_golden.py  ->  tests/golden/eval_helpers.npz
  src/loss.py:12-21 EPE, src/postpro.py:5-24 calc_vorticity, :27-52 de_vort.
float64 flows, so that the vectors do not depend on the numpy version's scalar promotion rules (see src/postpro.py)."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    post = _load("/root/reference/src/postpro.py", "ref_postpro")
    loss = _load("/root/reference/src/loss.py", "ref_loss")
    rng = np.random.default_rng(20261018)
    out = {}
    for i, (shape, calib) in enumerate([((17, 23), 1.0), ((8, 5), 0.37), ((3, 3), 2.5)]):
        flow = rng.standard_normal(shape + (2,))
        out[f"flow{i}"] = flow
        out[f"calib{i}"] = np.float64(calib)
        for nm, fn in (("vort", post.calc_vorticity), ("devort", post.de_vort)):
            for j, a in enumerate(fn(flow, calib)):
                out[f"{nm}{i}_{j}"] = np.asarray(a, dtype=np.float64)
    a = torch.from_numpy(rng.standard_normal((3, 2, 12, 20)).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal((3, 2, 12, 20)).astype(np.float32))
    out["epe_a"], out["epe_b"] = a.numpy(), b.numpy()
    out["epe_mean"] = np.float32(loss.EPE(a, b, True).item())
    out["epe_sum"] = np.float32(loss.EPE(a, b, False).item())
    np.savez_compressed(os.path.join(HERE, "eval_helpers.npz"), **out)


if __name__ == "__main__":
    main()
