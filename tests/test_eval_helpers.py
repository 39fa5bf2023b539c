"""Evaluation helpers: `src.loss.EPE`, `src.postpro.calc_vorticity` / `de_vort` against golden vectors of
the unmodified reference functions (tests/golden/make_eval_golden.py) and, where /root/reference is mounted, against the
reference live."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
from src.loss import EPE  # noqa: E402
from src.postpro import calc_vorticity, de_vort  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "eval_helpers.npz")
REF = "/root/reference/src"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_epe_golden():
    g = np.load(GOLD)
    a, b = torch.from_numpy(g["epe_a"]), torch.from_numpy(g["epe_b"])
    assert abs(EPE(a, b).item() - float(g["epe_mean"])) <= 1e-6 * float(g["epe_mean"])
    assert abs(EPE(a, b, mean=False).item() - float(g["epe_sum"])) <= 1e-6 * float(g["epe_sum"])
    assert EPE(a, a).item() == 0.0
    # a uniform 3-4-5 offset is an end-point error of exactly 5 px everywhere
    t = a.clone()
    t[:, 0] += 3.0
    t[:, 1] -= 4.0
    assert abs(EPE(a, t).item() - 5.0) < 1e-5


def test_vorticity_golden():
    g = np.load(GOLD)
    for i in range(3):
        flow, calib = g[f"flow{i}"], float(g[f"calib{i}"])
        for nm, fn in (("vort", calc_vorticity), ("devort", de_vort)):
            got = fn(flow, calib)
            assert len(got) == 3
            for j, a in enumerate(got):
                ref = g[f"{nm}{i}_{j}"]
                assert a.shape == ref.shape and a.dtype == np.float64
                # float64 in, float64 out: de_vort keeps the reference's order of additions (bit-exact), calc_vorticity's
                # shifted sums differ from scipy's convolution order by rounding only
                tol = 0.0 if nm == "devort" else 4e-15
                assert np.abs(a - ref).max() <= tol, (nm, i, j, np.abs(a - ref).max())


def test_vorticity_known_answers():
    # rigid rotation u = -w y, v = w x (image rows = y): dv/dx = w, du/dy = -w -> the stencils give 2w / 0 in the interior
    h, w_, om = 9, 11, 0.25
    y, x = np.mgrid[0:h, 0:w_].astype(np.float64)
    flow = np.stack([-om * y, om * x], -1)
    vort, uy, vx = de_vort(flow)
    assert np.allclose(vx[1:-1, 1:-1], om) and np.allclose(uy[1:-1, 1:-1], om) and np.allclose(vort[1:-1, 1:-1], 0.0)
    v2, shear, normal = calc_vorticity(flow)
    assert np.allclose(v2[1:-1, 1:-1], 0.0) and np.allclose(shear[1:-1, 1:-1], 2 * om) and np.allclose(normal, -shear)
    # calib scales the derivatives
    assert np.allclose(de_vort(flow, 0.5)[2], 2 * vx)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_against_reference_live():
    post, loss = _load(os.path.join(REF, "postpro.py"), "ref_postpro"), _load(os.path.join(REF, "loss.py"), "ref_loss")
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        for shape in ((16, 16), (5, 9)):
            flow = rng.standard_normal(shape + (2,)).astype(dt)
            for calib in (1.0, 0.37):
                for fn_ref, fn in ((post.calc_vorticity, calc_vorticity), (post.de_vort, de_vort)):
                    for a, b in zip(fn_ref(flow, calib), fn(flow, calib)):
                        # float32 flows: the reference's scalar division runs in float32 under numpy 2 and in float64 under
                        # the numpy 1.x it was written for (ours): one float32 ulp
                        tol = 5e-7 * max(1.0, np.abs(a).max()) if dt == np.float32 else 4e-15
                        assert np.abs(np.asarray(a, np.float64) - b).max() <= tol
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(2, 2, 8, 8, generator=g), torch.randn(2, 2, 8, 8, generator=g)
    for mean in (True, False):
        assert abs(loss.EPE(a, b, mean).item() - EPE(a, b, mean).item()) <= 1e-6 * abs(loss.EPE(a, b, mean).item())
