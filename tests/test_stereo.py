"""Stereo-PIV post-processing: oracle vs the golden vectors of the unmodified reference functions
(tests/golden/stereo.npz, made by tests/golden/make_stereo_golden.py), and the fused GPU operators vs the oracle, bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import stereo_oracle as SO

DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stereo.npz"))


def test_oracle_nl_trans_matches_reference_bitwise(gold):
    nx, ny = SO.nl_trans(gold["fl"][:, :, 0], gold["fl"][:, :, 1], gold["AL"])
    assert nx.dtype == np.float32
    assert np.array_equal(nx, gold["nl_x"]) and np.array_equal(ny, gold["nl_y"])


def test_oracle_willert_matches_reference(gold):
    """The golden vector was produced under numpy >= 2, where the reference's willert promotes to float64; the oracle
    restates it with the float32 semantics of the reference's pinned numpy 1.17: equal to float32 round-off."""
    out = SO.willert([gold["fl"], gold["fr"]], gold["theta"], gold["beta"])
    assert out.dtype == np.float32 and out.shape == gold["willert"].shape
    scale = np.abs(gold["willert"]).max()
    assert np.abs(out - gold["willert"]).max() <= 2e-6 * scale


@pytest.mark.gpu
def test_nl_trans_gpu_bitwise(gold):
    from pivlfn import ops
    from stereo.dewarp import nl_trans
    x, y = gold["fl"][:, :, 0], gold["fl"][:, :, 1]
    ref = SO.nl_trans(x, y, gold["AL"])
    nx, ny = ops.nl_trans(torch.from_numpy(x.copy()).to(DEV), torch.from_numpy(y.copy()).to(DEV), gold["AL"])
    assert np.array_equal(nx.cpu().numpy(), ref[0]) and np.array_equal(ny.cpu().numpy(), ref[1])
    a, b = nl_trans(x, y, list(gold["AL"]))                      # numpy in -> numpy out, like the reference
    assert isinstance(a, np.ndarray) and np.array_equal(a, gold["nl_x"]) and np.array_equal(b, gold["nl_y"])


@pytest.mark.gpu
@pytest.mark.parametrize("calib", [None, 0.37])
def test_stereo_2d3c_gpu_bitwise(gold, calib):
    from pivlfn import ops
    fl, fr = gold["fl"], gold["fr"]
    fps = 15
    ref = SO.willert([SO.stereo_cal(fl, gold["AL"], fps, calib), SO.stereo_cal(fr, gold["AR"], fps, calib)],
                     gold["theta"], gold["beta"])
    nchw = lambda f: torch.from_numpy(np.ascontiguousarray(f.transpose(2, 0, 1))[None]).to(DEV)
    two = lambda t: torch.cat([t, t + 1.0])                      # batch of 2: second sample shifted
    out = ops.stereo_2d3c(two(nchw(fl)), two(nchw(fr)), gold["AL"], gold["AR"], calib, fps, gold["theta"], gold["beta"])
    assert out.shape == (2, fl.shape[0], fl.shape[1], 3) and out.dtype == torch.float32
    assert np.array_equal(out[0].cpu().numpy(), ref)
    ref2 = SO.willert([SO.stereo_cal(fl + 1.0, gold["AL"], fps, calib), SO.stereo_cal(fr + 1.0, gold["AR"], fps, calib)],
                      gold["theta"], gold["beta"])
    assert np.array_equal(out[1].cpu().numpy(), ref2)


@pytest.mark.gpu
def test_willert_dropin_and_stereo_estimate(gold):
    from pivlfn import synth
    from src.models import piv_liteflownet
    from stereo.vel3d import willert
    from stereo_run import _stereo_cal, camera_angles, stereo_estimate
    ref = SO.willert([gold["fl"], gold["fr"]], gold["theta"], gold["beta"])
    out = willert([gold["fl"], gold["fr"]], list(gold["theta"]), list(gold["beta"]))
    assert isinstance(out, np.ndarray) and np.array_equal(out, ref)
    cal = _stereo_cal(gold["fl"], list(gold["AL"]), 10, 0.5)
    assert np.array_equal(cal, SO.stereo_cal(gold["fl"], gold["AL"], 10, 0.5))
    theta, beta = camera_angles([30.0], [5.0, 4.0])
    assert theta[0] < 0 < theta[1] and np.isclose(beta[1], np.deg2rad(4.0))
    # whole step: two estimate() calls + fused post-processing == the same pieces run one by one
    net = piv_liteflownet(synth.synthetic_state_dict("piv", 0), 1).to(DEV)
    a, b, _ = synth.particle_batch(1, 64, 64, 3, "uniform")
    c, d, _ = synth.particle_batch(1, 64, 64, 4, "shear")
    coeff = {"Left": list(gold["AL"]), "Right": list(gold["AR"]), "calib": 2.0}
    uvw = stereo_estimate(net, a.to(DEV), b.to(DEV), c.to(DEV), d.to(DEV), coeff, [30.0, 32.0], [5.0, 4.0], fps=7, calib=1.0)
    from inference import estimate
    fl = estimate(net, a.to(DEV), b.to(DEV))
    fr = estimate(net, c.to(DEV), d.to(DEV))
    theta, beta = camera_angles([30.0, 32.0], [5.0, 4.0])
    ref = SO.willert([SO.stereo_cal(fl, coeff["Left"], 7, 0.5), SO.stereo_cal(fr, coeff["Right"], 7, 0.5)], theta, beta)
    assert uvw.shape == (1, 64, 64, 3) and np.array_equal(uvw[0].cpu().numpy(), ref)
