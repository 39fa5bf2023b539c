"""Golden vectors of the stereo post-processing from the UNMODIFIED reference functions (build container only):
/root/reference/stereo/vel3d.py::willert and stereo/dewarp.py::nl_trans (imported with cv2 / scipy / matplotlib stubbed:
those imports serve the calibration GUI, not these two functions).  Run:  python tests/golden/make_stereo_golden.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PIVLFN_REFERENCE", "/root/reference")


def load_reference():
    for name in ("cv2", "matplotlib", "matplotlib.pyplot", "scipy", "scipy.optimize", "scipy.interpolate", "scipy.ndimage",
                 "pandas", "imutils", "skimage", "skimage.feature"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "stereo" or k.startswith("stereo.")}
    sys.path.insert(0, REF)
    try:
        import importlib
        vel3d = importlib.import_module("stereo.vel3d")
        try:
            dewarp = importlib.import_module("stereo.dewarp")
        except Exception:
            dewarp = None
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "stereo" or k.startswith("stereo.")]:
            sys.modules["_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return vel3d, dewarp


def inputs():
    rng = np.random.default_rng(7)
    H, W = 24, 40
    fl = (rng.standard_normal((H, W, 2)) * 3.0).astype(np.float32)
    fr = (fl + rng.standard_normal((H, W, 2)) * 0.5).astype(np.float32)
    # near-identity rational maps: x' = (x + small quadratic) / (1 + small), y' likewise
    def coeff(seed):
        r = np.random.default_rng(seed)
        A = np.zeros(24)
        A[0], A[8] = 1.0, 1.0          # new_x ~ x / 1
        A[13], A[20] = 1.0, 1.0        # new_y ~ y / 1
        A += r.standard_normal(24) * 1e-3
        return [float(a) for a in A]
    theta = [-np.deg2rad(30.0), np.deg2rad(32.0)]
    beta = [-np.deg2rad(5.0), np.deg2rad(4.0)]
    return fl, fr, coeff(1), coeff(2), theta, beta


if __name__ == "__main__":
    vel3d, dewarp = load_reference()
    fl, fr, AL, AR, theta, beta = inputs()
    out = {"fl": fl, "fr": fr, "AL": np.array(AL), "AR": np.array(AR), "theta": np.array(theta), "beta": np.array(beta),
           "numpy_version": np.array(np.__version__)}
    out["willert"] = np.asarray(vel3d.willert([fl, fr], theta, beta))          # float64 under numpy >= 2
    if dewarp is not None:
        nx, ny = dewarp.nl_trans(fl[:, :, 0], fl[:, :, 1], AL)
        out["nl_x"], out["nl_y"] = np.asarray(nx), np.asarray(ny)
    np.savez_compressed(os.path.join(HERE, "stereo.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
