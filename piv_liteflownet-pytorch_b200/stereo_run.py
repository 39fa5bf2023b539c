"""Drop-in for the flow path of the reference's ``stereo_run.py`` on the B200: two ``estimate`` calls (left / right camera,
stereo_run.py:79,99), the 24-coefficient mapping of both flows (``_stereo_cal`` :153-163) and Willert's 2D3C recombination
(:85,145) -- the last two fused into one GPU kernel, so the flows never leave the device before the 3-band ``.flo`` is
written.  The reference's CLI plumbing around it (hard-coded debug argv :168-175, DataLoader drivers) is not reproduced.
"""
import os
import sys
from typing import List, Optional, Sequence

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from inference import estimate  # noqa: E402
from pivlfn import ops  # noqa: E402


def camera_angles(theta_deg: Sequence[float], alpha_deg: Sequence[float]):
    """stereo_run.py:111-118: degrees -> signed radians, the LEFT camera angles negative.  Returns (theta, beta)."""
    theta, beta = [], []
    for i in range(2):
        sign = (-1) ** (i + 1)
        a = alpha_deg[0] if len(alpha_deg) == 1 else alpha_deg[i]
        t = theta_deg[0] if len(theta_deg) == 1 else theta_deg[i]
        beta.append(sign * np.deg2rad(a))
        theta.append(sign * np.deg2rad(t))
    return theta, beta


def _stereo_cal(flow, A, fps: float, calibrate: Optional[float] = None):
    """stereo_run.py:153-163 for one camera.  flow: (H, W, 2) CUDA tensor or numpy array; returns the same kind."""
    from stereo.dewarp import nl_trans
    nx, ny = nl_trans(flow[:, :, 0], flow[:, :, 1], A)
    if isinstance(nx, np.ndarray):
        out = np.dstack([nx, ny])
        return out * np.float32(calibrate) * np.float32(fps) if calibrate else out
    out = torch.stack([nx, ny], dim=2)
    return out * calibrate * fps if calibrate else out


def stereo_estimate(net, left1, left2, right1, right2, coeffdict: dict, theta_deg: Sequence[float] = (45.0, 45.0),
                    alpha_deg: Sequence[float] = (0.0, 0.0), fps: float = 1, calib: Optional[float] = None) -> torch.Tensor:
    """One time step of stereo_run.direct_process (:60-88): images [B,3,H,W] on the device -> [B,H,W,3] (U, V, W) on the
    device.  ``coeffdict``: the mapping-coefficient json (keys "Left", "Right", optional "calib")."""
    theta, beta = camera_angles(theta_deg, alpha_deg)
    cal = (calib / coeffdict["calib"]) if (calib and "calib" in coeffdict) else None       # stereo_run.py:121-124
    fl = estimate(net, left1, left2, tensor=True)
    fr = estimate(net, right1, right2, tensor=True)
    return ops.stereo_2d3c(fl, fr, coeffdict["Left"], coeffdict["Right"], cal, fps, theta, beta)
