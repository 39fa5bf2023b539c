"""Drop-in for the reference's ``stereo/vel3d.py``: ``willert`` (stereo/vel3d.py:4-24), on the GPU."""
import numpy as np
import torch

from pivlfn import ops


def willert(flow, theta, beta):
    """Willert (1997) recombination.  ``flow``: [left, right], each (H, W, 2) float32 -- numpy arrays (result: numpy
    (H, W, 3)) or CUDA tensors (result: CUDA tensor); ``theta`` / ``beta``: the two signed camera angles in radians,
    index 0 = left camera, 1 = right camera."""
    as_numpy = isinstance(flow[0], np.ndarray)
    fl = [torch.as_tensor(np.ascontiguousarray(f, dtype=np.float32)).cuda() if isinstance(f, np.ndarray) else f for f in flow]
    # (H, W, 2) -> [1, 2, H, W], the layout of the fused operator
    nchw = [f.permute(2, 0, 1).contiguous()[None] for f in fl]
    out = ops.stereo_2d3c(nchw[0], nchw[1], None, None, None, 1.0, theta, beta)[0]
    return out.cpu().numpy() if as_numpy else out
