"""Drop-in for the part of the reference's ``stereo/dewarp.py`` that lies on the stereo_run.py flow path: ``nl_trans``
(stereo/dewarp.py:255-270), on the GPU.  Calibration / GUI helpers of that file are out of scope."""
import numpy as np
import torch

from pivlfn import ops


def nl_trans(x, y, A):
    """24-coefficient rational-quadratic map (stereo/dewarp.py:255-270).  ``x``, ``y``: float32 arrays of equal shape --
    CUDA tensors (result: CUDA tensors) or numpy arrays (uploaded, result: numpy arrays, like the reference)."""
    as_numpy = isinstance(x, np.ndarray)
    xt = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda() if as_numpy else x
    yt = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).cuda() if isinstance(y, np.ndarray) else y
    nx, ny = ops.nl_trans(xt, yt, A)
    if as_numpy:
        return nx.cpu().numpy(), ny.cpu().numpy()
    return nx, ny
