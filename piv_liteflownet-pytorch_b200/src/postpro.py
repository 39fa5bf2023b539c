"""Host-side post-processing of a flow field, the functions of the reference's ``src/postpro.py``: ``calc_vorticity`` (:5-24,
Sobel derivatives with a symmetric border, returns vorticity / shear / normal strain) and ``de_vort`` (:27-52, the same
stencil with an edge-replicated border, returns vorticity, du/dy, dv/dx).  Flow ``[H, W, 2]`` as ``estimate`` / the .flo
files hold it, ``calib`` = physical length of one pixel.  Both are restated as whole-array shifted sums (no scipy, no
per-pixel Python loop); float64 like the reference."""
import numpy as np

__all__ = ["calc_vorticity", "de_vort"]


def _sobel(field: np.ndarray, pad_mode: str, dtype=np.float64):
    """Sobel sums of a 2-D field with a one-pixel border of the given numpy pad mode:
    gx[i, j] = sum_r w_r * (f[i+r, j+1] - f[i+r, j-1]),  gy[i, j] = sum_c w_c * (f[i+1, j+c] - f[i-1, j+c]),  w = (1, 2, 1).
    The sums run left to right in ``dtype`` (None = the field's own floating type)."""
    field = np.asarray(field)
    if dtype is None:
        dtype = field.dtype if np.issubdtype(field.dtype, np.floating) else np.float64
    p = np.pad(field.astype(dtype, copy=False), 1, mode=pad_mode)
    h, w = field.shape
    c = lambda dy, dx: p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    two = dtype(2) if isinstance(dtype, type) else np.dtype(dtype).type(2)
    gx = (c(1, 1) + two * c(0, 1) + c(-1, 1)) - (c(1, -1) + two * c(0, -1) + c(-1, -1))
    gy = (c(1, -1) + two * c(1, 0) + c(1, 1)) - (c(-1, -1) + two * c(-1, 0) + c(-1, 1))
    return gx, gy


def calc_vorticity(flow: np.ndarray, calib: float = 1.0):
    u, v = flow[:, :, 0], flow[:, :, 1]
    # the reference convolves (kernel flipped) u with -K^T and v with K, K = [[1,0,-1],[2,0,-2],[1,0,-1]] / (8 calib),
    # symmetric border: that is dv/dx by the Sobel x-difference and du/dy by the NEGATED Sobel y-difference
    gx_v, _ = _sobel(v, "symmetric")
    _, gy_u = _sobel(u, "symmetric")
    dv = gx_v / (8.0 * calib)
    du = -gy_u / (8.0 * calib)
    return dv - du, dv + du, -(dv + du)


def de_vort(flow: np.ndarray, calib: float = 1.0):
    u, v = flow[:, :, 0], flow[:, :, 1]
    # the reference's per-pixel loop adds numpy scalars of the flow's own dtype (float32 for a .flo) in this order and
    # divides by the Python float 8 * calib -- a float64 division under the numpy 1.x scalar rules it was written for
    gx_v, _ = _sobel(v, "edge", None)
    _, gy_u = _sobel(u, "edge", None)
    vx = gx_v.astype(np.float64) / (8 * calib)
    uy = (-gy_u).astype(np.float64) / (8 * calib)
    return vx - uy, uy, vx
