"""CPU oracle of the stereo-PIV post-processing (TEST INFRASTRUCTURE): numpy restatements of ``nl_trans``
(/root/reference/stereo/dewarp.py:255-270), ``_stereo_cal`` (stereo_run.py:153-163) and ``willert``
(stereo/vel3d.py:4-24) with the dtype behaviour of the reference's pinned numpy 1.17 (requirements.txt:9) made explicit:
a Python float / np.float64 scalar that meets a float32 array is rounded to float32 and the operation runs in float32;
scalar-with-scalar arithmetic stays float64.  (Run under numpy >= 2, the reference's own willert promotes to float64; the
two agree to ~1e-7 relative, checked in tests/test_stereo.py when /root/reference is mounted.)"""
import numpy as np

f32 = np.float32


def nl_trans(x, y, A):
    x, y = np.asarray(x, f32), np.asarray(y, f32)
    A = [f32(a) for a in A]

    def poly(a):
        # stereo/dewarp.py:263-264: a0*x + a1*y + a2 + a3*x**2 + a4*y**2 + a5*x*y, left to right
        return a[0] * x + a[1] * y + a[2] + a[3] * (x * x) + a[4] * (y * y) + a[5] * x * y

    return poly(A[0:6]) / poly(A[6:12]), poly(A[12:18]) / poly(A[18:24])


def stereo_cal(flow, A, fps, calibrate=None):
    """stereo_run.py:153-163; flow: (H, W, 2)."""
    nx, ny = nl_trans(flow[:, :, 0], flow[:, :, 1], A)
    out = np.dstack([nx, ny])
    if calibrate:
        out = out * f32(calibrate) * f32(fps)
    return out


def willert(flow, theta, beta):
    """stereo/vel3d.py:4-24; flow: [left, right] of (H, W, 2) float32; returns (H, W, 3) float32."""
    u = [np.asarray(f[:, :, 0], f32) for f in flow]
    v = [np.asarray(f[:, :, 1], f32) for f in flow]
    t0, t1 = np.tan(np.float64(theta[0])), np.tan(np.float64(theta[1]))
    b0, b1 = np.tan(np.float64(beta[0])), np.tan(np.float64(beta[1]))
    dt, db = f32(t0 - t1), f32(b1 - b0)
    u3 = (u[1] * f32(t0) - u[0] * f32(t1)) / dt
    v3 = (v[0] + v[1]) / f32(2) + (u[1] - u[0]) * db / dt / f32(2)
    w3 = (u[1] - u[0]) / dt
    return np.dstack([u3, v3, w3]).astype(f32)
