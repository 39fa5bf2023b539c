"""A minimal stand-in for the CuPy 7 API surface that the reference's ``src/correlation.py`` uses (``:5,278-280``):

    @cupy.util.memoize(for_each_device=True)
    cupy.cuda.compile_with_cache(source).get_function(name)(grid=..., block=..., args=[...], shared_mem=...)

CuPy is not installed in this image (and current CuPy releases dropped both entry points).  With this package on
``sys.path`` the reference's files run UNMODIFIED on a GPU: the kernel source the reference generates is compiled with
NVRTC (``cuda.bindings.nvrtc``) for the current device and launched through the CUDA driver API on torch's current stream
with the reference's own launch geometry.  On a machine without a GPU the package imports fine and only compilation
fails (the CPU baseline replaces the correlation call, see oracle/ref_run.py).

TEST / BASELINE INFRASTRUCTURE: never imported by the product package.
"""
from . import cuda, util  # noqa: F401

__version__ = "7.2.0+pivlfn.stub"
