"""Import the UNMODIFIED reference model (``/root/reference/src/models.py``) on the CPU.

TEST INFRASTRUCTURE (build container only: ``/root/reference`` does not exist on the GPU
box).  Used by ``tests/golden/make_golden.py`` to generate golden vectors and by
``tests/test_oracle_golden.py`` (when the mount is present) to pin ``oracle/lfn_oracle.py``.

Two shims, both outside the reference files:
  1. a stub module named ``cupy`` (``src/correlation.py:5,278-280`` needs the name at import
     time) and a pure-torch correlation patched over ``src.models.FunctionCorrelation``
     (the reference has no CPU branch, ``src/correlation.py:339-340``);
  2. ``torch.Tensor.cuda`` made a no-op while the reference runs, because ``backwarp``
     forces ``.cuda()`` on its grid (``src/models.py:27``).
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("PIVLFN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "models.py"))


def _stub_cupy():
    if "cupy" in sys.modules:
        return
    cupy = types.ModuleType("cupy")
    util = types.ModuleType("cupy.util")

    def memoize(for_each_device=False):
        def deco(fn):
            return fn
        return deco

    util.memoize = memoize
    cuda = types.ModuleType("cupy.cuda")

    def compile_with_cache(*a, **k):  # pragma: no cover - never reached on CPU
        raise RuntimeError("cupy stub: no CUDA compilation in the CPU container")

    cuda.compile_with_cache = compile_with_cache
    cupy.util, cupy.cuda = util, cuda
    sys.modules["cupy"] = cupy
    sys.modules["cupy.util"] = util
    sys.modules["cupy.cuda"] = cuda


def load_reference_models(corr_fn):
    """Returns the reference's ``src.models`` module (imported under the name ``_ref_src.models``
    so it cannot shadow this repo's own drop-in ``src`` package)."""
    _stub_cupy()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    sys.path.insert(0, REF_ROOT)
    try:
        models = importlib.import_module("src.models")
        corr = importlib.import_module("src.correlation")
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            sys.modules["_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    models.FunctionCorrelation = lambda tensorFirst, tensorSecond, intStride: corr_fn(
        tensorFirst, tensorSecond, intStride)
    models.backwarp_tensorGrid.clear()
    return models, corr


@contextlib.contextmanager
def cpu_cuda_noop():
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
