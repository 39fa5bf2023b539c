"""Import the UNMODIFIED reference model (``src/models.py`` + ``src/correlation.py``).

TEST / BASELINE INFRASTRUCTURE.  The files come from the read-only mount ``/root/reference`` (build container) or from the
byte-identical copy that ``baseline/install_ref.py`` leaves in the git-ignored ``baseline/_ref/reference`` (which travels
to the GPU box).  Used by ``tests/golden/make_golden.py`` to generate golden vectors, by ``tests/test_oracle_golden.py``
to pin ``oracle/lfn_oracle.py``, by the GPU parity tests (the reference's own CUDA path through the cupy stand-in) and by
``bench.py``'s reference arms.

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

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_MOUNT = os.environ.get("PIVLFN_REFERENCE", "/root/reference")
_INSTALLED = os.path.join(_REPO, "baseline", "_ref", "reference")      # byte-identical copy made by baseline/install_ref.py
REF_ROOT = _MOUNT if os.path.isfile(os.path.join(_MOUNT, "src", "models.py")) else _INSTALLED
_CUPY_STUB = os.path.join(_REPO, "baseline", "cupy_stub")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "models.py"))


def _stub_cupy():
    """The reference imports ``cupy`` at module level (src/correlation.py:5); CuPy is not in this image.  The stand-in
    (baseline/cupy_stub) implements the two entry points the reference calls on top of NVRTC + the driver API."""
    if "cupy" in sys.modules:
        return
    sys.path.insert(0, _CUPY_STUB)
    try:
        importlib.import_module("cupy")
    finally:
        sys.path.remove(_CUPY_STUB)


def load_reference_models(corr_fn=None):
    """Returns the reference's UNMODIFIED ``src.models`` and ``src.correlation`` modules (imported under the names
    ``_ref_src.*`` so they cannot shadow this repo's own drop-in ``src`` package).

    corr_fn given (CPU use): ``src.models.FunctionCorrelation`` is replaced by it, because the reference has no CPU branch
    (src/correlation.py:339-340).  corr_fn None (GPU use): nothing is patched -- the reference's own CUDA kernels run through
    the cupy stand-in."""
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
    if corr_fn is not None:
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
