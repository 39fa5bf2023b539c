"""Drop-in for the reference's ``src/correlation.py`` (forward only).

``FunctionCorrelation(tensorFirst, tensorSecond, intStride)`` (src/correlation.py:411-412) and
``ModuleCorrelation`` (:417-424) keep their names, argument meaning and error behaviour
(:297-298 ``assert`` contiguity, :339-340 ``NotImplementedError`` for CPU tensors), but run ONE
sm_100a kernel (``pivlfn_corr_nchw`` in libpivlfn.so) instead of CuPy-JIT'd rearrange x2 +
updateOutput + three memsets.  The backward kernels (:106-234, :348-405) are training-only and out
of scope: asking for a gradient raises instead of silently returning none.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pivlfn import ops  # noqa: E402


class _FunctionCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, first, second, intStride):
        assert first.is_contiguous() is True
        assert second.is_contiguous() is True
        if first.is_cuda is not True or second.is_cuda is not True:
            raise NotImplementedError()
        assert first.shape == second.shape and first.dim() == 4
        return ops.corr_nchw(first, second, int(intStride))

    @staticmethod
    def backward(ctx, gradOutput):
        raise NotImplementedError("pivlfn implements the forward pass only (correlation backward is out of scope)")


def FunctionCorrelation(tensorFirst, tensorSecond, intStride):
    return _FunctionCorrelation.apply(tensorFirst, tensorSecond, intStride)


class ModuleCorrelation(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, tensorFirst, tensorSecond, intStride):
        return _FunctionCorrelation.apply(tensorFirst, tensorSecond, intStride)
