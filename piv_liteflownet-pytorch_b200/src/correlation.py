"""Drop-in for the reference's ``src/correlation.py``.

``FunctionCorrelation(tensorFirst, tensorSecond, intStride)`` (src/correlation.py:411-412) and
``ModuleCorrelation`` (:417-424) keep their names, argument meaning and error behaviour
(:297-298 ``assert`` contiguity, :339-340 ``NotImplementedError`` for CPU tensors), but run ONE
sm_100a kernel (``pivlfn_corr_nchw`` in libpivlfn.so) instead of CuPy-JIT'd rearrange x2 +
updateOutput + three memsets.  ``backward`` (:348-405, kernels :106-234) returns the same gradients from
``pivlfn_corr_backward_nchw`` (training is outside the accelerated path; the operator stays differentiable so that the
reference's trainer keeps working against it).
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
        ctx.save_for_backward(first, second)
        ctx.intStride = int(intStride)
        return ops.corr_nchw(first, second, int(intStride))

    @staticmethod
    def backward(ctx, gradOutput):
        first, second = ctx.saved_tensors
        assert gradOutput.is_contiguous() is True
        gradFirst, gradSecond = ops.corr_backward_nchw(first, second, gradOutput, ctx.intStride, ctx.needs_input_grad[0],
                                                       ctx.needs_input_grad[1])
        return gradFirst, gradSecond, None


def FunctionCorrelation(tensorFirst, tensorSecond, intStride):
    return _FunctionCorrelation.apply(tensorFirst, tensorSecond, intStride)


class ModuleCorrelation(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, tensorFirst, tensorSecond, intStride):
        return _FunctionCorrelation.apply(tensorFirst, tensorSecond, intStride)
