"""Run the REFERENCE's own CUDA correlation kernels (TEST INFRASTRUCTURE, GPU box).

The binaries are ``oracle/_ref/*.cubin``, compiled by ``oracle/build_ref.py`` from the kernel strings of the
reference's ``src/correlation.py:9-104`` after the reference's own ``cupy_kernel()`` templating.  This file restates
only the host side of ``_FunctionCorrelation.forward`` (``src/correlation.py:285-344``): three zero-filled buffers and
three launches with the reference's grids/blocks, through the CUDA driver API (cuda-python) on torch's current stream.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Tuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_funcs: Dict[Tuple[str, Tuple[int, ...]], object] = {}


def manifest():
    p = os.path.join(REF, "manifest.json")
    return json.load(open(p)) if os.path.isfile(p) else None


def shapes():
    m = manifest()
    return [tuple(e["shape"]) for e in m["entries"]] if m else []


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _function(kind: str, shape: Tuple[int, ...]):
    from cuda.bindings import driver as drv
    key = (kind, shape)
    if key not in _funcs:
        m = manifest()
        ent = next(e for e in m["entries"] if tuple(e["shape"]) == shape)
        data = open(os.path.join(REF, ent["files"][kind]), "rb").read()
        mod = _check(drv.cuModuleLoadData(data))
        _funcs[key] = (_check(drv.cuModuleGetFunction(mod, ("kernel_Correlation_" + kind).encode())), mod)
    return _funcs[key][0]


def _launch(fn, grid, block, smem, n: int, ptrs):
    from cuda.bindings import driver as drv
    args = [np.array([n], dtype=np.int32)] + [np.array([p], dtype=np.uint64) for p in ptrs]
    argv = np.array([a.ctypes.data for a in args], dtype=np.uint64)
    stream = torch.cuda.current_stream().cuda_stream
    _check(drv.cuLaunchKernel(fn, grid[0], grid[1], grid[2], block[0], block[1], block[2], smem, stream,
                              argv.ctypes.data, 0))


def reference_correlation(first: torch.Tensor, second: torch.Tensor, stride: int) -> torch.Tensor:
    """``FunctionCorrelation`` exactly as the reference's CUDA branch computes it (src/correlation.py:288-337)."""
    assert first.is_cuda and first.is_contiguous() and second.is_contiguous()
    B, C, H, W = first.shape
    shape = (B, C, H, W, int(stride))
    rbot0 = first.new_zeros([B, H + 6 * stride, W + 6 * stride, C])                    # :288-291
    rbot1 = first.new_zeros([B, H + 6 * stride, W + 6 * stride, C])
    out = first.new_zeros([B, 49, int(math.ceil(H / stride)), int(math.ceil(W / stride))])   # :300-301
    torch.cuda.current_stream().synchronize()       # make sure torch's primary context is current and buffers exist
    n = H * W
    f_re = _function("rearrange", shape)
    _launch(f_re, (int((n + 16 - 1) / 16), C, B), (16, 1, 1), 0, n, [first.data_ptr(), rbot0.data_ptr()])     # :305-313
    _launch(f_re, (int((n + 16 - 1) / 16), C, B), (16, 1, 1), 0, n, [second.data_ptr(), rbot1.data_ptr()])    # :315-324
    n = out.shape[1] * out.shape[2] * out.shape[3]
    _launch(_function("updateOutput", shape), (out.shape[3], out.shape[2], out.shape[0]), (32, 1, 1), C * 4, n,
            [rbot0.data_ptr(), rbot1.data_ptr(), out.data_ptr()])                                            # :326-337
    torch.cuda.current_stream().synchronize()
    return out
