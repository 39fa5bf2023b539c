"""``cupy.cuda.compile_with_cache`` (CuPy 7) on top of NVRTC + the CUDA driver API (cuda-python)."""
import re

import numpy as np


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"cupy stub: CUDA / NVRTC error {err}")
    return res[1] if len(res) == 2 else res[1:]


class _Function:
    def __init__(self, module, name, source):
        from cuda.bindings import driver as drv
        self._fn = _check(drv.cuModuleGetFunction(module, name.encode()))
        self._module = module
        # parameter kinds from the kernel's own signature: pointers are 64-bit, everything else here is `int`
        m = re.search(r"__global__\s+void\s+" + re.escape(name) + r"\s*\(([^)]*)\)", source, flags=re.S)
        if m is None:
            raise RuntimeError(f"cupy stub: kernel {name} not found in the source")
        self._kinds = []
        for p in m.group(1).split(","):
            p = p.strip()
            if "*" in p:
                self._kinds.append(np.uint64)
            elif re.search(r"\bfloat\b", p):
                self._kinds.append(np.float32)
            else:
                self._kinds.append(np.int32)

    def __call__(self, grid, block, args, shared_mem=0, stream=None):
        import torch
        from cuda.bindings import driver as drv
        if len(args) != len(self._kinds):
            raise TypeError("cupy stub: wrong number of kernel arguments")
        vals = [np.array([0 if a is None else a], dtype=k) for a, k in zip(args, self._kinds)]     # None -> NULL pointer
        argv = np.array([v.ctypes.data for v in vals], dtype=np.uint64)
        grid = tuple(grid) + (1,) * (3 - len(grid))
        block = tuple(block) + (1,) * (3 - len(block))
        st = torch.cuda.current_stream().cuda_stream
        _check(drv.cuLaunchKernel(self._fn, int(grid[0]), int(grid[1]), int(grid[2]), int(block[0]), int(block[1]), int(block[2]),
                                  int(shared_mem), st, argv.ctypes.data, 0))


class _Module:
    def __init__(self, source):
        import torch
        from cuda.bindings import driver as drv
        from cuda.bindings import nvrtc
        torch.cuda.current_stream().synchronize()          # torch's primary context is current from here on
        major, minor = torch.cuda.get_device_capability()
        arch = f"sm_{major}{minor}" + ("a" if major >= 9 else "")
        prog = _check(nvrtc.nvrtcCreateProgram(source.encode(), b"kernel.cu", 0, [], []))
        opts = [f"--gpu-architecture={arch}".encode(), b"--std=c++14"]
        res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if int(res[0]) != 0:
            n = _check(nvrtc.nvrtcGetProgramLogSize(prog))
            log = b" " * n
            nvrtc.nvrtcGetProgramLog(prog, log)
            raise RuntimeError("cupy stub: NVRTC compilation failed:\n" + log.decode(errors="replace"))
        n = _check(nvrtc.nvrtcGetCUBINSize(prog))
        cubin = b" " * n
        _check(nvrtc.nvrtcGetCUBIN(prog, cubin))
        self._source = source
        self._module = _check(drv.cuModuleLoadData(cubin))

    def get_function(self, name):
        return _Function(self._module, name, self._source)


def compile_with_cache(source, options=(), arch=None, cache_dir=None, extra_source=None):
    return _Module(source)
