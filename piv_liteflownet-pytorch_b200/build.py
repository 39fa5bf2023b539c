"""Build libpivlfn.so (all CUDA kernels + the C ABI) in-tree for sm_100a with nvcc.

    python piv_liteflownet-pytorch_b200/build.py [--force]

No GPU is needed (nvcc cross-compiles).  The library links cudart statically and resolves the
one driver entry point it needs (cuTensorMapEncodeTiled) at run time through
cudaGetDriverEntryPoint, so it also loads on a machine without a CUDA driver.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "pivlfn", "libpivlfn.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/pivlfn.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()

    def fresh():
        return os.path.isfile(OUT) and os.path.isfile(stamp) and open(stamp).read() == dig

    if not force and fresh():
        return OUT
    if not os.path.isfile(NVCC):
        if os.path.isfile(OUT):
            return OUT          # GPU box without a changed source tree: use the shipped library
        raise RuntimeError("nvcc not found and no prebuilt libpivlfn.so")
    # N torchrun ranks import the package at the same moment: one of them builds, the others wait on the lock and find
    # the stamp fresh.  Objects and the library are written under private names and renamed into place, so nobody can
    # dlopen or link a half-written file.
    import fcntl
    with open(os.path.join(OBJ, "lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return OUT
            tag = ".%d.tmp" % os.getpid()

            def cc(src):
                obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
                r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj + tag], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
                os.replace(obj + tag, obj)
                return obj, r.stderr

            with ThreadPoolExecutor(max_workers=8) as ex:
                res = list(ex.map(cc, sources()))
            if verbose:
                for _, log in res:
                    sys.stderr.write(log)
            r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT + tag,
                                *[o for o, _ in res], "-cudart", "static"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n" + r.stderr)
            os.replace(OUT + tag, OUT)
            with open(stamp + tag, "w") as fh:
                fh.write(dig)
            os.replace(stamp + tag, stamp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
