"""Compile the REFERENCE's own CUDA correlation kernels into ``oracle/_ref/*.cubin`` (TEST INFRASTRUCTURE).

The reference keeps its kernels as Python strings (``/root/reference/src/correlation.py:9-104``) that its own
``cupy_kernel()`` (``:237-273``) specialises per tensor shape (``SIZE_n(tensor)`` and ``{{intStride}}`` become
literals) before CuPy JIT-compiles them.  CuPy is not in this image, so this recipe

  1. imports the reference module where it lies (``/root/reference``, with the ``cupy`` name stubbed),
  2. calls the reference's OWN ``cupy_kernel()`` for each shape in ``SHAPES`` (shape-carrying CPU tensors are all it
     needs: it only reads ``.size()``),
  3. pipes the resulting source through ``nvcc -cubin -arch=sm_100a`` (temporary file outside the repo) and
  4. writes only binaries + a manifest (shapes, file names) into ``oracle/_ref/`` -- git-ignored, not gpurun-ignored,
     so the cubins travel to the GPU box where ``oracle/ref_cuda.py`` launches them with the reference's launch
     geometry (``src/correlation.py:305-337``).

No reference source is copied into the repository.  Run:  python oracle/build_ref.py
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

def model_shapes(B, H, W, lowest=1):
    """FunctionCorrelation call shapes of one forward at frame size HxW (src/models.py:173-184: stride 2 at levels 1-3,
    stride 1 at 4-6; matching features 64,64,64,96,128,192 channels)."""
    ch = {1: 64, 2: 64, 3: 64, 4: 96, 5: 128, 6: 192}
    return [(B, ch[l], H >> (l - 1), W >> (l - 1), 2 if l < 4 else 1) for l in range(lowest, 7)]


# (B, C, H, W, stride): operator-level cases (odd sizes at stride 2 -> Ho = ceil(H/2), batch > 1, few channels), then every
# correlation shape of whole forwards: PIV 1x128x128, PIV 2x64x96, Hui 1x64x128 (parity), PIV 1x1024x1024 and Hui 16x448x1024
# (BASELINE configs[2] after estimate()'s resize: timing of the reference's CUDA path)
SHAPES = sorted(set([
    (2, 64, 32, 32, 2),
    (1, 64, 33, 35, 2),
    (1, 64, 6, 7, 2),
    (2, 96, 16, 24, 1),
    (1, 32, 9, 5, 1),
    (3, 64, 64, 96, 2),
] + model_shapes(1, 128, 128) + model_shapes(2, 64, 96) + model_shapes(1, 64, 128, lowest=2)
  + model_shapes(1, 1024, 1024) + model_shapes(16, 448, 1024, lowest=2)))


def shape_tag(B, C, H, W, s):
    return f"b{B}c{C}h{H}w{W}s{s}"


def build(force: bool = False) -> str:
    """Returns the manifest path.  Without ``/root/reference`` (GPU box) the shipped binaries are used as they are."""
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_import as R
    manifest = os.path.join(OUT, "manifest.json")
    if not R.available() or not os.path.isfile(NVCC):
        return manifest
    import torch
    if not force and os.path.isfile(manifest):
        have = json.load(open(manifest))
        if have.get("shapes") == [list(s) for s in SHAPES] and all(
                os.path.isfile(os.path.join(OUT, f)) for e in have["entries"] for f in e["files"].values()):
            return manifest
    os.makedirs(OUT, exist_ok=True)
    _, corr = R.load_reference_models(lambda *a: None)
    entries = []
    with tempfile.TemporaryDirectory() as tmp:
        for (B, C, H, W, s) in SHAPES:
            first = torch.empty(B, C, H, W)
            rbot = torch.empty(B, H + 6 * s, W + 6 * s, C)                      # src/correlation.py:288-291
            top = torch.empty(B, 49, -(-H // s), -(-W // s))                    # :300-301
            srcs = {
                "rearrange": corr.cupy_kernel("kernel_Correlation_rearrange",
                                              {"intStride": s, "input": first, "output": rbot}),
                "updateOutput": corr.cupy_kernel("kernel_Correlation_updateOutput",
                                                 {"intStride": s, "rbot0": rbot, "rbot1": rbot, "top": top}),
            }
            files = {}
            for name, text in srcs.items():
                cu = os.path.join(tmp, f"{name}_{shape_tag(B, C, H, W, s)}.cu")
                with open(cu, "w") as fh:
                    fh.write(text)
                out = os.path.join(OUT, f"{name}_{shape_tag(B, C, H, W, s)}.cubin")
                r = subprocess.run([NVCC, "-cubin", "-arch=sm_100a", "-O3", "-o", out, cu], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for the reference kernel {name}:\n{r.stderr}")
                files[name] = os.path.basename(out)
            entries.append({"shape": [B, C, H, W, s], "files": files})
    with open(manifest, "w") as fh:
        json.dump({"shapes": [list(s) for s in SHAPES], "entries": entries,
                   "source": "compiled from the kernel strings of /root/reference/src/correlation.py:9-104 through the "
                             "reference's own cupy_kernel() templating; nvcc -cubin -arch=sm_100a"}, fh, indent=1)
    return manifest


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
