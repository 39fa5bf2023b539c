"""Drop-in names of the flow I/O helpers of the reference's ``src/utils_plot.py`` (``read_flow`` :26-73, ``write_flow``
:120-158, ``flowname_modifier`` :310-318) over ``pivlfn.io`` (which documents the Middlebury ``.flo`` wire format).
The plotting helpers of that file (quiver, colour wheel) are visualisation and out of scope."""
import os

import numpy as np

from pivlfn import io as _io

TAG_FLOAT = _io.FLO_MAGIC
TAG_STRING = 'PIEH'


def read_flow(filename, use_stereo: bool = False) -> np.ndarray:
    """[H, W, 2] float32, or [H, W, 3] for the stereo 2D3C files (``use_stereo``).  Accepts a path or an open binary file;
    malformed input raises AssertionError like the reference."""
    return _io.read_flo(filename, bands=3 if use_stereo else 2)


def write_flow(flow: np.ndarray, filename: str):
    _io.write_flo(filename, flow)


def flowname_modifier(indir: str, outdir: str, ext: str = '_out.flo', pair: bool = True) -> str:
    """Output path for the flow of input ``indir``: ``<outdir>/<stem><ext>``, where for ``pair`` inputs (``x_img1.tif``) the
    ``_img1`` part of the stem is dropped."""
    stem = os.path.splitext(os.path.basename(indir))[0]
    if pair and '_' in stem:
        stem = stem[:stem.rfind('_')]
    return os.path.join(outdir, stem + ext)
