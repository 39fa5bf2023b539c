"""Drop-in for the flow I/O part of the reference's ``src/utils_plot.py``: Middlebury ``.flo`` files
(``read_flow`` :26-73, ``write_flow`` :120-158, ``flowname_modifier`` :310-318).  Wire format: float32 tag 202021.25
("PIEH"), int32 width, int32 height, then height*width*bands float32 in HWC order (bands = 2, or 3 for stereo 2D3C).
The plotting helpers of that file (quiver, colour wheel) are visualisation and out of scope."""
import io
import os

import numpy as np

TAG_STRING = 'PIEH'
TAG_FLOAT = 202021.25


def read_flow(filename, use_stereo: bool = False) -> np.ndarray:
    if not isinstance(filename, io.BufferedReader):
        if not isinstance(filename, str):
            raise AssertionError("Input [{p}] is not a string".format(p=filename))
        if not os.path.isfile(filename):
            raise AssertionError("Path [{p}] does not exist".format(p=filename))
        if not filename.split('.')[-1] == 'flo':
            raise AssertionError("File extension [flo] required, [{f}] given".format(f=filename.split('.')[-1]))
        flo = open(filename, 'rb')
    else:
        flo = filename
    with flo:
        tag = np.frombuffer(flo.read(4), np.float32, count=1)[0]
        if not TAG_FLOAT == tag:
            raise AssertionError("Wrong Tag [{t}]".format(t=tag))
        width = int(np.frombuffer(flo.read(4), np.int32, count=1)[0])
        if not (0 < width < 100000):
            raise AssertionError("Illegal width [{w}]".format(w=width))
        height = int(np.frombuffer(flo.read(4), np.int32, count=1)[0])
        if not (0 < height < 100000):
            raise AssertionError("Illegal height [{h}]".format(h=height))
        bands = 3 if use_stereo else 2
        data = np.frombuffer(flo.read(bands * width * height * 4), np.float32, count=bands * width * height)
    return data.reshape(height, width, bands).copy()


def write_flow(flow: np.ndarray, filename: str):
    assert type(filename) is str, "file is not str (%r)" % str(filename)
    assert filename[-4:] == '.flo', "file ending is not .flo (%r)" % filename[-4:]
    height, width, n_bands = flow.shape
    assert n_bands == 2 or n_bands == 3, "Number of bands = %r != (2 or 3)" % n_bands
    with open(filename, 'wb') as f:
        np.array([TAG_FLOAT], dtype=np.float32).tofile(f)
        np.array([width], dtype=np.int32).tofile(f)
        np.array([height], dtype=np.int32).tofile(f)
        np.ascontiguousarray(flow, dtype=np.float32).tofile(f)


def flowname_modifier(indir: str, outdir: str, ext: str = '_out.flo', pair: bool = True) -> str:
    out_name = os.path.splitext(os.path.basename(indir))[0]
    if pair:
        out_name = str(out_name.rsplit('_', 1)[0]) + ext
    else:
        out_name += ext
    return os.path.join(outdir, out_name)
