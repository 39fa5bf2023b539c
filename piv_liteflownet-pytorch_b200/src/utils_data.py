"""Drop-in names of the two input helpers of the reference's ``src/utils_data.py`` that the batch driver uses
(``image_files_from_folder`` :12-33, ``read_gen`` :46-56), over ``pivlfn.io``."""
import os
from typing import List, Tuple

import numpy as np

from pivlfn import io as _io


def image_files_from_folder(folder: str, pair: bool = True, upper: bool = True, n_images: int = -1, start_at: int = 0,
                            extensions: Tuple[str, ...] = _io.IMAGE_SUFFIXES) -> List[str]:
    """Image files of ``folder`` (only the ``*_img1.*`` ones when ``pair``), grouped by extension in the given order."""
    names = sorted(os.listdir(folder))
    found: List[str] = []
    for e in extensions:
        for suffix in ((e, e.upper()) if upper else (e,)):
            tail = ('_img1.' if pair else '.') + suffix
            found.extend(os.path.join(folder, n) for n in names if n.endswith(tail))
    stop = None if n_images < 0 else start_at + n_images
    return found[start_at:stop]


def read_gen(file_name: str, im_extensions: Tuple[str, ...] = tuple('.' + s for s in _io.IMAGE_SUFFIXES)):
    """A PIL RGB image for image files, an array for ``.bin`` / ``.raw`` / ``.flo``, ``[]`` otherwise."""
    kind = os.path.splitext(file_name)[-1]
    if kind in im_extensions:
        import PIL.Image
        return PIL.Image.open(file_name).convert('RGB')
    if kind == '.flo':
        return _io.read_flo(file_name)
    if kind in ('.bin', '.raw'):
        return np.load(file_name)
    return []
