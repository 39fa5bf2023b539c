"""Drop-in for the two input helpers of the reference's ``src/utils_data.py`` that the batch driver uses:
``image_files_from_folder`` (:12-33) and ``read_gen`` (:46-56)."""
import os
from glob import glob
from typing import List, Tuple

import numpy as np


def image_files_from_folder(folder: str, pair: bool = True, upper: bool = True, n_images: int = -1, start_at: int = 0,
                            extensions: Tuple[str, ...] = ('jpg', 'jpeg', 'png', 'bmp', 'tif', 'ppm')) -> List[str]:
    img_files: List[str] = []
    for ext in extensions:
        pattern = f'*_img1.{ext}' if pair else f'*.{ext}'
        img_files += sorted(glob(os.path.join(folder, pattern)))
        if upper:
            pattern = f'*_img1.{ext.upper()}' if pair else f'*.{ext.upper()}'
            img_files += sorted(glob(os.path.join(folder, pattern)))
    return img_files[start_at:] if n_images < 0 else img_files[start_at:start_at + n_images]


def read_gen(file_name: str, im_extensions: Tuple[str, ...] = ('.jpg', '.jpeg', '.png', '.bmp', '.tif', '.ppm')):
    ext = os.path.splitext(file_name)[-1]
    if ext in im_extensions:
        import PIL.Image
        return PIL.Image.open(file_name).convert('RGB')
    if ext in ('.bin', '.raw'):
        return np.load(file_name)
    if ext == '.flo':
        from .utils_plot import read_flow
        return read_flow(file_name)
    return []
