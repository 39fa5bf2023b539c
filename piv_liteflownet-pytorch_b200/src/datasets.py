"""Drop-in for the ``Run`` dataset of the reference's ``src/datasets.py`` (:438-487), the input contract of ``run.py``:
RGB float [0,1] ``[3,H,W]`` image pairs, either ``*_img1.ext`` / ``*_img2.ext`` files (``is_pair``) or n sequential frames
giving n-1 pairs.  The training datasets of that file (HDF5 / LMDB / augmentation) are out of scope."""
import os
from typing import List, Tuple

import numpy as np
import torch

from .utils_data import image_files_from_folder, read_gen


def _to_tensor(im) -> torch.Tensor:
    a = np.asarray(im)
    if a.ndim == 2:
        a = a[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))
    return t.to(torch.float32).div(255.0) if t.dtype == torch.uint8 else t.to(torch.float32)


class Run(torch.utils.data.Dataset):
    def __init__(self, root: str, is_pair: bool = True, n_images: int = -1, start_at: int = 0) -> None:
        if not os.path.isdir(root):
            raise ValueError(f"Input image directory is NOT found! '{root}'")
        file_list = image_files_from_folder(root, pair=is_pair, n_images=n_images, start_at=start_at, upper=False)
        self.image_list, self.name_list = [], []
        prev_file = None
        for file in file_list:
            if is_pair:
                imbase, imext = os.path.splitext(os.path.basename(str(file)))
                fbase = imbase.rsplit('_', 1)[0]
                img1, img2 = file, os.path.join(root, str(fbase) + '_img2' + imext)
            else:
                if prev_file is None:
                    prev_file = file
                    continue
                img1, img2 = prev_file, file
                fbase = os.path.splitext(os.path.basename(str(img1)))[0]
                prev_file = file
            if not os.path.isfile(img1) or not os.path.isfile(img2):
                continue
            self.image_list.append([img1, img2])
            self.name_list.append(fbase)
        self.size = len(self.name_list)

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, index: int) -> Tuple[List[torch.Tensor], str]:
        index = index % self.size
        img1 = read_gen(self.image_list[index][0])
        img2 = read_gen(self.image_list[index][1])
        return [_to_tensor(img1), _to_tensor(img2)], self.name_list[index]
