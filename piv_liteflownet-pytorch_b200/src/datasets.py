"""Drop-in names of the inference datasets of the reference's ``src/datasets.py``: ``Run`` (:438-487), the input contract of
``run.py`` -- RGB float [0,1] ``[3,H,W]`` image pairs from ``*_img1.ext`` / ``*_img2.ext`` files (``is_pair``) or from n
sequential frames (n-1 pairs) -- and ``InferenceEval`` (:491-564), image pairs with their ground-truth ``.flo`` for
evaluation.  Both sit on ``pivlfn.io`` (pair discovery, decoding); the batch driver itself uses ``pivlfn.io.BatchReader``,
which feeds whole batches through pinned memory.  The training datasets of that file (HDF5 / LMDB / augmentation) are out of scope."""
import json
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from pivlfn import io as _io


def _chw_float(rgb_u8: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(rgb_u8.transpose(2, 0, 1))).to(torch.float32).div_(255.0)


class Run(torch.utils.data.Dataset):
    def __init__(self, root: str, is_pair: bool = True, n_images: int = -1, start_at: int = 0) -> None:
        self.index = _io.PairIndex(root, is_pair, n_images, start_at)
        self.image_list = [[p.first, p.second] for p in self.index]
        self.name_list = [p.stem for p in self.index]
        self.size = len(self.index)

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, index: int) -> Tuple[List[torch.Tensor], str]:
        p = self.index[index % self.size]
        return [_chw_float(_io.decode_rgb(p.first)), _chw_float(_io.decode_rgb(p.second))], p.stem


def _center_crop(a: np.ndarray, size) -> np.ndarray:
    """flow_transforms.Crop(size, crop_type='center') of the reference (src/flow_transforms.py:285-371): ``size[0]`` is the crop
    HEIGHT and ``size[1]`` the crop WIDTH.  (``InferenceEval`` derives its size from PIL's (width, height), src/datasets.py:533-537,
    so for non-square frames the reference crops with the two swapped; mirrored, not corrected.)  A crop larger than the frame
    needs padding, which the reference refuses when no padding colour is given."""
    ch, cw = int(size[0]), int(size[1])
    h, w = a.shape[:2]
    if ch > h or cw > w:
        raise RuntimeError("flow_transforms.Crop() need padding while padding argument is None\n")
    y0, x0 = (h - ch) // 2, (w - cw) // 2
    return a[y0:y0 + ch, x0:x0 + cw]


class InferenceEval(torch.utils.data.Dataset):
    """Every ``<stem>_flow.flo`` (any ``*.flo``) under ``root`` that has ``<stem>_img1.*`` / ``<stem>_img2.*`` next to it; or the
    ``set_type`` list of a json file.  Items: ``([img1, img2], [flow])`` as CHW float tensors, centre-cropped to
    ``inference_size`` (width, height) -- or to the largest multiple of 64 when that is negative or the frames are not
    multiples of 64 (src/datasets.py:533-537).  Files with 'test' in their path are skipped like in the reference."""

    def __init__(self, inference_size: Tuple = (-1, -1), root: str = '', set_type: Optional[str] = None) -> None:
        self.render_size = list(inference_size)
        suffix = os.path.splitext(root)[1]
        if suffix and set_type is not None:
            if suffix != '.json':
                raise ValueError(f'Only json format is currently supported! Change the input path ({root}).')
            with open(root) as fh:
                flows = list(json.load(fh)[set_type])
        else:
            flows = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(root) for f in fs if f.endswith('.flo'))
        self.image_list: List[List[str]] = []
        self.flow_list: List[str] = []
        for flo in flows:
            if 'test' in flo:
                continue
            stem = os.path.splitext(flo)[0]
            stem = stem[:stem.rfind('_')] if '_' in os.path.basename(stem) else stem
            for e in _io.IMAGE_SUFFIXES:
                a, b = f'{stem}_img1.{e}', f'{stem}_img2.{e}'
                if os.path.isfile(a) and os.path.isfile(b) and os.path.isfile(flo):
                    self.image_list.append([a, b])
                    self.flow_list.append(flo)
                    break
        self.size = len(self.image_list)
        self.frame_size = None
        if self.size:
            h, w, _ = _io.decode_rgb(self.image_list[0][0]).shape
            self.frame_size = (w, h)                           # PIL's (width, height)
            if self.render_size[0] < 0 or self.render_size[1] < 0 or w % 64 or h % 64:
                self.render_size = [(w // 64) * 64, (h // 64) * 64]

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, index: int):
        index = index % self.size
        imgs = [_chw_float(_center_crop(_io.decode_rgb(p), self.render_size)) for p in self.image_list[index]]
        flow = _center_crop(_io.read_flo(self.flow_list[index]), self.render_size)
        return imgs, [torch.from_numpy(np.ascontiguousarray(flow.transpose(2, 0, 1)))]
