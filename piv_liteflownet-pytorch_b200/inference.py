"""Drop-in for the reference's ``inference.py`` entry points on the forward-pass hot path:
``estimate(net, img1, img2, tensor=False)`` (inference.py:30-67) and ``Inference.parser`` (:202-213).
The video / webcam / quiver-plot drivers of the reference (:81-200) need cv2 GUI, imutils and matplotlib
and are outside the accelerated path.

Both bilinear resizes of ``estimate`` (to a multiple of 32 before the network, back to the input size
after it, with the u,v rescale of :60-61 folded in) run in ``pivlfn_resize_bilinear_nchw``; when the
input is already a multiple of 32 they are the identity and are skipped.
"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pivlfn import ops  # noqa: E402
from src.models import hui_liteflownet, piv_liteflownet  # noqa: E402,F401


def estimate(net: torch.nn.Module, img1: torch.Tensor, img2: torch.Tensor, tensor: bool = False):
    assert (img1.size(2) == img2.size(2))
    assert (img1.size(3) == img2.size(3))
    input_width, input_height = img1.size(3), img1.size(2)
    adaptive_width = int(math.floor(math.ceil(input_width / 32.0) * 32.0))
    adaptive_height = int(math.floor(math.ceil(input_height / 32.0) * 32.0))
    scale_width = float(input_width) / float(adaptive_width)
    scale_height = float(input_height) / float(adaptive_height)
    same = (adaptive_width == input_width and adaptive_height == input_height)
    with torch.set_grad_enabled(False):
        net.eval()
        if same:
            # interpolate to the same size is the identity, but the reference's forward then mutates the
            # interpolate OUTPUT, not the caller's tensor: keep that by handing the network a copy
            tensor_im1, tensor_im2 = img1.contiguous().clone(), img2.contiguous().clone()
        else:
            tensor_im1 = ops.resize_bilinear(img1.contiguous(), adaptive_height, adaptive_width)
            tensor_im2 = ops.resize_bilinear(img2.contiguous(), adaptive_height, adaptive_width)
        tensor_raw_output = net(tensor_im1, tensor_im2)
        oh, ow = tensor_raw_output.shape[2], tensor_raw_output.shape[3]
        if (oh, ow) == (input_height, input_width):
            tensor_flow = tensor_raw_output
        else:
            tensor_flow = ops.resize_bilinear(tensor_raw_output, input_height, input_width, scale_width, scale_height)
    if tensor:
        return tensor_flow.detach()
    out = tensor_flow.squeeze().permute(1, 2, 0).cpu().numpy()
    eng = getattr(net, "_engine", None)
    if eng is not None and eng.check_range(wait=True):
        # the (small) forward above left the fp16 range, which its deferred check only reports now that the stream has been
        # synchronised anyway: the engine has switched to tf32c, repeat once
        return estimate(net, img1, img2, tensor=False)
    return out


class Inference:
    """The constructor and the ``parser`` static method of the reference's ``Inference`` class (inference.py:70-79,202-213).
    Like the reference, ``device`` defaults to 'cpu' -- and like the reference's own CPU path (its correlation raises
    NotImplementedError, src/correlation.py:339-340) a CPU call fails loudly: pass ``device='cuda'``."""

    def __init__(self, net, netname=None, output_dir='./outputs', device='cpu'):
        self.netname = 'test' if netname is None else os.path.splitext(os.path.basename(netname))[0]
        self.default = os.path.join(output_dir, self.netname)
        self.device = device if torch.cuda.is_available() else 'cpu'
        self.net = net

    @staticmethod
    def parser(net, im1, im2, device='cpu'):
        assert im1.size == im2.size
        tensor_im1 = _to_tensor(im1).to(device)
        tensor_im2 = _to_tensor(im2).to(device)
        C, H, W = tensor_im1.size()
        return estimate(net, tensor_im1.view(1, C, H, W), tensor_im2.view(1, C, H, W))


def _to_tensor(im) -> torch.Tensor:
    """torchvision.transforms.ToTensor for a PIL image or an HxW / HxWxC uint8 array: [C,H,W] fp32 in [0,1]."""
    a = np.asarray(im)
    if a.ndim == 2:
        a = a[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))
    return t.to(torch.float32).div(255.0) if t.dtype == torch.uint8 else t.to(torch.float32)
