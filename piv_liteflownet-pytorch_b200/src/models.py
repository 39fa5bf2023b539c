"""Drop-in for the reference's ``src/models.py`` (forward pass).

Same public names (``__all__`` at src/models.py:8 plus the two classes and ``backwarp``), same
constructor arguments, same ``state_dict`` keys/shapes in the same order (strict ``load_state_dict``
works with the reference's ``*.paramOnly`` files), same ``forward(tensorFirst, tensorSecond)``
contract including the in-place mean subtraction of the caller's tensors (src/models.py:321-323) and
the training-mode list-of-levels return (:365-367, :709-713).  The arithmetic runs in hand-written
sm_100a CUDA (libpivlfn.so) through ``pivlfn.model.Engine``; the ``torch.nn`` parameters here are only
containers.  CUDA only: a CPU tensor raises ``NotImplementedError`` exactly like the reference's
correlation op does.
"""
import os
import sys
from collections import OrderedDict
from typing import List, Optional, Tuple, Union

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pivlfn import ops  # noqa: E402
from pivlfn.arch import ModelCfg, param_specs  # noqa: E402
from pivlfn.model import Engine  # noqa: E402
from .correlation import FunctionCorrelation  # noqa: E402,F401  (re-exported like src/models.py:6)

__all__ = ['hui_liteflownet', 'piv_liteflownet']

_HUI_MEAN = (0.411618, 0.434631, 0.454253, 0.410782, 0.433645, 0.452793)


def backwarp(tensorInput: torch.Tensor, tensorFlow: torch.Tensor) -> torch.Tensor:
    """src/models.py:20-35: out[b,c,y,x] = bilinear(in[b,c], x + flow[b,0,y,x], y + flow[b,1,y,x]), zeros outside.
    NCHW in / NCHW out like the reference (the layout change is host-side plumbing around the NHWC kernel)."""
    if not tensorInput.is_cuda:
        raise NotImplementedError()
    B, C, H, W = tensorInput.shape
    cp = (C + 3) & ~3
    x = torch.zeros((B, H, W, cp), device=tensorInput.device, dtype=torch.float32)
    x[..., :C] = tensorInput.permute(0, 2, 3, 1)
    fl = tensorFlow.permute(0, 2, 3, 1).contiguous()
    y = torch.empty_like(x)
    ops.warp(ops.view(x, 0, C), fl, 1.0, ops.view(y, 0, C), B, H, W)
    return y[..., :C].permute(0, 3, 1, 2).contiguous()


class _Node(torch.nn.Module):
    """Parameter container node; children/parameters are attached by name."""


class _LiteFlowNetBase(torch.nn.Module):
    _version_id = 1

    def __init__(self, starting_scale, lowest_level, rgb_mean):
        super().__init__()
        rgb_mean = list(rgb_mean)
        self.MEAN = [rgb_mean[:3], rgb_mean[3:]]
        self.lowest_level = int(lowest_level)
        self.PLEVELS = 6
        self.SCALEFACTOR = [float(starting_scale) / (2.0 ** level) for level in range(self.PLEVELS + 1)]
        self._cfg = ModelCfg(type(self).__name__, self._version_id, float(starting_scale), int(lowest_level),
                             tuple(float(m) for m in rgb_mean))
        gen = torch.Generator().manual_seed(0)
        for name, shape in param_specs(self._cfg).items():
            node = self
            parts = name.split('.')
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            if parts[-1] == 'weight':
                fan_in = shape[1] * shape[2] * shape[3]
                bound = 1.0 / max(fan_in, 1) ** 0.5      # kaiming_uniform(a=sqrt(5)) bound of torch.nn.Conv2d
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            else:
                t = torch.zeros(shape)
            node.register_parameter(parts[-1], torch.nn.Parameter(t))
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self.precision = None      # None -> env PIVLFN_PRECISION or the fp32-equivalent default

    # -- engine cache: rebuilt whenever a parameter is replaced, modified in place or moved -------------------
    def _params_key(self):
        return tuple((p.data_ptr(), p._version, p.device) for p in self.parameters()) + (self.precision,)

    def engine(self) -> Engine:
        key = self._params_key()
        if self._engine is None or key != self._engine_key:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise NotImplementedError("pivlfn: the model must be on a CUDA device (no CPU path)")
            self._engine = Engine(self._cfg, {k: v for k, v in self.state_dict().items()}, dev, self.precision)
            self._engine_key = key
        return self._engine

    def forward(self, img1: torch.Tensor, img2: torch.Tensor):
        if not img1.is_cuda:
            raise NotImplementedError()
        if self.training:
            im_shape = (img1.shape[2], img1.shape[3])
            _, levels = self.engine().forward(img1, img2, return_levels=True)
            return self._training_output(levels, im_shape)
        return self.engine().forward(img1, img2)

    def _training_output(self, levels, im_shape):
        return levels


class LiteFlowNet(_LiteFlowNetBase):
    """src/models.py:39-370."""
    _version_id = 1

    def __init__(self, starting_scale: int = 40, lowest_level: int = 2,
                 rgb_mean: Union[Tuple[float, ...], List[float]] = _HUI_MEAN) -> None:
        super().__init__(starting_scale, lowest_level, rgb_mean)


class LiteFlowNet2(_LiteFlowNetBase):
    """src/models.py:373-716."""
    _version_id = 2

    def __init__(self, starting_scale: int = 40, lowest_level: int = 3,
                 rgb_mean: Union[Tuple[float, ...], List[float]] = _HUI_MEAN) -> None:
        super().__init__(starting_scale, lowest_level, rgb_mean)

    def _training_output(self, levels, im_shape):
        # src/models.py:709-713: LiteFlowNet2 appends the last flow upsampled to the input size
        levels.append([ops.resize_bilinear(levels[-1][2], im_shape[0], im_shape[1])])
        return levels


def hui_liteflownet(params: Optional[OrderedDict] = None, version: int = 1):
    """src/models.py:719-740."""
    if version == 1:
        model = LiteFlowNet(rgb_mean=_HUI_MEAN)
    elif version == 2:
        model = LiteFlowNet2()
    else:
        raise ValueError(f'Wrong input of model version (input = {version})! Choose between version 1 or 2 only!')
    if params is not None:
        model.load_state_dict(params)
    return model


def piv_liteflownet(params: Optional[OrderedDict] = None, version: int = 1):
    """src/models.py:743-766."""
    if version == 1:
        model = LiteFlowNet(starting_scale=10, lowest_level=1,
                            rgb_mean=(0.173935, 0.180594, 0.192608, 0.172978, 0.179518, 0.191300))
    elif version == 2:
        model = LiteFlowNet2(starting_scale=10, lowest_level=2,
                             rgb_mean=(0.194286, 0.190633, 0.191766, 0.194220, 0.190595, 0.191701))
    else:
        raise ValueError(f'Wrong input of model version (input = {version})! Choose between version 1 or 2 only!')
    if params is not None:
        model.load_state_dict(params)
    return model
