"""CPU oracle for the PIV-LiteFlowNet / LiteFlowNet forward pass.

THIS FILE IS TEST INFRASTRUCTURE.  It is a plain fp32 restatement, on the CPU,
of the algorithm in the reference's ``src/models.py`` / ``src/correlation.py``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline) -- never as part of the shipped CUDA path.

Pinning: the restatement is checked (``tests/test_oracle_golden.py``) against
golden vectors produced by importing the *unmodified* reference
``/root/reference/src/models.py`` in the build container
(``tests/golden/make_golden.py``), and the literal correlation emulation below
is checked on the GPU box against cubins compiled from the reference's own CUDA
kernel strings (``oracle/build_ref.py`` -> ``oracle/_ref/``).  The reference ships
no tests and its pretrained weights are absent from the mount, so the only
upstream golden vector (``images/demo/DNS_turbulence_out.flo``) cannot be
reproduced: beyond the vectors above, parity is otherwise unpinned.

The arithmetic that lives in third-party code (torch 1.4 / cuDNN convolution,
``grid_sample``, ``interpolate``, ``unfold``) is restated with the same torch
ops on the CPU in fp32; the custom CUDA correlation is restated from its
kernel source.

All tensors are NCHW fp32, exactly as in the reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LRELU = 0.1  # negative_slope everywhere (src/models.py:72,178)
KSIZE = [0, 7, 7, 5, 5, 3, 3]  # per-level kernel size (src/models.py:161,205,225)

MODEL_CFG = {
    # name: (class, starting_scale, lowest_level, rgb_mean)   src/models.py:729-730,754-758
    "piv": ("LiteFlowNet", 10.0, 1,
            (0.173935, 0.180594, 0.192608, 0.172978, 0.179518, 0.191300)),
    "hui": ("LiteFlowNet", 40.0, 2,
            (0.411618, 0.434631, 0.454253, 0.410782, 0.433645, 0.452793)),
    "piv2": ("LiteFlowNet2", 10.0, 2,
             (0.194286, 0.190633, 0.191766, 0.194220, 0.190595, 0.191701)),
    "hui2": ("LiteFlowNet2", 40.0, 3,
             (0.411618, 0.434631, 0.454253, 0.410782, 0.433645, 0.452793)),
}


def scalefactor(starting_scale: float) -> List[float]:
    """SCALEFACTOR[l] = starting_scale / 2**l  (src/models.py:58-63)."""
    return [float(starting_scale) / (2.0 ** l) for l in range(7)]


def lrelu(x: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(x, LRELU)


# --------------------------------------------------------------------------
# backwarp  (src/models.py:20-35)
# --------------------------------------------------------------------------
def backwarp(inp: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """Bilinear backward warp, zeros outside, built exactly like the reference:
    a linspace(-1,1) grid plus the flow divided by ((W-1)/2, (H-1)/2), sampled
    with ``grid_sample(align_corners=True)``.  (The reference caches the grid in
    a process-global dict and forces ``.cuda()``; neither is arithmetic.)"""
    B, _, H, W = flow.shape
    hor = torch.linspace(-1.0, 1.0, W, dtype=flow.dtype, device=flow.device).view(1, 1, 1, W).expand(B, -1, H, -1)
    ver = torch.linspace(-1.0, 1.0, H, dtype=flow.dtype, device=flow.device).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([hor, ver], 1)
    nflow = torch.cat([flow[:, 0:1] / ((inp.shape[3] - 1.0) / 2.0),
                       flow[:, 1:2] / ((inp.shape[2] - 1.0) / 2.0)], 1)
    return F.grid_sample(inp, (grid + nflow).permute(0, 2, 3, 1), mode="bilinear",
                         padding_mode="zeros", align_corners=True)


def backwarp_direct(inp: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """Same operator written as explicit pixel-space bilinear taps
    (out[b,c,y,x] = bilinear(inp[b,c], x+u, y+v), zero outside); used to show
    the CUDA kernels' formulation equals the grid_sample one to fp32 round-off."""
    B, C, H, W = inp.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=inp.dtype), torch.arange(W, dtype=inp.dtype), indexing="ij")
    sx = xs[None] + flow[:, 0]
    sy = ys[None] + flow[:, 1]
    x0 = torch.floor(sx)
    y0 = torch.floor(sy)
    ax = sx - x0
    ay = sy - y0
    out = torch.zeros_like(inp)
    flat = inp.reshape(B, C, H * W)
    for oy, ox, w in ((0, 0, (1 - ay) * (1 - ax)), (0, 1, (1 - ay) * ax),
                      (1, 0, ay * (1 - ax)), (1, 1, ay * ax)):
        yy = (y0 + oy).long()
        xx = (x0 + ox).long()
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, H * W).expand(-1, C, -1)
        val = torch.gather(flat, 2, idx).reshape(B, C, H, W)
        out = out + val * (w * ok.to(inp.dtype))[:, None]
    return out


# --------------------------------------------------------------------------
# correlation  (src/correlation.py:9-34 rearrange, :36-104 updateOutput, :285-344 launch)
# --------------------------------------------------------------------------
def correlation(first: torch.Tensor, second: torch.Tensor, stride: int) -> torch.Tensor:
    """out[b,(dy+3)*7+(dx+3),y,x] = (1/C) sum_c f1[b,c,y*s,x*s] * f2[b,c,y*s+dy*s,x*s+dx*s],
    dy,dx in [-3,3], zeros outside, out size ceil(H/s) x ceil(W/s); channel index has dx
    fastest (``top_channel % 7`` is the x offset, src/correlation.py:73-74) and the sum is
    divided by C (:98-100).  Vectorised pad-and-shift formulation."""
    B, C, H, W = first.shape
    s = int(stride)
    Ho, Wo = int(math.ceil(H / s)), int(math.ceil(W / s))
    pad = 3 * s
    f2p = F.pad(second, (pad, pad, pad, pad))
    f1s = first[:, :, ::s, ::s]
    out = first.new_zeros(B, 49, Ho, Wo)
    for iy in range(7):
        for ix in range(7):
            oy, ox = iy * s, ix * s  # = pad + (iy-3)*s
            win = f2p[:, :, oy:oy + (Ho - 1) * s + 1:s, ox:ox + (Wo - 1) * s + 1:s]
            out[:, iy * 7 + ix] = (f1s * win).sum(1) / float(C)
    return out


def correlation_literal(first: np.ndarray, second: np.ndarray, stride: int) -> np.ndarray:
    """Literal emulation (numpy, fp32) of the two reference kernels including their padded
    NHWC intermediate ``rbot`` and the 32-lane partial-sum order of
    ``kernel_Correlation_updateOutput`` (lane l accumulates channels l, l+32, ... in order,
    then lane 0 adds the 32 partials serially, then divides by C).  Small inputs only."""
    first = np.ascontiguousarray(first, dtype=np.float32)
    second = np.ascontiguousarray(second, dtype=np.float32)
    B, C, H, W = first.shape
    s = int(stride)
    Hp, Wp = H + 6 * s, W + 6 * s
    # kernel_Correlation_rearrange: NCHW -> zero padded NHWC (src/correlation.py:28-32)
    rbot0 = np.zeros((B, Hp, Wp, C), np.float32)
    rbot1 = np.zeros((B, Hp, Wp, C), np.float32)
    rbot0[:, 3 * s:3 * s + H, 3 * s:3 * s + W, :] = first.transpose(0, 2, 3, 1)
    rbot1[:, 3 * s:3 * s + H, 3 * s:3 * s + W, :] = second.transpose(0, 2, 3, 1)
    Ho, Wo = int(math.ceil(H / s)), int(math.ceil(W / s))
    top = np.zeros((B, 49, Ho, Wo), np.float32)
    nlane = 32
    for b in range(B):
        for by in range(Ho):
            for bx in range(Wo):
                x1 = (bx + 3) * s  # src/correlation.py:48-49
                y1 = (by + 3) * s
                patch = rbot0[b, y1, x1, :]
                for tc in range(49):
                    s2o = (tc % 7 - 3) * s
                    s2p = (tc // 7 - 3) * s
                    other = rbot1[b, y1 + s2p, x1 + s2o, :]
                    prod = patch * other  # fp32 products
                    partial = np.zeros(nlane, np.float32)
                    for ch0 in range(0, C, nlane):
                        seg = prod[ch0:ch0 + nlane]
                        partial[:seg.shape[0]] = partial[:seg.shape[0]] + seg
                    tot = np.float32(0.0)
                    for l in range(nlane):
                        tot = np.float32(tot + partial[l])
                    top[b, tc, by, bx] = tot / np.float32(C)
    return top


# --------------------------------------------------------------------------
# network pieces, driven by a reference-layout state_dict
# --------------------------------------------------------------------------
def _conv(x, sd, prefix, stride=1, padding=0, act=True):
    y = F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)
    return lrelu(y) if act else y


def features(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> List[torch.Tensor]:
    """NetC (src/models.py:66-116): six feature maps, every conv followed by LeakyReLU(0.1)."""
    f1 = _conv(x, sd, "NetC.conv1.0", 1, 3)
    t = _conv(f1, sd, "NetC.conv2.0", 2, 1)
    t = _conv(t, sd, "NetC.conv2.2", 1, 1)
    f2 = _conv(t, sd, "NetC.conv2.4", 1, 1)
    t = _conv(f2, sd, "NetC.conv3.0", 2, 1)
    f3 = _conv(t, sd, "NetC.conv3.2", 1, 1)
    t = _conv(f3, sd, "NetC.conv4.0", 2, 1)
    f4 = _conv(t, sd, "NetC.conv4.2", 1, 1)
    f5 = _conv(f4, sd, "NetC.conv5.0", 2, 1)
    f6 = _conv(f5, sd, "NetC.conv6.0", 2, 1)
    return [f1, f2, f3, f4, f5, f6]


def _chain(x, sd, prefix, idxs, last_k):
    """conv_M / conv_S: 3x3 conv+lrelu ..., then a KxK 32->2 flow head without activation
    (src/models.py:154-163,197-207; LiteFlowNet2: :487-500,534-548)."""
    for i in idxs[:-1]:
        x = _conv(x, sd, f"{prefix}.{i}", 1, 1)
    return _conv(x, sd, f"{prefix}.{idxs[-1]}", 1, last_k // 2, act=False)


def _seq_idxs(sd, prefix):
    return sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")})


def matching(sd, i, level, scale, f1, f2, xflow, corr_fn=correlation):
    """NetE-M (src/models.py:165-187)."""
    p = f"NetE_M.{i}"
    if xflow is not None:
        xflow = F.conv_transpose2d(xflow, sd[p + ".upConv_M.weight"], None, stride=2, padding=1, groups=2)
        f2 = backwarp(f2, xflow * scale)
    if level >= 4:  # upCorr_M is None (src/models.py:148-149)
        c = lrelu(corr_fn(f1, f2, 1))
    else:
        c = F.conv_transpose2d(lrelu(corr_fn(f1, f2, 2)), sd[p + ".upCorr_M.weight"], None,
                               stride=2, padding=1, groups=49)
    out = _chain(c, sd, p + ".conv_M", _seq_idxs(sd, p + ".conv_M"), KSIZE[level])
    return out + (xflow if xflow is not None else 0.0)


def subpixel(sd, i, level, scale, f1, f2, xflow):
    """NetE-S (src/models.py:209-217)."""
    p = f"NetE_S.{i}"
    f2w = backwarp(f2, xflow * scale)
    out = _chain(torch.cat([f1, f2w, xflow], 1), sd, p + ".conv_S", _seq_idxs(sd, p + ".conv_S"), KSIZE[level])
    return out + xflow


def regularization(sd, i, level, scale, img1, img2, feat1, xflow_s):
    """NetE-R (src/models.py:274-303)."""
    p = f"NetE_R.{i}"
    K = KSIZE[level]
    B = xflow_s.shape[0]
    rm = xflow_s - xflow_s.view(B, 2, -1).mean(2, True).view(B, 2, 1, 1)
    warp = backwarp(img2, xflow_s * scale)
    norm = (img1 - warp).pow(2.0).sum(1, True).sqrt()
    feat = _conv(feat1, sd, p + ".moduleFeat.0", 1, 0) if level < 5 else feat1
    x = torch.cat([norm, rm, feat], 1)
    for j in (0, 2, 4, 6, 8, 10):
        x = _conv(x, sd, f"{p}.conv_R.{j}", 1, 1)
    if level < 5:  # separable (K,1) then (1,K), no activation (src/models.py:252-261)
        x = _conv(x, sd, p + ".conv_dist_R.0", 1, (K // 2, 0), act=False)
        x = _conv(x, sd, p + ".conv_dist_R.1", 1, (0, K // 2), act=False)
    else:
        x = _conv(x, sd, p + ".conv_dist_R.0", 1, K // 2, act=False)
    negsq = x.pow(2.0).neg()
    dist = (negsq - negsq.max(1, True)[0]).exp()
    div = dist.sum(1, True).reciprocal()
    ux = F.unfold(xflow_s[:, 0:1], kernel_size=K, stride=1, padding=(K - 1) // 2).view_as(dist)
    uy = F.unfold(xflow_s[:, 1:2], kernel_size=K, stride=1, padding=(K - 1) // 2).view_as(dist)
    fx = F.conv2d(dist * ux, sd[p + ".moduleScaleX.weight"], sd[p + ".moduleScaleX.bias"]) * div
    fy = F.conv2d(dist * uy, sd[p + ".moduleScaleY.weight"], sd[p + ".moduleScaleY.bias"]) * div
    return torch.cat([fx, fy], 1)


def forward(sd: Dict[str, torch.Tensor], img1: torch.Tensor, img2: torch.Tensor, model: str = "piv",
            corr_fn=correlation, return_levels: bool = False):
    """LiteFlowNet.forward / LiteFlowNet2.forward in eval mode (src/models.py:319-370, 660-716).
    Does NOT mutate its inputs (the reference subtracts the mean in place, :321-323)."""
    _, start, lowest, mean = MODEL_CFG[model]
    sf = scalefactor(start)
    m1 = torch.tensor(mean[:3], dtype=img1.dtype, device=img1.device).view(1, 3, 1, 1)
    m2 = torch.tensor(mean[3:], dtype=img1.dtype, device=img1.device).view(1, 3, 1, 1)
    img1 = img1 - m1
    img2 = img2 - m2
    feat1 = features(sd, img1)
    feat2 = features(sd, img2)
    im1, im2 = [img1], [img2]
    for l in range(1, 6):  # image pyramid (src/models.py:336-343)
        size = (feat1[l].shape[2], feat1[l].shape[3])
        im1.append(F.interpolate(im1[-1], size=size, mode="bilinear", align_corners=False))
        im2.append(F.interpolate(im2[-1], size=size, mode="bilinear", align_corners=False))
    levels = list(range(lowest, 7))
    xflow = None
    trace = []
    n_ext = len({k.split(".")[1] for k in sd if k.startswith("NetC_ext.")})
    for i in reversed(range(len(levels))):
        level = levels[i]
        idx = level - 1  # 0-based index into the feature lists (the reference's ``pyr_level``)
        if idx < 2:  # src/models.py:353-355; NetC_ext[idx-1] wraps to [-1] when idx == 0
            e = (idx - 1) % n_ext
            f1 = _conv(feat1[idx], sd, f"NetC_ext.{e}.conv_ext.0", 1, 0)
            f2 = _conv(feat2[idx], sd, f"NetC_ext.{e}.conv_ext.0", 1, 0)
        else:
            f1, f2 = feat1[idx], feat2[idx]
        xm = matching(sd, i, level, sf[level], f1, f2, xflow, corr_fn)
        xs = subpixel(sd, i, level, sf[level], f1, f2, xm)
        xflow = regularization(sd, i, level, sf[level], im1[idx], im2[idx], feat1[idx], xs)
        trace.append((level, xm, xs, xflow))
    out = xflow * sf[1]
    return (out, trace) if return_levels else out


def estimate(sd, img1: torch.Tensor, img2: torch.Tensor, model: str = "piv") -> torch.Tensor:
    """inference.estimate (inference.py:30-67) with tensor=True semantics."""
    assert img1.shape[2:] == img2.shape[2:]
    H, W = img1.shape[2], img1.shape[3]
    Wa = int(math.floor(math.ceil(W / 32.0) * 32.0))
    Ha = int(math.floor(math.ceil(H / 32.0) * 32.0))
    a = F.interpolate(img1, size=(Ha, Wa), mode="bilinear", align_corners=False)
    b = F.interpolate(img2, size=(Ha, Wa), mode="bilinear", align_corners=False)
    with torch.no_grad():
        raw = forward(sd, a, b, model)
    flow = F.interpolate(raw, size=(H, W), mode="bilinear", align_corners=False)
    flow[:, 0] *= float(W) / float(Wa)
    flow[:, 1] *= float(H) / float(Ha)
    return flow
