"""Forward pass of LiteFlowNet / LiteFlowNet2 (PIV and Hui variants) as a launch plan over the C ABI.

``Engine`` owns (i) the weights repacked once into the kernels' layouts, (ii) one workspace of NHWC
fp32 buffers per input shape and (iii) optionally a CUDA graph of the whole forward.  ``forward``
replaces ``LiteFlowNet.forward`` (src/models.py:319-370) / ``LiteFlowNet2.forward`` (:660-716) of the
reference: same inputs ([B,3,H,W] fp32 NCHW in [0,1], mean-subtracted IN PLACE like the reference),
same output ([B,2,H',W'] fp32 NCHW flow in pixels).

Data layout in HBM.  Every activation is NHWC fp32.  The reference's ``torch.cat`` calls disappear:
each concatenated tensor is ONE buffer whose channel slices are written in place by their producers
(``View`` = pointer + channel count + pixel pitch):

    Sbuf[l] = [ f1 (Cm) | backwarp(f2) (Cm) | flow_M (2) | pad ]       input of conv_S  (src/models.py:216)
    Rbuf[l] = [ feat (Cr) | err (1) | flow_S - mean (2) | pad ]        input of conv_R  (src/models.py:280)
    (pixel pitch rounded up to 8 floats = 32 bytes)

Rbuf's channel order differs from the reference's (err, rm, feat): the first conv_R weight is permuted
at pack time instead, so that every slice stays 16-byte aligned.

There is no CPU path: everything below launches kernels from libpivlfn.so.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from .arch import (CFGS, CONV_R, DIST_CH, KSIZE, LEVEL_FEAT_CH, MATCH_FEAT_CH, NETC, NETC_LEVEL_END, ModelCfg,
                   param_specs)
from .ops import View, view

# conv back-ends
SIMT = "simt"          # exact fp32 FFMA on the CUDA cores
TC_TF32 = "tf32"       # tcgen05 kind::tf32, one pass
TC_3XTF32 = "3xtf32"   # tcgen05 kind::tf32, error-compensated three passes (fp32-equivalent)
TC_TF32C = "tf32c"     # tf32 main product + the two low-order products in bf16 (fp32-equivalent, 2/3 of the MMA work)
TC_F16C = "f16c"       # all three products in fp16 on split operands (fp32-equivalent, 1/2 of the MMA work; activations
                       # must stay inside the fp16 range -- checked on the device, falls back to tf32c otherwise)
PRECISIONS = (SIMT, TC_TF32, TC_3XTF32, TC_TF32C, TC_F16C)
DEFAULT_PRECISION = TC_F16C
PASSES = {TC_TF32: 1, TC_3XTF32: 3, TC_TF32C: 2, TC_F16C: 4}


def _r4(c: int) -> int:
    return (c + 3) & ~3


def _r8(c: int) -> int:
    """Pixel pitch of the concat buffers: a multiple of 8 floats, so that every row starts on a 32-byte sector and the
    convolution epilogues can use 256-bit stores into the slices (measured: 2.5x faster than 16-byte stores at a
    528-byte pitch)."""
    return (c + 7) & ~7


def tf32_round(w: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away, like cvt.rna.tf32.f32) of fp32 to the 10-bit-mantissa TF32 grid."""
    i = w.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return torch.where(torch.isfinite(w), r, w)


@dataclass
class ConvW:
    """One convolution's packed weights."""
    cin: int
    cout: int
    kh: int
    kw: int
    stride: int
    w_simt: torch.Tensor                    # [KH*KW*Cin, CoutP]
    bias: Optional[torch.Tensor]
    w_hi: Optional[torch.Tensor] = None     # [CoutP16, KH*KW, CinP32] TF32 split for the tcgen05 kernel
    w_lo: Optional[torch.Tensor] = None
    w_c16: Optional[torch.Tensor] = None    # [2, CoutP16, KH*KW, CinP32] bf16: bf16(w), bf16(w - w_hi)  (precision tf32c)
    w_f16: Optional[torch.Tensor] = None    # stage_image of [2, CoutP16, KH*KW, CinP32] fp16: f16(w), f16((w - f16(w)) * 2048)  (precision f16c)

    w_s2: Optional[torch.Tensor] = None     # 3x3 stride-2 layers restated over the four input parities: fp16 pack (2 or 3 tiles)
    s2_passes: int = 0                      # of [CoutP16, 4, 4*Cin] for pivlfn_conv_s2_tc, and its kernel mode (4 or 5)
    w_f16s: Optional[torch.Tensor] = None   # stage_image of [3, CoutP16, KH*KW, CinP32] fp16, W = 256 w: f16(W), f16(W - f16(W)), f16(f16(W) / 2048)
                                            # (single-accumulator variant of f16c, packed for Cout > 64 only)
    w_f8: Optional[torch.Tensor] = None     # the P16 pipeline's weight image (see _pack_f8; bytes): f16(S w) and the e5m2 correction tile
                                            # + a trailer with 1 / S; for a 3x3 stride-2 layer: of the parity restatement

    def to_(self, dev) -> "ConvW":
        for f in ("w_simt", "bias", "w_hi", "w_lo", "w_c16", "w_f16", "w_s2", "w_f16s", "w_f8"):
            t = getattr(self, f)
            if t is not None:
                setattr(self, f, t.to(dev))
        return self

    def passes_for(self, passes: int) -> int:
        """The kernel mode of this layer for an engine-level mode: f16c (4) runs its single-accumulator variant (5) on the
        wide layers, where [main | corr] accumulators for two stacked tiles would fill all of TMEM."""
        return 5 if (passes == 4 and self.w_f16s is not None) else passes

    def pack16(self, passes: int) -> Optional[torch.Tensor]:
        return {2: self.w_c16, 4: self.w_f16, 5: self.w_f16s}.get(passes)
    stem: bool = False                      # 7x7 3->32 packed as [32, 7, 32] for pivlfn_conv_stem_tc


def pack_conv(w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1, cin_pad: int = 0,
              in_perm: Optional[torch.Tensor] = None, tc: bool = True) -> ConvW:
    """w: [Cout,Cin,KH,KW] (torch.nn.Conv2d layout) -> kernel layouts.  ``in_perm`` reorders input channels
    (new channel j reads old channel in_perm[j]); ``cin_pad`` appends zero input channels."""
    w = w.detach().to(torch.float32)
    if in_perm is not None:
        w = w[:, in_perm]
    if cin_pad:
        w = torch.cat([w, w.new_zeros(w.shape[0], cin_pad, w.shape[2], w.shape[3])], 1)
    cout, cin, kh, kw = w.shape
    coutp = _r4(cout)
    ws = w.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout)
    if coutp != cout:
        ws = torch.cat([ws, ws.new_zeros(ws.shape[0], coutp - cout)], 1)
    cw = ConvW(cin, cout, kh, kw, stride, ws.contiguous(), None if b is None else b.detach().to(torch.float32).contiguous())
    if tc and tc_eligible(cin, cout, kh, kw, stride):
        cinp = (cin + 31) // 32 * 32
        coutp = (cout + 15) // 16 * 16
        wt = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
        wt = torch.nn.functional.pad(wt, (0, cinp - cin, 0, 0, 0, coutp - cout))
        cw.w_hi, cw.w_lo = _split_tf32(wt)
        cw.w_c16 = _pack_c16(wt, cw.w_hi)
        cw.w_f16 = stage_image(_pack_f16(wt))
        if coutp > 64 and stride == 1 and cw.w_f16 is not None and os.environ.get("PIVLFN_F16_SINGLE", "1") != "0":
            cw.w_f16s = stage_image(_pack_f16_single(wt))
        if stride == 1 and coutp <= 128:
            cw.w_f8 = _pack_f8(wt)
        if stride == 2 and kh == 3 and kw == 3 and cin % 32 == 0 and os.environ.get("PIVLFN_S2_HALO", "1") != "0":
            w2 = _restate_s2(w)                                      # [cout, 4, 4*cin]
            w2 = torch.nn.functional.pad(w2, (0, 0, 0, 0, 0, coutp - cout))
            if coutp > 64:
                cw.w_s2, cw.s2_passes = stage_image(_pack_f16_single(w2)), 5
            else:
                cw.w_s2, cw.s2_passes = stage_image(_pack_f16(w2)), 4
            if coutp <= 128:
                cw.w_f8 = _pack_f8(w2)
    return cw


def _restate_s2(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] stride-2 weights -> [Cout, 4 taps, 4*Cin] of the equivalent 2x2-tap convolution over the four
    pixel-parity phases of the input: input row 2y + ky - 1 = 2(y + by) + py with (by, py) = (-1, 1), (0, 0), (0, 1) for
    ky = 0, 1, 2; tap = (by + 1) * 2 + (bx + 1), channel = (py * 2 + px) * Cin + c."""
    cout, cin = w.shape[0], w.shape[1]
    m = {0: (0, 1), 1: (1, 0), 2: (1, 1)}                            # k -> (block-tap index, parity)
    w2 = w.new_zeros(cout, 4, 4, cin)
    for ky in range(3):
        ty, py = m[ky]
        for kx in range(3):
            tx, px = m[kx]
            w2[:, ty * 2 + tx, py * 2 + px, :] = w[:, :, ky, kx]
    return w2.reshape(cout, 4, 4 * cin)


def _pack_c16(wt: torch.Tensor, w_hi: torch.Tensor) -> torch.Tensor:
    return torch.stack([wt.to(torch.bfloat16), (wt - w_hi).to(torch.bfloat16)]).contiguous()


def stage_image(pack: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """[parts, CoutP, ntaps, CinP] fp16 operand tiles -> the exact shared-memory image of the kernel's weight ring, so that one
    ring stage is ONE contiguous bulk copy (a tensor-map box of 64-byte rows streams at 29 B/clk per SM, a contiguous copy at
    47: tools/ubench/tma_tile_rate.cu): [chunk of 32 channels][tap][part][cout row][64 bytes], with the 64B swizzle of the
    MMA operand layout applied to the four 16-byte units of every row (unit u of row r sits at u ^ ((r >> 1) & 3); every
    tile starts on a 512-byte boundary in shared memory)."""
    if pack is None:
        return None
    parts, coutp, ntaps, cinp = pack.shape
    assert cinp % 32 == 0 and coutp % 8 == 0
    x = pack.reshape(parts, coutp, ntaps, cinp // 32, 4, 8).permute(3, 2, 0, 1, 4, 5)       # (chunk, tap, part, row, unit, elem)
    r = torch.arange(coutp, device=pack.device)
    u = torch.arange(4, device=pack.device)
    src = (u[None, :] ^ ((r[:, None] >> 1) & 3)).view(1, 1, 1, coutp, 4, 1).expand(x.shape)
    return torch.gather(x, 4, src).contiguous()


def _pack_f16(wt: torch.Tensor) -> torch.Tensor:
    """w = f16(w) + 2^-11 * f16((w - f16(w)) * 2^11) up to ~2^-24 |w|."""
    hi = wt.to(torch.float16)
    if not torch.isfinite(hi).all():
        return None                      # a weight outside the fp16 range: the engine falls back to tf32c
    lo = ((wt - hi.to(torch.float32)) * 2048.0).to(torch.float16)
    return torch.stack([hi, lo]).contiguous()


def _pack_f16_single(wt: torch.Tensor) -> Optional[torch.Tensor]:
    """Three fp16 tiles of W = 256 w whose products with the activation pair (a_hi, a_lo' = (a - a_hi) * 2^11) all carry
    the same scale: W_hi = f16(W), W_lo = f16(W - W_hi) (unscaled: |W_lo| <= 2^-12 |W| stays a normal fp16 number for
    |w| >= 1e-3, and below that its absolute error is 2^-25 of W's unit), W_hi2 = f16(W_hi * 2^-11)."""
    W = wt * 256.0
    hi = W.to(torch.float16)
    if not torch.isfinite(hi).all():
        return None
    lo = (W - hi.to(torch.float32)).to(torch.float16)
    hi2 = (hi.to(torch.float32) / 2048.0).to(torch.float16)
    return torch.stack([hi, lo, hi2]).contiguous()


P16_MODE = 6            # pivlfn_conv_p16's product scheme: f16 main product + fp8 corrections (csrc/p16.cuh)
W8 = torch.float8_e4m3fn        # the weights' correction tile: scaled per layer, so e4m3's range suffices and its extra mantissa bit
                                # is free (tools/sim_precision.py: flow error -30 % against e5m2 on both sides)


def f8_scale(wt: torch.Tensor) -> float:
    """The power of two S that puts the largest |w| S of a layer into [8192, 16384): W_hi = f16(w S) is finite whatever the
    weights' magnitude, and the two correction tiles (w S / 2048 in [4, 8) and w S - W_hi <= 8 for the largest weight) stay inside the
    normal range of e4m3 (>= 2^-6) for every weight down to 2^-9 of the layer's largest; smaller ones lose correction bits whose
    absolute weight is below 2^-21 of the layer's largest weight."""
    m = float(wt.abs().max())
    if not math.isfinite(m) or m == 0.0:
        return 1.0
    return 2.0 ** max(-40, min(40, math.floor(math.log2(16384.0 / m))))


def _pack_f8(wt: torch.Tensor) -> Optional[torch.Tensor]:
    """[CoutP, ntaps, CinP] fp32 -> the weight image of pivlfn_conv_p16: stage_image of two operand tiles, both viewed as fp16
    [CoutP, ntaps, CinP] (64-byte rows per 32-channel chunk), plus a 16-byte trailer [1 / S, S, 0, 0] (fp32) that the kernel's
    epilogue reads.  With W = S w (S = f8_scale, a power of two per layer):
      tile 0: W_hi = f16(W), the B operand of a_hi * W_hi (kind::f16, K = 16: the row's bytes [0,32) and [32,64));
      tile 1: per 16-channel K step 32 e4m3 bytes [e4m3(W / 2048) x 16 | e4m3(W - W_hi) x 16], the B operand of the single
              fp8 MMA (K = 32) whose A operand is the activation's [lo8 x 16 | hi8 x 16] block (csrc/p16.cuh):
              sum lo8 * W 2^-11 + hi8 * W_lo  =  the two correction products of the split."""
    coutp, ntaps, cinp = wt.shape
    if not torch.isfinite(wt).all():
        return None
    S = f8_scale(wt)
    W = wt * S
    hi = W.to(torch.float16)
    c_lo = (W / 2048.0).to(W8).view(torch.uint8).reshape(coutp, ntaps, cinp // 16, 16)
    c_hi = (W - hi.to(torch.float32)).to(W8).view(torch.uint8).reshape(coutp, ntaps, cinp // 16, 16)
    corr = torch.cat([c_lo, c_hi], dim=-1).reshape(coutp, ntaps, cinp * 2).contiguous().view(torch.float16)
    img = stage_image(torch.stack([hi, corr]).contiguous())
    trailer = torch.tensor([1.0 / S, S, 0.0, 0.0], dtype=torch.float32, device=img.device).view(torch.uint8)
    return torch.cat([img.reshape(-1).view(torch.uint8), trailer])


def pack_stem(w: torch.Tensor, b: torch.Tensor) -> ConvW:
    """NetC.conv1 [32,3,7,7] -> [32, 7 (ky), 32 (kx*4 + c)] for the overlapping-window tensor-core stem."""
    w = w.detach().to(torch.float32)
    cw = pack_conv(w, b, 1, cin_pad=1, tc=False)
    wt = torch.zeros(32, 7, 8, 4, device=w.device)
    wt[:, :, :7, :3] = w.permute(0, 2, 3, 1)           # [cout, ky, kx, c]
    cw.w_hi, cw.w_lo = _split_tf32(wt.reshape(32, 7, 32))
    cw.w_c16 = _pack_c16(wt.reshape(32, 7, 32), cw.w_hi)
    cw.w_f16 = stage_image(_pack_f16(wt.reshape(32, 7, 32)))
    cw.stem = True
    return cw


def _split_tf32(wt: torch.Tensor):
    hi = tf32_round(wt)
    lo = tf32_round(wt - hi)
    return hi.contiguous(), lo.contiguous()


def tc_eligible(cin: int, cout: int, kh: int, kw: int, stride: int) -> bool:
    """Convolutions with at least 8 input channels and at most 128 output channels run on the tensor cores (any odd
    kernel up to 7x7, stride 1 or 2)."""
    return stride in (1, 2) and cout <= 128 and cin >= 8 and kh <= 7 and kw <= 7


class Engine:
    """Weights + workspaces + launch plan for one model."""

    def __init__(self, cfg: ModelCfg, state_dict: Dict[str, torch.Tensor], device: torch.device,
                 precision: str = None, use_graph: bool = None):
        if device.type != "cuda":
            raise NotImplementedError("pivlfn runs on CUDA devices only (no CPU path)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = device
        self.precision = precision or os.environ.get("PIVLFN_PRECISION", DEFAULT_PRECISION)
        if self.precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.use_graph = (os.environ.get("PIVLFN_GRAPH", "1") != "0") if use_graph is None else use_graph
        # fp16 range check of precision f16c: "sync" = read the device flag after every forward and repeat it in tf32c on a
        # hit (one stream sync per forward); "deferred" = read it asynchronously and switch at the NEXT call (an out-of-range
        # forward returns non-finite values, never silently saturated ones); "auto" (default) = sync for large forwards,
        # where the sync is free, deferred for small ones (run.py-style single pairs), where it would be a throughput cliff
        rc = os.environ.get("PIVLFN_RANGE_CHECK", "auto")
        self.range_check = {"1": "auto", "0": "off"}.get(rc, rc)
        if self.range_check not in ("auto", "sync", "deferred", "off"):
            raise ValueError("PIVLFN_RANGE_CHECK must be auto, sync, deferred or 0")
        # precision f16c runs the P16 pipeline (activations kept in HBM as the fp16 pairs the MMAs read: plan16.py);
        # PIVLFN_P16=0 selects the older fp32-activation plan whose kernels split the operands in shared memory
        self.p16 = self.precision == TC_F16C and os.environ.get("PIVLFN_P16", "1") != "0"
        self.flag = torch.zeros(1, device=device, dtype=torch.int32)          # P16 producers raise it (never cleared by them)
        self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._pending: Optional[torch.cuda.Event] = None
        # flow heads (KxK, 32 -> 2): "simt" = exact-fp32 CUDA-core kernel (default), "pairs" = 1x1 tensor-core convolution to
        # tap planes + gather-sum, "conv" = the generic convolution path
        self.head_mode = os.environ.get("PIVLFN_HEAD", "simt")
        self.sf = cfg.scalefactor
        self.w: Dict[str, ConvW] = {}
        self.raw: Dict[str, torch.Tensor] = {}
        self._plans: Dict[Tuple[int, int, int], "Plan"] = {}
        self.launches = 0          # kernels launched (graph replays counted by their captured node count)
        self._pack(state_dict)
        def unpacked(cw):          # a tensor-core layer whose fp16 operand tiles could not be built (weights outside the fp16 range)
            if cw.w_hi is None:
                return False
            return cw.w_f16 is None or (self.p16 and cw.w_f8 is None and not cw.stem and cw.cout <= 128 and (cw.stride == 1 or cw.w_s2 is not None))
        if self.precision == TC_F16C and any(unpacked(cw) for cw in self.w.values()):
            import warnings
            warnings.warn("pivlfn: a weight lies outside the fp16 range; precision 'f16c' replaced by 'tf32c'")
            self.precision = TC_TF32C
            self.p16 = False

    # ---------------------------------------------------------------------------------------------
    def _pack(self, sd: Dict[str, torch.Tensor]):
        dev = self.device
        cfg = self.cfg
        specs = param_specs(cfg)
        for k, shp in specs.items():
            if k not in sd:
                raise KeyError(f"missing parameter {k}")
            if tuple(sd[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: expected {tuple(shp)}, got {tuple(sd[k].shape)}")
        # Packing runs on the HOST (a few dozen small torch CPU ops per layer) and every packed tensor is uploaded once:
        # on the device the same repacking was ~1000 tiny ATen kernels per model, which drowned this library's own
        # kernels in the driver's launch list of smoke().
        g = lambda k: sd[k].detach().to(device="cpu", dtype=torch.float32)
        use_tc = self.precision != SIMT

        def conv(key, stride=1, cin_pad=0, in_perm=None):
            self.w[key] = pack_conv(g(key + ".weight"), g(key + ".bias") if key + ".bias" in sd else None,
                                    stride, cin_pad, in_perm, tc=use_tc)

        for seq, idx, cin, cout, k, st in NETC:
            if cin == 3 and use_tc:
                self.w[f"NetC.{seq}.{idx}"] = pack_stem(g(f"NetC.{seq}.{idx}.weight"), g(f"NetC.{seq}.{idx}.bias"))
            else:
                conv(f"NetC.{seq}.{idx}", st, cin_pad=1 if cin == 3 else 0)
        if self.p16:
            # 128 -> 192 (stride 2): two 96-channel halves for the tensor-core kernel (at most 128 accumulator columns)
            w6, b6 = g("NetC.conv6.0.weight"), g("NetC.conv6.0.bias")
            self.w["NetC.conv6.0#a"] = pack_conv(w6[:96], b6[:96], 2, tc=True)
            self.w["NetC.conv6.0#b"] = pack_conv(w6[96:], b6[96:], 2, tc=True)
        for e in range(cfg.n_ext):
            conv(f"NetC_ext.{e}.conv_ext.0")
        nh = len(cfg.head)
        for i, lv in enumerate(cfg.levels):
            for j in range(nh + 1):
                conv(f"NetE_M.{i}.conv_M.{2 * j}")
                conv(f"NetE_S.{i}.conv_S.{2 * j}")
            for key in (f"NetE_M.{i}.conv_M.{2 * nh}", f"NetE_S.{i}.conv_S.{2 * nh}"):
                w = g(key + ".weight")                                       # [2, cin, K, K]
                if self.p16:
                    # tensor-core flow head: 1xK convolution to the 2K row channels (ky*2 + co), summed over ky by
                    # pivlfn_head_rows_sum (plan16.py)
                    K = w.shape[2]
                    self.w[key + "#rows"] = pack_conv(w.permute(2, 0, 1, 3).reshape(2 * K, w.shape[1], 1, K), None, 1, tc=True)
                    # the transposed split (Kx1 convolution to the column channels kx*2 + co + pivlfn_head_cols_sum): narrower
                    # halo tiles, used by the untiled plan; the row form keeps the tiled plan's vertical halo bookkeeping simple
                    self.w[key + "#cols"] = pack_conv(w.permute(3, 0, 1, 2).reshape(2 * K, w.shape[1], K, 1), None, 1, tc=True)
                if self.head_mode == "simt" and w.shape[1] == 32:
                    # exact-fp32 CUDA-core flow head (pivlfn_flow_head): weights as [K*K][32][2]
                    self.raw[key + "#head"] = w.permute(2, 3, 1, 0).reshape(-1, 32, 2).contiguous()
                elif self.head_mode == "pairs" and use_tc and KSIZE[lv] >= 5:
                    # flow head restated as a 1x1 convolution to 2*K*K channels (row = tap*2 + co) + a gather-sum
                    w2 = w.permute(2, 3, 0, 1).reshape(-1, w.shape[1], 1, 1)   # [(K*K*2), cin, 1, 1]
                    self.w[key + "#pairs"] = pack_conv(w2, None, 1, tc=True)
            for nm in ("upConv_M", "upCorr_M"):
                key = f"NetE_M.{i}.{nm}.weight"
                if key in sd:
                    self.raw[key] = g(key).reshape(-1, 16).contiguous()
            p = f"NetE_R.{i}"
            if lv < 5:
                conv(p + ".moduleFeat.0")
            cr = 128 if lv < 5 else LEVEL_FEAT_CH[lv]
            perm = torch.cat([torch.arange(3, 3 + cr), torch.arange(0, 3)])
            conv(p + ".conv_R.0", in_perm=perm, cin_pad=1)
            for j in range(1, len(CONV_R)):
                conv(f"{p}.conv_R.{2 * j}")
            conv(p + ".conv_dist_R.0")
            if lv < 5:
                conv(p + ".conv_dist_R.1")
            for nm in ("moduleScaleX", "moduleScaleY"):
                self.raw[f"{p}.{nm}.weight"] = g(f"{p}.{nm}.weight").reshape(-1).contiguous()
                self.raw[f"{p}.{nm}.bias"] = g(f"{p}.{nm}.bias").reshape(-1).contiguous()
        # ---- upload ------------------------------------------------------------------------------------------------
        for cw in self.w.values():
            cw.to_(dev)
        self.raw = {k: v.to(dev) for k, v in self.raw.items()}

    # ---------------------------------------------------------------------------------------------
    MAX_PLANS = 4       # workspaces + graphs kept alive per engine (least recently used ones are dropped)

    def plan(self, B: int, H: int, W: int) -> "Plan":
        key = (B, H, W)
        p = self._plans.pop(key, None)
        if p is None:
            while len(self._plans) >= self.MAX_PLANS:
                # a directory of mixed-size frames or many odd last batches must not grow device memory without bound
                self._plans.pop(next(iter(self._plans)))
            with torch.cuda.device(self.device):
                if self.p16:
                    from .plan16 import Plan16
                    p = Plan16(self, B, H, W)
                else:
                    p = Plan(self, B, H, W)
        self._plans[key] = p            # (re)inserted last: dict order is the LRU order
        return p

    def forward(self, img1: torch.Tensor, img2: torch.Tensor, return_levels: bool = False):
        if not (img1.is_cuda and img2.is_cuda):
            raise NotImplementedError("pivlfn: CUDA tensors required (there is no CPU path)")
        assert img1.dtype == torch.float32 and img2.dtype == torch.float32
        assert img1.dim() == 4 and img1.shape[1] == 3 and img1.shape == img2.shape
        if img1.device != self.device or img2.device != self.device:
            raise RuntimeError(f"pivlfn: inputs on {img1.device}, model on {self.device}")
        B, _, H, W = img1.shape
        if H % 32 or W % 32:
            raise RuntimeError(f"pivlfn: H and W must be multiples of 32 (got {H}x{W}); use estimate() which resizes "
                               "like the reference (inference.py:39-49)")
        # strided inputs are accepted like the reference does; the in-place mean subtraction still lands in the caller's
        # tensors (mutate_inputs writes through a strided copy_)
        c1, c2 = img1.contiguous(), img2.contiguous()
        with torch.cuda.device(self.device):     # launches, the current stream and the range flag all belong to self.device
            res = self._forward_on_device(c1, c2, B, H, W, return_levels)
        if c1 is not img1:
            img1.copy_(c1)
        if c2 is not img2:
            img2.copy_(c2)
        return res

    def _fall_back(self, what: str):
        import warnings
        warnings.warn(f"pivlfn: an activation left the fp16 range in precision 'f16c' ({what}); this model now runs in 'tf32c'")
        self.precision = TC_TF32C
        self.p16 = False
        self._plans.clear()
        self._pending = None
        self.flag.zero_()

    def check_range(self, wait: bool = True) -> bool:
        """Resolve a deferred fp16-range check (P16 pipeline).  Returns True when the forward(s) since the last check left
        the range: their output contains non-finite values and this engine has switched to 'tf32c'."""
        if self._pending is None:
            return False
        if not wait and not self._pending.query():
            return False
        self._pending.synchronize()
        self._pending = None
        if int(self._flag_host[0]) != 0:
            self._fall_back("detected after the forward returned: its output is not finite")
            return True
        return False

    def _forward_on_device(self, img1, img2, B, H, W, return_levels):
        if self.p16:
            return self._forward_p16(img1, img2, B, H, W, return_levels)
        plan = self.plan(B, H, W)
        check = self.precision == TC_F16C and self.range_check != "off"
        if check:
            # the flag is sticky and process-wide per device: a stale hit (a direct ops call, another engine, the tiled
            # path) must not downgrade THIS engine, so it is cleared before the forward it is meant to judge
            if int(self.lib.pivlfn_f16_range_flag_clear(torch.cuda.current_stream().cuda_stream)) != 0:
                raise _lib.PivlfnError("pivlfn_f16_range_flag_clear: CUDA error")
        res = plan.run(img1, img2, return_levels, mutate_inputs=False)
        if check:
            # the fp32-activation f16c kernels convert activations to fp16 pairs in shared memory: a value outside the fp16
            # range raises a sticky device flag (never a silent saturation).  One stream sync + 4-byte read per forward;
            # on a hit this engine switches to tf32c for good and the forward is repeated from the caller's (still
            # unmodified) images.
            torch.cuda.current_stream().synchronize()
            flag = int(self.lib.pivlfn_f16_range_flag(1))
            if flag < 0:
                raise _lib.PivlfnError("pivlfn_f16_range_flag: CUDA error")
            if flag:
                self._fall_back("forward repeated")
                plan = self.plan(B, H, W)
                res = plan.run(img1, img2, return_levels, mutate_inputs=False)
        plan.mutate_inputs(img1, img2)
        return res

    def _forward_p16(self, img1, img2, B, H, W, return_levels):
        """P16 pipeline: every producer of a P16 tensor raises ``self.flag`` (owned by this engine) when a value is not
        finite in fp16."""
        mode = self.range_check
        if mode == "auto":
            mode = "sync" if B * H * W >= (1 << 20) else "deferred"
        if self.check_range(wait=False):                      # a deferred check of an earlier forward came back positive
            return self._forward_on_device(img1, img2, B, H, W, return_levels)
        plan = self.plan(B, H, W)
        res = plan.run(img1, img2, return_levels, mutate_inputs=False)
        if mode == "sync":
            if int(self.flag.item()) != 0:                    # .item(): one stream sync + 4-byte read
                self._fall_back("forward repeated")
                return self._forward_on_device(img1, img2, B, H, W, return_levels)
        elif mode == "deferred":
            if self._pending is None or self._pending.query():
                self._flag_host.copy_(self.flag, non_blocking=True)
                self._pending = torch.cuda.Event()
                self._pending.record()
        plan.mutate_inputs(img1, img2)
        return res


class Plan:
    """Workspace and launch sequence for one (B, H, W)."""

    def __init__(self, eng: Engine, B: int, H: int, W: int):
        self.eng, self.B, self.H, self.W = eng, B, H, W
        cfg = eng.cfg
        dev = eng.device
        E = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        Z = lambda *shape: torch.zeros(shape, device=dev, dtype=torch.float32)
        self.hw = {l: (H >> (l - 1), W >> (l - 1)) for l in range(1, 7)}
        N2 = 2 * B
        # static inputs (graph replays need fixed addresses)
        self.in1, self.in2 = E(B, 3, H, W), E(B, 3, H, W)
        self.img = {1: E(N2, H, W, 4)}
        self.img_pad = Z(N2, H, W + 8, 4) if eng.precision != SIMT else None     # zero border written once
        for l in range(2, 7):
            self.img[l] = E(N2, *self.hw[l], 4)
        # NetC intermediates and features
        self.netc_out: List[torch.Tensor] = []
        lvl = 1
        for seq, idx, cin, cout, k, st in NETC:
            if st == 2:
                lvl += 1
            self.netc_out.append(E(N2, *self.hw[lvl], cout))
        self.feat = {l: self.netc_out[NETC_LEVEL_END[l]] for l in range(1, 7)}
        self.lv: Dict[int, dict] = {}
        nh = len(cfg.head)
        for l in cfg.levels:
            h, w = self.hw[l]
            cm, cr = MATCH_FEAT_CH[l], (128 if l < 5 else LEVEL_FEAT_CH[l])
            s = 2 if l < 4 else 1
            d = dict(
                f2=E(B, h, w, cm) if l <= 2 else None,
                flowU=E(B, h, w, 2) if l != 6 else None,
                corr=Z(B, (h + s - 1) // s, (w + s - 1) // s, 52),
                corrU=Z(B, h, w, 52) if l < 4 else None,
                Sbuf=Z(B, h, w, _r8(2 * cm + 4)),
                Rbuf=Z(B, h, w, _r8(cr + 4)),
                flowM=E(B, h, w, 2), flowS=E(B, h, w, 2), flowR=E(B, h, w, 2),
                partial=E(B, ops.flow_mean_parts(), 2),
                dist=E(B, h, w, _r4(DIST_CH[l])), dist0=E(B, h, w, _r4(DIST_CH[l])) if l < 5 else None,
            )
            d["planes"] = (E(KSIZE[l] * KSIZE[l], B * h * w, 2)
                           if (eng.head_mode == "pairs" and eng.precision != SIMT and KSIZE[l] >= 5 and w >= 8) else None)
            widths = sorted(set(cfg.head) | set(CONV_R))
            d["t"] = {c: [E(B, h, w, c), E(B, h, w, c)] for c in widths}   # ping-pong per width
            self.lv[l] = d
        lo = cfg.lowest_level
        self.out = E(B, 2, *self.hw[lo])
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_launches = 0
        self._warm = 0

    # ---------------------------------------------------------------------------------------------
    def _conv(self, key: str, x: View, n: int, h: int, w: int, y: View, lrelu: bool = True,
              res: Optional[View] = None):
        eng = self.eng
        cw = eng.w[key]
        assert x.C == cw.cin and y.C == cw.cout, (key, x.C, cw.cin, y.C, cw.cout)
        passes = cw.passes_for(PASSES.get(eng.precision, 1))
        c16 = cw.pack16(passes)
        if eng.precision == TC_F16C and cw.w_s2 is not None and w // 2 >= 8 and res is None:
            ops.conv_s2_tc(x, n, h, w, cw.w_s2, cw.bias, y, lrelu, cw.s2_passes)
        elif cw.stem and eng.precision != SIMT:
            ops.conv_stem_tc(self.img_pad, n, h, w, cw.w_hi, cw.w_lo, cw.bias, y, lrelu, passes, c16)
        elif cw.w_hi is not None and eng.precision != SIMT:
            ops.conv_tc(x, n, h, w, cw.w_hi, cw.w_lo, cw.bias, y, cw.kh, cw.kw, lrelu, passes, res, c16, cw.stride)
        else:
            ops.conv_simt(x, n, h, w, cw.w_simt, cw.bias, y, cw.kh, cw.kw, cw.stride, lrelu, res)

    def _chain(self, prefix: str, idxs: List[int], x: View, l: int, res: View, out: View, out2: Optional[View] = None):
        """conv_M / conv_S: 3x3 conv + LeakyReLU ..., then the KxK 32->2 flow head plus residual flow."""
        B = self.B
        h, w = self.hw[l]
        t = self.lv[l]["t"]
        used: Dict[int, int] = {}
        for j in idxs[:-1]:
            c = self.eng.w[f"{prefix}.{j}"].cout
            k = used.get(c, 0)
            used[c] = k ^ 1
            y = view(t[c][k])
            self._conv(f"{prefix}.{j}", x, B, h, w, y)
            x = y
        key = f"{prefix}.{idxs[-1]}"
        pk = self.eng.w.get(key + "#pairs")
        hw_ = self.eng.raw.get(key + "#head")
        if hw_ is not None:
            ops.flow_head(x, B, h, w, hw_, self.eng.w[key].bias, res, out, KSIZE[l], out2)
            return
        if pk is not None and w >= 8 and self.lv[l].get("planes") is not None:
            K = KSIZE[l]
            passes = PASSES.get(self.eng.precision, 1)
            ops.conv1x1_pairs_tc(x, B, h, w, pk.w_hi, pk.w_lo, pk.pack16(passes), self.lv[l]["planes"], K * K, passes)
            ops.flow_head_sum(self.lv[l]["planes"], K, self.eng.w[key].bias, res, out, B, h, w)
        else:
            self._conv(key, x, B, h, w, out, lrelu=False, res=res)
        if out2 is not None:
            ops.copy(out, out2, B * h * w)

    def launch_all(self):
        """Enqueue the whole forward on the current stream (inputs already in self.in1 / self.in2)."""
        eng, cfg, B = self.eng, self.eng.cfg, self.B
        N2 = 2 * B
        H, W = self.H, self.W
        ops.prep_images(self.in1, self.in2, self.img[1], cfg.mean, self.img_pad)
        for l in range(2, 7):
            ops.avgpool2(self.img[l - 1], self.img[l])
        # ---- NetC on both images at once (shared weights): batch 2B ----------------------------------------
        x = view(self.img[1])
        lvl = 1
        for i, (seq, idx, cin, cout, k, st) in enumerate(NETC):
            hi, wi = self.hw[lvl]
            if st == 2:
                lvl += 1
            y = view(self.netc_out[i])
            self._conv(f"NetC.{seq}.{idx}", x, N2, hi, wi, y)
            x = y
        nh = len(cfg.head)
        head_idx = [2 * j for j in range(nh + 1)]
        rconv_idx = [2 * j for j in range(len(CONV_R))]
        xflow = None
        for i in reversed(range(len(cfg.levels))):
            l = cfg.levels[i]
            d = self.lv[l]
            h, w = self.hw[l]
            cm = MATCH_FEAT_CH[l]
            cr = 128 if l < 5 else LEVEL_FEAT_CH[l]
            s = 2 if l < 4 else 1
            scale = eng.sf[l]
            feat = self.feat[l]                      # [2B,h,w,Cf]
            feat1, feat2 = feat[:B], feat[B:]
            S_f1, S_f2w, S_fl = view(d["Sbuf"], 0, cm), view(d["Sbuf"], cm, cm), view(d["Sbuf"], 2 * cm, 2)
            if l <= 2:
                # NetC_ext (src/models.py:353-355): list index idx = l-1 uses NetC_ext[idx-1] (wraps to [-1])
                e = (l - 2) % cfg.n_ext
                self._conv(f"NetC_ext.{e}.conv_ext.0", view(feat1), B, h, w, S_f1)
                self._conv(f"NetC_ext.{e}.conv_ext.0", view(feat2), B, h, w, view(d["f2"]))
                f2 = view(d["f2"])
            else:
                ops.copy(view(feat1), S_f1, B * h * w)
                f2 = view(feat2)
            # ---- Matching (src/models.py:165-187) ------------------------------------------------------------
            if xflow is not None:
                ops.deconv4x4s2_dw(view(xflow), B, h // 2, w // 2, eng.raw[f"NetE_M.{i}.upConv_M.weight"], view(d["flowU"]))
                flowU = d["flowU"]
            else:
                flowU = None
            ops.corr_nhwc(S_f1, f2, flowU, scale, view(d["corr"], 0, 49), B, h, w, s, lrelu=True)
            if l < 4:
                ops.deconv4x4s2_dw(view(d["corr"], 0, 49), B, (h + 1) // 2, (w + 1) // 2,
                                   eng.raw[f"NetE_M.{i}.upCorr_M.weight"], view(d["corrU"], 0, 49))
                cin = view(d["corrU"], 0, 49)
            else:
                cin = view(d["corr"], 0, 49)
            # flow_M is written twice by the flow head: dense (flowM) and into its slice of the Subpixel concat buffer
            self._chain(f"NetE_M.{i}.conv_M", head_idx, cin, l, view(flowU) if flowU is not None else None,
                        view(d["flowM"]), S_fl)
            # ---- Subpixel (src/models.py:209-217) ------------------------------------------------------------
            ops.warp(f2, d["flowM"], scale, S_f2w, B, h, w)
            self._chain(f"NetE_S.{i}.conv_S", head_idx, view(d["Sbuf"], 0, 2 * cm + 2), l, view(d["flowM"]),
                        view(d["flowS"]))
            # ---- Regularization (src/models.py:274-303) ------------------------------------------------------
            ops.flow_mean(d["flowS"], d["partial"])
            ops.reg_input(self.img[l][:B], self.img[l][B:], d["flowS"], scale, d["partial"], view(d["Rbuf"], cr, 3))
            if l < 5:
                self._conv(f"NetE_R.{i}.moduleFeat.0", view(feat1), B, h, w, view(d["Rbuf"], 0, cr))
            else:
                ops.copy(view(feat1), view(d["Rbuf"], 0, cr), B * h * w)
            x = view(d["Rbuf"], 0, cr + 4)
            t = d["t"]
            used: Dict[int, int] = {}
            for j in rconv_idx:
                c = eng.w[f"NetE_R.{i}.conv_R.{j}"].cout
                k = used.get(c, 0)
                used[c] = k ^ 1
                y = view(t[c][k])
                self._conv(f"NetE_R.{i}.conv_R.{j}", x, B, h, w, y)
                x = y
            dc = DIST_CH[l]
            if l < 5:
                self._conv(f"NetE_R.{i}.conv_dist_R.0", x, B, h, w, view(d["dist0"], 0, dc), lrelu=False)
                self._conv(f"NetE_R.{i}.conv_dist_R.1", view(d["dist0"], 0, dc), B, h, w, view(d["dist"], 0, dc), lrelu=False)
            else:
                self._conv(f"NetE_R.{i}.conv_dist_R.0", x, B, h, w, view(d["dist"], 0, dc), lrelu=False)
            p = f"NetE_R.{i}"
            last = (l == cfg.lowest_level)
            ops.reg_tail(view(d["dist"], 0, dc), d["flowS"], eng.raw[p + ".moduleScaleX.weight"],
                         eng.raw[p + ".moduleScaleX.bias"], eng.raw[p + ".moduleScaleY.weight"],
                         eng.raw[p + ".moduleScaleY.bias"], d["flowR"], self.out if last else None, eng.sf[1], KSIZE[l])
            xflow = d["flowR"]

    # ---------------------------------------------------------------------------------------------
    def run_static(self):
        """Run the forward on the static input buffers (self.in1/self.in2 -> self.out)."""
        eng = self.eng
        if not eng.use_graph:
            c0 = ops.launch_count()
            self.launch_all()
            eng.launches += ops.launch_count() - c0
            return
        if self.graph is None:
            if self._warm < 1:
                # one eager run first: surfaces launch errors outside capture and warms lazy module loading
                keep1, keep2 = self.in1.clone(), self.in2.clone()
                c0 = ops.launch_count()
                self.launch_all()
                eng.launches += ops.launch_count() - c0
                torch.cuda.current_stream().synchronize()
                self._warm = 1
                self.in1.copy_(keep1)
                self.in2.copy_(keep2)
            g = torch.cuda.CUDAGraph()
            c0 = ops.launch_count()
            # thread_local: only THIS thread's calls are checked while the stream is capturing.  The batch driver's reader /
            # writer threads pin host memory and wait on events at any time (pivlfn.io), which in the default global mode
            # invalidates a capture that happens to be in progress (cudaErrorStreamCaptureInvalidated); the cyclic garbage
            # collector is paused for the same reason (it may destroy an old plan's graph and free its pool mid-capture).
            import gc
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                # an explicit capture stream on THIS engine's device: torch's default capture stream is created once per process
                # on whatever device captured first, and entering it switches the current device -- a second engine on another
                # GPU would record its kernels (with its own device's pointers) on the first GPU's stream
                with torch.cuda.graph(g, stream=torch.cuda.Stream(device=eng.device), capture_error_mode="thread_local"):
                    self.launch_all()
            finally:
                if gc_was_on:
                    gc.enable()
            self.graph_launches = ops.launch_count() - c0
            self.graph = g
        self.graph.replay()
        eng.launches += self.graph_launches

    def mutate_inputs(self, img1: torch.Tensor, img2: torch.Tensor):
        """The reference subtracts the per-channel mean from the CALLER's tensors (src/models.py:321-323)."""
        ops.copy_dense(self.in1, img1)
        ops.copy_dense(self.in2, img2)

    def run(self, img1: torch.Tensor, img2: torch.Tensor, return_levels: bool = False, mutate_inputs: bool = True):
        # staged with a kernel, not cudaMemcpy (see ops.copy_dense)
        ops.copy_dense(img1, self.in1)
        ops.copy_dense(img2, self.in2)
        self.run_static()
        if mutate_inputs:
            self.mutate_inputs(img1, img2)
        out = torch.empty_like(self.out)
        ops.copy_dense(self.out, out)
        if not return_levels:
            return out
        levels = []
        for i in reversed(range(len(self.eng.cfg.levels))):
            d = self.lv[self.eng.cfg.levels[i]]
            levels.append([d[k].permute(0, 3, 1, 2).contiguous() for k in ("flowM", "flowS", "flowR")])
        return out, levels
