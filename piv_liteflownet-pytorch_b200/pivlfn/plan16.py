"""Launch plan of the P16 pipeline (the default for precision ``f16c``): the forward of ``pivlfn.model.Plan`` with every
activation that feeds a convolution kept in HBM as the operands the tensor cores consume (fp16 value, e5m2 residual, e5m2
value: csrc/p16.cuh).

Same reference semantics (src/models.py:319-370 / :660-716), same buffers-as-concats idea as ``Plan``:

    Sbuf[l] = [ f1 (Cm) | flow_M (2) + 14 zero ]                          P16, input of conv_S   (src/models.py:216); the
              middle Cm channels of the reference's concat, backwarp(f2), exist only inside conv_S.0's shared memory
    Rbuf[l] = [ feat (Cr) | err, rm_u, rm_v + 13 zero ]                   P16, input of conv_R   (src/models.py:280)

Channel slices start at multiples of 16 (one P16 group = 64 bytes per pixel).  What stays fp32 NHWC: the images, the
flows, the second image's NetC_ext features at levels 1-2 (read only by the cost volume and the backwarp), the
half-resolution cost volume in front of upCorr_M and the flow-head planes (the distance maps of the regularisation stay in TMEM).

Differences from ``Plan`` besides the format:
  * all convolutions (stride-2 NetC layers, the 192-channel conv6 as two 96-channel halves, tiny levels) run in ONE kernel
    family, ``pivlfn_conv_p16``: no operand split in shared memory, 16 epilogue warps;
  * the Subpixel backwarp (src/models.py:214) is fused into its consumer: conv_S.0's gather warps sample f2 and write the MMA
    operand tile directly (``pivlfn_conv_p16_warp``), so the warped features never reach HBM (PIVLFN_FUSE_WARP=0: separate
    ``pivlfn_warp_p16`` kernel writing a slice of Sbuf);
  * the KxK 32 -> 2 flow heads run on the tensor cores as a Kx1 convolution to 2K column planes + a K-column gather-sum
    (``pivlfn_head_cols_sum``), which also writes flow_M's P16 group into Sbuf;
  * the Regularization tail (src/models.py:279-300) runs in the epilogue of the last conv_dist_R layer
    (``pivlfn_conv_p16_tail``; PIVLFN_FUSE_TAIL=0: fp32 distance maps + ``pivlfn_reg_tail``).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops
from .arch import CONV_R, DIST_CH, KSIZE, LEVEL_FEAT_CH, MATCH_FEAT_CH, NETC, NETC_LEVEL_END
from .model import P16_MODE, Plan
from .ops import OUT_F32, OUT_P16, OUT_PLANES, View, view


def _r16(c: int) -> int:
    return (c + 15) & ~15


def _r4(c: int) -> int:
    return (c + 3) & ~3


class Plan16(Plan):
    """Workspace and launch sequence for one (B, H, W) in the P16 format."""

    def __init__(self, eng, B: int, H: int, W: int):     # noqa: D401  (does not call Plan.__init__: different buffers)
        self.eng, self.B, self.H, self.W = eng, B, H, W
        cfg = eng.cfg
        dev = eng.device
        E = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        Z = lambda *shape: torch.zeros(shape, device=dev, dtype=torch.float32)
        self.hw = {l: (H >> (l - 1), W >> (l - 1)) for l in range(1, 7)}
        self.fuse_warp = os.environ.get("PIVLFN_FUSE_WARP", "1") != "0"
        self.fuse_tail = os.environ.get("PIVLFN_FUSE_TAIL", "1") != "0"
        N2 = 2 * B
        self.in1, self.in2 = E(B, 3, H, W), E(B, 3, H, W)
        self.img = {1: E(N2, H, W, 4)}
        self.img_pad = Z(N2, H, W + 8, 4)                # zero border written once
        for l in range(2, 7):
            self.img[l] = E(N2, *self.hw[l], 4)
        # NetC intermediates and features: P16, channel counts are multiples of 16 (32, 64, 96, 128, 192)
        self.netc_out: List[torch.Tensor] = []
        lvl = 1
        for seq, idx, cin, cout, k, st in NETC:
            if st == 2:
                lvl += 1
            self.netc_out.append(E(N2, *self.hw[lvl], _r16(cout)))
        self.feat = {l: self.netc_out[NETC_LEVEL_END[l]] for l in range(1, 7)}
        self.lv: Dict[int, dict] = {}
        for l in cfg.levels:
            h, w = self.hw[l]
            cm, cr = MATCH_FEAT_CH[l], (128 if l < 5 else LEVEL_FEAT_CH[l])
            s = 2 if l < 4 else 1
            K = KSIZE[l]
            d = dict(
                f2=E(B, h, w, cm) if l <= 2 else None,                                   # fp32 NHWC
                flowU=E(B, h, w, 2) if l != 6 else None,
                corr=Z(B, (h + s - 1) // s, (w + s - 1) // s, 52) if l < 4 else None,    # fp32, in front of upCorr_M
                corrU=Z(B, h, w, 64),                                                    # P16: input of conv_M
                Sbuf=Z(B, h, w, (cm if self.fuse_warp else 2 * cm) + 16),
                Rbuf=Z(B, h, w, cr + 16),
                flowM=E(B, h, w, 2), flowS=E(B, h, w, 2), flowR=E(B, h, w, 2),
                partial=E(B, ops.flow_mean_parts(), 2),
                dist=E(B, h, w, _r4(DIST_CH[l])),                                        # fp32: input of the tail
                dist0=Z(B, h, w, _r16(DIST_CH[l])) if l < 5 else None,                   # P16
                planes=E(K, B * h * w, 2),                                               # fp32 row planes of the flow heads
            )
            widths = sorted(set(cfg.head) | set(CONV_R))
            d["t"] = {c: [E(B, h, w, _r16(c)), E(B, h, w, _r16(c))] for c in widths}     # P16 ping-pong per width
            self.lv[l] = d
        lo = cfg.lowest_level
        self.out = E(B, 2, *self.hw[lo])
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_launches = 0
        self._warm = 0

    # ---------------------------------------------------------------------------------------------
    def _conv(self, key: str, x: View, n: int, h: int, w: int, y: View, lrelu: bool = True, out_fmt: int = OUT_P16,
              cin: Optional[int] = None):
        """h, w: INPUT size."""
        eng = self.eng
        cw = eng.w[key]
        cin = cw.cin if cin is None else cin
        assert x.C >= _r16(cin) or x.C == cin, (key, x.C, cin)
        assert cw.w_f8 is not None, key
        ops.conv_p16(x, n, h, w, cin, cw.w_f8, P16_MODE, cw.bias, y, cw.cout, cw.kh, cw.kw, cw.stride, lrelu, out_fmt, 0, eng.flag)

    def _chain(self, prefix: str, idxs: List[int], x: View, l: int, res: Optional[torch.Tensor], out: torch.Tensor,
               out_p16: Optional[View] = None, warp=None):
        """conv_M / conv_S: 3x3 conv + LeakyReLU ..., then the KxK 32->2 flow head plus residual flow.
        warp = (src view, src_p16, flow, scale, c0, n): channels [c0, c0 + n) of the first layer's input are
        backwarp(src, scale * flow), gathered inside the kernel."""
        eng, B = self.eng, self.B
        h, w = self.hw[l]
        d = self.lv[l]
        t = d["t"]
        used: Dict[int, int] = {}
        for j in idxs[:-1]:
            cw = eng.w[f"{prefix}.{j}"]
            c = cw.cout
            k = used.get(c, 0)
            used[c] = k ^ 1
            y = view(t[c][k])
            if warp is not None and j == idxs[0]:
                src, src_p16, flow, scale, c0, n = warp
                w_img, mode = cw.w_f8, P16_MODE
                ops.conv_p16_warp(x, B, h, w, cw.cin, w_img, mode, cw.bias, y, cw.cout, cw.kh, cw.kw, True, src, src_p16, flow, scale,
                                  c0, n, eng.flag)
            else:
                self._conv(f"{prefix}.{j}", x, B, h, w, y)
            x = y
        key = f"{prefix}.{idxs[-1]}"
        K = KSIZE[l]
        rw = eng.w[key + "#cols"]                       # Kx1 convolution to the 2K column channels (kx*2 + co)
        ops.conv_p16(x, B, h, w, rw.cin, rw.w_f8, P16_MODE, None, view(d["planes"].view(1, K, B * h * w, 2)), 2 * K, K, 1, 1, False,
                     OUT_PLANES, 2 * B * h * w, eng.flag)
        ops.head_cols_sum(d["planes"], K, eng.w[key].bias, res, out, out_p16, B, h, w, eng.flag)

    def launch_all(self):
        """Enqueue the whole forward on the current stream (inputs already in self.in1 / self.in2)."""
        eng, cfg, B = self.eng, self.eng.cfg, self.B
        N2 = 2 * B
        fl = eng.flag
        ops.prep_images(self.in1, self.in2, self.img[1], cfg.mean, self.img_pad)
        for l in range(2, 7):
            ops.avgpool2(self.img[l - 1], self.img[l])
        # ---- NetC on both images at once (shared weights): batch 2B ----------------------------------------
        x = None
        lvl = 1
        for i, (seq, idx, cin, cout, k, st) in enumerate(NETC):
            hi, wi = self.hw[lvl]
            if st == 2:
                lvl += 1
            y = view(self.netc_out[i])
            key = f"NetC.{seq}.{idx}"
            if i == 0:
                ops.conv_stem_p16(self.img_pad, N2, hi, wi, eng.w[key].w_f16, eng.w[key].bias, y, True, fl)
            elif cout > 128:
                # 128 -> 192: two 96-channel halves (the accumulator tile holds at most 128 columns)
                for half, off in (("#a", 0), ("#b", 96)):
                    self._conv(key + half, x, N2, hi, wi, view(self.netc_out[i], off, 96))
            else:
                self._conv(key, x, N2, hi, wi, y)
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
            feat = self.feat[l]                      # [2B,h,w,Cf] P16
            feat1, feat2 = feat[:B], feat[B:]
            fuse = self.fuse_warp
            S_f1, S_fl = view(d["Sbuf"], 0, cm), view(d["Sbuf"], cm if fuse else 2 * cm, 16)
            if l <= 2:
                # NetC_ext (src/models.py:353-355): list index idx = l-1 uses NetC_ext[idx-1] (wraps to [-1])
                e = (l - 2) % cfg.n_ext
                self._conv(f"NetC_ext.{e}.conv_ext.0", view(feat1), B, h, w, S_f1)
                self._conv(f"NetC_ext.{e}.conv_ext.0", view(feat2), B, h, w, view(d["f2"]), out_fmt=OUT_F32)
                f2, f2_p16 = view(d["f2"]), False
            else:
                ops.copy(view(feat1), S_f1, B * h * w)
                f2, f2_p16 = view(feat2), True
            # ---- Matching (src/models.py:165-187) ------------------------------------------------------------
            if xflow is not None:
                ops.deconv4x4s2_dw(view(xflow), B, h // 2, w // 2, eng.raw[f"NetE_M.{i}.upConv_M.weight"], view(d["flowU"]))
                flowU = d["flowU"]
            else:
                flowU = None
            if l < 4:
                ops.corr_p16(S_f1, True, f2, f2_p16, flowU, scale, view(d["corr"], 0, 49), False, B, h, w, cm, s, True, fl)
                ops.deconv4x4s2_dw_p16(view(d["corr"], 0, 49), B, (h + 1) // 2, (w + 1) // 2, 49,
                                       eng.raw[f"NetE_M.{i}.upCorr_M.weight"], view(d["corrU"]), fl)
            else:
                ops.corr_p16(S_f1, True, f2, f2_p16, flowU, scale, view(d["corrU"]), True, B, h, w, cm, s, True, fl)
            self._chain(f"NetE_M.{i}.conv_M", head_idx, view(d["corrU"]), l, flowU, d["flowM"], S_fl)
            # ---- Subpixel (src/models.py:209-217) ------------------------------------------------------------
            if fuse:
                self._chain(f"NetE_S.{i}.conv_S", head_idx, view(d["Sbuf"]), l, d["flowM"], d["flowS"],
                            warp=(f2, f2_p16, d["flowM"], scale, cm, cm))
            else:
                ops.warp_p16(f2, f2_p16, d["flowM"], scale, view(d["Sbuf"], cm, cm), B, h, w, cm, fl)
                self._chain(f"NetE_S.{i}.conv_S", head_idx, view(d["Sbuf"]), l, d["flowM"], d["flowS"])
            # ---- Regularization (src/models.py:274-303) ------------------------------------------------------
            ops.flow_mean(d["flowS"], d["partial"])
            ops.reg_input_p16(self.img[l][:B], self.img[l][B:], d["flowS"], scale, d["partial"], view(d["Rbuf"], cr, 16), fl)
            if l < 5:
                self._conv(f"NetE_R.{i}.moduleFeat.0", view(feat1), B, h, w, view(d["Rbuf"], 0, cr))
            else:
                ops.copy(view(feat1), view(d["Rbuf"], 0, cr), B * h * w)
            x = view(d["Rbuf"])
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
            p = f"NetE_R.{i}"
            last = (l == cfg.lowest_level)
            tail = (d["flowS"], eng.raw[p + ".moduleScaleX.weight"], eng.raw[p + ".moduleScaleX.bias"], eng.raw[p + ".moduleScaleY.weight"],
                    eng.raw[p + ".moduleScaleY.bias"], d["flowR"], self.out if last else None, eng.sf[1])
            if l < 5:
                self._conv(f"NetE_R.{i}.conv_dist_R.0", x, B, h, w, view(d["dist0"]), lrelu=False)
                dkey, dx = f"NetE_R.{i}.conv_dist_R.1", view(d["dist0"])
            else:
                dkey, dx = f"NetE_R.{i}.conv_dist_R.0", x
            cw = eng.w[dkey]
            if self.fuse_tail:
                # the distances never leave TMEM: softmax(-d^2), unfold, ScaleX / ScaleY and the division run in the conv's epilogue
                ops.conv_p16_tail(dx, B, h, w, cw.cin, cw.w_f8, cw.bias, cw.kh, cw.kw, KSIZE[l], *tail)
            else:
                self._conv(dkey, dx, B, h, w, view(d["dist"]), lrelu=False, out_fmt=OUT_F32)
                ops.reg_tail(view(d["dist"], 0, dc), *tail, KSIZE[l])
            xflow = d["flowR"]
