"""One very large PIV frame tiled by rows across GPUs (BASELINE.json configs[4]).

The reference has no tiling at all ("1024 x 1024 takes too much memory!", inference.py:227).  Here rank r of P owns
rows [r*H/P, (r+1)*H/P) of the frame at every pyramid level and runs the SAME kernels as the single-GPU plan on a local
buffer that carries E halo rows above and below its slab:

* after an operator with vertical reach r the outermost r halo rows of its output are stale; validity is tracked
  statically per tensor and halo rows are re-fetched from the row neighbours (NCCL send/recv of contiguous NHWC row
  blocks over NVLink) only when the next consumer needs more valid rows than are left -- a chain of 3x3 convolutions
  exchanges once every ~E layers, not once per layer;
* beyond the true frame edge the halo rows are kept at zero (the zero padding of conv / correlation / backwarp / unfold);
* the only global coupling of the network, the per-sample flow mean of the regularisation (src/models.py:275), is a
  2-float all-reduce per level;
* coarse levels whose slabs would be thinner than the halo are replicated on every rank (all-gather of a few rows).

The backwarp reach is data dependent; ``warp_reach`` rows are provisioned and the bound is verified after the run.
``LoopbackGroup`` runs P ranks in lock-step inside one process (tests on a single GPU); ``DistGroup`` uses
torch.distributed (one process per GPU, backend nccl).  Batch size is 1 (left / right camera pairs are independent
units and go through pair sharding).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import ops
from .arch import CONV_R, DIST_CH, KSIZE, LEVEL_FEAT_CH, MATCH_FEAT_CH, NETC, NETC_LEVEL_END
from .model import P16_MODE, PASSES, SIMT, Engine, _r8
from .ops import View, view


class TT:
    """Tensor handle of the tiled plan: an NHWC buffer [N, rows, W, C] at pyramid level ``level``.
    ``tiled``: rows = own + 2E (local slab + halos) else the full frame (replicated level).
    ``valid``: number of halo rows on each side that currently hold exact data (E after an exchange)."""

    def __init__(self, t: torch.Tensor, level: int, tiled: bool, name: str):
        self.t, self.level, self.tiled, self.name = t, level, tiled, name
        self.valid = 0


@dataclass
class Step:
    kind: str                   # "op" | "exchange" | "allreduce" | "allgather" | "zero"
    fn: Optional[Callable] = None
    tt: Optional[TT] = None
    src: Optional[torch.Tensor] = None
    dst: Optional[torch.Tensor] = None


class TiledPlan:
    def __init__(self, eng: Engine, H: int, W: int, rank: int, world: int, halo: int = 16, warp_reach: int = 8):
        if H % (32 * world) or W % 32:
            raise ValueError("tiled mode: H must be a multiple of 32*world and W of 32")
        if halo % 2 or halo < 8 or warp_reach + 6 > halo:
            raise ValueError("halo must be even, >= 8 and >= warp_reach + 6 (stride-2 cost volume reach)")
        self.eng, self.H, self.W, self.rank, self.world = eng, H, W, rank, world
        self.E, self.wr = halo, warp_reach
        self.own = {l: (H // world) >> (l - 1) for l in range(1, 7)}
        self.Hl = {l: H >> (l - 1) for l in range(1, 7)}
        self.Wl = {l: W >> (l - 1) for l in range(1, 7)}
        # levels whose slab is at least one halo thick are tiled; coarser levels are replicated
        self.Lt = max([l for l in range(1, 7) if self.own[l] >= halo and self.own[l] % 2 == 0] or [0])
        if self.Lt < eng.cfg.lowest_level:
            raise ValueError("frame too small to tile at this world size / halo")
        self.top, self.bottom = rank == 0, rank == world - 1
        self.steps: List[Step] = []
        self.handles: List[TT] = []
        self.warp_flows: List[Tuple[TT, int]] = []
        dev = eng.device
        # every buffer starts zeroed: halo rows that no operator ever writes are still READ by the operators (their results
        # land in rows nobody consumes), and uninitialised memory there could hold non-finite bit patterns that raise the
        # fp16 range flag of precision f16c
        self._Z = lambda *shape: torch.zeros(shape, device=dev, dtype=torch.float32)
        self._E = self._Z
        self._build()

    # ---- geometry -------------------------------------------------------------------------------------------------
    def is_tiled(self, l: int) -> bool:
        return l <= self.Lt

    def rows(self, l: int) -> int:
        return self.own[l] + 2 * self.E if self.is_tiled(l) else self.Hl[l]

    def new(self, name: str, l: int, C: int, N: int = 1, zero: bool = False) -> TT:
        t = (self._Z if zero else self._E)(N, self.rows(l), self.Wl[l], C)
        h = TT(t, l, self.is_tiled(l), name)
        if not h.tiled:
            h.valid = 1 << 30
        self.handles.append(h)
        return h

    # ---- step recording -------------------------------------------------------------------------------------------
    def _need(self, tt: TT, need: int):
        """Make sure ``need`` halo rows of tt are exact before the next operator reads them."""
        if tt.tiled and tt.valid < need:
            assert need <= self.E, (tt.name, need)
            self.steps.append(Step("exchange", tt=tt))
            for h in self.handles:                    # channel slices share buffers: whole rows travel
                if h.t.data_ptr() == tt.t.data_ptr() and h.valid >= 0:
                    h.valid = self.E
            tt.valid = self.E

    def _op(self, fn: Callable, out: TT, valid: int):
        self.steps.append(Step("op", fn=fn))
        if out.tiled:
            out.valid = max(0, min(valid, self.E))
            self.steps.append(Step("zero", tt=out))        # recorded on every rank (identical step lists); no-op inside

    def alias(self, tt: TT, name: str) -> TT:
        """A second handle on the same buffer (another channel slice) with its own validity."""
        h = TT(tt.t, tt.level, tt.tiled, name)
        h.valid = tt.valid
        self.handles.append(h)
        return h

    @staticmethod
    def _img(tt: TT, n: int, r0: int = 0, nr: Optional[int] = None) -> torch.Tensor:
        """Rows [r0, r0+nr) of image n as a dense [1, nr, W, C] tensor (contiguous because N == 1)."""
        nr = tt.t.shape[1] - r0 if nr is None else nr
        return tt.t[n:n + 1, r0:r0 + nr]

    # ---- operators (same-level) -----------------------------------------------------------------------------------
    def conv(self, key: str, x: TT, xo: int, xc: int, y: TT, yo: int, lrelu: bool = True, res: Optional[TT] = None,
             stem_pad: Optional[TT] = None):
        eng = self.eng
        cw = eng.w[key]
        r = cw.kh // 2
        self._need(x, r)
        if res is not None:
            self._need(res, 0)
        N, hh, ww = x.t.shape[0], x.t.shape[1], x.t.shape[2]
        passes = cw.passes_for(PASSES.get(eng.precision, 1))
        c16 = cw.pack16(passes)

        def run():
            for n in range(N):
                xv = view(self._img(x, n), xo, xc)
                yv = view(self._img(y, n), yo, cw.cout)
                rv = view(self._img(res, n), 0, cw.cout) if res is not None else None
                if cw.stem and eng.precision != SIMT:
                    ops.conv_stem_tc(self._img(stem_pad, n), 1, hh, ww, cw.w_hi, cw.w_lo, cw.bias, yv, lrelu, passes, c16)
                elif cw.w_hi is not None and eng.precision != SIMT:
                    ops.conv_tc(xv, 1, hh, ww, cw.w_hi, cw.w_lo, cw.bias, yv, cw.kh, cw.kw, lrelu, passes, rv, c16)
                else:
                    ops.conv_simt(xv, 1, hh, ww, cw.w_simt, cw.bias, yv, cw.kh, cw.kw, 1, lrelu, rv)
        v = x.valid - r
        if res is not None:
            v = min(v, res.valid)
        self._op(run, y, v)

    def copy(self, x: TT, xo: int, C: int, y: TT, yo: int):
        def run():
            for n in range(x.t.shape[0]):
                a, b = self._img(x, n), self._img(y, n)
                ops.copy(view(a, xo, C), view(b, yo, C), a.shape[1] * a.shape[2])
        self._op(run, y, x.valid)

    def _steers_warp(self, flow: TT, out_valid: int):
        """Remember a flow tensor that steers a backwarp (cost volume: flowU, Subpixel: flowM, brightness error: flowS)
        together with the halo depth down to which the warped result is consumed: the displacement bound is checked on
        exactly those rows after every run (``check_bounds``)."""
        if flow.tiled:
            self.warp_flows.append((flow, max(0, min(out_valid, self.E))))

    def warp(self, f2: TT, C: int, flow: TT, scale: float, y: TT, yo: int):
        self._need(f2, self.wr)
        self._need(flow, 0)
        self._steers_warp(flow, min(flow.valid, f2.valid - self.wr))

        def run():
            a, fl, b = self._img(f2, 0), self._img(flow, 0), self._img(y, 0)
            ops.warp(view(a, 0, C), fl, scale, view(b, yo, C), 1, a.shape[1], a.shape[2])
        self._op(run, y, min(flow.valid, f2.valid - self.wr))

    # ---- the plan -------------------------------------------------------------------------------------------------
    def _down_views(self, x: TT, y: TT, n: int, pad_rows: int):
        """Input / output row windows of a stride-2 operator from level l (x) to level l+1 (y).
        Returns (x tensor, y tensor, rows of y to all-gather or None)."""
        E = self.E
        if x.tiled and y.tiled:
            # local output row j <-> local input row 2j - E (+ tap): input window from row 0, output from row E/2
            hin = x.t.shape[1]
            return self._img(x, n), self._img(y, n, E // 2, hin // 2), None
        if not x.tiled:
            return self._img(x, n), self._img(y, n), None
        raise AssertionError("tiled -> replicated goes through _down_gather")

    def down(self, kind: str, x: TT, y: TT, key: Optional[str] = None, C: int = 4):
        """Stride-2 operator across levels: 'pool' (2x2 mean, src/models.py:336-343) or 'conv' (3x3 s2 NetC conv)."""
        eng = self.eng
        N = x.t.shape[0]
        E = self.E
        if x.tiled and not y.tiled:
            # boundary: compute the owned rows of the coarser level (+1 junk row each side) and all-gather them
            self._need(x, 2)
            own2 = self.own[y.level]
            tmp = self._E(N, own2 + 2, self.Wl[y.level], y.t.shape[3])

            def run():
                for n in range(N):
                    xin = self._img(x, n, E - 2, 2 * own2 + 4)
                    self._down_call(kind, key, xin, tmp[n:n + 1], C)
            self.steps.append(Step("op", fn=run))
            self.steps.append(Step("allgather", src=tmp, dst=y.t))
            return
        if kind == "conv":
            self._need(x, 1)

        def run2():
            for n in range(N):
                a, b, _ = self._down_views(x, y, n, 1)
                self._down_call(kind, key, a, b, C)
        v = (x.valid - 1) // 2 if kind == "conv" else x.valid // 2
        self._op(run2, y, min(v, E // 2) if y.tiled else v)

    def _down_call(self, kind, key, a: torch.Tensor, b: torch.Tensor, C: int):
        if kind == "pool":
            ops.avgpool2(a, b)
        else:
            cw = self.eng.w[key]
            ops.conv_simt(view(a, 0, cw.cin), 1, a.shape[1], a.shape[2], cw.w_simt, cw.bias, view(b, 0, cw.cout), 3, 3, 2, True)

    def up(self, x: TT, C: int, w: torch.Tensor, y: TT, yo: int = 0, xo: int = 0):
        """Depthwise ConvTranspose 4x4 s2 from level l+1 (x) to level l (y) (src/models.py:144-145,151-152)."""
        E = self.E
        self._need(x, 1)
        if x.tiled and y.tiled:
            j0, i0 = E // 2, 0
            hin = min(x.t.shape[1] - j0, (y.t.shape[1] - i0) // 2)
            v = max(0, min(2 * (x.valid - 1), E - 2))
        elif not x.tiled and y.tiled:
            R0 = self.rank * self.own[y.level]
            j0 = max(0, (R0 - E) // 2)
            i0 = 2 * j0 - R0 + E
            hin = min(x.t.shape[1] - j0, (y.t.shape[1] - i0) // 2)
            v = E - 2
        else:
            j0, i0, hin, v = 0, 0, x.t.shape[1], 1 << 30

        def run():
            a = self._img(x, 0, j0, hin)
            b = self._img(y, 0, i0, 2 * hin)
            ops.deconv4x4s2_dw(view(a, xo, C), 1, hin, a.shape[2], w, view(b, yo, C))
        self._op(run, y, v)

    def _build(self):
        eng, cfg, E = self.eng, self.eng.cfg, self.E
        tc = eng.precision != SIMT
        # ---- inputs: every rank is handed its slab of both mean-free images incl. halo rows (zero outside the frame)
        self.in1 = self._E(1, 3, self.rows(1), self.W)
        self.in2 = self._E(1, 3, self.rows(1), self.W)
        img: Dict[int, TT] = {1: self.new("img1", 1, 4, N=2)}
        img_pad = None
        if tc:
            img_pad = TT(self._Z(2, self.rows(1), self.W + 8, 4), 1, True, "img_pad")
            img_pad.valid = E

        def prep():
            ops.prep_images(self.in1, self.in2, img[1].t, (0.0,) * 6, img_pad.t if img_pad is not None else None)
        self.steps.append(Step("op", fn=prep))
        img[1].valid = E                                           # inputs arrive with exact halos
        for l in range(2, 7):
            img[l] = self.new(f"img{l}", l, 4, N=2)
            self.down("pool", img[l - 1], img[l])
        # ---- NetC on both images ------------------------------------------------------------------------------------
        feats: Dict[int, TT] = {}
        x, lvl = img[1], 1
        for i, (seq, idx, cin, cout, k, st) in enumerate(NETC):
            key = f"NetC.{seq}.{idx}"
            if st == 2:
                lvl += 1
                y = self.new(key, lvl, cout, N=2)
                self.down("conv", x, y, key)
            else:
                y = self.new(key, lvl, cout, N=2)
                self.conv(key, x, 0, x.t.shape[3], y, 0, stem_pad=img_pad)
            x = y
            if NETC_LEVEL_END[lvl] == i:
                feats[lvl] = y
        nh = len(cfg.head)
        head_idx = [2 * j for j in range(nh + 1)]
        xflow: Optional[TT] = None
        self.flows: Dict[int, TT] = {}
        for i in reversed(range(len(cfg.levels))):
            l = cfg.levels[i]
            cm = MATCH_FEAT_CH[l]
            cr = 128 if l < 5 else LEVEL_FEAT_CH[l]
            s = 2 if l < 4 else 1
            scale = eng.sf[l]
            feat = feats[l]
            f1 = TT(feat.t[0:1], l, feat.tiled, f"feat1_{l}"); f1.valid = feat.valid
            f2r = TT(feat.t[1:2], l, feat.tiled, f"feat2_{l}"); f2r.valid = feat.valid
            self.handles += [f1, f2r]
            im1 = TT(img[l].t[0:1], l, img[l].tiled, f"im1_{l}"); im1.valid = img[l].valid
            im2 = TT(img[l].t[1:2], l, img[l].tiled, f"im2_{l}"); im2.valid = img[l].valid
            self.handles += [im1, im2]
            Sbuf = self.new(f"Sbuf{l}", l, _r8(2 * cm + 4), zero=True)
            S_f2w, S_fl = self.alias(Sbuf, f"S_f2w{l}"), self.alias(Sbuf, f"S_fl{l}")
            if l <= 2:
                e = (l - 2) % cfg.n_ext
                self.conv(f"NetC_ext.{e}.conv_ext.0", f1, 0, LEVEL_FEAT_CH[l], Sbuf, 0)
                f2 = self.new(f"f2e{l}", l, cm)
                self.conv(f"NetC_ext.{e}.conv_ext.0", f2r, 0, LEVEL_FEAT_CH[l], f2, 0)
            else:
                self.copy(f1, 0, cm, Sbuf, 0)
                f2 = f2r
            # ---- Matching ------------------------------------------------------------------------------------
            flowU = None
            if xflow is not None:
                flowU = self.new(f"flowU{l}", l, 2)
                self.up(xflow, 2, eng.raw[f"NetE_M.{i}.upConv_M.weight"], flowU)
            corr_rows = (self.rows(l) + s - 1) // s
            corr = TT(self._Z(1, corr_rows, (self.Wl[l] + s - 1) // s, 52), l, self.is_tiled(l), f"corr{l}")
            self.handles.append(corr)
            self._need(Sbuf, 0)
            self._need(f2, 3 * s + (self.wr if flowU is not None else 0))
            if flowU is not None:
                self._need(flowU, 3 * s)

            def run_corr(Sbuf=Sbuf, f2=f2, flowU=flowU, corr=corr, cm=cm, s=s, scale=scale):
                a, b = self._img(Sbuf, 0), self._img(f2, 0)
                ops.corr_nhwc(view(a, 0, cm), view(b, 0, cm), self._img(flowU, 0) if flowU is not None else None, scale,
                              view(corr.t, 0, 49), 1, a.shape[1], a.shape[2], s, True)
            self.steps.append(Step("op", fn=run_corr))
            vin = min(Sbuf.valid, f2.valid - 3 * s - (self.wr if flowU is not None else 0),
                      flowU.valid - 3 * s if flowU is not None else 1 << 30)
            if flowU is not None:
                self._steers_warp(flowU, flowU.valid)
            if s == 2:
                corrU = self.new(f"corrU{l}", l, 52, zero=True)

                def run_upc(corr=corr, corrU=corrU, i=i):
                    ops.deconv4x4s2_dw(view(corr.t, 0, 49), 1, corr.t.shape[1], corr.t.shape[2],
                                       eng.raw[f"NetE_M.{i}.upCorr_M.weight"], view(corrU.t, 0, 49))
                # half-resolution cost volume rows 2 apart: valid rows shrink by 2 through the 4x4 up-convolution
                self._op(run_upc, corrU, vin - 2)
                cin = corrU
            else:
                corr.valid = max(0, min(vin, E)) if corr.tiled else 1 << 30
                if corr.tiled:
                    self.steps.append(Step("zero", tt=corr))
                cin = corr
            flowM = self.new(f"flowM{l}", l, 2)
            self._chain(f"NetE_M.{i}.conv_M", head_idx, cin, 49, l, flowU, flowM)
            # ---- Subpixel ------------------------------------------------------------------------------------
            self.warp(f2, cm, flowM, scale, S_f2w, cm)
            self.copy(flowM, 0, 2, S_fl, 2 * cm)
            Sall = self.alias(Sbuf, f"Sall{l}")
            Sall.valid = min(Sbuf.valid, S_f2w.valid, S_fl.valid)
            flowS = self.new(f"flowS{l}", l, 2)
            self._chain(f"NetE_S.{i}.conv_S", head_idx, Sall, 2 * cm + 2, l, flowM, flowS)
            # ---- Regularization ------------------------------------------------------------------------------
            partial = self._E(1, ops.flow_mean_parts(), 2)
            r0, nr = (E, self.own[l]) if flowS.tiled else (0, self.Hl[l])
            frac = float(self.rows(l)) / float(self.Hl[l])     # reg_input divides by the LOCAL pixel count

            def run_mean(flowS=flowS, partial=partial, r0=r0, nr=nr):
                ops.flow_mean(self._img(flowS, 0, r0, nr).contiguous(), partial)
            self.steps.append(Step("op", fn=run_mean))
            if flowS.tiled:
                self.steps.append(Step("allreduce", src=partial))
                self.steps.append(Step("op", fn=lambda partial=partial, frac=frac: partial.mul_(frac)))
            Rbuf = self.new(f"Rbuf{l}", l, _r8(cr + 4), zero=True)
            R_in = self.alias(Rbuf, f"R_in{l}")
            self._need(im2, self.wr)
            self._need(im1, 0)
            self._need(flowS, 0)

            def run_ri(im1=im1, im2=im2, flowS=flowS, partial=partial, Rbuf=Rbuf, cr=cr, scale=scale):
                ops.reg_input(self._img(im1, 0), self._img(im2, 0), self._img(flowS, 0), scale, partial,
                              view(self._img(Rbuf, 0), cr, 3))
            self._steers_warp(flowS, min(im1.valid, flowS.valid, im2.valid - self.wr))
            self._op(run_ri, R_in, min(im1.valid, flowS.valid, im2.valid - self.wr))
            if l < 5:
                self.conv(f"NetE_R.{i}.moduleFeat.0", f1, 0, LEVEL_FEAT_CH[l], Rbuf, 0)
            else:
                self.copy(f1, 0, cr, Rbuf, 0)
            Rall = self.alias(Rbuf, f"Rall{l}")
            Rall.valid = min(Rbuf.valid, R_in.valid)
            x = Rall
            xc = cr + 4
            for j in range(len(CONV_R)):
                key = f"NetE_R.{i}.conv_R.{2 * j}"
                y = self.new(key, l, eng.w[key].cout)
                self.conv(key, x, 0, xc, y, 0)
                x, xc = y, eng.w[key].cout
            dc = DIST_CH[l]
            dist = self.new(f"dist{l}", l, (dc + 3) & ~3)
            if l < 5:
                dist0 = self.new(f"dist0{l}", l, (dc + 3) & ~3)
                self.conv(f"NetE_R.{i}.conv_dist_R.0", x, 0, xc, dist0, 0, lrelu=False)
                self.conv(f"NetE_R.{i}.conv_dist_R.1", dist0, 0, dc, dist, 0, lrelu=False)
            else:
                self.conv(f"NetE_R.{i}.conv_dist_R.0", x, 0, xc, dist, 0, lrelu=False)
            flowR = self.new(f"flowR{l}", l, 2)
            K = KSIZE[l]
            last = l == cfg.lowest_level
            if last:
                self.out_local = self._E(1, 2, self.rows(l), self.Wl[l])
            self._need(flowS, K // 2)
            self._need(dist, 0)
            p = f"NetE_R.{i}"

            def run_tail(dist=dist, flowS=flowS, flowR=flowR, dc=dc, K=K, last=last, p=p):
                ops.reg_tail(view(self._img(dist, 0), 0, dc), self._img(flowS, 0), eng.raw[p + ".moduleScaleX.weight"],
                             eng.raw[p + ".moduleScaleX.bias"], eng.raw[p + ".moduleScaleY.weight"],
                             eng.raw[p + ".moduleScaleY.bias"], self._img(flowR, 0), self.out_local if last else None,
                             eng.sf[1], K)
            self._op(run_tail, flowR, min(dist.valid, flowS.valid - K // 2))
            self.flows[l] = flowS
            xflow = flowR

    def _chain(self, prefix: str, idxs: List[int], x: TT, xc: int, l: int, res: Optional[TT], out: TT):
        for j in idxs[:-1]:
            key = f"{prefix}.{j}"
            y = self.new(key, l, self.eng.w[key].cout)
            self.conv(key, x, 0, xc, y, 0)
            x, xc = y, self.eng.w[key].cout
        self.conv(f"{prefix}.{idxs[-1]}", x, 0, xc, out, 0, lrelu=False, res=res)

    # ---- execution ------------------------------------------------------------------------------------------------
    def load_inputs(self, img1: torch.Tensor, img2: torch.Tensor):
        """img1, img2: the full [1,3,H,W] images in [0,1] (any device); this rank's slab + halo rows are cut out here and
        the per-channel mean is subtracted (src/models.py:321-323); rows outside the frame stay zero, which is the zero
        padding the first convolution sees."""
        E, own = self.E, self.own[1]
        R0 = self.rank * own
        mean = self.eng.cfg.mean
        for dst, src, m in ((self.in1, img1, mean[:3]), (self.in2, img2, mean[3:])):
            dst.zero_()
            g0, g1 = max(0, R0 - E), min(self.H, R0 + own + E)
            mm = torch.tensor(m, device=dst.device, dtype=torch.float32).view(1, 3, 1, 1)
            dst[:, :, g0 - (R0 - E):g1 - (R0 - E)] = src[:, :, g0:g1].to(dst.device) - mm

    def owned_output(self) -> torch.Tensor:
        l = self.eng.cfg.lowest_level
        return self.out_local[:, :, self.E:self.E + self.own[l]]

    def local_bounds(self) -> torch.Tensor:
        """[largest vertical displacement (px at its level, + 1 px bilinear footprint) that steered a backwarp on rows this
        rank's results depend on, fp16-range flag of this device].  Device tensor of 2 floats; reduced with MAX over the
        ranks by the group's ``run``."""
        m = torch.zeros((), device=self.eng.device, dtype=torch.float32)
        for f, v in self.warp_flows:
            E, own = self.E, self.own[f.level]
            rows = f.t[:, E - v:E + own + v, :, 1]
            m = torch.maximum(m, rows.abs().max() * self.eng.sf[f.level] + 1.0)
        flag = 0.0
        if self.eng.precision == "f16c":
            torch.cuda.current_stream().synchronize()
            flag = float(int(self.eng.lib.pivlfn_f16_range_flag(1)) != 0)
        return torch.stack([m, torch.full((), flag, device=self.eng.device)])

    def check_bounds(self, reduced: torch.Tensor):
        """Raise -- on EVERY rank, ``reduced`` being the MAX over ranks of ``local_bounds`` -- when a seam could be wrong."""
        m, flag = float(reduced[0]), float(reduced[1])
        if not (m <= self.wr):          # also catches NaN
            raise TiledBoundsError(f"tiled mode: a vertical displacement of {m - 1.0:.2f} px (+1 px bilinear footprint) "
                                   f"exceeds warp_reach={self.wr}; rebuild the plan with a larger halo / warp_reach")
        if flag:
            raise TiledBoundsError("tiled mode: an activation left the fp16 range in precision 'f16c'; rebuild the plan "
                                   "from an engine with precision 'tf32c'")

    def clear_range_flag(self):
        if self.eng.precision == "f16c":
            self.eng.lib.pivlfn_f16_range_flag_clear(torch.cuda.current_stream().cuda_stream)


class TiledBoundsError(RuntimeError):
    """The data-dependent assumptions of a tiled run (backwarp reach, fp16 range) did not hold: the result is discarded."""


def _zero_outside(tt: TT, plan: TiledPlan):
    E, own = plan.E, plan.own[tt.level]
    if plan.top:
        tt.t[:, :E].zero_()
    if plan.bottom:
        tt.t[:, E + own:].zero_()


class LoopbackGroup:
    """P ranks run in lock-step inside ONE process on one device (tests): communication steps are tensor copies."""

    def __init__(self, plans: List[TiledPlan]):
        self.plans = plans
        n = len(plans[0].steps)
        assert all(len(p.steps) == n for p in plans), "ranks must record identical step sequences"

    def run(self):
        P = len(self.plans)
        self.plans[0].clear_range_flag()
        self._steps()
        red = torch.stack([p.local_bounds() for p in self.plans]).max(0).values
        for p in self.plans:
            p.check_bounds(red)

    def _steps(self):
        P = len(self.plans)
        for k in range(len(self.plans[0].steps)):
            kind = self.plans[0].steps[k].kind
            assert all(p.steps[k].kind == kind for p in self.plans)
            if kind == "op":
                for p in self.plans:
                    p.steps[k].fn()
            elif kind == "zero":
                for p in self.plans:
                    _zero_outside(p.steps[k].tt, p)
            elif kind == "exchange":
                for r, p in enumerate(self.plans):
                    tt, E, own = p.steps[k].tt, p.E, p.own[p.steps[k].tt.level]
                    if r > 0:
                        up = self.plans[r - 1].steps[k].tt
                        tt.t[:, :E].copy_(up.t[:, own:own + E])
                    if r < P - 1:
                        dn = self.plans[r + 1].steps[k].tt
                        tt.t[:, E + own:].copy_(dn.t[:, E:2 * E])
            elif kind == "allreduce":
                tot = sum(p.steps[k].src for p in self.plans)
                for p in self.plans:
                    p.steps[k].src.copy_(tot)
            elif kind == "allgather":
                for p in self.plans:
                    dst = p.steps[k].dst
                    o2 = dst.shape[1] // P
                    for r, q in enumerate(self.plans):
                        dst[:, r * o2:(r + 1) * o2].copy_(q.steps[k].src[:, 1:1 + o2])


class DistGroup:
    """One process per GPU: NCCL send/recv of halo row blocks with the row neighbours, all-reduce, all-gather."""

    def __init__(self, plan: TiledPlan, use_graph: Optional[bool] = None):
        import os
        import torch.distributed as dist
        self.plan, self.dist = plan, dist
        self.use_graph = (os.environ.get("PIVLFN_TILED_GRAPH", "1") != "0") if use_graph is None else use_graph
        self._segments = None            # [(kind, payload)]: ("graph", CUDAGraph) for a run of local steps, ("comm", Step)
        self._warm = False

    def run(self):
        """One forward; raises TiledBoundsError on all ranks together when any rank saw a displacement beyond the
        provisioned backwarp reach or an fp16 range overflow (2-float MAX all-reduce).

        The ~170 kernel launches of a rank are captured into CUDA graphs, one per run of LOCAL steps between two communication
        steps (about 40 segments); the NCCL calls themselves stay eager -- capturing the P2P groups into the graphs as well
        was tried and dead-locked at 2 ranks.  At 8 ranks a rank's share of the GPU work is ~20 ms, against which ~200 eager
        ctypes launches from Python were a quarter of the wall time."""
        p = self.plan
        p.clear_range_flag()
        if not self.use_graph:
            self._steps()
        else:
            if not self._warm:
                self._steps()                    # eager once: lazy module loading, NCCL communicator set-up
                torch.cuda.current_stream().synchronize()
                self._warm = True
            if self._segments is None:
                self._compile()
            for kind, payload in self._segments:
                if kind == "graph":
                    payload.replay()
                else:
                    self._comm(payload)
        red = p.local_bounds()
        self.dist.all_reduce(red, op=self.dist.ReduceOp.MAX)
        p.check_bounds(red)

    def _compile(self):
        """Split the step list at the communication steps and capture every local run into its own CUDA graph (each rank
        captures only its own kernels: no other rank is involved while a capture is in progress)."""
        import gc
        p = self.plan
        segs, run = [], []

        def flush():
            if not run:
                return
            steps = list(run)
            run.clear()
            g = torch.cuda.CUDAGraph()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, stream=torch.cuda.Stream(device=p.eng.device), capture_error_mode="thread_local"):
                    for st in steps:
                        if st.kind == "op":
                            st.fn()
                        else:
                            _zero_outside(st.tt, p)
            finally:
                if gc_was_on:
                    gc.enable()
            segs.append(("graph", g))

        for st in p.steps:
            if st.kind in ("op", "zero"):
                run.append(st)
            else:
                flush()
                segs.append(("comm", st))
        flush()
        self._segments = segs

    def _steps(self):
        p = self.plan
        for st in p.steps:
            if st.kind == "op":
                st.fn()
            elif st.kind == "zero":
                _zero_outside(st.tt, p)
            else:
                self._comm(st)

    def _comm(self, st):
        dist, p = self.dist, self.plan
        E = p.E
        if True:
            if st.kind == "exchange":
                t, own = st.tt.t, p.own[st.tt.level]
                reqs = []
                bufs = []
                for n in range(t.shape[0]):
                    if p.rank > 0:
                        reqs.append(dist.P2POp(dist.isend, t[n, E:2 * E], p.rank - 1))
                        reqs.append(dist.P2POp(dist.irecv, t[n, :E], p.rank - 1))
                    if p.rank < p.world - 1:
                        reqs.append(dist.P2POp(dist.isend, t[n, own:own + E], p.rank + 1))
                        reqs.append(dist.P2POp(dist.irecv, t[n, E + own:], p.rank + 1))
                if reqs:
                    for w in dist.batch_isend_irecv(reqs):
                        w.wait()
            elif st.kind == "allreduce":
                dist.all_reduce(st.src)
            elif st.kind == "allgather":
                o2 = st.dst.shape[1] // p.world
                for n in range(st.src.shape[0]):
                    piece = st.src[n, 1:1 + o2].contiguous()
                    dist.all_gather_into_tensor(st.dst[n], piece)


# =================================================================================================================
# P16 pipeline (engines with precision f16c, see plan16.py): the same row-slab plan over the P16 kernels
# =================================================================================================================
def _r16(c: int) -> int:
    return (c + 15) & ~15


class TiledPlan16(TiledPlan):
    """TiledPlan for an engine that runs the P16 pipeline: activations that feed convolutions are P16 (channel counts rounded
    up to 16 words per pixel, csrc/p16.cuh), every convolution is ``pivlfn_conv_p16`` (incl. the stride-2 NetC layers and
    the 192-channel conv6), the flow heads run as 1xK row planes + ``pivlfn_head_rows_sum``.  Halo validity tracking, exchange
    and zeroing rules are inherited unchanged: a P16 row block is sent as the 4-byte words it is stored in, and all-zero
    words are (hi, lo') = (0, 0)."""

    def conv16(self, key: str, x: TT, xo: int, y: TT, yo: int, lrelu: bool = True, out_fmt: int = ops.OUT_P16,
               cin: Optional[int] = None, yc: Optional[int] = None):
        """stride-1 convolution x[:, :, :, xo:] -> y[:, :, :, yo:] on the whole local buffer (valid rows shrink by KH // 2)."""
        eng = self.eng
        cw = eng.w[key]
        cin = cw.cin if cin is None else cin
        r = cw.kh // 2
        self._need(x, r)
        N, hh, ww = x.t.shape[0], x.t.shape[1], x.t.shape[2]
        w_img, mode = cw.w_f8, P16_MODE
        yc = (_r16(cw.cout) if out_fmt == ops.OUT_P16 else y.t.shape[3] - yo) if yc is None else yc

        def run():
            for n in range(N):
                xv = view(self._img(x, n), xo, _r16(cin))
                yv = view(self._img(y, n), yo, yc)
                ops.conv_p16(xv, 1, hh, ww, cin, w_img, mode, cw.bias, yv, cw.cout, cw.kh, cw.kw, 1, lrelu, out_fmt, 0, eng.flag)
        self._op(run, y, x.valid - r)

    def down(self, kind: str, x: TT, y: TT, key: Optional[str] = None, C: int = 4):
        if kind == "pool":
            return super().down(kind, x, y, key, C)
        # 3x3 stride-2 NetC convolution on P16 buffers (conv6: two 96-channel halves)
        eng, N, E = self.eng, x.t.shape[0], self.E
        halves = [(key, 0)] if key in eng.w and eng.w[key].w_f8 is not None else [(key + "#a", 0), (key + "#b", 96)]

        def call(a: torch.Tensor, b: torch.Tensor):
            for k, off in halves:
                cw = eng.w[k]
                ops.conv_p16(view(a, 0, _r16(cw.cin)), 1, a.shape[1], a.shape[2], cw.cin, cw.w_f8, P16_MODE, cw.bias,
                             view(b, off, _r16(cw.cout)), cw.cout, 3, 3, 2, True, ops.OUT_P16, 0, eng.flag)

        if x.tiled and not y.tiled:
            self._need(x, 2)
            own2 = self.own[y.level]
            tmp = self._E(N, own2 + 2, self.Wl[y.level], y.t.shape[3])

            def run():
                for n in range(N):
                    call(self._img(x, n, E - 2, 2 * own2 + 4), tmp[n:n + 1])
            self.steps.append(Step("op", fn=run))
            self.steps.append(Step("allgather", src=tmp, dst=y.t))
            return
        self._need(x, 1)

        def run2():
            for n in range(N):
                a, b, _ = self._down_views(x, y, n, 1)
                call(a, b)
        v = (x.valid - 1) // 2
        self._op(run2, y, min(v, E // 2) if y.tiled else v)

    def _head(self, key: str, x: TT, l: int, res: Optional[TT], out: TT, out16: Optional[TT], out16_off: int, planes: torch.Tensor):
        """KxK 32 -> 2 flow head: 1xK convolution to K row planes, then the K-row gather-sum (+ bias + residual flow)."""
        eng = self.eng
        K = KSIZE[l]
        rw = eng.w[key + "#rows"]
        self._need(x, 0)
        if res is not None:
            self._need(res, 0)
        hh, ww = x.t.shape[1], x.t.shape[2]

        def run():
            ops.conv_p16(view(self._img(x, 0), 0, _r16(rw.cin)), 1, hh, ww, rw.cin, rw.w_f8, P16_MODE, None, view(planes.view(1, K, hh * ww, 2)),
                         2 * K, 1, K, 1, False, ops.OUT_PLANES, 2 * hh * ww, eng.flag)
            ops.head_rows_sum(planes, K, eng.w[key].bias, self._img(res, 0) if res is not None else None, self._img(out, 0),
                              view(self._img(out16, 0), out16_off, 16) if out16 is not None else None, 1, hh, ww, eng.flag)
        v = x.valid - K // 2
        if res is not None:
            v = min(v, res.valid)
        self._op(run, out, v)
        if out16 is not None:
            out16.valid = out.valid

    def _chain16(self, prefix: str, idxs: List[int], x: TT, l: int, res: Optional[TT], out: TT, out16: Optional[TT] = None,
                 out16_off: int = 0, planes: Optional[torch.Tensor] = None):
        for j in idxs[:-1]:
            key = f"{prefix}.{j}"
            y = self.new(key, l, _r16(self.eng.w[key].cout))
            self.conv16(key, x, 0, y, 0)
            x = y
        self._head(f"{prefix}.{idxs[-1]}", x, l, res, out, out16, out16_off, planes)

    def _build(self):
        eng, cfg, E = self.eng, self.eng.cfg, self.E
        fl = eng.flag
        self.in1 = self._E(1, 3, self.rows(1), self.W)
        self.in2 = self._E(1, 3, self.rows(1), self.W)
        img: Dict[int, TT] = {1: self.new("img1", 1, 4, N=2)}
        img_pad = TT(self._Z(2, self.rows(1), self.W + 8, 4), 1, True, "img_pad")
        img_pad.valid = E

        def prep():
            ops.prep_images(self.in1, self.in2, img[1].t, (0.0,) * 6, img_pad.t)
        self.steps.append(Step("op", fn=prep))
        img[1].valid = E
        for l in range(2, 7):
            img[l] = self.new(f"img{l}", l, 4, N=2)
            self.down("pool", img[l - 1], img[l])
        # ---- NetC on both images ------------------------------------------------------------------------------------
        feats: Dict[int, TT] = {}
        x, lvl = img[1], 1
        for i, (seq, idx, cin, cout, k, st) in enumerate(NETC):
            key = f"NetC.{seq}.{idx}"
            if st == 2:
                lvl += 1
                y = self.new(key, lvl, _r16(cout), N=2)
                self.down("conv", x, y, key)
            elif i == 0:
                y = self.new(key, lvl, 32, N=2)
                cw = eng.w[key]
                self._need(img_pad, 3)
                hh, ww = y.t.shape[1], y.t.shape[2]

                def run_stem(y=y, cw=cw, hh=hh, ww=ww):
                    for n in range(2):
                        ops.conv_stem_p16(self._img(img_pad, n), 1, hh, ww, cw.w_f16, cw.bias, view(self._img(y, n)), True, fl)
                self._op(run_stem, y, img_pad.valid - 3)
            else:
                y = self.new(key, lvl, _r16(cout), N=2)
                self.conv16(key, x, 0, y, 0)
            x = y
            if NETC_LEVEL_END[lvl] == i:
                feats[lvl] = y
        nh = len(cfg.head)
        head_idx = [2 * j for j in range(nh + 1)]
        xflow: Optional[TT] = None
        self.flows: Dict[int, TT] = {}
        for i in reversed(range(len(cfg.levels))):
            l = cfg.levels[i]
            cm = MATCH_FEAT_CH[l]
            cr = 128 if l < 5 else LEVEL_FEAT_CH[l]
            s = 2 if l < 4 else 1
            scale = eng.sf[l]
            K = KSIZE[l]
            feat = feats[l]
            f1 = TT(feat.t[0:1], l, feat.tiled, f"feat1_{l}"); f1.valid = feat.valid
            f2r = TT(feat.t[1:2], l, feat.tiled, f"feat2_{l}"); f2r.valid = feat.valid
            self.handles += [f1, f2r]
            im1 = TT(img[l].t[0:1], l, img[l].tiled, f"im1_{l}"); im1.valid = img[l].valid
            im2 = TT(img[l].t[1:2], l, img[l].tiled, f"im2_{l}"); im2.valid = img[l].valid
            self.handles += [im1, im2]
            planes = self._E(K, self.rows(l) * self.Wl[l], 2)
            Sbuf = self.new(f"Sbuf{l}", l, 2 * cm + 16, zero=True)            # [f1 | backwarp(f2) | flow_M group]
            S_f2w, S_fl = self.alias(Sbuf, f"S_f2w{l}"), self.alias(Sbuf, f"S_fl{l}")
            if l <= 2:
                e = (l - 2) % cfg.n_ext
                self.conv16(f"NetC_ext.{e}.conv_ext.0", f1, 0, Sbuf, 0)
                f2 = self.new(f"f2e{l}", l, cm)                                # fp32 NHWC: read by the cost volume and the warp only
                self.conv16(f"NetC_ext.{e}.conv_ext.0", f2r, 0, f2, 0, out_fmt=ops.OUT_F32)
                f2_p16 = False
            else:
                self.copy(f1, 0, cm, Sbuf, 0)
                f2, f2_p16 = f2r, True
            # ---- Matching ------------------------------------------------------------------------------------
            flowU = None
            if xflow is not None:
                flowU = self.new(f"flowU{l}", l, 2)
                self.up(xflow, 2, eng.raw[f"NetE_M.{i}.upConv_M.weight"], flowU)
            self._need(Sbuf, 0)
            self._need(f2, 3 * s + (self.wr if flowU is not None else 0))
            if flowU is not None:
                self._need(flowU, 3 * s)
                self._steers_warp(flowU, flowU.valid)
            vin = min(Sbuf.valid, f2.valid - 3 * s - (self.wr if flowU is not None else 0),
                      flowU.valid - 3 * s if flowU is not None else 1 << 30)
            corrU = self.new(f"corrU{l}", l, 64, zero=True)                    # P16: input of conv_M
            if s == 2:
                corr_rows = (self.rows(l) + 1) // 2
                corr = self._Z(1, corr_rows, (self.Wl[l] + 1) // 2, 52)

                def run_corr(Sbuf=Sbuf, f2=f2, f2_p16=f2_p16, flowU=flowU, corr=corr, corrU=corrU, cm=cm, scale=scale, i=i):
                    a, b = self._img(Sbuf, 0), self._img(f2, 0)
                    ops.corr_p16(view(a, 0, cm), True, view(b, 0, cm), f2_p16, self._img(flowU, 0) if flowU is not None else None, scale,
                                 view(corr, 0, 49), False, 1, a.shape[1], a.shape[2], cm, 2, True, fl)
                    ops.deconv4x4s2_dw_p16(view(corr, 0, 49), 1, corr.shape[1], corr.shape[2], 49,
                                           eng.raw[f"NetE_M.{i}.upCorr_M.weight"], view(corrU.t), fl)
                # half-resolution cost volume rows 2 apart: valid rows shrink by 2 through the 4x4 up-convolution
                self._op(run_corr, corrU, vin - 2)
            else:
                def run_corr1(Sbuf=Sbuf, f2=f2, f2_p16=f2_p16, flowU=flowU, corrU=corrU, cm=cm, scale=scale):
                    a, b = self._img(Sbuf, 0), self._img(f2, 0)
                    ops.corr_p16(view(a, 0, cm), True, view(b, 0, cm), f2_p16, self._img(flowU, 0) if flowU is not None else None, scale,
                                 view(corrU.t), True, 1, a.shape[1], a.shape[2], cm, 1, True, fl)
                self._op(run_corr1, corrU, vin)
            flowM = self.new(f"flowM{l}", l, 2)
            self._chain16(f"NetE_M.{i}.conv_M", head_idx, corrU, l, flowU, flowM, S_fl, 2 * cm, planes)
            # ---- Subpixel ------------------------------------------------------------------------------------
            self._need(f2, self.wr)
            self._need(flowM, 0)
            self._steers_warp(flowM, min(flowM.valid, f2.valid - self.wr))

            def run_warp(f2=f2, f2_p16=f2_p16, flowM=flowM, Sbuf=Sbuf, cm=cm, scale=scale):
                a, b = self._img(f2, 0), self._img(Sbuf, 0)
                ops.warp_p16(view(a, 0, cm), f2_p16, self._img(flowM, 0), scale, view(b, cm, cm), 1, a.shape[1], a.shape[2], cm, fl)
            self._op(run_warp, S_f2w, min(flowM.valid, f2.valid - self.wr))
            Sall = self.alias(Sbuf, f"Sall{l}")
            Sall.valid = min(Sbuf.valid, S_f2w.valid, S_fl.valid)
            flowS = self.new(f"flowS{l}", l, 2)
            self._chain16(f"NetE_S.{i}.conv_S", head_idx, Sall, l, flowM, flowS, None, 0, planes)
            # ---- Regularization ------------------------------------------------------------------------------
            partial = self._E(1, ops.flow_mean_parts(), 2)
            r0, nr = (E, self.own[l]) if flowS.tiled else (0, self.Hl[l])
            frac = float(self.rows(l)) / float(self.Hl[l])     # reg_input divides by the LOCAL pixel count

            def run_mean(flowS=flowS, partial=partial, r0=r0, nr=nr):
                ops.flow_mean(self._img(flowS, 0, r0, nr).contiguous(), partial)
            self.steps.append(Step("op", fn=run_mean))
            if flowS.tiled:
                self.steps.append(Step("allreduce", src=partial))
                self.steps.append(Step("op", fn=lambda partial=partial, frac=frac: partial.mul_(frac)))
            Rbuf = self.new(f"Rbuf{l}", l, cr + 16, zero=True)
            R_in = self.alias(Rbuf, f"R_in{l}")
            self._need(im2, self.wr)
            self._need(im1, 0)
            self._need(flowS, 0)

            def run_ri(im1=im1, im2=im2, flowS=flowS, partial=partial, Rbuf=Rbuf, cr=cr, scale=scale):
                ops.reg_input_p16(self._img(im1, 0), self._img(im2, 0), self._img(flowS, 0), scale, partial,
                                  view(self._img(Rbuf, 0), cr, 16), fl)
            self._steers_warp(flowS, min(im1.valid, flowS.valid, im2.valid - self.wr))
            self._op(run_ri, R_in, min(im1.valid, flowS.valid, im2.valid - self.wr))
            if l < 5:
                self.conv16(f"NetE_R.{i}.moduleFeat.0", f1, 0, Rbuf, 0)
            else:
                self.copy(f1, 0, cr, Rbuf, 0)
            Rall = self.alias(Rbuf, f"Rall{l}")
            Rall.valid = min(Rbuf.valid, R_in.valid)
            x = Rall
            for j in range(len(CONV_R)):
                key = f"NetE_R.{i}.conv_R.{2 * j}"
                y = self.new(key, l, _r16(eng.w[key].cout))
                self.conv16(key, x, 0, y, 0)
                x = y
            dc = DIST_CH[l]
            dist = self.new(f"dist{l}", l, (dc + 3) & ~3)
            if l < 5:
                dist0 = self.new(f"dist0{l}", l, _r16(dc), zero=True)
                self.conv16(f"NetE_R.{i}.conv_dist_R.0", x, 0, dist0, 0, lrelu=False)
                self.conv16(f"NetE_R.{i}.conv_dist_R.1", dist0, 0, dist, 0, lrelu=False, out_fmt=ops.OUT_F32)
            else:
                self.conv16(f"NetE_R.{i}.conv_dist_R.0", x, 0, dist, 0, lrelu=False, out_fmt=ops.OUT_F32)
            flowR = self.new(f"flowR{l}", l, 2)
            last = l == cfg.lowest_level
            if last:
                self.out_local = self._E(1, 2, self.rows(l), self.Wl[l])
            self._need(flowS, K // 2)
            self._need(dist, 0)
            p = f"NetE_R.{i}"

            def run_tail(dist=dist, flowS=flowS, flowR=flowR, dc=dc, K=K, last=last, p=p):
                ops.reg_tail(view(self._img(dist, 0), 0, dc), self._img(flowS, 0), eng.raw[p + ".moduleScaleX.weight"],
                             eng.raw[p + ".moduleScaleX.bias"], eng.raw[p + ".moduleScaleY.weight"],
                             eng.raw[p + ".moduleScaleY.bias"], self._img(flowR, 0), self.out_local if last else None,
                             eng.sf[1], K)
            self._op(run_tail, flowR, min(dist.valid, flowS.valid - K // 2))
            self.flows[l] = flowS
            xflow = flowR

    # the engine-owned range flag replaces the library-global one of the fp32-activation kernels
    def local_bounds(self) -> torch.Tensor:
        m = torch.zeros((), device=self.eng.device, dtype=torch.float32)
        for f, v in self.warp_flows:
            E, own = self.E, self.own[f.level]
            m = torch.maximum(m, f.t[:, E - v:E + own + v, :, 1].abs().max() * self.eng.sf[f.level] + 1.0)
        return torch.stack([m, self.eng.flag[0].to(torch.float32)])

    def clear_range_flag(self):
        self.eng.flag.zero_()


def make_tiled_plan(eng: Engine, H: int, W: int, rank: int, world: int, halo: int = 16, warp_reach: int = 8) -> TiledPlan:
    """The tiled plan that matches the engine's pipeline (P16 for precision f16c, else the fp32-activation kernels)."""
    cls = TiledPlan16 if getattr(eng, "p16", False) else TiledPlan
    return cls(eng, H, W, rank, world, halo=halo, warp_reach=warp_reach)
