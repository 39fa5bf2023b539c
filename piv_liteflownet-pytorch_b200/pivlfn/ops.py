"""Thin torch-tensor wrappers over the C ABI (include/pivlfn.h).

PyTorch supplies device memory and the current stream; all arithmetic happens in libpivlfn.so.
Every wrapper insists on CUDA fp32 tensors: there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise NotImplementedError("pivlfn operators run on CUDA tensors only (no CPU path)")
        if t.dtype != torch.float32:
            raise TypeError("pivlfn operators are fp32")


@dataclass
class View:
    """C channels of a dense NHWC buffer [N,H,W,ld], starting at channel ``off``."""
    base: torch.Tensor
    off: int
    C: int

    @property
    def ptr(self) -> int:
        return self.base.data_ptr() + 4 * self.off

    @property
    def ld(self) -> int:
        return self.base.shape[-1]

    @property
    def N(self):
        return self.base.shape[0]

    @property
    def H(self):
        return self.base.shape[1]

    @property
    def W(self):
        return self.base.shape[2]

    def torch(self) -> torch.Tensor:
        return self.base[..., self.off:self.off + self.C]


def view(t: torch.Tensor, off: int = 0, C: Optional[int] = None) -> View:
    assert t.dim() == 4 and t.is_contiguous()
    return View(t, off, t.shape[-1] - off if C is None else C)


# ------------------------------------------------------------------------------------------------
def corr_nchw(first: torch.Tensor, second: torch.Tensor, stride: int) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(first, second)
    B, Cc, H, W = first.shape
    out = torch.empty((B, 49, int(math.ceil(H / stride)), int(math.ceil(W / stride))),
                      device=first.device, dtype=torch.float32)
    _lib.check(lib.pivlfn_corr_nchw(first.data_ptr(), second.data_ptr(), out.data_ptr(),
                                    B, Cc, H, W, int(stride), _stream()), "corr_nchw")
    return out


def corr_backward_nchw(first: torch.Tensor, second: torch.Tensor, grad_out: torch.Tensor, stride: int, need_first: bool,
                       need_second: bool):
    lib = _lib.load()
    _need_cuda(first, second, grad_out)
    B, Cc, H, W = first.shape
    g1 = torch.empty_like(first) if need_first else None
    g2 = torch.empty_like(first) if need_second else None
    _lib.check(lib.pivlfn_corr_backward_nchw(first.data_ptr(), second.data_ptr(), grad_out.data_ptr(),
                                             g1.data_ptr() if g1 is not None else None,
                                             g2.data_ptr() if g2 is not None else None, B, Cc, H, W, int(stride), _stream()),
               "corr_backward_nchw")
    return g1, g2


def prep_images(img1, img2, out, mean6, out_pad=None):
    lib = _lib.load()
    _need_cuda(img1, img2, out, out_pad)
    B, _, H, W = img1.shape
    arr = (C.c_float * 6)(*[float(m) for m in mean6])
    _lib.check(lib.pivlfn_prep_images(img1.data_ptr(), img2.data_ptr(), out.data_ptr(),
                                      out_pad.data_ptr() if out_pad is not None else None, B, H, W, arr, _stream()),
               "prep_images")


def avgpool2(x: torch.Tensor, out: torch.Tensor):
    lib = _lib.load()
    N, H, W, Cc = x.shape
    _lib.check(lib.pivlfn_avgpool2(x.data_ptr(), out.data_ptr(), N, H, W, Cc, _stream()), "avgpool2")


def copy(src: View, dst: View, npix: int):
    lib = _lib.load()
    assert src.C == dst.C
    _lib.check(lib.pivlfn_copy_nhwc(src.ptr, src.ld, dst.ptr, dst.ld, npix, src.C, _stream()), "copy_nhwc")


def copy_dense(src: torch.Tensor, dst: torch.Tensor):
    """dst <- src for two dense fp32 CUDA tensors of equal size, by a KERNEL (float4 lanes), not cudaMemcpy: the copy
    engines stay free for host <-> device traffic of neighbouring batches (pivlfn.pipeline)."""
    lib = _lib.load()
    n = src.numel()
    assert dst.numel() == n and src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype == torch.float32
    if n % 4 or (src.data_ptr() & 15) or (dst.data_ptr() & 15):
        _lib.check(lib.pivlfn_copy_nhwc(src.data_ptr(), 1, dst.data_ptr(), 1, n, 1, _stream()), "copy_dense")
    else:
        _lib.check(lib.pivlfn_copy_nhwc(src.data_ptr(), 4, dst.data_ptr(), 4, n // 4, 4, _stream()), "copy_dense")


def conv_simt(x: View, N, H, W, w, bias, y: View, KH, KW, stride, lrelu, res: Optional[View] = None):
    lib = _lib.load()
    _lib.check(lib.pivlfn_conv_simt(x.ptr, x.ld, N, H, W, x.C, w.data_ptr(),
                                    bias.data_ptr() if bias is not None else None,
                                    y.ptr, y.ld, y.C, KH, KW, stride, int(lrelu),
                                    res.ptr if res is not None else None, res.ld if res is not None else 0,
                                    _stream()), "conv_simt")


def conv_tc(x: View, N, H, W, w_hi, w_lo, bias, y: View, KH, KW, lrelu, passes, res: Optional[View] = None, w_c16=None,
            stride: int = 1):
    lib = _lib.load()
    _lib.check(lib.pivlfn_conv_tc(x.ptr, x.ld, N, H, W, x.C, w_hi.data_ptr(),
                                  w_lo.data_ptr() if w_lo is not None else None,
                                  w_c16.data_ptr() if w_c16 is not None else None,
                                  bias.data_ptr() if bias is not None else None,
                                  y.ptr, y.ld, y.C, KH, KW, int(stride), int(lrelu),
                                  res.ptr if res is not None else None, res.ld if res is not None else 0,
                                  int(passes), _stream()), "conv_tc")


def conv_s2_tc(x: View, N, H, W, w16, bias, y: View, lrelu, passes):
    """3x3 stride-2 convolution in the fp16 modes; H, W are the INPUT size."""
    lib = _lib.load()
    _lib.check(lib.pivlfn_conv_s2_tc(x.ptr, x.ld, N, H, W, x.C, w16.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     y.ptr, y.ld, y.C, int(lrelu), int(passes), _stream()), "conv_s2_tc")


def conv1x1_pairs_tc(x: View, N, H, W, w_hi, w_lo, w_c16, planes: torch.Tensor, npair: int, passes: int):
    lib = _lib.load()
    _lib.check(lib.pivlfn_conv1x1_pairs_tc(x.ptr, x.ld, N, H, W, x.C, w_hi.data_ptr(),
                                           w_lo.data_ptr() if w_lo is not None else None,
                                           w_c16.data_ptr() if w_c16 is not None else None,
                                           planes.data_ptr(), int(npair), int(passes), _stream()), "conv1x1_pairs_tc")


def flow_head_sum(planes: torch.Tensor, K: int, bias, res: Optional[View], out: View, N, H, W):
    lib = _lib.load()
    _lib.check(lib.pivlfn_flow_head_sum(planes.data_ptr(), int(K), bias.data_ptr() if bias is not None else None,
                                        res.ptr if res is not None else None, res.ld if res is not None else 0,
                                        out.ptr, out.ld, N, H, W, _stream()), "flow_head_sum")


def flow_head(x: View, N, H, W, w, bias, res: Optional[View], out: View, K: int, out2: Optional[View] = None):
    lib = _lib.load()
    _lib.check(lib.pivlfn_flow_head(x.ptr, x.ld, N, H, W, x.C, w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                    res.ptr if res is not None else None, res.ld if res is not None else 0,
                                    out.ptr, out.ld, out2.ptr if out2 is not None else None,
                                    out2.ld if out2 is not None else 0, int(K), _stream()), "flow_head")


def conv_stem_tc(img_pad: torch.Tensor, N, H, W, w_hi, w_lo, bias, y: View, lrelu, passes, w_c16=None):
    lib = _lib.load()
    _lib.check(lib.pivlfn_conv_stem_tc(img_pad.data_ptr(), N, H, W, w_hi.data_ptr(),
                                       w_lo.data_ptr() if w_lo is not None else None,
                                       w_c16.data_ptr() if w_c16 is not None else None,
                                       bias.data_ptr() if bias is not None else None,
                                       y.ptr, y.ld, int(lrelu), int(passes), _stream()), "conv_stem_tc")


def deconv4x4s2_dw(x: View, N, H, W, w, y: View):
    lib = _lib.load()
    _lib.check(lib.pivlfn_deconv4x4s2_dw(x.ptr, x.ld, w.data_ptr(), y.ptr, y.ld, N, H, W, x.C, _stream()),
               "deconv4x4s2_dw")


def warp(x: View, flow: torch.Tensor, scale: float, y: View, N, H, W):
    lib = _lib.load()
    _lib.check(lib.pivlfn_warp_nhwc(x.ptr, x.ld, flow.data_ptr(), float(scale), y.ptr, y.ld, N, H, W, x.C,
                                    _stream()), "warp_nhwc")


def corr_nhwc(f1: View, f2: View, flow: Optional[torch.Tensor], scale: float, out: View, N, H, W, stride, lrelu=True):
    lib = _lib.load()
    _lib.check(lib.pivlfn_corr_nhwc(f1.ptr, f1.ld, f2.ptr, f2.ld, flow.data_ptr() if flow is not None else None,
                                    float(scale), out.ptr, out.ld, N, H, W, f1.C, int(stride), int(lrelu),
                                    _stream()), "corr_nhwc")


def flow_mean_parts() -> int:
    return _lib.load().pivlfn_flow_mean_parts()


def flow_mean(flow: torch.Tensor, partial: torch.Tensor):
    lib = _lib.load()
    N, H, W, _ = flow.shape
    _lib.check(lib.pivlfn_flow_mean(flow.data_ptr(), partial.data_ptr(), N, H, W, _stream()), "flow_mean")


def reg_input(img1, img2, flow, scale, partial, out: View):
    lib = _lib.load()
    N, H, W, _ = flow.shape
    _lib.check(lib.pivlfn_reg_input(img1.data_ptr(), img2.data_ptr(), flow.data_ptr(), float(scale),
                                    partial.data_ptr(), out.ptr, out.ld, N, H, W, _stream()), "reg_input")


def reg_tail(dist: View, flow_in, wx, bx, wy, by, flow_out, out_nchw, final_scale, K):
    lib = _lib.load()
    N, H, W, _ = flow_in.shape
    _lib.check(lib.pivlfn_reg_tail(dist.ptr, dist.ld, flow_in.data_ptr(), wx.data_ptr(), bx.data_ptr(),
                                   wy.data_ptr(), by.data_ptr(), flow_out.data_ptr(),
                                   out_nchw.data_ptr() if out_nchw is not None else None, float(final_scale),
                                   K, N, H, W, _stream()), "reg_tail")


# ---- P16 pipeline (include/pivlfn.h, "the P16 pipeline") -----------------------------------------------------------------
OUT_P16, OUT_F32, OUT_PLANES = 0, 1, 2


def _flag(flag: Optional[torch.Tensor]):
    return flag.data_ptr() if flag is not None else None


def p16_encode(x: View, y: View, npix: int, flag: Optional[torch.Tensor] = None):
    _lib.check(_lib.load().pivlfn_p16_encode(x.ptr, x.ld, x.C, y.ptr, y.ld, npix, _flag(flag), _stream()), "p16_encode")


def p16_decode(x: View, C: int, y: View, npix: int):
    _lib.check(_lib.load().pivlfn_p16_decode(x.ptr, x.ld, C, y.ptr, y.ld, npix, _stream()), "p16_decode")


def conv_p16(x: View, N, H, W, cin, w_img, mode, bias, y: View, cout, KH, KW, stride=1, lrelu=True, out_fmt=OUT_P16,
             plane_stride=0, flag: Optional[torch.Tensor] = None):
    """H, W: input size.  x: P16 view holding ceil16(cin) words per pixel."""
    _lib.check(_lib.load().pivlfn_conv_p16(x.ptr, x.ld, N, H, W, int(cin), w_img.data_ptr(), int(mode),
                                           bias.data_ptr() if bias is not None else None, y.ptr, y.ld, int(cout),
                                           int(KH), int(KW), int(stride), int(lrelu), int(out_fmt), int(plane_stride),
                                           _flag(flag), _stream()), "conv_p16")


def conv_p16_warp(x: View, N, H, W, cin, w_img, mode, bias, y: View, cout, KH, KW, lrelu, wsrc: View, wsrc_p16: bool,
                  wflow: torch.Tensor, wscale: float, wc0: int, wn: int, flag: Optional[torch.Tensor] = None):
    """conv_p16 whose input channels [wc0, wc0 + wn) are backwarp(wsrc, wscale * wflow), gathered inside the kernel; x holds the
    remaining cin - wn channels contiguously."""
    _lib.check(_lib.load().pivlfn_conv_p16_warp(x.ptr, x.ld, N, H, W, int(cin), w_img.data_ptr(), int(mode),
                                                bias.data_ptr() if bias is not None else None, y.ptr, y.ld, int(cout),
                                                int(KH), int(KW), int(lrelu), wsrc.ptr, wsrc.ld, int(wsrc_p16),
                                                wflow.data_ptr(), float(wscale), int(wc0), int(wn), _flag(flag), _stream()),
               "conv_p16_warp")


def conv_p16_tail(x: View, N, H, W, cin, w_img, bias, KH, KW, K, flow_in, wx, bx, wy, by, flow_out, out_nchw, final_scale):
    """conv_dist (mode 4, cout = K*K, no activation) + the Regularization tail (see reg_tail) in one launch."""
    _lib.check(_lib.load().pivlfn_conv_p16_tail(x.ptr, x.ld, N, H, W, int(cin), w_img.data_ptr(),
                                                bias.data_ptr() if bias is not None else None, int(KH), int(KW), int(K),
                                                flow_in.data_ptr(), wx.data_ptr(), bx.data_ptr(), wy.data_ptr(), by.data_ptr(),
                                                flow_out.data_ptr(), out_nchw.data_ptr() if out_nchw is not None else None,
                                                float(final_scale), _stream()), "conv_p16_tail")


def conv_stem_p16(img_pad: torch.Tensor, N, H, W, w_img, bias, y: View, lrelu=True, flag: Optional[torch.Tensor] = None):
    _lib.check(_lib.load().pivlfn_conv_stem_p16(img_pad.data_ptr(), N, H, W, w_img.data_ptr(),
                                                bias.data_ptr() if bias is not None else None, y.ptr, y.ld, int(lrelu),
                                                _flag(flag), _stream()), "conv_stem_p16")


def corr_p16(f1: View, f1_p16: bool, f2: View, f2_p16: bool, flow: Optional[torch.Tensor], scale: float, out: View,
             out_p16: bool, N, H, W, C, stride, lrelu=True, flag: Optional[torch.Tensor] = None):
    _lib.check(_lib.load().pivlfn_corr_p16(f1.ptr, f1.ld, int(f1_p16), f2.ptr, f2.ld, int(f2_p16),
                                           flow.data_ptr() if flow is not None else None, float(scale), out.ptr, out.ld,
                                           int(out_p16), N, H, W, int(C), int(stride), int(lrelu), _flag(flag), _stream()),
               "corr_p16")


def warp_p16(x: View, in_p16: bool, flow: torch.Tensor, scale: float, y: View, N, H, W, C, flag: Optional[torch.Tensor] = None):
    _lib.check(_lib.load().pivlfn_warp_p16(x.ptr, x.ld, int(in_p16), flow.data_ptr(), float(scale), y.ptr, y.ld, N, H, W,
                                           int(C), _flag(flag), _stream()), "warp_p16")


def deconv4x4s2_dw_p16(x: View, N, H, W, C, w, y: View, flag: Optional[torch.Tensor] = None):
    _lib.check(_lib.load().pivlfn_deconv4x4s2_dw_p16(x.ptr, x.ld, w.data_ptr(), y.ptr, y.ld, N, H, W, int(C), _flag(flag),
                                                     _stream()), "deconv4x4s2_dw_p16")


def reg_input_p16(img1, img2, flow, scale, partial, out: View, flag: Optional[torch.Tensor] = None):
    N, H, W, _ = flow.shape
    _lib.check(_lib.load().pivlfn_reg_input_p16(img1.data_ptr(), img2.data_ptr(), flow.data_ptr(), float(scale),
                                                partial.data_ptr(), out.ptr, out.ld, N, H, W, _flag(flag), _stream()),
               "reg_input_p16")


def head_rows_sum(planes: torch.Tensor, K: int, bias, res: Optional[torch.Tensor], out: torch.Tensor,
                  out_p16: Optional[View], N, H, W, flag: Optional[torch.Tensor] = None):
    _lib.check(_lib.load().pivlfn_head_rows_sum(planes.data_ptr(), int(K), bias.data_ptr() if bias is not None else None,
                                                res.data_ptr() if res is not None else None, out.data_ptr(),
                                                out_p16.ptr if out_p16 is not None else None,
                                                out_p16.ld if out_p16 is not None else 0, N, H, W, _flag(flag), _stream()),
               "head_rows_sum")


def head_cols_sum(planes: torch.Tensor, K: int, bias, res: Optional[torch.Tensor], out: torch.Tensor,
                  out_p16: Optional[View], N, H, W, flag: Optional[torch.Tensor] = None):
    """head_rows_sum for the transposed split: planes kx = Kx1 convolution outputs, summed at horizontal offsets."""
    _lib.check(_lib.load().pivlfn_head_cols_sum(planes.data_ptr(), int(K), bias.data_ptr() if bias is not None else None,
                                                res.data_ptr() if res is not None else None, out.data_ptr(),
                                                out_p16.ptr if out_p16 is not None else None,
                                                out_p16.ld if out_p16 is not None else 0, N, H, W, _flag(flag), _stream()),
               "head_cols_sum")


def resize_bilinear(x: torch.Tensor, Ho: int, Wo: int, mul_even: float = 1.0, mul_odd: float = 1.0) -> torch.Tensor:
    """[B,C,H,W] -> [B,C,Ho,Wo], bilinear, align_corners=False (inference.py:46-49,57-61)."""
    lib = _lib.load()
    _need_cuda(x)
    assert x.dim() == 4 and x.is_contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, Ho, Wo), device=x.device, dtype=torch.float32)
    _lib.check(lib.pivlfn_resize_bilinear_nchw(x.data_ptr(), out.data_ptr(), B * Cc, H, W, Ho, Wo,
                                               float(mul_even), float(mul_odd), _stream()), "resize_bilinear_nchw")
    return out


def nl_trans(x: torch.Tensor, y: torch.Tensor, A):
    """stereo/dewarp.py:255-270 on CUDA float32 tensors of equal shape; A: 24 coefficients (host)."""
    lib = _lib.load()
    _need_cuda(x, y)
    assert x.shape == y.shape and len(A) == 24
    x, y = x.contiguous(), y.contiguous()
    nx, ny = torch.empty_like(x), torch.empty_like(y)
    arr = (C.c_float * 24)(*[float(a) for a in A])
    _lib.check(lib.pivlfn_nl_trans(x.data_ptr(), y.data_ptr(), arr, nx.data_ptr(), ny.data_ptr(), x.numel(), _stream()),
               "nl_trans")
    return nx, ny


def stereo_2d3c(flow_left: torch.Tensor, flow_right: torch.Tensor, A_left, A_right, calib, fps, theta, beta) -> torch.Tensor:
    """stereo_run._stereo_cal (:153-163) on both camera flows + willert (stereo/vel3d.py:4-24), fused.
    flow_*: [B,2,H,W] CUDA float32 (estimate(..., tensor=True)); A_*: 24 mapping coefficients each, or None for willert
    only; calib: None or the real-length calibration factor; theta, beta: signed camera angles in radians (left, right).
    Returns [B,H,W,3] float32 (U, V, W)."""
    lib = _lib.load()
    _need_cuda(flow_left, flow_right)
    assert flow_left.shape == flow_right.shape and flow_left.dim() == 4 and flow_left.shape[1] == 2
    fl, fr = flow_left.contiguous(), flow_right.contiguous()
    B, _, H, W = fl.shape
    out = torch.empty((B, H, W, 3), device=fl.device, dtype=torch.float32)
    al = (C.c_float * 24)(*[float(a) for a in A_left]) if A_left is not None else None
    ar = (C.c_float * 24)(*[float(a) for a in A_right]) if A_right is not None else None
    tt = [math.tan(float(t)) for t in theta]
    tb = [math.tan(float(b)) for b in beta]
    _lib.check(lib.pivlfn_stereo_2d3c(fl.data_ptr(), fr.data_ptr(), al, ar, int(calib is not None),
                                      float(calib) if calib is not None else 1.0, float(fps), tt[0], tt[1], tb[0], tb[1],
                                      out.data_ptr(), B, H, W, _stream()), "stereo_2d3c")
    return out


def launch_count() -> int:
    return int(_lib.load().pivlfn_launch_count())
