"""ctypes binding of libpivlfn.so (the C ABI declared in include/pivlfn.h)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpivlfn.so")

_p, _i, _f, _ll, _d = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double

# name -> (restype, argtypes); mirrors include/pivlfn.h one to one (checked by tests/test_abi.py)
PROTOTYPES = {
    "pivlfn_abi_version": (_i, []),
    "pivlfn_device_is_sm100": (_i, []),
    "pivlfn_launch_count": (_ll, []),
    "pivlfn_corr_nchw": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pivlfn_corr_backward_nchw": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pivlfn_prep_images": (_i, [_p, _p, _p, _p, _i, _i, _i, C.POINTER(_f), _p]),
    "pivlfn_avgpool2": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "pivlfn_conv_simt": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "pivlfn_conv_tc": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "pivlfn_f16_range_flag": (_i, [_i]),
    "pivlfn_f16_range_flag_clear": (_i, [_p]),
    "pivlfn_conv_s2_tc": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "pivlfn_conv1x1_pairs_tc": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p]),
    "pivlfn_flow_head": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _i, _i, _p]),
    "pivlfn_flow_head_sum": (_i, [_p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "pivlfn_conv_stem_tc": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "pivlfn_deconv4x4s2_dw": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pivlfn_warp_nhwc": (_i, [_p, _i, _p, _f, _p, _i, _i, _i, _i, _i, _p]),
    "pivlfn_corr_nhwc": (_i, [_p, _i, _p, _i, _p, _f, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pivlfn_flow_mean_parts": (_i, []),
    "pivlfn_flow_mean": (_i, [_p, _p, _i, _i, _i, _p]),
    "pivlfn_reg_input": (_i, [_p, _p, _p, _f, _p, _p, _i, _i, _i, _i, _p]),
    "pivlfn_reg_tail": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _p]),
    "pivlfn_copy_nhwc": (_i, [_p, _i, _p, _i, _ll, _i, _p]),
    "pivlfn_p16_encode": (_i, [_p, _i, _i, _p, _i, _ll, _p, _p]),
    "pivlfn_p16_decode": (_i, [_p, _i, _i, _p, _i, _ll, _p]),
    "pivlfn_conv_p16": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _ll, _p, _p]),
    "pivlfn_conv_p16_tail": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _f, _p]),
    "pivlfn_conv_p16_warp": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _f, _i, _i, _p, _p]),
    "pivlfn_conv_stem_p16": (_i, [_p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p]),
    "pivlfn_corr_p16": (_i, [_p, _i, _i, _p, _i, _i, _p, _f, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "pivlfn_warp_p16": (_i, [_p, _i, _i, _p, _f, _p, _i, _i, _i, _i, _i, _p, _p]),
    "pivlfn_deconv4x4s2_dw_p16": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "pivlfn_reg_input_p16": (_i, [_p, _p, _p, _f, _p, _p, _i, _i, _i, _i, _p, _p]),
    "pivlfn_head_rows_sum": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "pivlfn_head_cols_sum": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "pivlfn_nl_trans": (_i, [_p, _p, C.POINTER(_f), _p, _p, _ll, _p]),
    "pivlfn_stereo_2d3c": (_i, [_p, _p, C.POINTER(_f), C.POINTER(_f), _i, _f, _f, _d, _d, _d, _d, _p, _i, _i, _i, _p]),
    "pivlfn_resize_bilinear_nchw": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _f, _p]),
}

_lib = None


class PivlfnError(RuntimeError):
    pass


def load():
    """Load (building first if the in-tree library is missing and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    # (re)build in-tree when the sources changed since the last build (no-op when the stamp matches; a box
    # without nvcc uses the shipped library as is)
    import importlib.util
    spec = importlib.util.spec_from_file_location("pivlfn_build", os.path.join(os.path.dirname(_HERE), "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    if not os.path.isfile(LIB_PATH):
        raise PivlfnError("libpivlfn.so is missing: run `python piv_liteflownet-pytorch_b200/build.py` "
                          "(there is no CPU / eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str):
    if code == 0:
        return
    if code == -1:
        raise AssertionError(f"{what}: invalid shape/stride/alignment argument (PIVLFN_EINVAL)")
    if code == -2:
        raise NotImplementedError(f"{what}: unsupported configuration (PIVLFN_EUNSUPPORTED)")
    if code == -3:
        raise PivlfnError(f"{what}: CUDA driver entry point unavailable (PIVLFN_EDRIVER)")
    raise PivlfnError(f"{what}: CUDA error {code}")
