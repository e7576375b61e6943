"""ctypes binding of csrc/liblcr.so — the C-ABI declared in include/lcr.h.

The library is loaded from the repo tree (never from site-packages) so that the driver's
"which .so did the process load" check sees it.  Loading fails LOUDLY: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

LCR_MAX_ANCHORS = 16
LCR_MAX_LEVELS = 8
LCR_MAX_TOPK = 8192
LCR_MAX_NMS_BOXES = 32768
# LcrStatus (include/lcr.h)
LCR_OK, LCR_ERR_INVALID_ARG, LCR_ERR_CAPACITY, LCR_ERR_WORKSPACE, LCR_ERR_ALIGNMENT, LCR_ERR_CUDA, LCR_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)


class LcrRpnLevel(C.Structure):
    _fields_ = [
        ("objectness", C.c_void_p),
        ("deltas", C.c_void_p),
        ("anchors", C.c_void_p),
        ("h", C.c_int),
        ("w", C.c_int),
        ("stride", C.c_int),
        ("reserved", C.c_int),
        ("base_anchors", C.c_float * (LCR_MAX_ANCHORS * 4)),
    ]


class LcrRpnCfg(C.Structure):
    _fields_ = [
        ("num_anchors", C.c_int),
        ("pre_nms_top_n", C.c_int),
        ("score_thresh", C.c_float),
        ("score_strict", C.c_int),
        ("min_size", C.c_float),
        ("img_h", C.c_int),
        ("img_w", C.c_int),
        ("topk_on_sigmoid", C.c_int),
        ("decode_weights", C.c_float * 4),
        ("xform_clip", C.c_float),
    ]


class LcrFeatLevel(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("N", C.c_int),
        ("H", C.c_int),
        ("W", C.c_int),
        ("reserved", C.c_int),
        ("sn", C.c_int64),
        ("sc", C.c_int64),
        ("sh", C.c_int64),
        ("sw", C.c_int64),
        ("spatial_scale", C.c_float),
        ("reserved2", C.c_int),
    ]


# name -> (restype, argtypes); also the list tests/test_abi_symbols.py checks against include/lcr.h
SIGNATURES = {
    "lcr_version": (C.c_int, []),
    "lcr_error_string": (C.c_char_p, [C.c_int]),
    "lcr_last_cuda_error": (C.c_int, []),
    "lcr_launch_count": (C.c_uint64, []),
    "lcr_set_tuning": (C.c_int, [C.c_char_p, C.c_char_p]),
    "lcr_anchors_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int, C.c_void_p]),
    "lcr_clip_boxes_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "lcr_filter_small_boxes_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "lcr_box_decode_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_f32p, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p]),
    "lcr_rpn_select_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "lcr_rpn_select_f32": (C.c_int, [C.POINTER(LcrRpnLevel), C.c_int, C.c_int, C.POINTER(LcrRpnCfg), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcr_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "lcr_nms_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_float,
                              C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcr_gather_kept_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcr_level_map_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                    C.c_void_p, C.c_void_p]),
    "lcr_roi_align_fwd_f32": (C.c_int, [C.POINTER(LcrFeatLevel), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lcr_roi_align_bwd_f32": (C.c_int, [C.c_void_p, C.POINTER(LcrFeatLevel), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "lcr_nchw_to_nhwc_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "lcr_nhwc_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "lcr_paste_masks_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_uint8, C.c_void_p, C.c_void_p]),
    "lcr_paste_masks_tv_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "lcr_rpn_concat_levels_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcr_pack_records_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lcr_box_iou_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "lcr_box_iou_max_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcr_match_boxes_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcr_mask_targets_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]),
    "lcr_mask_tail_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lcr_mask_region_counts_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class LcrError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the .so is missing or stale and nvcc is available) liblcr.so."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing and _build.is_stale():
        _build.build()          # serialised across processes by a file lock (8 torchrun ranks may arrive here together)
    if not os.path.exists(path):
        raise LcrError(f"{path} is missing: build it with __graft_entry__.build() — there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export what lcr.h declares
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.lcr_error_string(rc).decode()
        extra = f" (cudaError {lib.lcr_last_cuda_error()})" if rc == -5 else ""
        raise LcrError(f"liblcr {what}: {msg}{extra}")


def set_tuning(key: str, value=None) -> None:
    """Override (value=None: reset) one of the library's LCR_* tuning switches at run time; the environment itself is only
    read once, when the library first consults a switch."""
    check(load().lcr_set_tuning(key.encode(), None if value is None else str(value).encode()), "set_tuning")


class tuning:
    """Context manager: ``with tuning(LCR_ROI_FWD="staged"): ...`` sets switches and restores the previous state."""

    def __init__(self, **switches):
        self.switches = switches

    def __enter__(self):
        for k, v in self.switches.items():
            set_tuning(k, v)
        return self

    def __exit__(self, *exc):
        for k in self.switches:
            env = os.environ.get(k)
            set_tuning(k, env)
        return False


def launch_count() -> int:
    return int(load().lcr_launch_count())
