"""Loader of csrc/lcr_torch.so — the thin torch extension (C++ autograd node of RoIAlign, nms) above the C-ABI.

The extension only removes Python call overhead on the small, launch-latency-bound shapes; it computes nothing
itself (every kernel lives in liblcr.so).  ``LCR_TORCH_EXT=0`` keeps the ctypes path (A/B runs, tests)."""
from __future__ import annotations

import importlib.util
import os

from . import _lib
from . import build as _build

_mod = None
_tried = False


def load():
    """Returns the extension module, or None when it is disabled or cannot be built/loaded here (the ctypes path over
    the same liblcr.so kernels is then used — never a CPU fallback)."""
    global _mod, _tried
    if _tried:
        return _mod
    _tried = True
    if os.environ.get("LCR_TORCH_EXT", "1") == "0":
        return None
    _lib.load()  # liblcr.so first: fails loudly when the CUDA library itself is missing
    if _build.ext_is_stale():
        try:
            _build.build_torch_ext()
        except Exception:  # no compiler on this box: keep the ctypes path
            if not os.path.exists(_build.EXT_PATH):
                return None
    import torch  # noqa: F401  (libtorch symbols must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("lcr_torch", _build.EXT_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.lcr_version() == _lib.load().lcr_version()
    _mod = mod
    return mod
