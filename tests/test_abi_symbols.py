"""CPU: the C-ABI library loads and exports every symbol include/lcr.h declares; host-side argument
validation works without a GPU (no compute launches here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "lcr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcr_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from livecell_instance_segmentation_b200 import _lib
    return _lib.load()


def test_header_and_binding_agree():
    from livecell_instance_segmentation_b200 import _lib
    assert declared_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.lcr_version() == 100
    assert b"workspace" in lib.lcr_error_string(-3)
    assert os.path.dirname(lib._name).endswith(os.path.join("livecell-instance-segmentation_b200", "csrc"))   # in-tree


def test_struct_layouts_match_header():
    from livecell_instance_segmentation_b200 import _lib
    assert ctypes.sizeof(_lib.LcrRpnLevel) == 3 * 8 + 4 * 4 + _lib.LCR_MAX_ANCHORS * 16
    assert ctypes.sizeof(_lib.LcrRpnCfg) == 13 * 4
    assert ctypes.sizeof(_lib.LcrFeatLevel) == 64
    hdr = open(os.path.join(ROOT, "include", "lcr.h")).read()
    assert f"#define LCR_MAX_ANCHORS {_lib.LCR_MAX_ANCHORS} " in hdr
    assert f"#define LCR_MAX_TOPK {_lib.LCR_MAX_TOPK} " in hdr


def test_host_side_argument_errors(lib):
    """Errors are detected on the host before any launch (SURVEY §8b error convention)."""
    assert lib.lcr_anchors_f32(None, 4, 4, 4, None, 9, None) == -1
    base = (ctypes.c_float * 4)(0, 0, 1, 1)
    assert lib.lcr_anchors_f32(ctypes.c_void_p(256), 4, 4, 4, base, 99, None) == -2      # A > LCR_MAX_ANCHORS
    assert lib.lcr_clip_boxes_f32(None, 0, 1.0, 1.0, None) == 0                            # K == 0 is a no-op
    assert lib.lcr_clip_boxes_f32(ctypes.c_void_p(4), 3, 1.0, 1.0, None) == -4             # misaligned
    assert lib.lcr_nms_f32(ctypes.c_void_p(256), None, None, None, 1, 40000, 0.5, 0.0, 0, 10,
                           ctypes.c_void_p(256), ctypes.c_void_p(256), ctypes.c_void_p(256), 1 << 20, None) == -2
    assert lib.lcr_nms_f32(ctypes.c_void_p(256), None, None, None, 1, 100, 0.5, 0.0, 0, 10,
                           ctypes.c_void_p(256), ctypes.c_void_p(256), None, 0, None) == -3
    assert lib.lcr_nms_workspace_bytes(64, 2048) >= 64 * 2048 * 64 * 4
    assert lib.lcr_rpn_select_workspace_bytes(64, 1, 2000) >= 2 * 64 * 2000 * 8
    assert lib.lcr_paste_masks_u8(None, None, None, 0, 28, 520, 704, 0.5, 255, None, None) == 0
    assert lib.lcr_paste_masks_u8(None, None, None, 3, 28, 520, 704, 0.5, 255, None, None) == -1


def test_ops_fail_loudly_without_cuda():
    import torch
    from livecell_instance_segmentation_b200 import ops
    from livecell_instance_segmentation_b200._lib import LcrError
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(LcrError):
        ops.anchors(4, 4, 4, ops.base_anchors(), "cpu")
    with pytest.raises(LcrError):
        ops.paste_masks(torch.zeros(1, 28, 28), torch.zeros(1, 4), 8, 8)
    with pytest.raises(LcrError):
        ops.nms_batched(torch.zeros(1, 4, 4), torch.zeros(1, 4), 0.5, post_n=4)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing in the product package may import, load or link it."""
    pkg = os.path.join(ROOT, "livecell-instance-segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in ("import oracle", "from oracle", "liblcr_oracle", "orc_"):
                    assert pat not in src, (f, pat)


def test_torch_extension_builds_loads_and_rejects_cpu_tensors():
    """csrc/lcr_torch.so (C++ autograd node of RoIAlign + nms over the C-ABI): builds with plain g++ against the
    installed torch, links liblcr.so in-tree, and refuses CPU tensors like every other entry point."""
    import torch
    from livecell_instance_segmentation_b200 import _ext, build, roi_align as ra
    from livecell_instance_segmentation_b200._lib import LcrError
    ext = _ext.load()
    assert ext is not None and os.path.dirname(build.EXT_PATH) == os.path.dirname(build.LIB_PATH)
    assert ext.lcr_version() == 100
    for name in ("roi_align", "roi_align_list", "nms"):
        assert hasattr(ext, name)
    with pytest.raises(RuntimeError):
        ext.nms(torch.zeros(3, 4), torch.zeros(3), 0.5)
    with pytest.raises(RuntimeError):
        ext.roi_align(torch.zeros(1, 4, 8, 8), torch.zeros(2, 5), 0.25, 7, 7, 2, False)
    with pytest.raises(LcrError):
        ra.nms(torch.zeros(3, 4), torch.zeros(3), 0.5)
    with pytest.raises(LcrError):
        ra.roi_align(torch.zeros(1, 4, 8, 8), [torch.zeros(2, 4)], (7, 7), 0.25, 2)
