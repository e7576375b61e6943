"""GPU parity: anchors / clip / min-size / decode / level map."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


def test_anchors_bit_exact(ops, golden, oracle):
    from gpu_util import N, DEV
    g = golden("anchors")
    base = ops.base_anchors(tuple(g["sizes"]), tuple(g["ratios"]))
    assert np.array_equal(N(base), oracle.base_anchors(tuple(g["sizes"]), tuple(g["ratios"])))
    assert np.array_equal(N(ops.anchors(5, 7, 4, base, DEV)), g["small"])
    full = N(ops.anchors(130, 176, 4, base, DEV))
    assert np.array_equal(sha(full), g["full_sha256"])
    assert np.array_equal(full, oracle.anchors(130, 176, 4, N(base)))
    assert np.array_equal(sha(N(ops.anchors(56, 75, 4, base, DEV))), g["tile_sha256"])
    # other strides / anchor counts against the oracle
    b3 = np.array([[-8, -4, 8, 4], [-4, -8, 4, 8], [-6.5, -6.5, 6.5, 6.5]], np.float32)
    assert np.array_equal(N(ops.anchors(33, 44, 16, b3, DEV)), oracle.anchors(33, 44, 16, b3))


def test_anchor_generator_shim(golden):
    from gpu_util import N, DEV
    from livecell_instance_segmentation_b200.src.components.anchor_generator import AnchorGenerator
    g = AnchorGenerator()
    assert g.num_anchors_per_location == 9 and g.sizes == (32, 64, 128)
    a = g.generate_anchors((5, 7), 4, DEV)
    assert a.dtype.is_floating_point and tuple(a.shape) == (5 * 7 * 9, 4)
    assert np.array_equal(N(a), golden("anchors")["small"])


def test_clip_filter_decode(ops, golden, oracle):
    from gpu_util import N, T
    g = golden("box_utils")
    boxes = T(g["boxes"])
    ret = ops.clip_boxes_(boxes, 520, 704)
    assert ret.data_ptr() == boxes.data_ptr()          # in place, same tensor (box_utils.py:32-37)
    assert np.array_equal(N(boxes), g["clipped"])
    assert np.array_equal(N(ops.filter_small_boxes(boxes, 10)), g["keep10"])
    assert np.array_equal(N(ops.filter_small_boxes(boxes, 5)), g["keep5"])
    d1 = N(ops.box_decode(T(g["deltas"]), T(g["anchors"]), (1, 1, 1, 1)))
    d10 = N(ops.box_decode(T(g["deltas"]), T(g["anchors"]), (10, 10, 5, 5)))
    np.testing.assert_allclose(d1, g["decoded_w1"], rtol=1e-5, atol=1e-4)   # 1e-5 relative (north star)
    np.testing.assert_allclose(d10, g["decoded_w10"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(d1, oracle.box_decode(g["deltas"], g["anchors"]), rtol=1e-5, atol=1e-4)
    # round trip with the reference's encode_boxes
    rt = N(ops.box_decode(T(g["encoded"]), T(g["anchors"]), (1, 1, 1, 1)))
    np.testing.assert_allclose(rt, g["gt"], rtol=2e-5, atol=2e-4)
    dc = N(ops.box_decode(T(g["deltas"]), T(g["anchors"]), (1, 1, 1, 1), img_size=(520, 704)))
    np.testing.assert_allclose(dc, oracle.clip_boxes(d1, 520, 704), rtol=1e-5, atol=1e-4)


def test_box_utils_shims(golden):
    from gpu_util import N, T
    from livecell_instance_segmentation_b200.src.utils import box_utils as bu
    g = golden("box_utils")
    b = T(g["boxes"])
    out = bu.clip_boxes_to_image(b, (520, 704))
    assert out is b and np.array_equal(N(b), g["clipped"])
    assert np.array_equal(N(bu.filter_small_boxes(b, 10)), g["keep10"])
    np.testing.assert_allclose(N(bu.encode_boxes(T(g["gt"]), T(g["anchors"]))), g["encoded"], rtol=1e-6, atol=1e-6)
    assert bu.clip_boxes_to_image(T(np.zeros((0, 4), np.float32)), (5, 5)).shape == (0, 4)
    assert bu.filter_small_boxes(T(np.zeros((0, 4), np.float32)), 3).shape == (0,)


def test_level_map(ops, golden, oracle, synth):
    from gpu_util import N, T
    g = golden("roi_align")
    assert np.array_equal(N(ops.level_map(T(g["ms_boxes"]))), g["ms_levels"])
    lv = N(ops.level_map(T(g["lm_boxes"])))
    mism = int((lv != g["lm_levels"]).sum())
    # levels are exact except possibly at fp log2 boundaries ("report")
    assert mism == 0, f"{mism} level mismatches of {lv.size}"
    rois = synth.make_rois(500, 9, mode="fpn")
    assert np.array_equal(N(ops.level_map(T(rois))), oracle.level_map(rois))
