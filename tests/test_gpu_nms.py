"""GPU parity: greedy NMS keep lists, bit-exact vs golden torchvision outputs, the
oracle and torchvision's own CUDA op on the same device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


def gpu_nms(boxes, scores, thr, cpu_threshold=True):
    """cpu_threshold=True: the CPU op's rule (fp32 IoU vs the double threshold) — what the CPU-generated golden vectors
    and the oracle follow; False = the drop-in's default, torchvision's CUDA rule (threshold rounded to fp32)."""
    from gpu_util import N, T
    from livecell_instance_segmentation_b200.roi_align import nms, nms_cpu_rule
    out = (nms_cpu_rule if cpu_threshold else nms)(T(boxes), T(scores), thr)
    assert out.dtype == torch.int64
    return N(out)


def test_golden_keep_lists(golden):
    g = golden("nms")
    for n in (250, 2000):
        for thr in (0.4, 0.5, 0.7):
            assert np.array_equal(gpu_nms(g[f"boxes_{n}"], g[f"scores_{n}"], thr), g[f"keep_{n}_{int(thr * 10)}"]), (n, thr)


def test_known_answer_probes(golden):
    g = golden("nms")
    assert np.array_equal(gpu_nms(g["tie_boxes"], g["tie_scores"], 0.5), g["tie_keep"])     # ties: lower index first
    assert np.array_equal(gpu_nms(g["eq_boxes"], g["eq_scores"], 0.4), g["eq_keep_04"])     # IoU == 0.4f suppressed
    assert np.array_equal(gpu_nms(g["eq_boxes"], g["eq_scores"], 0.5), g["eq_keep_05"])
    assert np.array_equal(gpu_nms(g["zero_boxes"], g["zero_scores"], 0.4), g["zero_keep"])  # NaN IoU never suppresses
    assert np.array_equal(gpu_nms(g["nan_boxes"], g["nan_scores"], 0.4), g["nan_keep"])     # NaN score first
    assert gpu_nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)


def test_threshold_rule_matches_torchvision_cuda(golden):
    """for a pair whose fp32 IoU equals float32(thr) exactly, torchvision's CPU op suppresses
    (0.4f > 0.4 as doubles) and its CUDA op, which the reference runs on a GPU, does not (threshold rounded to fp32).
    The drop-in's default must follow the CUDA op on this device; cpu_threshold=True keeps the CPU rule."""
    import torchvision
    from gpu_util import N, T
    g = golden("nms")
    for name in ("eq", "tie", "zero", "nan"):
        b, s = g[f"{name}_boxes"], g[f"{name}_scores"]
        for thr in (0.4, 0.5):
            tv = N(torchvision.ops.nms(T(b), T(s), thr))
            assert np.array_equal(gpu_nms(b, s, thr, cpu_threshold=False), tv), (name, thr)
    # the two rules really differ on the exact-tie probe (otherwise this test pins nothing)
    cpu_rule = gpu_nms(g["eq_boxes"], g["eq_scores"], 0.4, cpu_threshold=True)
    cuda_rule = gpu_nms(g["eq_boxes"], g["eq_scores"], 0.4, cpu_threshold=False)
    assert np.array_equal(cpu_rule, g["eq_keep_04"]) and not np.array_equal(cpu_rule, cuda_rule)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 64, 65, 257, 1000, 2047, 2048, 2049, 5000])
def test_sizes_vs_oracle_and_torchvision(oracle, synth, n):
    import torchvision
    from gpu_util import N, T
    rng = np.random.RandomState(n)
    boxes = synth.make_rois(n, 200 + n)[:, 1:]
    scores = rng.permutation(n).astype(np.float32) / n
    for thr in (0.3, 0.5):
        mine = gpu_nms(boxes, scores, thr)
        assert np.array_equal(mine, oracle.nms(boxes, scores, thr)), (n, thr)
        tv = N(torchvision.ops.nms(T(boxes), T(scores), thr))    # the reference's CUDA path on this device
        assert np.array_equal(mine, tv), (n, thr)
    if n > 40:                                                   # a block of tied scores: stable order (oracle rule)
        tied = scores.copy()
        tied[rng.choice(n, n // 8, replace=False)] = 0.5
        assert np.array_equal(gpu_nms(boxes, tied, 0.5), oracle.nms(boxes, tied, 0.5)), n


def test_crowded_duplicates(oracle):
    """Many identical / nested boxes: long suppression chains across chunk boundaries."""
    rng = np.random.RandomState(11)
    base = np.array([100, 100, 160, 150], np.float32)
    boxes = base[None, :] + rng.randint(-3, 4, size=(1500, 4)).astype(np.float32)
    boxes[::7] += 300
    scores = rng.permutation(1500).astype(np.float32)
    for thr in (0.4, 0.7, 0.9):
        assert np.array_equal(gpu_nms(boxes, scores, thr), oracle.nms(boxes, scores, thr))


@pytest.mark.parametrize("n", [64, 700, 2048])
def test_suppression_chain_depth(oracle, n):
    """A staircase where every box overlaps only its successor above the threshold and scores fall along it: the keep
    decision of box i depends on box i-1 (kept, dropped, kept, ...), the deepest dependency chain there is — the
    parallel resolve needs its maximum number of rounds (one per 32-box block)."""
    step = 10.0
    x = np.arange(n, dtype=np.float32) * step
    boxes = np.stack([x, np.zeros(n, np.float32), x + 40.0, np.full(n, 40.0, np.float32)], axis=1)   # IoU(i, i+1) = 0.6, (i, i+2) = 0.33
    scores = np.arange(n, 0, -1).astype(np.float32)
    for thr in (0.5, 0.3):
        ref = oracle.nms(boxes, scores, thr)
        assert np.array_equal(gpu_nms(boxes, scores, thr), ref), (n, thr)
    rng = np.random.RandomState(n)   # the same staircase in shuffled input order
    perm = rng.permutation(n)
    assert np.array_equal(gpu_nms(boxes[perm], scores[perm], 0.5), oracle.nms(boxes[perm], scores[perm], 0.5))


def test_batched_segments_counts_presorted_postn(ops, oracle, synth):
    from gpu_util import N, T
    S, stride = 5, 300
    rng = np.random.RandomState(2)
    boxes = np.stack([synth.make_rois(stride, 300 + s)[:, 1:] for s in range(S)])
    scores = np.stack([np.sort(rng.rand(stride).astype(np.float32))[::-1] for _ in range(S)])
    counts = np.array([300, 0, 1, 33, 257], np.int32)
    # presorted (scores=None) with device-side counts and truncation
    keep, kc = ops.nms_batched(T(boxes), None, 0.4, post_n=40, counts=T(counts))
    keep, kc = N(keep), N(kc)
    for s in range(S):
        ref = oracle.nms(boxes[s, : counts[s]], None, 0.4, post_n=40)
        assert kc[s] == len(ref) and np.array_equal(keep[s, : kc[s]], ref), s
    # unsorted scores + score threshold (the box_scores > 0.4 filter of custom_maskrcnn.py:185)
    sc2 = rng.rand(S, stride).astype(np.float32)
    keep, kc = ops.nms_batched(T(boxes), T(sc2), 0.5, post_n=stride, counts=T(counts), score_thresh=0.4)
    keep, kc = N(keep), N(kc)
    for s in range(S):
        ref = oracle.nms(boxes[s, : counts[s]], sc2[s, : counts[s]], 0.5, score_thresh=0.4, use_score_thresh=True)
        assert kc[s] == len(ref) and np.array_equal(keep[s, : kc[s]], ref), s
    # equivalent reference formulation: filter first, then nms, then map back
    s = 0
    m = sc2[s] > 0.4
    idx = np.nonzero(m)[0]
    ref = idx[oracle.nms(boxes[s][m], sc2[s][m], 0.5)]
    assert np.array_equal(keep[s, : kc[s]], ref)


def test_batched_nms_categories(oracle, synth):
    import torchvision
    from gpu_util import N, T
    from livecell_instance_segmentation_b200.roi_align import batched_nms
    n = 800
    rng = np.random.RandomState(4)
    boxes = synth.make_rois(n, 77)[:, 1:]
    scores = rng.permutation(n).astype(np.float32) / n
    cat = rng.randint(0, 4, size=n).astype(np.int64)
    mine = N(batched_nms(T(boxes), T(scores), T(cat), 0.5))
    assert np.array_equal(mine, oracle.nms(boxes, scores, 0.5, category=cat))
    tv = N(torchvision.ops.boxes._batched_nms_vanilla(T(boxes), T(scores), T(cat), 0.5))
    assert np.array_equal(mine, tv)
