"""GPU parity: RoIAlign forward / backward: within 1e-5 relative (fp32),
against the golden outputs of the compiled torchvision CPU op, the oracle, and torchvision's CUDA op."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("roi_cpu_coords")]   # oracle / golden = torchvision's CPU-op coordinate rule

RTOL = 1e-5   # north star: "RoIAlign forward/backward ... within 1e-5 relative error in fp32"


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


CASES = [("p7", 7, 2, False), ("p14", 14, 2, False), ("p7a", 7, 2, True), ("p7ad", 7, 0, False)]


@pytest.mark.parametrize("tag,P,sr,al", CASES)
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_forward_golden(ops, golden, tag, P, sr, al, layout):
    from gpu_util import N, T, nhwc, assert_close_rel
    g = golden("roi_align")
    feat = T(g["feat"])
    if layout == "nhwc":
        feat = nhwc(feat)                      # TMA fast path for sr == 2, generic kernel otherwise
    out = ops.roi_align_fwd([feat], [0.25], T(g["rois"]), None, (P, P), sr, al)
    assert_close_rel(N(out), g[f"out_{tag}"], RTOL)


@pytest.mark.parametrize("tag,P,sr,al", CASES)
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_backward_golden(ops, golden, tag, P, sr, al, layout):
    from gpu_util import N, T, assert_close_rel
    g = golden("roi_align")
    shape = g["feat"].shape
    gin = torch.full(shape, 7.0, device="cuda:0")           # must be zero-filled by the callee
    if layout == "nhwc":
        gin = gin.contiguous(memory_format=torch.channels_last)
    ops.roi_align_bwd(T(g[f"gout_{tag}"]), [gin], [0.25], T(g["rois"]), None, sr, al, zero_grad=True)
    assert_close_rel(N(gin), g[f"gin_{tag}"], RTOL)


def test_module_and_autograd(golden):
    from gpu_util import N, T, assert_close_rel
    from livecell_instance_segmentation_b200.roi_align import RoIAlign
    g = golden("roi_align")
    op = RoIAlign(output_size=(7, 7), spatial_scale=1.0 / 4.0, sampling_ratio=2)      # custom_maskrcnn.py:48-50
    assert sum(p.numel() for p in op.parameters()) == 0 and len(list(op.buffers())) == 0
    # reference call form: feature_map[b:b+1], [proposals]
    out = op(T(g["feat"][:1]), [T(g["rois"][:, 1:])])
    assert_close_rel(N(out), g["out_listform"], RTOL)
    f = T(g["feat"]).requires_grad_(True)
    y = op(f, T(g["rois"]))
    y.backward(T(g["gout_p7"]))
    assert_close_rel(N(y), g["out_p7"], RTOL)
    assert_close_rel(N(f.grad), g["gin_p7"], RTOL)
    assert op(T(g["feat"][:1]), [T(np.zeros((0, 4), np.float32))]).shape == (0, 8, 7, 7)


@pytest.mark.parametrize("P", [7, 14])
def test_c1_shape_vs_oracle_and_torchvision(ops, oracle, synth, P):
    """Level-0 map of a 704x520 frame, 256 channels, anchor-shaped + edge-case RoIs."""
    import torchvision
    from gpu_util import N, T, nhwc, assert_close_rel
    C, H, W, K = 256, 130, 176, 96
    feat = synth.make_features(1, C, H, W, seed=5)
    rois = synth.make_rois(K, 6, edge_cases=True)
    ref = oracle.roi_align_fwd(feat, rois, P, P, 0.25, 2, False)
    ft = T(feat)
    fast = ops.roi_align_fwd([nhwc(ft)], [0.25], T(rois), None, (P, P), 2, False)
    gen = ops.roi_align_fwd([ft], [0.25], T(rois), None, (P, P), 2, False)
    assert_close_rel(N(fast), ref, RTOL)
    assert_close_rel(N(gen), ref, RTOL)
    # torchvision's CUDA op deviates from its CPU op by up to ~2e-5 relative: nvcc contracts its sample-coordinate
    # expression into an FMA, and one ulp of a coordinate near 130 (1.5e-5) moves a bilinear weight by as much.  With the
    # matching coordinate rule (the library default) the CUDA op is reproduced to fp32 rounding; the oracle can do both.
    tv = torchvision.ops.roi_align(ft, T(rois), (P, P), 0.25, 2, False)
    assert_close_rel(N(fast), N(tv), 1e-4)
    fast_cuda = ops.roi_align_fwd([nhwc(ft)], [0.25], T(rois), None, (P, P), 2, False, cpu_coords=False)
    gen_cuda = ops.roi_align_fwd([ft], [0.25], T(rois), None, (P, P), 2, False, cpu_coords=False)
    assert_close_rel(N(fast_cuda), N(tv), 2e-6)
    assert_close_rel(N(gen_cuda), N(tv), 2e-6)
    assert_close_rel(N(fast_cuda), oracle.roi_align_fwd(feat, rois, P, P, 0.25, 2, False, cuda_coords=True), RTOL)
    # backward: all-ones grad, sum(grad_in) == sum(grad_out) for in-range RoIs
    inr = synth.make_rois(64, 7)
    gout = torch.ones((64, C, P, P), device="cuda:0")
    gin = torch.empty((1, C, H, W), device="cuda:0").contiguous(memory_format=torch.channels_last)
    ops.roi_align_bwd(gout, [gin], [0.25], T(inr), None, 2, False)
    assert abs(float(gin.double().sum()) - 64 * C * P * P) < 1e-3 * 64 * C * P * P
    rng = np.random.RandomState(8)
    go = rng.standard_normal((K, C, P, P)).astype(np.float32)
    ref_b = oracle.roi_align_bwd(go, rois, (1, C, H, W), 0.25, 2, False)
    gin2 = torch.empty((1, C, H, W), device="cuda:0").contiguous(memory_format=torch.channels_last)
    ops.roi_align_bwd(T(go), [gin2], [0.25], T(rois), None, 2, False)
    assert_close_rel(N(gin2), ref_b, RTOL)
    gin3 = torch.empty((1, C, H, W), device="cuda:0")
    ops.roi_align_bwd(T(go), [gin3], [0.25], T(rois), None, 2, False)     # generic (NCHW) path
    assert_close_rel(N(gin3), ref_b, RTOL)


def test_padding_rois_and_batch(ops, oracle, synth):
    from gpu_util import N, T, nhwc, assert_close_rel
    feat = synth.make_features(3, 64, 20, 24, seed=12)
    rois = synth.make_rois(40, 13, img_h=80, img_w=96, batch=3)
    rois[5, 0] = -1.0
    rois[17, 0] = -1.0
    ref = oracle.roi_align_fwd(feat, rois)
    out = ops.roi_align_fwd([nhwc(T(feat))], [0.25], T(rois), None, (7, 7), 2, False)
    assert_close_rel(N(out), ref, RTOL)
    assert float(out[5].abs().max()) == 0.0 and float(out[17].abs().max()) == 0.0
    out2 = ops.roi_align_fwd([T(feat)], [0.25], T(rois), None, (7, 7), 2, False)
    assert_close_rel(N(out2), ref, RTOL)


def test_multiscale_golden(golden, synth):
    from gpu_util import N, T, nhwc, assert_close_rel
    from livecell_instance_segmentation_b200.roi_align import MultiScaleRoIAlign
    g = golden("roi_align")
    feats = [T(synth.make_features(1, 8, 32 >> i, 40 >> i, seed=40 + i)) for i in range(4)]
    ms = MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    out = ms({str(i): f for i, f in enumerate(feats)}, [T(g["ms_boxes"])], [(128, 160)])
    assert_close_rel(N(out), g["ms_out"], RTOL)
    out2 = ms({str(i): nhwc(f) for i, f in enumerate(feats)}, [T(g["ms_boxes"])], [(128, 160)])
    assert_close_rel(N(out2), g["ms_out"], RTOL)


@pytest.mark.parametrize("shape", [(2, 37, 19, 23), (3, 256, 33, 44), (2, 72, 10, 70), (1, 4, 2, 2), (2, 128, 64, 64)])
def test_layout_helpers(ops, synth, shape):
    """both transpose kernels: odd shapes (4-byte path) and C, H*W multiples of 4 (16-byte path, partial 64x64 tiles)"""
    from gpu_util import N, T
    n, c, h, w = shape
    x = synth.make_features(n, c, h, w, seed=3)
    y = ops.to_nhwc(T(x))
    assert y.stride() == (h * w * c, 1, w * c, c)
    assert np.array_equal(N(y), x)
    z = ops.to_nchw(y)
    assert z.is_contiguous() and np.array_equal(N(z), x)


@pytest.mark.parametrize("P", [7, 14])
def test_multilevel_fpn_rois_vs_oracle(ops, oracle, synth, P):
    """MultiScaleRoIAlign semantics (TV:ops/poolers.py:147-227) at the C5 geometry: 4 FPN levels of a 704x520 frame,
    64 channels, RoIs of 12-400 px assigned by the level mapper, box (7x7) and mask (14x14) pooling, forward and
    backward against the oracle — one launch for all levels."""
    from gpu_util import N, T, nhwc, assert_close_rel
    shapes, scales, C, K = [(130, 176), (65, 88), (33, 44), (17, 22)], [0.25, 0.125, 0.0625, 0.03125], 64, 300
    feats = [synth.make_features(1, C, h, w, seed=60 + i) for i, (h, w) in enumerate(shapes)]
    rois = synth.make_rois(K, 61, mode="fpn", edge_cases=True)
    lv_ref = oracle.level_map(rois[:, 1:])
    lvl = ops.level_map(T(rois), 2, 5, 224.0, 4)
    assert np.array_equal(N(lvl), lv_ref) and len(set(lv_ref.tolist())) >= 3          # the RoIs really spread over levels
    ref = oracle.multiscale_roi_align_fwd(feats, scales, rois, lv_ref, P, P, 2, False)
    out = ops.roi_align_fwd([nhwc(T(f)) for f in feats], scales, T(rois), lvl, (P, P), 2, False)
    assert_close_rel(N(out), ref, RTOL)
    out2 = ops.roi_align_fwd([T(f) for f in feats], scales, T(rois), lvl, (P, P), 2, False)     # NCHW: generic kernel
    assert_close_rel(N(out2), ref, RTOL)
    rng = np.random.RandomState(62)
    go = rng.standard_normal((K, C, P, P)).astype(np.float32)
    grads = [torch.empty((1, C, h, w), device="cuda:0").contiguous(memory_format=torch.channels_last) for h, w in shapes]
    ops.roi_align_bwd(T(go), grads, scales, T(rois), lvl, 2, False, zero_grad=True)
    for l, (h, w) in enumerate(shapes):
        sel = lv_ref == l
        ref_b = oracle.roi_align_bwd(go[sel], rois[sel], (1, C, h, w), scales[l], 2, False) if sel.any() else np.zeros((1, C, h, w), np.float32)
        assert_close_rel(N(grads[l]), ref_b, RTOL)


@pytest.mark.parametrize("C,K,mode,levels", [(256, 600, "anchor", 1), (256, 37, "anchor", 1), (256, 3000, "fpn", 4), (128, 300, "fpn", 1),
                                              (64, 129, "anchor", 1), (192, 64, "fpn", 2)])
def test_staged_forward_is_bit_identical_to_the_warp_kernel(ops, synth, oracle, tune, C, K, mode, levels):
    """roi_fwd_staged_kernel (window rows staged through shared memory by bulk copies, producer warp + pooling warps)
    performs the arithmetic of roi_fwd_warp_kernel operation for operation: same bits, for staged RoIs, for the wide RoIs
    its consumers gather themselves, for padding rows and for every edge case; and both stay within 1e-5 of the oracle."""
    from gpu_util import N, T, nhwc, assert_close_rel
    B, H, W = 2, 520, 704
    rois = synth.make_rois(K, 300 + K, img_h=H, img_w=W, mode=mode, batch=B, edge_cases=True)
    rois[K // 2, 0] = -1.0                                           # a padding row (batch index -1): zero tile
    feats, scales = [], []
    for l in range(levels):
        h, w = -(-H // (4 << l)), -(-W // (4 << l))
        feats.append(synth.make_features(B, C, h, w, seed=40 + l))
        scales.append(1.0 / (4 << l))
    lv = None
    if levels > 1:
        lv = oracle.level_map(rois[:, 1:], 2, 1 + levels, 224.0, 4).astype(np.int32)
    fd = [nhwc(T(f)) for f in feats]
    lvd = None if lv is None else T(lv)
    tune(LCR_ROI_FWD="warp")                                         # the sample-walk warp kernel
    ref = N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False))
    tune(LCR_ROI_FWD="staged")
    got = N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False))
    assert np.array_equal(got, ref)                                  # 8 pooling warps of 32 channels per CTA
    tune(LCR_ROI_STAGED_WARPS="4")                                   # 4 pooling warps of 64 channels
    assert np.array_equal(N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False)), ref)
    tune(LCR_ROI_STAGED_WARPS=None)
    tune(LCR_ROI_FWD="staged_direct")                                # same kernel, nothing staged
    assert np.array_equal(N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False)), ref)
    tune(LCR_ROI_RPC="1", LCR_ROI_FWD="staged")
    assert np.array_equal(N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False)), ref)
    tune(LCR_ROI_RPC="37")                                           # long CTAs: ring wrap-around, descriptor reuse
    assert np.array_equal(N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False)), ref)
    if K <= 600:
        live = rois[:, 0] >= 0
        if levels == 1:
            want = oracle.roi_align_fwd(feats[0], rois[live], 7, 7, scales[0], 2, False)
        else:
            want = oracle.multiscale_roi_align_fwd(feats, scales, rois[live], lv[live], 7, 7, 2, False)
        assert_close_rel(got[live], want, RTOL)
        assert not got[~live].any()


@pytest.mark.parametrize("path", ["warp", "staged", "cta", "generic_nchw"])
def test_c64_reference_fixture_elementwise(ops, golden, synth, oracle, tune, path):
    """tests/golden/roi_align_c64.npz: torchvision's compiled CPU op on a 64-channel map (make_golden.py:gen_roi_align) — the
    reference-run fixture that reaches the warp-item, staged and per-CTA fast paths (C % 64 == 0, channels_last) and the generic
    NCHW kernel; forward AND backward, checked ELEMENTWISE against the pooled magnitude (gpu_util.assert_close_elementwise),
    not only in max-norm."""
    from gpu_util import N, T, nhwc, assert_close_elementwise, assert_close_rel
    g = golden("roi_align_c64")
    s_feat, s_rois, s_g = [int(v) for v in g["seeds"]]
    feat = synth.make_features(1, 64, 40, 48, seed=s_feat)
    rois = g["rois"]
    assert np.array_equal(rois, synth.make_rois(16, s_rois, img_h=160, img_w=192, edge_cases=True))
    gout = np.random.RandomState(s_g).standard_normal((16, 64, 7, 7)).astype(np.float32)
    if path in ("staged", "cta"):
        tune(LCR_ROI_FWD=path, LCR_ROI_BWD=path)
    fd = T(feat) if path == "generic_nchw" else nhwc(T(feat))
    out = N(ops.roi_align_fwd([fd], [0.25], T(rois), None, (7, 7), 2, False))
    mag = oracle.roi_align_fwd(np.abs(feat), rois, 7, 7, 0.25, 2, False)
    worst, plain = assert_close_elementwise(out, g["out"], mag, RTOL, f"forward[{path}]")
    assert_close_rel(out, g["out"], RTOL)
    gin = torch.full((1, 64, 40, 48), 3.0, device="cuda:0")
    if path != "generic_nchw":
        gin = gin.contiguous(memory_format=torch.channels_last)
    ops.roi_align_bwd(T(gout), [gin], [0.25], T(rois), None, 2, False, zero_grad=True)
    gmag = oracle.roi_align_bwd(np.abs(gout), rois, feat.shape, 0.25, 2, False)
    bw, bplain = assert_close_elementwise(N(gin), g["gin"], gmag, RTOL, f"backward[{path}]")
    print(f"[roi_align c64 {path}] forward worst err/magnitude {worst:.2e} (plain relative {plain:.2e}); backward {bw:.2e} ({bplain:.2e})")


def test_abi_rejects_non_dense_zero_fill_and_ignores_bad_levels(ops, synth):
    """VERDICT r01 hygiene: lcr_roi_align_bwd_f32 with zero_grad memsets N*C*H*W floats from `data`, so a level that is
    not one dense run must be refused (LCR_ERR_INVALID_ARG) instead of zeroing its neighbours; a roi_level outside [0, L) or a
    batch index >= N makes the RoI padding (zero rows forward, ignored backward) instead of an out-of-range access."""
    import ctypes as C
    from gpu_util import N, T, nhwc
    from livecell_instance_segmentation_b200 import _lib
    lib = _lib.load()
    big = torch.full((1, 8, 12, 32), 5.0, device="cuda:0")
    view = big[:, :, :, ::2]                                         # every other column: not dense
    rois = T(synth.make_rois(6, 3, img_h=48, img_w=64))
    gout = torch.ones((6, 8, 7, 7), device="cuda:0")
    rc = lib.lcr_roi_align_bwd_f32(gout.data_ptr(), ops._feat_levels([view], [0.25]), 1, 8, rois.data_ptr(), None, 6, 7, 7, 2, 0, 1,
                                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == _lib.LCR_ERR_INVALID_ARG and bool((big == 5.0).all())
    feat = nhwc(T(synth.make_features(2, 64, 12, 16, seed=4)))
    r = synth.make_rois(8, 5, img_h=48, img_w=64, batch=2)
    lv = np.zeros(8, np.int32)
    lv[1], lv[2] = 7, -3                                             # levels that do not exist
    r[3, 0] = 2.0                                                    # batch index past the end
    out = N(ops.roi_align_fwd([feat], [0.25], T(r), T(lv), (7, 7), 2, False))
    assert not out[[1, 2, 3]].any() and out[[0, 4, 5, 6, 7]].any(axis=(1, 2, 3)).all()
    gin = torch.zeros_like(feat)
    ops.roi_align_bwd(T(np.ones_like(out)), [gin], [0.25], T(r), T(lv), 2, False, zero_grad=True)
    good = np.ones(8, bool)
    good[[1, 2, 3]] = False
    gref = torch.zeros_like(feat)
    ops.roi_align_bwd(T(np.ones_like(out[good])), [gref], [0.25], T(r[good]), T(lv[good]), 2, False, zero_grad=True)
    assert torch.allclose(gin, gref, rtol=0, atol=1e-5)


@pytest.mark.parametrize("variant", ["rm", "rm1", "rmp", "team", None])
@pytest.mark.parametrize("K,mode,levels", [(600, "anchor", 1), (41, "anchor", 1), (2500, "fpn", 4)])
def test_row_major_forward(ops, synth, oracle, tune, variant, K, mode, levels):
    """roi_fwd_rm_kernel (row program: distinct window rows with pre-added y weights into three rotating accumulator sets;
    rm = two rows' loads in flight per phase, rm1 = one): within 1e-5 of the oracle ELEMENTWISE (pooled-magnitude bound) and
    within fp32 rounding of the warp kernel; edge cases, padding rows, multi-level lists and RoIs that fall back to the
    sample walk (tiny / tall / wide) included."""
    from gpu_util import N, T, nhwc, assert_close_elementwise
    B, H, W, C = 2, 520, 704, 256
    rois = synth.make_rois(K, 700 + K, img_h=H, img_w=W, mode=mode, batch=B, edge_cases=True)
    rois[K // 3, 0] = -1.0
    rois[K // 2, 1:] = [300.0, 200.0, 306.0, 204.0]                  # 1.5 x 1 feature px: bins far below one pixel (fallback)
    rois[K // 2 + 1, 1:] = [10.0, 5.0, 60.0, 515.0]                  # 128 feature rows tall (fallback: > 32 rows)
    feats, scales = [], []
    for l in range(levels):
        h, w = -(-H // (4 << l)), -(-W // (4 << l))
        feats.append(synth.make_features(B, C, h, w, seed=80 + l))
        scales.append(1.0 / (4 << l))
    lv = None if levels == 1 else oracle.level_map(rois[:, 1:], 2, 1 + levels, 224.0, 4).astype(np.int32)
    fd = [nhwc(T(f)) for f in feats]
    lvd = None if lv is None else T(lv)
    tune(LCR_ROI_FWD="warp")
    ref = N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False))
    tune(LCR_ROI_FWD=variant)
    # into a NaN-filled buffer: a RoI the kernel skips must not pass on whatever the allocator left there (the persistent
    # team kernel once ended a team at its first exhausted claim and left RoIs claimed out of order unwritten)
    canvas = torch.full((K, C, 7, 7), float("nan"), device="cuda:0")
    got = N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False, out=canvas))
    assert np.isfinite(got).all()
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-6 * scale, np.abs(got - ref).max() / scale
    tune(LCR_ROI_IPW="1")
    canvas.fill_(float("nan"))
    again = N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False, out=canvas))
    if not np.array_equal(again, got):
        d = np.abs(again - got).reshape(K, 256, 49)
        bad = np.nonzero(d.reshape(K, -1).max(1))[0]
        k = bad[0]
        ch = np.nonzero(d[k].max(1))[0]
        tune(LCR_ROI_FWD="rm1")
        rm1 = N(ops.roi_align_fwd(fd, scales, T(rois), lvd, (7, 7), 2, False))
        info = [(int(kk), bool(np.array_equal(got[kk], rm1[kk])), bool(np.array_equal(again[kk], rm1[kk])), bool(np.array_equal(got[kk], ref[kk])),
                 bool(np.array_equal(again[kk], ref[kk]))) for kk in bad[:8]]
        raise AssertionError(f"rerun differs: (k, got==rm1, again==rm1, got==ref, again==ref) {info} max {d.max():.3e} rois {bad.tolist()[:8]} channels {ch[:8].tolist()} (n={len(ch)}) bins "
                             f"{np.nonzero(d[k].max(0))[0].tolist()} roi {rois[k]} again-ref {np.abs(again - ref).max():.3e} got-ref {np.abs(got - ref).max():.3e}")
    if K <= 600:
        live = rois[:, 0] >= 0
        want = oracle.roi_align_fwd(feats[0], rois[live], 7, 7, scales[0], 2, False)
        mag = oracle.roi_align_fwd(np.abs(feats[0]), rois[live], 7, 7, scales[0], 2, False)
        assert_close_elementwise(got[live], want, mag, RTOL, f"forward[{variant}]")
        assert not got[~live].any()


def test_default_forward_dispatch(ops, synth, tune):
    """Which forward kernel an unset LCR_ROI_FWD selects: the pipelined row-program kernel for the reference's pooler shape on
    a serving batch (7x7, 256 channels, K >= 8 x SMs, >= 4 frames) — bit-identical to LCR_ROI_FWD=rmp —, the sample-walk warp
    kernel otherwise (short lists, fewer frames) — bit-identical to LCR_ROI_FWD=warp."""
    from gpu_util import N, T, nhwc
    C, H, W = 256, 520, 704
    for B, K, want in ((4, 2400, "rmp"), (2, 2400, "warp"), (4, 600, "warp")):
        rois = synth.make_rois(K, 900 + K + B, img_h=H, img_w=W, mode="anchor", batch=B, edge_cases=True)
        feat = nhwc(T(synth.make_features(B, C, H // 4, W // 4, seed=90 + B)))
        out = {}
        for v in (None, "rmp", "warp"):
            tune(LCR_ROI_FWD=v)
            canvas = torch.full((K, C, 7, 7), float("nan"), device="cuda:0")
            out[v] = N(ops.roi_align_fwd([feat], [0.25], T(rois), None, (7, 7), 2, False, out=canvas))
            assert np.isfinite(out[v]).all()
        other = "warp" if want == "rmp" else "rmp"
        assert np.array_equal(out[None], out[want]), (B, K, want)
        assert not np.array_equal(out[None], out[other]), (B, K, "the two kernels round differently: the check above would be vacuous")
        assert np.abs(out["rmp"] - out["warp"]).max() <= 2e-6 * np.abs(out["warp"]).max()


@pytest.mark.parametrize("P", [7, 14])
def test_default_coordinate_rule_reproduces_torchvision_cuda(ops, oracle, synth, P):
    """The library default rounds the sample coordinates as torchvision's CUDA kernel does (what the reference runs on a
    GPU): forward and backward agree with torchvision's CUDA op to fp32 rounding on white-noise features at BASELINE C5 sizes
    (where its CPU and CUDA ops differ from each other by ~1.7e-5 of the range), single- and multi-level, aligned and not."""
    import torchvision
    from gpu_util import N, T, nhwc
    from livecell_instance_segmentation_b200 import ops as o
    prev = o.set_roi_coord_rule("cuda")
    try:
        C, H, W, K = 256, 130, 176, 2048
        g = torch.Generator(device="cuda:0").manual_seed(3)
        feat = torch.randn((1, C, H, W), generator=g, device="cuda:0")
        fn = nhwc(feat)
        rois = T(synth.make_rois(K, 77, mode="anchor", edge_cases=True))
        gout = torch.randn((K, C, P, P), generator=g, device="cuda:0")
        for aligned in (False, True):
            fr = feat.clone().requires_grad_(True)
            tv = torchvision.ops.roi_align(fr, rois, (P, P), 0.25, 2, aligned)
            tv.backward(gout)
            scale, gscale = float(tv.abs().max()), float(fr.grad.abs().max())
            for src in (fn, feat):                                  # fast NHWC kernels and the generic strided kernels
                out = o.roi_align_fwd([src], [0.25], rois, None, (P, P), 2, aligned)
                assert float((out - tv).abs().max()) <= 1e-6 * scale, (aligned, float((out - tv).abs().max()) / scale)
                gin = torch.empty_like(src)
                o.roi_align_bwd(gout, [gin], [0.25], rois, None, 2, aligned, zero_grad=True)
                assert float((gin - fr.grad).abs().max()) <= 4e-6 * gscale, (aligned, float((gin - fr.grad).abs().max()) / gscale)
            if not aligned:                                          # the other rule really is the other op
                cpu_rule = o.roi_align_fwd([fn], [0.25], rois, None, (P, P), 2, False, cpu_coords=True)
                assert float((cpu_rule - tv).abs().max()) > 3e-6 * scale
        # the drop-in module and the C++ autograd node use the default rule
        from livecell_instance_segmentation_b200.roi_align import RoIAlign
        fr = feat.clone().requires_grad_(True)
        y = RoIAlign((P, P), 0.25, 2)(fr, [rois[:200, 1:]])
        tv = torchvision.ops.roi_align(feat, [rois[:200, 1:]], (P, P), 0.25, 2, False)
        assert float((y - tv).abs().max()) <= 1e-6 * float(tv.abs().max())
    finally:
        o.set_roi_coord_rule(prev)


@pytest.mark.parametrize("P,sr,al", [((7, 7), 2, False), ((14, 14), 2, True), ((5, 3), 3, False), ((7, 7), 0, False), ((2, 9), 0, True)])
def test_planes_kernel_is_the_generic_kernel_bit_for_bit(ops, oracle, synth, tune, P, sr, al):
    """Maps the NHWC fast path does not take (NCHW as the reference's FPN emits them; any pooled size / sampling ratio) are
    pooled by roi_fwd_planes_kernel (per-RoI tap tables in shared memory, one CTA per RoI x channel chunk): same arithmetic
    as the one-thread-per-output generic kernel (LCR_ROI_FWD=generic) -> identical bits, on one level and on a pyramid, with
    short channel chunks (K small), dead / degenerate RoIs, and adaptive sampling grids wider than the table (in-place taps)."""
    from gpu_util import N, T, assert_close_rel
    C = 40
    feats = [synth.make_features(2, C, 60, 72, seed=11), synth.make_features(2, C, 30, 36, seed=12)]
    scales = [0.25, 0.125]
    for K in (9, 700):
        rois = synth.make_rois(K, 40 + K, img_h=240, img_w=288, batch=2, edge_cases=K >= 8)
        rois[-1, 1:] = [1.0, 2.0, 287.0, 239.0]                      # adaptive grid: ceil(59 / 7) = 9 samples per bin > table
        lvl = (np.arange(K) % 2).astype(np.int32)
        lvl[-1] = 0
        for fl, sc, lv in ((feats[:1], scales[:1], None), (feats, scales, T(lvl))):
            got = ops.roi_align_fwd([T(f) for f in fl], sc, T(rois), lv, P, sr, al)
            tune(LCR_ROI_FWD="generic")
            ref = ops.roi_align_fwd([T(f) for f in fl], sc, T(rois), lv, P, sr, al)
            tune(LCR_ROI_FWD=None)
            assert torch.equal(got, ref)
    # and against the oracle (one level, square pooling: the oracle's signature)
    if P[0] == P[1]:
        rois = synth.make_rois(64, 3, img_h=240, img_w=288, batch=2, edge_cases=True)
        got = ops.roi_align_fwd([T(feats[0])], [0.25], T(rois), None, P, sr, al)
        assert_close_rel(N(got), oracle.roi_align_fwd(feats[0], rois, P[0], P[1], 0.25, sr, al), RTOL)
