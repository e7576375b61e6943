// roi_align.cu — RoIAlign forward / backward (a9, a10) and multi-level pooling (a11), one launch.
//
// Replaces torchvision::roi_align / _roi_align_backward as called by
// RoIAlign((7,7), 0.25, sampling_ratio=2) at src/custom_maskrcnn.py:48-50,:120,:177 and, with
// several levels + roi_level, MultiScaleRoIAlign (TV:ops/poolers.py:147-227).
// Arithmetic spec: SURVEY.md App. B.1 / B.2 (restated in oracle/lcr_oracle.c).
//
// torchvision's kernel is one thread per output element doing 4 samples x 4 scattered 4-byte NCHW
// loads (16 L1 transactions per 4 output bytes).  The B200 design here:
//
//   * NHWC feature maps (channels_last memory): a pixel is one contiguous C*4-byte vector, so a
//     warp's load is a full coalesced line and lanes map to channels.
//   * Separable bilinear pooling.  With sampling_ratio 2 the 2P x 2P sample grid of a RoI is a
//     tensor product, so out = Wy * F_window * Wx^T.  A thread owns VEC channels, walks the
//     DISTINCT rows of the RoI window once (row cache of two T-vectors), and inside a row loads
//     every DISTINCT column once (the a[]/b[] slots below): feature bytes read per RoI are the
//     unique window, not 16 taps per output.  fp32 FMAs are issued as packed FFMA2 (sm_100).
//   * Output staging: the [channels][P*P] tile of a RoI (50 176 B) is assembled in shared memory
//     and leaves with ONE TMA bulk store (cp.async.bulk.global.shared::cta, SASS UBLKCP) — the
//     dominant traffic (the output) is written as full lines by the copy engine.
//   * Backward is the transpose: the grad_out tile arrives with one TMA bulk load
//     (cp.async.bulk.shared::cluster.global + mbarrier), a thread accumulates per distinct window
//     pixel and issues one vector red.global.add per (pixel, channel pair): lanes hit consecutive
//     channels, so a warp's atomics coalesce into full 256-byte lines at L2.
//
// Roofline: HBM.  Algorithmic bytes fwd = 4*K*C*P*P (out) + unique feature bytes + 20*K.
// Generic kernels (any strides / sampling ratio / pooled size) back every other configuration.
#include <stdlib.h>

#include <atomic>
#include <string.h>

#include "common.cuh"

namespace lcr {

struct LvParam {
  float* data;
  int N, H, W;
  long long sn, sc, sh, sw;
  float scale;
};

struct RoiParams {
  LvParam lv[LCR_MAX_LEVELS];
  const float* rois;
  const int* roi_level;
  int L, C, K, PH, PW, sr, aligned;
  int cpu_coords;  // 1: sample coordinates rounded like torchvision's CPU op (every operation rounded); 0: like its CUDA op
                   // (nvcc contracts start + bin_index * bin_size, and roi * scale - 0.5, into FMAs)
};

// ------------------------------------------------------------------------------------------------
// RoI geometry, operation-for-operation the oracle's roi_geometry()/make_tap() (no contraction).
// ------------------------------------------------------------------------------------------------
struct RoiGeom {
  float sw, sh, bw, bh;
  int gw, gh;
  int b, lvl;
  bool live;
  int cpu_coords;
};

__device__ __forceinline__ RoiGeom roi_geom(const RoiParams& p, int k) {
  RoiGeom g;
  const float* r = p.rois + (size_t)k * 5;
  const float bf = __ldg(r);
  g.lvl = p.roi_level ? __ldg(p.roi_level + k) : 0;
  g.b = (int)bf;
  g.live = (bf >= 0.f) && g.lvl >= 0 && g.lvl < p.L && g.b < p.lv[g.lvl < 0 || g.lvl >= p.L ? 0 : g.lvl].N;
  const float scale = p.lv[g.live ? g.lvl : 0].scale;
  const float off = p.aligned ? 0.5f : 0.0f;
  float ew, eh;
  if (p.cpu_coords) {
    g.sw = __fsub_rn(__fmul_rn(__ldg(r + 1), scale), off);
    g.sh = __fsub_rn(__fmul_rn(__ldg(r + 2), scale), off);
    ew = __fsub_rn(__fmul_rn(__ldg(r + 3), scale), off);
    eh = __fsub_rn(__fmul_rn(__ldg(r + 4), scale), off);
  } else {  // roi * spatial_scale - offset as one FMA (identical when offset == 0)
    g.sw = __fmaf_rn(__ldg(r + 1), scale, -off);
    g.sh = __fmaf_rn(__ldg(r + 2), scale, -off);
    ew = __fmaf_rn(__ldg(r + 3), scale, -off);
    eh = __fmaf_rn(__ldg(r + 4), scale, -off);
  }
  float rw = __fsub_rn(ew, g.sw), rh = __fsub_rn(eh, g.sh);
  if (!p.aligned) {
    rw = fmaxf(rw, 1.0f);
    rh = fmaxf(rh, 1.0f);
  }
  g.bh = __fdiv_rn(rh, (float)p.PH);
  g.bw = __fdiv_rn(rw, (float)p.PW);
  g.gh = p.sr > 0 ? p.sr : (int)ceilf(__fdiv_rn(rh, (float)p.PH));
  g.gw = p.sr > 0 ? p.sr : (int)ceilf(__fdiv_rn(rw, (float)p.PW));
  g.cpu_coords = p.cpu_coords;
  return g;
}

// position of sample i of bin pbin: start + pbin*bin + (i+0.5)*bin/grid
// cpu_coords: every operation rounded (torchvision's CPU op); otherwise start + pbin * bin is one FMA, as nvcc compiles
// `roi_start + ph * bin_size + (iy + .5f) * bin_size / grid` in torchvision's CUDA kernel.  The two differ by an ulp of the
// coordinate (~4e-6 feature px at x ~ 100), i.e. by up to ~2e-5 of the output range on white-noise features.
__device__ __forceinline__ float sample_pos(float start, int pbin, float bin, int i, int grid, int cpu_coords) {
  const float head = cpu_coords ? __fadd_rn(start, __fmul_rn((float)pbin, bin)) : __fmaf_rn((float)pbin, bin, start);
  return __fadd_rn(head, __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
}

// 1-D half of make_tap(): neighbours lo/hi and weights (w_lo = 1 - frac, w_hi = frac).
__device__ __forceinline__ bool axis_tap(float pos, int size, int& lo, int& hi, float& w_lo, float& w_hi) {
  if (pos < -1.0f || pos > (float)size) {
    lo = hi = 0;
    w_lo = w_hi = 0.f;
    return false;
  }
  if (pos <= 0.f) pos = 0.f;
  lo = (int)pos;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    pos = (float)lo;
  } else {
    hi = lo + 1;
  }
  w_hi = __fsub_rn(pos, (float)lo);
  w_lo = __fsub_rn(1.0f, w_hi);
  return true;
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: one thread per output element, arbitrary strides / pooled size / sampling ratio.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) roi_fwd_generic_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out) {
  const size_t total = (size_t)p.K * p.C * p.PH * p.PW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % p.PW);
    const int ph = (int)((idx / p.PW) % p.PH);
    const int c = (int)((idx / ((size_t)p.PW * p.PH)) % p.C);
    const int k = (int)(idx / ((size_t)p.PW * p.PH * p.C));
    const RoiGeom g = roi_geom(p, k);
    float acc = 0.f;
    if (g.live) {
      const LvParam& lv = p.lv[g.lvl];
      const float* f = lv.data + (size_t)g.b * lv.sn + (size_t)c * lv.sc;
      for (int iy = 0; iy < g.gh; ++iy) {
        int yl, yh;
        float wyl, wyh;
        if (!axis_tap(sample_pos(g.sh, ph, g.bh, iy, g.gh, p.cpu_coords), lv.H, yl, yh, wyl, wyh)) continue;
        for (int ix = 0; ix < g.gw; ++ix) {
          int xl, xh;
          float wxl, wxh;
          if (!axis_tap(sample_pos(g.sw, pw, g.bw, ix, g.gw, p.cpu_coords), lv.W, xl, xh, wxl, wxh)) continue;
          float v = __fmul_rn(__fmul_rn(wyl, wxl), __ldg(f + yl * lv.sh + xl * lv.sw));
          v = __fadd_rn(v, __fmul_rn(__fmul_rn(wyl, wxh), __ldg(f + yl * lv.sh + xh * lv.sw)));
          v = __fadd_rn(v, __fmul_rn(__fmul_rn(wyh, wxl), __ldg(f + yh * lv.sh + xl * lv.sw)));
          v = __fadd_rn(v, __fmul_rn(__fmul_rn(wyh, wxh), __ldg(f + yh * lv.sh + xh * lv.sw)));
          acc = __fadd_rn(acc, v);
        }
      }
      const int cnt = g.gh * g.gw;
      acc = __fdiv_rn(acc, (float)(cnt > 1 ? cnt : 1));
    }
    out[idx] = acc;
  }
}

__global__ void __launch_bounds__(256) roi_bwd_generic_kernel(const __grid_constant__ RoiParams p, const float* __restrict__ gout) {
  const size_t total = (size_t)p.K * p.C * p.PH * p.PW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % p.PW);
    const int ph = (int)((idx / p.PW) % p.PH);
    const int c = (int)((idx / ((size_t)p.PW * p.PH)) % p.C);
    const int k = (int)(idx / ((size_t)p.PW * p.PH * p.C));
    const RoiGeom g = roi_geom(p, k);
    if (!g.live) continue;
    const LvParam& lv = p.lv[g.lvl];
    float* f = lv.data + (size_t)g.b * lv.sn + (size_t)c * lv.sc;
    const float go = __ldg(gout + idx);
    const int cnt = g.gh * g.gw;
    const float count = (float)(cnt > 1 ? cnt : 1);
    for (int iy = 0; iy < g.gh; ++iy) {
      int yl, yh;
      float wyl, wyh;
      if (!axis_tap(sample_pos(g.sh, ph, g.bh, iy, g.gh, p.cpu_coords), lv.H, yl, yh, wyl, wyh)) continue;
      for (int ix = 0; ix < g.gw; ++ix) {
        int xl, xh;
        float wxl, wxh;
        if (!axis_tap(sample_pos(g.sw, pw, g.bw, ix, g.gw, p.cpu_coords), lv.W, xl, xh, wxl, wxh)) continue;
        atomicAdd(f + yl * lv.sh + xl * lv.sw, __fdiv_rn(__fmul_rn(go, __fmul_rn(wyl, wxl)), count));
        atomicAdd(f + yl * lv.sh + xh * lv.sw, __fdiv_rn(__fmul_rn(go, __fmul_rn(wyl, wxh)), count));
        atomicAdd(f + yh * lv.sh + xl * lv.sw, __fdiv_rn(__fmul_rn(go, __fmul_rn(wyh, wxl)), count));
        atomicAdd(f + yh * lv.sh + xh * lv.sw, __fdiv_rn(__fmul_rn(go, __fmul_rn(wyh, wxh)), count));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Forward for maps the NHWC fast path does not take (NCHW as the reference's FPN emits them, any strides / pooled size /
// sampling ratio): one CTA per (RoI, channel chunk) with the RoI's axis taps in shared memory.
//
// roi_fwd_generic_kernel recomputes the RoI geometry (5 loads, 4 divisions) and every sample's position and taps
// (2 divisions + the clamping of axis_tap per sample and axis) in EVERY output thread: ~16 divisions per output element
// next to its 16 loads.  Here the PH*gh + PW*gw taps of a RoI are computed once per CTA, by as many threads, and an
// output element is 4 table reads per sample next to its loads.  Thread -> (channel, ph, pw) with pw fastest: the output
// leaves as full lines, and on an NCHW map the lanes of a warp read neighbouring columns of the same few plane rows.
// The arithmetic is the generic kernel's, operation for operation: results are bit-identical to it (and to the oracle's
// restatement of the CPU op with LCR_ROI_CPU_COORDS).  RoIs with more than kPlaneTaps samples on an axis (adaptive
// sampling_ratio on a huge RoI) compute their taps in place, like the generic kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kPlaneTaps = 64;
struct __align__(16) PlaneTap {
  int lo, hi;  // lo < 0: the sample lies outside the map and contributes nothing
  float w_lo, w_hi;
};

__device__ __forceinline__ PlaneTap plane_tap(float start, float bin, int grid, int size, int u, int cpu_coords) {
  const int b = u / grid;
  int lo, hi;
  float wl, wh;
  const bool ok = axis_tap(sample_pos(start, b, bin, u - b * grid, grid, cpu_coords), size, lo, hi, wl, wh);
  return PlaneTap{ok ? lo : -1, hi, wl, wh};
}

// SR = 2: the reference's sampling ratio with the sample loops unrolled (4 samples, 16 loads in flight per output); SR = 0: any.
template <int SR>
__global__ void __launch_bounds__(256) roi_fwd_planes_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int cchunk) {
  __shared__ PlaneTap s_x[kPlaneTaps], s_y[kPlaneTaps];
  const int chunks = (p.C + cchunk - 1) / cchunk;
  const int k = (int)(blockIdx.x / (unsigned)chunks);
  const int c0 = (int)(blockIdx.x - (unsigned)k * (unsigned)chunks) * cchunk;
  const int nc = min(cchunk, p.C - c0);
  const int PP = p.PH * p.PW;
  const int total = nc * PP;
  float* const o = out + ((size_t)k * p.C + c0) * PP;
  const RoiGeom g = roi_geom(p, k);  // block-uniform
  const int gh = SR ? SR : g.gh, gw = SR ? SR : g.gw;
  if (!g.live || gh <= 0 || gw <= 0) {
    for (int i = threadIdx.x; i < total; i += blockDim.x) o[i] = 0.f;
    return;
  }
  const LvParam& lv = p.lv[g.lvl];
  const bool tabled = (long long)p.PH * gh <= kPlaneTaps && (long long)p.PW * gw <= kPlaneTaps;
  if (tabled) {
    const int ny = p.PH * gh, nx = p.PW * gw;
    for (int t = threadIdx.x; t < ny + nx; t += blockDim.x) {
      if (t < ny) s_y[t] = plane_tap(g.sh, g.bh, gh, lv.H, t, p.cpu_coords);
      else s_x[t - ny] = plane_tap(g.sw, g.bw, gw, lv.W, t - ny, p.cpu_coords);
    }
    __syncthreads();
  }
  const int cnt = gh * gw;
  const float count = (float)(cnt > 1 ? cnt : 1);
  const float* const fb = lv.data + (size_t)g.b * lv.sn + (size_t)c0 * lv.sc;
  // One output element: sum over the bin's gh x gw samples, in the order of torchvision's per-output kernel.
  auto pool = [&](const float* __restrict__ f, int ph, int pw, auto&& ytap, auto&& xtap) -> float {
    float acc = 0.f;
#pragma unroll
    for (int iy = 0; iy < gh; ++iy) {
      const PlaneTap ty = ytap(ph * gh + iy, iy);
      if (ty.lo < 0) continue;
      const float* const r0 = f + ty.lo * lv.sh;
      const float* const r1 = f + ty.hi * lv.sh;
#pragma unroll
      for (int ix = 0; ix < gw; ++ix) {
        const PlaneTap tx = xtap(pw * gw + ix, ix);
        if (tx.lo < 0) continue;
        float v = __fmul_rn(__fmul_rn(ty.w_lo, tx.w_lo), __ldg(r0 + tx.lo * lv.sw));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(ty.w_lo, tx.w_hi), __ldg(r0 + tx.hi * lv.sw)));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(ty.w_hi, tx.w_lo), __ldg(r1 + tx.lo * lv.sw)));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(ty.w_hi, tx.w_hi), __ldg(r1 + tx.hi * lv.sw)));
        acc = __fadd_rn(acc, v);
      }
    }
    return __fdiv_rn(acc, count);
  };
  if (SR > 0 && tabled && blockDim.x % PP == 0) {
    // The launch made the block a multiple of PH*PW threads: a thread keeps ONE bin — its 2 + 2 taps live in registers —
    // and walks the chunk's channels; the inner loop is 16 loads and their FMAs, no table reads and no index arithmetic
    // (ncu of the table-reading loop, K = 1024: 15.9 M of its LSU wavefronts were the four LDS.128 per output).
    const int bin = (int)threadIdx.x % PP, cstep = (int)blockDim.x / PP;
    const int ph = bin / p.PW, pw = bin - ph * p.PW;
    PlaneTap ty[SR > 0 ? SR : 1], tx[SR > 0 ? SR : 1];
#pragma unroll
    for (int j = 0; j < SR; ++j) {
      ty[j] = s_y[ph * SR + j];
      tx[j] = s_x[pw * SR + j];
    }
    for (int c = (int)threadIdx.x / PP; c < nc; c += cstep)
      o[c * PP + bin] = pool(fb + (size_t)c * lv.sc, ph, pw, [&](int, int j) { return ty[j]; }, [&](int, int j) { return tx[j]; });
    return;
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = i / PP;
    const int r = i - c * PP;
    const int ph = r / p.PW, pw = r - ph * p.PW;
    o[i] = pool(fb + (size_t)c * lv.sc, ph, pw,
                [&](int u, int) { return tabled ? s_y[u] : plane_tap(g.sh, g.bh, gh, lv.H, u, p.cpu_coords); },
                [&](int u, int) { return tabled ? s_x[u] : plane_tap(g.sw, g.bw, gw, lv.W, u, p.cpu_coords); });
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path (NHWC, sampling_ratio == 2, square P x P).
// ------------------------------------------------------------------------------------------------
enum : uint32_t { kSame = 0u, kShift = 1u, kNew = 2u, kModeMask = 3u, kBorder = 4u, kValid = 8u };

struct __align__(16) AxisSample {
  int off_lo, off_hi;   // element offsets (already multiplied by the axis stride)
  float w_lo, w_hi;
};

template <int P>
struct RoiTables {
  AxisSample xs[2 * P];
  AxisSample ys[2 * P];
  uint32_t xmode[2 * P];
  uint32_t ymode[2 * P];
  int lo[2][2 * P];
  int hi[2][2 * P];
};

// Threads t < 2P build both axis tables of the current RoI; thread 0 / 1 then derive the
// SAME / SHIFT / NEW reuse flags (a sequential scan over <= 2P samples).
template <int P>
__device__ __forceinline__ void build_tables(RoiTables<P>& tb, const RoiGeom& g, const LvParam& lv, int tid) {
  if (tid < 2 * P) {
    int lo, hi;
    float wl, wh;
    bool ok = axis_tap(sample_pos(g.sw, tid >> 1, g.bw, tid & 1, 2, g.cpu_coords), lv.W, lo, hi, wl, wh);
    tb.xs[tid] = AxisSample{(int)(lo * lv.sw), (int)(hi * lv.sw), wl, wh};
    tb.lo[0][tid] = ok ? lo : -1;
    tb.hi[0][tid] = hi;
    ok = axis_tap(sample_pos(g.sh, tid >> 1, g.bh, tid & 1, 2, g.cpu_coords), lv.H, lo, hi, wl, wh);
    tb.ys[tid] = AxisSample{(int)(lo * lv.sh), (int)(hi * lv.sh), wl, wh};
    tb.lo[1][tid] = ok ? lo : -1;
    tb.hi[1][tid] = hi;
  }
  __syncthreads();
  if (tid < 2) {
    uint32_t* mode = tid == 0 ? tb.xmode : tb.ymode;
    int c0 = -1, c1 = -1;
    for (int t = 0; t < 2 * P; ++t) {
      const int lo = tb.lo[tid][t], hi = tb.hi[tid][t];
      if (lo < 0) {
        mode[t] = 0u;  // invalid: contributes nothing, cache state unchanged
        continue;
      }
      uint32_t m = (lo == c0) ? kSame : ((lo == c1) ? kShift : kNew);
      if (hi == lo) m |= kBorder;
      mode[t] = m | kValid;
      c0 = lo;
      c1 = hi;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// T[pw] = sum over the two x-samples of bin pw of (w_lo * F[row][x_lo] + w_hi * F[row][x_hi]),
// every distinct column of the row loaded once.  `row` points at this thread's channel pair.
template <int P>
__device__ __forceinline__ void pool_row(const RoiTables<P>& tb, const float* __restrict__ row, float2 (&T)[P]) {
  float2 a[2 * P], b[2 * P];
  // phase 1: all loads of the row in flight together (predicated off for reused columns)
#pragma unroll
  for (int t = 0; t < 2 * P; ++t) {
    const uint32_t m = tb.xmode[t];
    const AxisSample s = tb.xs[t];
    a[t] = make_float2(0.f, 0.f);
    b[t] = make_float2(0.f, 0.f);
    if ((m & kValid) && (m & kModeMask) == kNew) a[t] = __ldg(reinterpret_cast<const float2*>(row + s.off_lo));
    if ((m & kValid) && (m & kModeMask) != kSame && !(m & kBorder)) b[t] = __ldg(reinterpret_cast<const float2*>(row + s.off_hi));
  }
  // phase 2: resolve reuse and accumulate
  float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
#pragma unroll
  for (int pw = 0; pw < P; ++pw) T[pw] = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 2 * P; ++t) {
    const uint32_t m = tb.xmode[t];
    if (m & kValid) {  // warp-uniform
      const uint32_t mode = m & kModeMask;
      const AxisSample s = tb.xs[t];
      if (mode == kNew) pa = a[t];
      else if (mode == kShift) pa = pb;
      if (mode != kSame) pb = (m & kBorder) ? pa : b[t];
      T[t >> 1] = ffma2(splat(s.w_lo), pa, T[t >> 1]);
      T[t >> 1] = ffma2(splat(s.w_hi), pb, T[t >> 1]);
    }
  }
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_smem_to_global(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Forward.  NT threads, each VEC=2 channels: a work item is (roi, channel group of 2*NT channels).
template <int P, int NT>
__global__ void __launch_bounds__(NT, 4) roi_fwd_nhwc_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out,
                                                          int groups) {
  constexpr int CPB = 2 * NT;
  constexpr int PP = P * P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);                                  // [CPB][PP]
  RoiTables<P>& tb = *reinterpret_cast<RoiTables<P>*>(smem_raw + sizeof(float) * CPB * PP);
  const int tid = threadIdx.x;
  const long long items = (long long)p.K * groups;

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int k = (int)(item / groups);
    const int cg = (int)(item - (long long)k * groups);
    const int c0 = cg * CPB;
    const int nch = min(CPB, p.C - c0);          // channels of this group (multiple of 4)
    const RoiGeom g = roi_geom(p, k);
    const LvParam& lv = p.lv[g.live ? g.lvl : 0];
    if (g.live) build_tables<P>(tb, g, lv, tid);  // g.live is block-uniform

    // the previous item's bulk store must have finished READING the tile before it is rewritten
    if (tid == 0) bulk_wait_read_all();
    __syncthreads();

    const int c = c0 + 2 * tid;
    if (2 * tid < nch) {
      float* my = tile + (size_t)(2 * tid) * PP;
      if (!g.live) {
#pragma unroll 7
        for (int j = 0; j < PP; ++j) {
          my[j] = 0.f;
          my[PP + j] = 0.f;
        }
      } else {
        const float* fb = lv.data + (size_t)g.b * lv.sn + c;  // sc == 1
        float2 T0[P], T1[P];
#pragma unroll
        for (int pw = 0; pw < P; ++pw) T0[pw] = T1[pw] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int ph = 0; ph < P; ++ph) {
          float2 acc[P];
#pragma unroll
          for (int pw = 0; pw < P; ++pw) acc[pw] = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t m = tb.ymode[2 * ph + i];
            if (m & kValid) {  // block-uniform
              const uint32_t mode = m & kModeMask;
              const AxisSample s = tb.ys[2 * ph + i];
              if (mode == kShift) {
#pragma unroll
                for (int pw = 0; pw < P; ++pw) T0[pw] = T1[pw];
              } else if (mode == kNew) {
                pool_row<P>(tb, fb + s.off_lo, T0);
              }
              if (mode != kSame) {
                if (m & kBorder) {
#pragma unroll
                  for (int pw = 0; pw < P; ++pw) T1[pw] = T0[pw];
                } else {
                  pool_row<P>(tb, fb + s.off_hi, T1);
                }
              }
#pragma unroll
              for (int pw = 0; pw < P; ++pw) {
                acc[pw] = ffma2(splat(s.w_lo), T0[pw], acc[pw]);
                acc[pw] = ffma2(splat(s.w_hi), T1[pw], acc[pw]);
              }
            }
          }
          // count = 4 for sampling_ratio 2: multiply by 0.25 is exact
#pragma unroll
          for (int pw = 0; pw < P; ++pw) {
            my[ph * P + pw] = acc[pw].x * 0.25f;
            my[PP + ph * P + pw] = acc[pw].y * 0.25f;
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_store_smem_to_global(out + ((size_t)k * p.C + c0) * PP, tile, (uint32_t)(nch * PP * sizeof(float)));
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all();
}

// ---- forward, warp items (P = 7 and P = 14) ------------------------------------------------------
// A work item belongs to ONE warp: no block barrier anywhere, every warp streams its own items
// (table build -> row walk -> tile in shared memory -> one bulk store), so the SM always has 16
// independent load streams in flight.
//   P = 7 : item = (RoI, 64 channels), lane = channel pair, each lane pools all 7 x-bins.
//   P = 14: item = (RoI, 32 channels); the two half-warps take x-bins 0-6 and 7-13 of the same 16
//           channel pairs, so a lane still owns 7 bins (same register budget) and a warp load still
//           covers whole 128-byte lines (16 lanes x 8 B per column).
//
// x direction, "bin-dense": the two samples of a bin touch a contiguous run of <= NB columns
// (NB = 3 for RoIs up to 14 feature px wide at P = 7, 4 up to 28).  The per-sample (lo, hi, w_lo,
// w_hi) taps are folded once per RoI into NB weights per bin, and a window row is pooled with a
// fixed, branch-free sequence of 7*NB vector loads + 7*NB packed FMAs (no reuse bookkeeping per
// sample; columns shared by neighbouring bins hit L1).  Wider RoIs take the per-sample path (NB = 0).
// y direction: the distinct rows are walked once with the two-row cache (SAME / SHIFT / NEW).
struct __align__(16) AxisTapB {
  uint32_t off_lo, off_hi;  // BYTE offsets (already multiplied by the axis stride)
  float w_lo, w_hi;
};

constexpr int kPadBins = 16;  // table capacity in x-bins (halves * bins per lane <= 16)

template <int P>
struct __align__(16) WarpTables {
  AxisTapB ys[2 * P];
  AxisTapB xs[2 * kPadBins];
  float4 xw[kPadBins];      // folded weights of the bin's columns xoff, xoff + sw, ... (<= 4 columns)
  uint32_t xoff[kPadBins];  // byte offset of the bin's first column
  int xfirst[kPadBins];     // its column index (backward: how far the accumulator window slides between bins)
  uint32_t ymode[2 * P];
  int lo[2][2 * P];      // scratch: neighbour indices per axis (lo = -1: sample contributes nothing)
  int hi[2][2 * P];
};

// Builds the tables of the warp's current RoI; returns the widest bin run in columns (0..4 -> bin
// path with NB = max(3, run); > 4 -> per-sample path).  All 32 lanes must call.
// Bins P .. nbins-1 are padding for lanes whose bin group is short (7 bins over two half-warps = 4 + 3): they alias
// the last real bin's columns with zero weights.
template <int P>
__device__ __forceinline__ int build_tables_warp(WarpTables<P>& tb, const RoiGeom& g, const LvParam& lv, int lane, int nbins) {
  static_assert(2 * P <= 32 && P < 31, "one lane per sample");
  __syncwarp();  // the previous item's readers are done with the tables
  if (lane < 2 * P) {
    int lo, hi;
    float wl, wh;
    bool ok = axis_tap(sample_pos(g.sw, lane >> 1, g.bw, lane & 1, 2, g.cpu_coords), lv.W, lo, hi, wl, wh);
    tb.xs[lane] = AxisTapB{(uint32_t)(lo * lv.sw) * 4u, (uint32_t)(hi * lv.sw) * 4u, wl, wh};
    tb.lo[0][lane] = ok ? lo : -1;
    tb.hi[0][lane] = hi;
    ok = axis_tap(sample_pos(g.sh, lane >> 1, g.bh, lane & 1, 2, g.cpu_coords), lv.H, lo, hi, wl, wh);
    tb.ys[lane] = AxisTapB{(uint32_t)(lo * lv.sh) * 4u, (uint32_t)(hi * lv.sh) * 4u, wl, wh};
    tb.lo[1][lane] = ok ? lo : -1;
    tb.hi[1][lane] = hi;
  }
  __syncwarp();
  int run = 0;
  {  // row reuse flags: sample t compares its rows with the previous VALID sample's (one ballot + two shuffles
     // instead of a 2P-step serial scan on one lane, which was half of the table build's latency)
    const int ylo = lane < 2 * P ? tb.lo[1][lane] : -1, yhi = lane < 2 * P ? tb.hi[1][lane] : 0;
    const uint32_t before = __ballot_sync(0xffffffffu, ylo >= 0) & ((1u << lane) - 1u);
    const int prev = before ? 31 - __clz(before) : lane;
    int c0 = __shfl_sync(0xffffffffu, ylo, prev), c1 = __shfl_sync(0xffffffffu, yhi, prev);
    if (!before) c0 = c1 = -1;
    if (lane < 2 * P) {
      uint32_t m = 0u;
      if (ylo >= 0) {
        m = (ylo == c0) ? kSame : ((ylo == c1) ? kShift : kNew);
        if (yhi == ylo) m |= kBorder;
        m |= kValid;
      }
      tb.ymode[lane] = m;
    }
  }
  if (lane < P) {  // fold the two samples of bin `lane` into weights over a run of columns
    const int l0 = tb.lo[0][2 * lane], h0 = tb.hi[0][2 * lane];
    const int l1 = tb.lo[0][2 * lane + 1], h1 = tb.hi[0][2 * lane + 1];
    float* w = reinterpret_cast<float*>(&tb.xw[lane]);
    w[0] = w[1] = w[2] = w[3] = 0.f;
    int first = 0;
    if (l0 >= 0 || l1 >= 0) {
      const int cmin = l0 < 0 ? l1 : (l1 < 0 ? l0 : min(l0, l1));
      const int cmax = l0 < 0 ? h1 : (l1 < 0 ? h0 : max(h0, h1));
      first = max(min(cmin, lv.W - 4), 0);  // keep first .. first+3 inside the row (W >= 4)
      run = cmax - first + 1;
      if (run <= 4) {
        if (l0 >= 0) {
          w[l0 - first] += tb.xs[2 * lane].w_lo;
          w[h0 - first] += tb.xs[2 * lane].w_hi;
        }
        if (l1 >= 0) {
          w[l1 - first] += tb.xs[2 * lane + 1].w_lo;
          w[h1 - first] += tb.xs[2 * lane + 1].w_hi;
        }
      }
    }
    tb.xoff[lane] = (uint32_t)(first * lv.sw) * 4u;
    tb.xfirst[lane] = first;
  }
  run = __reduce_max_sync(0xffffffffu, run);
  __syncwarp();
  if (lane >= P && lane < nbins) {
    tb.xw[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    tb.xoff[lane] = tb.xoff[P - 1];
    tb.xfirst[lane] = tb.xfirst[P - 1];
    tb.xs[2 * lane] = AxisTapB{0u, 0u, 0.f, 0.f};
    tb.xs[2 * lane + 1] = AxisTapB{0u, 0u, 0.f, 0.f};
  }
  __syncwarp();
  return run;
}

__device__ __forceinline__ float2 ldg_f2b(const char* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// One window row -> T[0..6] = this lane's 7 x-bins (bins xb0 .. xb0+6), x-pooled.
// CSW: compile-time column stride in elements (0 = use swb).
template <int P, int XB, int NB, int CSW>
__device__ __forceinline__ void pool_row_warp(const WarpTables<P>& tb, int xb0, const uint32_t (&xo)[XB],
                                              const float (&xw)[XB][NB > 0 ? NB : 1], const char* __restrict__ row, uint32_t swb,
                                              float2 (&T)[XB]) {
  const uint32_t cs = CSW ? (uint32_t)CSW * 4u : swb;
  if (NB > 0) {
    float2 v[XB][NB > 0 ? NB : 1];
#pragma unroll
    for (int pw = 0; pw < XB; ++pw) {
      const char* p = row + xo[pw];
#pragma unroll
      for (int j = 0; j < NB; ++j) v[pw][j] = ldg_f2b(p + j * cs);
    }
#pragma unroll
    for (int pw = 0; pw < XB; ++pw) {
      float2 t = __fmul2_rn(splat(xw[pw][0]), v[pw][0]);
#pragma unroll
      for (int j = 1; j < NB; ++j) t = ffma2(splat(xw[pw][j]), v[pw][j], t);
      T[pw] = t;
    }
  } else {
    float2 a[2 * XB], b[2 * XB];
#pragma unroll
    for (int t = 0; t < 2 * XB; ++t) {
      const AxisTapB s = tb.xs[2 * xb0 + t];  // samples that contribute nothing have offsets 0 and weights 0
      a[t] = ldg_f2b(row + s.off_lo);
      b[t] = ldg_f2b(row + s.off_hi);
    }
#pragma unroll
    for (int pw = 0; pw < XB; ++pw) T[pw] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 2 * XB; ++t) {
      const AxisTapB s = tb.xs[2 * xb0 + t];
      if (s.w_lo + s.w_hi != 0.f) {  // uniform per half-warp
        T[t >> 1] = ffma2(splat(s.w_lo), a[t], T[t >> 1]);
        T[t >> 1] = ffma2(splat(s.w_hi), b[t], T[t >> 1]);
      }
    }
  }
}

// Walks the distinct rows of the window and writes this lane's two channels x 7 x-bins of the tile.
// Every branch is warp-uniform and says so through a vote, so that ptxas keeps the uniform datapath.
template <int P, int XB, int NB, int CSW>
__device__ __forceinline__ void roi_warp_body(const WarpTables<P>& tb, int xb0, const char* __restrict__ fb, uint32_t swb,
                                              float* __restrict__ my) {
  constexpr int PP = P * P;
  constexpr uint32_t kAll = 0xffffffffu;
  float2 T0[XB], T1[XB], acc[XB];
  uint32_t xo[XB];  // bin column offsets stay in registers for the whole RoI
#pragma unroll
  for (int pw = 0; pw < XB; ++pw) {
    T0[pw] = T1[pw] = acc[pw] = make_float2(0.f, 0.f);
    xo[pw] = NB > 0 ? tb.xoff[xb0 + pw] : 0u;
  }
  // the bins' folded column weights stay in registers for the whole RoI (one broadcast LDS.128 per bin per ROW before:
  // 105 of the item's 375 shared-memory wavefronts, on the LSU data pipe that bounds the kernel)
  float xw[XB][NB > 0 ? NB : 1];
#pragma unroll
  for (int pw = 0; pw < XB; ++pw) {
    const float4 w = tb.xw[xb0 + pw];
    xw[pw][0] = w.x;
    if (NB > 1) xw[pw][1 % (NB > 0 ? NB : 1)] = w.y;
    if (NB > 2) xw[pw][2 % (NB > 0 ? NB : 1)] = w.z;
    if (NB > 3) xw[pw][3 % (NB > 0 ? NB : 1)] = w.w;
  }
  // P = 7, tile stores without bank conflicts: the lane's channel rows start 98 words apart, so lanes l and
  // l+16 share banks; the upper half-warp therefore stores its odd channel while the lower stores its even
  // one (49 words = 17 banks apart) and vice versa.  (P = 14: plain order.)
  const bool lower = P != 7 || (threadIdx.x & 16) == 0;
  float* const o_a = my + (lower ? 0 : PP);
  float* const o_b = my + (lower ? PP : 0);
  const int nmine = min(XB, P - xb0);  // real bins of this lane (the rest are padding)
  // The two-row cache without register copies: `flip` says which register set holds the sample's LOWER row.  When a
  // sample's lower row is the previous sample's upper row (kShift: every step of a tall RoI) the sets swap roles and
  // only the new upper row is pooled, into the set that held the old lower row — a T0 = T1 copy here cost 28 moves
  // per sample, 14 % of the kernel's instructions (ncu r01c source page).
  bool flip = false;
#pragma unroll 1
  for (int t = 0; t < 2 * P; ++t) {
    const uint32_t m = tb.ymode[t];
    if (__any_sync(kAll, m & kValid)) {
      const uint32_t mode = m & kModeMask;
      const AxisTapB s = tb.ys[t];
      // count = 4 for sampling_ratio 2; the exact factor 0.25 rides on the row weights (scaling by a power of two
      // commutes with every rounding of the FMA chain)
      const float2 wl = splat(s.w_lo * 0.25f), wh = splat(s.w_hi * 0.25f);
      const bool border = __any_sync(kAll, m & kBorder);
      const bool is_new = __any_sync(kAll, mode == kNew), is_shift = __any_sync(kAll, mode == kShift);
      auto y_step = [&](auto& LO, auto& HI) -> bool {  // returns true when the sets swapped roles
        if (is_new || (is_shift && !border)) pool_row_warp<P, XB, NB, CSW>(tb, xb0, xo, xw, fb + (is_new ? s.off_lo : s.off_hi), swb, LO);
        if (is_shift && !border) {  // lower row = HI (kept), upper row = LO (just pooled)
#pragma unroll
          for (int pw = 0; pw < XB; ++pw) {
            acc[pw] = ffma2(wl, HI[pw], acc[pw]);
            acc[pw] = ffma2(wh, LO[pw], acc[pw]);
          }
          return true;
        }
        if (is_shift) {  // bottom edge: both rows are the previous upper row
#pragma unroll
          for (int pw = 0; pw < XB; ++pw) LO[pw] = HI[pw];
        } else if (is_new) {
          if (border) {
#pragma unroll
            for (int pw = 0; pw < XB; ++pw) HI[pw] = LO[pw];
          } else {
            pool_row_warp<P, XB, NB, CSW>(tb, xb0, xo, xw, fb + s.off_hi, swb, HI);
          }
        }
#pragma unroll
        for (int pw = 0; pw < XB; ++pw) {
          acc[pw] = ffma2(wl, LO[pw], acc[pw]);
          acc[pw] = ffma2(wh, HI[pw], acc[pw]);
        }
        return false;
      };
      if (!flip) flip = y_step(T0, T1);
      else flip = !y_step(T1, T0);
    }
    if (t & 1) {  // bin row complete
      const int o = (t >> 1) * P;
#pragma unroll
      for (int pw = 0; pw < XB; ++pw) {
        const float e = acc[pw].x, f = acc[pw].y;
        if (pw < nmine) {
          o_a[o + pw] = lower ? e : f;
          o_b[o + pw] = lower ? f : e;
        }
        acc[pw] = make_float2(0.f, 0.f);
      }
    }
  }
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_store_smem_to_global_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}

template <int P, int XB>
struct WarpItem {
  static constexpr int kHalves = (P + XB - 1) / XB;       // x-bin groups per warp: 1 (P=7, XB=7) or 2 (P=14 XB=7; P=7 XB=4)
  static constexpr int kPairs = 32 / kHalves;             // channel pairs per warp item
  static constexpr int kChannels = 2 * kPairs;            // 64 or 32
  static constexpr int kTileFloats = kChannels * P * P;   // 12 544 B or 25 088 B
};

template <int P, int XB, int WARPS, int CSW>
__global__ void __launch_bounds__(WARPS * 32, P == 7 ? (XB == 7 ? 4 : 6) : 2)
    roi_fwd_warp_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int groups, int flags, int ipw) {
  using WI = WarpItem<P, XB>;
  static_assert(WI::kHalves <= 2 && WI::kHalves * XB <= kPadBins, "x-bins: one or two groups per warp");
  constexpr int PP = P * P;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);  // tells ptxas the value is warp-uniform
  const int lane = threadIdx.x & 31;
  const int pair = lane % WI::kPairs;          // channel pair of this lane inside the item
  const int xb0 = (lane / WI::kPairs) * XB;   // first x-bin of this lane
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * WI::kTileFloats;
  WarpTables<P>* tbs = reinterpret_cast<WarpTables<P>*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats);
  __shared__ int s_run[WARPS], s_b[WARPS], s_lvl[WARPS];
  float* my = tile + (size_t)(2 * pair) * PP + xb0;
  const uint64_t pol = l2_policy_evict_first();
  const long long items = (long long)p.K * groups;
  // Work order.  ipw > 0: CTA b owns the WARPS*ipw consecutive items starting at b*WARPS*ipw; the grid covers all
  // items and the hardware launches CTAs in index order as SMs drain, so the RoIs in flight are always neighbours
  // in the list (one or two frames: their feature maps stay L2-resident).  ipw == 0: persistent grid with a static
  // stride (kept for A/B runs; warps drift apart by many frames and the maps are re-read from HBM).
  //
  // Shared tables: when a RoI has exactly WARPS channel groups (C = 256), the CTA's ipw RoIs are shared by its warps
  // (warp w = channel group w), so the geometry / tap / fold tables are built ONCE per RoI (warp i builds RoI i) behind
  // the kernel's only __syncthreads instead of once per warp item (17 % of the executed instructions, ncu r01c).
  const bool shared_tables = groups == WARPS && ipw >= 1 && ipw <= WARPS && !(flags & 2);
  const int k0 = blockIdx.x * ipw;
  if (shared_tables) {
    if (warp < ipw && k0 + warp < p.K) {
      const RoiGeom g = roi_geom(p, k0 + warp);
      const bool live = __any_sync(kAll, g.live);
      int run = -1;
      if (live) run = build_tables_warp<P>(tbs[warp], g, p.lv[g.lvl], lane, WI::kHalves * XB);
      if (lane == 0) {
        s_run[warp] = run;
        s_b[warp] = g.b;
        s_lvl[warp] = live ? g.lvl : 0;
      }
    }
    __syncthreads();
  }
  const long long first = ipw > 0 ? (long long)blockIdx.x * WARPS * ipw + warp : (long long)blockIdx.x * WARPS + warp;
  const long long last = ipw > 0 ? min(items, ((long long)blockIdx.x + 1) * WARPS * ipw) : items;
  const long long stride = ipw > 0 ? WARPS : (long long)gridDim.x * WARPS;

  for (long long item = first; item < last; item += stride) {
    int k, cg, run = 0, gb, glvl;
    bool live;
    const WarpTables<P>* tbp;
    if (shared_tables) {  // block-uniform
      const int slot = (int)((item - first) / WARPS);
      k = k0 + slot;
      cg = warp;
      run = s_run[slot];
      live = __any_sync(kAll, run >= 0);
      gb = s_b[slot];
      glvl = s_lvl[slot];
      tbp = &tbs[slot];
    } else {
      k = (int)(item / groups);
      cg = (int)(item - (long long)k * groups);
      const RoiGeom g = roi_geom(p, k);
      live = __any_sync(kAll, g.live);
      gb = g.b;
      glvl = live ? g.lvl : 0;
      if (live) run = build_tables_warp<P>(tbs[warp], g, p.lv[glvl], lane, WI::kHalves * XB);
      tbp = &tbs[warp];
    }
    const WarpTables<P>& tb = *tbp;
    const int c0 = cg * WI::kChannels;
    const int nch = min(WI::kChannels, p.C - c0);  // multiple of 4
    const LvParam& lv = p.lv[glvl];
    // the previous item's bulk store must have finished READING the tile before it is rewritten
    if (lane == 0) bulk_wait_read_all();
    __syncwarp();
    if (!live) {
      for (int j = lane; j < WI::kTileFloats; j += 32) tile[j] = 0.f;
    } else {
      // lanes past the last channel pair of a short group redo the last pair (their tile rows are not stored)
      const char* fb = reinterpret_cast<const char*>(lv.data + (size_t)gb * lv.sn + c0 + min(2 * pair, nch - 2));  // sc == 1
      const uint32_t swb = (uint32_t)lv.sw * 4u;
      if (__all_sync(kAll, run <= 3)) roi_warp_body<P, XB, 3, CSW>(tb, xb0, fb, swb, my);
      else if (__all_sync(kAll, run == 4)) roi_warp_body<P, XB, 4, CSW>(tb, xb0, fb, swb, my);
      else roi_warp_body<P, XB, 0, CSW>(tb, xb0, fb, swb, my);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      float* dst = out + ((size_t)k * p.C + c0) * PP;
      const uint32_t bytes = (uint32_t)(nch * PP * sizeof(float));
      if (flags & 1) bulk_store_smem_to_global_hint(dst, tile, bytes, pol);
      else bulk_store_smem_to_global(dst, tile, bytes);
      bulk_commit();
    }
  }
  // Only the tile READS have to finish before the CTA's shared memory goes away; the writes complete on their own
  // before the grid does.  (wait_group without .read compiles to DEPBAR + CCTL.IVALL: every exiting CTA would
  // invalidate the SM's L1 under the three CTAs still gathering through it.)
  if (lane == 0) bulk_wait_read_all();
}

// ---- forward, row-major warp items (P = 7) -----------------------------------------------------------
// roi_fwd_warp_kernel walks the 14 y-samples of a RoI and keeps a two-row cache: every sample costs reuse bookkeeping
// (SAME / SHIFT / NEW votes, register-set flips, border copies: 2 600 instructions per item for ~1 000 of payload), and
// every distinct window row costs one exposed L2 round trip (its loads are issued, waited for and consumed before the next
// row's are issued).  Here the RoI's y taps are turned, once per RoI, into a ROW PROGRAM: the distinct window rows in
// ascending order, each with the (pre-added) weights it contributes to at most three consecutive bin rows.  A warp then
//   * pools ROWS (1 or 2) window rows per phase — with 2, both rows' loads are in flight together, halving the exposed
//     round trips per item at the cost of 42 more registers (12 instead of 16 warps per SM);
//   * adds each x-pooled row into three accumulator sets acc0..acc2 = bin rows base, base+1, base+2 with its three weights
//     (21 FFMA2, no mode logic); when a row's first bin row is past `base`, the finished bin rows are written to the tile and
//     the sets rotate.
// Rows whose weights span more than three bin rows (bins under ~0.7 feature px: RoIs under ~20 image px), windows taller
// than 32 rows or wider than NB = 4 columns per bin take roi_warp_body as before.  Pre-adding the taps of a shared row
// changes the rounding of the y sum by an ulp: results agree with roi_fwd_warp_kernel to ~1e-7 relative, not bit for bit.
struct __align__(16) RowOp {
  uint32_t off_pa;   // byte offset of the row inside the map, first bin row in the low 3 bits (rows are >= 16 bytes apart)
  float w0, w1, w2;  // weights (count 1/4 folded in) for bin rows pa, pa + 1, pa + 2
};

// Builds the row program of the warp's current RoI into tb.xs (free once the bins are folded: run <= 4) and returns the number
// of rows (0: every sample is outside the map -> an all-zero tile), or -1 when the RoI does not qualify.  All lanes call.
__device__ __forceinline__ int build_row_program(WarpTables<7>& tb, uint32_t shb, int lane, int* span_out = nullptr) {
  constexpr uint32_t kAll = 0xffffffffu;
  const int ylo = lane < 14 ? tb.lo[1][lane] : -1, yhi = lane < 14 ? tb.hi[1][lane] : -1;
  const int rmin = __reduce_min_sync(kAll, ylo >= 0 ? ylo : 0x7fffffff);
  const int rmax = __reduce_max_sync(kAll, ylo >= 0 ? yhi : -1);
  if (rmax < 0) return 0;
  if (rmax - rmin >= 32) return -1;
  const int r = rmin + lane;
  // Row r's weights for the (at most three) bin rows it feeds, accumulated in sample order exactly like a per-bin-row sum
  // w[ph] += w_lo / 4; w[ph] += w_hi / 4 would be (zero terms are exact no-ops): pa = first bin row with a non-zero term,
  // pb = last.  No per-lane array: nvcc puts a dynamically read w[7] into local memory (LDL/STL on the table build's
  // critical path).
  int pa = 7, pb = -1;
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;
#pragma unroll
  for (int t = 0; t < 14; ++t) {
    const int lo = tb.lo[1][t], hi = tb.hi[1][t];
    if (lo >= 0) {
      const AxisTapB s = tb.ys[t];
      const float a = lo == r ? s.w_lo * 0.25f : 0.f;
      const float b = hi == r ? s.w_hi * 0.25f : 0.f;
      if (a != 0.f || b != 0.f) {
        pa = min(pa, t >> 1);
        pb = t >> 1;
      }
      const int d = (t >> 1) - pa;
      w0 += d == 0 ? a : 0.f;
      w0 += d == 0 ? b : 0.f;
      w1 += d == 1 ? a : 0.f;
      w1 += d == 1 ? b : 0.f;
      w2 += d == 2 ? a : 0.f;
      w2 += d == 2 ? b : 0.f;
    }
  }
  const bool used = pb >= 0 && r <= rmax;
  if (!__all_sync(kAll, !used || pb - pa <= 2)) return -1;
  if (span_out) *span_out = __reduce_max_sync(kAll, used ? pb - pa : 0);  // 1: every row feeds at most two bin rows
  const uint32_t mask = __ballot_sync(kAll, used);
  // first bin rows must not decrease along the rows (they cannot for x2 >= x1 geometry; checked, not assumed)
  const uint32_t before = mask & ((1u << lane) - 1u);
  const int prev_pa = __shfl_sync(kAll, pa, before ? 31 - __clz(before) : lane);
  if (!__all_sync(kAll, !used || !before || prev_pa <= pa)) return -1;
  if (used) {
    RowOp op;
    op.off_pa = ((uint32_t)r * shb) | (uint32_t)pa;
    op.w0 = w0;
    op.w1 = w1;
    op.w2 = w2;
    reinterpret_cast<RowOp*>(tb.xs)[__popc(before)] = op;
  }
  __syncwarp();
  return __popc(mask);
}

template <int NB, int ROWS, int CSW>
__device__ __forceinline__ void roi_warp_body_rm(const WarpTables<7>& tb, int nrows, const char* __restrict__ fb, uint32_t swb,
                                                 float* __restrict__ my) {
  constexpr int P = 7, PP = 49;
  const uint32_t cs = CSW ? (uint32_t)CSW * 4u : swb;
  const RowOp* rows = reinterpret_cast<const RowOp*>(tb.xs);
  uint32_t xo[P];
  float2 acc0[P], acc1[P], acc2[P];
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
    xo[pw] = tb.xoff[pw];
    acc0[pw] = acc1[pw] = acc2[pw] = make_float2(0.f, 0.f);
  }
  // (the folded x weights are re-read from shared memory per bin, one broadcast LDS.128 each: holding all 21-28 of them in
  // registers next to two rows of loaded columns does not fit the 168 registers that 12 warps per SM allow — the spill
  // stores of just-loaded columns serialised the two rows' loads again)
  const bool lower = (threadIdx.x & 16) == 0;  // conflict-free tile stores, see roi_warp_body
  float* const o_a = my + (lower ? 0 : PP);
  float* const o_b = my + (lower ? PP : 0);
  int base = 0;
  auto retire = [&](int upto) {  // bin rows base .. upto-1 are complete: write them out and rotate the accumulator sets
    while (base < upto) {
      const int o = base * P;
#pragma unroll
      for (int pw = 0; pw < P; ++pw) {
        const float e = acc0[pw].x, f = acc0[pw].y;
        o_a[o + pw] = lower ? e : f;
        o_b[o + pw] = lower ? f : e;
        acc0[pw] = acc1[pw];
        acc1[pw] = acc2[pw];
        acc2[pw] = make_float2(0.f, 0.f);
      }
      ++base;
    }
  };
  auto load_row = [&](const RowOp& op, float2 (&v)[P][NB]) {
    const char* row = fb + (op.off_pa & ~7u);
#pragma unroll
    for (int pw = 0; pw < P; ++pw) {
      const char* q = row + xo[pw];
#pragma unroll
      for (int j = 0; j < NB; ++j) v[pw][j] = ldg_f2b(q + j * cs);
    }
  };
  auto apply_row = [&](const RowOp& op, const float2 (&v)[P][NB]) {
    retire((int)(op.off_pa & 7u));
    const float2 w0 = splat(op.w0), w1 = splat(op.w1), w2 = splat(op.w2);
#pragma unroll
    for (int pw = 0; pw < P; ++pw) {
      const float4 xw4 = tb.xw[pw];
      const float xw[4] = {xw4.x, xw4.y, xw4.z, xw4.w};
      float2 t = __fmul2_rn(splat(xw[0]), v[pw][0]);
#pragma unroll
      for (int j = 1; j < NB; ++j) t = ffma2(splat(xw[j]), v[pw][j], t);
      acc0[pw] = ffma2(w0, t, acc0[pw]);
      acc1[pw] = ffma2(w1, t, acc1[pw]);
      acc2[pw] = ffma2(w2, t, acc2[pw]);
    }
  };
#pragma unroll 1
  for (int i = 0; i < nrows; i += ROWS) {
    const RowOp ra = rows[i];
    float2 va[P][NB];
    load_row(ra, va);
    if (ROWS == 2 && i + 1 < nrows) {  // warp-uniform: both rows' loads are in flight before either is consumed
      const RowOp rb = rows[i + 1];
      float2 vb[P][NB];
      load_row(rb, vb);
      apply_row(ra, va);
      apply_row(rb, vb);
    } else {
      apply_row(ra, va);
    }
  }
  retire(P);
}

template <int ROWS, int CSW>
__global__ void __launch_bounds__(128, ROWS == 2 ? 3 : 4)
    roi_fwd_rm_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int flags, int ipw) {
  using WI = WarpItem<7, 7>;
  constexpr int WARPS = 4, PP = 49;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * WI::kTileFloats;
  WarpTables<7>* tbs = reinterpret_cast<WarpTables<7>*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats);
  __shared__ int s_run[WARPS], s_b[WARPS], s_lvl[WARPS], s_rows[WARPS];
  float* my = tile + (size_t)(2 * lane) * PP;
  const uint64_t pol = l2_policy_evict_first();
  // C == 256: the CTA's ipw RoIs are shared by its four warps (warp = 64-channel group); warp i builds RoI i's tables
  const int k0 = blockIdx.x * ipw;
  if (warp < ipw && k0 + warp < p.K) {
    const RoiGeom g = roi_geom(p, k0 + warp);
    const bool live = __any_sync(kAll, g.live);
    int run = -1, nrows = -1;
    if (live) {
      const LvParam& lv = p.lv[g.lvl];
      run = build_tables_warp<7>(tbs[warp], g, lv, lane, 7);
      if (run <= 4) nrows = build_row_program(tbs[warp], (uint32_t)lv.sh * 4u, lane);
    }
    if (lane == 0) {
      s_run[warp] = run;
      s_b[warp] = g.b;
      s_lvl[warp] = live ? g.lvl : 0;
      s_rows[warp] = nrows;
    }
  }
  __syncthreads();
  for (int slot = 0; slot < ipw && k0 + slot < p.K; ++slot) {
    const int k = k0 + slot;
    const int run = s_run[slot], nrows = s_rows[slot];
    const WarpTables<7>& tb = tbs[slot];
    const int c0 = warp * WI::kChannels;
    const LvParam& lv = p.lv[s_lvl[slot]];
    if (lane == 0) bulk_wait_read_all();  // the previous item's bulk store has finished reading the tile
    __syncwarp();
    if (run < 0) {
      for (int j = lane; j < WI::kTileFloats; j += 32) tile[j] = 0.f;
    } else {
      const char* fb = reinterpret_cast<const char*>(lv.data + (size_t)s_b[slot] * lv.sn + c0 + 2 * lane);
      const uint32_t swb = (uint32_t)lv.sw * 4u;
      if (nrows >= 0 && run <= 3) roi_warp_body_rm<3, ROWS, CSW>(tb, nrows, fb, swb, my);
      else if (nrows >= 0) roi_warp_body_rm<4, 1, CSW>(tb, nrows, fb, swb, my);
      else if (run <= 3) roi_warp_body<7, 7, 3, CSW>(tb, 0, fb, swb, my);
      else if (run == 4) roi_warp_body<7, 7, 4, CSW>(tb, 0, fb, swb, my);
      else roi_warp_body<7, 7, 0, CSW>(tb, 0, fb, swb, my);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      float* dst = out + ((size_t)k * p.C + c0) * PP;
      const uint32_t bytes = (uint32_t)(WI::kChannels * PP * sizeof(float));
      if (flags & 1) bulk_store_smem_to_global_hint(dst, tile, bytes, pol);
      else bulk_store_smem_to_global(dst, tile, bytes);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read_all();
}

// A RoI that certainly has no row program (decided from its geometry alone, before any table is built): bins at least
// 4.5 px wide put a bin's two samples (half a bin apart, two columns each) on five or more columns; a window over 34 px tall
// spans more than 32 rows.  Only for RoIs that the map does not clip (a clipped one may still qualify: built and decided there).
__device__ __forceinline__ bool row_program_hopeless(const RoiGeom& g, const LvParam& lv) {
  return (g.bw >= 4.5f && g.sw >= 0.f && g.sw + 7.f * g.bw <= (float)lv.W - 1.f) ||
         (g.bh * 7.f >= 34.f && g.sh >= 0.f && g.sh + 7.f * g.bh <= (float)lv.H - 1.f);
}

// The y half of build_tables_warp (what build_row_program reads), for a warp that builds a RoI's row program while another
// warp builds the RoI's x tables.
__device__ __forceinline__ void build_y_taps(WarpTables<7>& tb, const RoiGeom& g, const LvParam& lv, int lane) {
  __syncwarp();
  if (lane < 14) {
    int lo, hi;
    float wl, wh;
    const bool ok = axis_tap(sample_pos(g.sh, lane >> 1, g.bh, lane & 1, 2, g.cpu_coords), lv.H, lo, hi, wl, wh);
    tb.ys[lane] = AxisTapB{(uint32_t)(lo * lv.sh) * 4u, (uint32_t)(hi * lv.sh) * 4u, wl, wh};
    tb.lo[1][lane] = ok ? lo : -1;
    tb.hi[1][lane] = hi;
  }
  __syncwarp();
}

// ---- forward, row program with PIPELINED rows (P = 7, C == 256): LCR_ROI_FWD=rmp ---------------------------------
// ncu of roi_fwd_warp_kernel / roi_fwd_rm_kernel on the bench list (profiles/r02b_*): 39-45 % of all warp samples are
// long-scoreboard stalls, two thirds of them on the FIRST FMUL2 of a window row — a warp issues a row's 21-28 loads, has
// nothing else in flight, and sits through one loaded L2 round trip (~900 cycles) per row, ~12 rows per item.  Here the
// row loop is software-pipelined WITHOUT extra registers: as soon as a bin's columns of row i have been multiplied into the
// bin's x-pooled value, the same registers are re-loaded with the bin's columns of row i+1, so the next row's loads are in
// flight during this row's remaining FMAs, the accumulator rotation and the tile stores, and a warp's wait per row shrinks
// from (round trip) to (round trip - row compute).  The wait for the previous item's bulk store (tile reuse) moves from
// the top of the item to the first tile write.
template <int NB, int NACC, int CSW, int XWR>
__device__ __forceinline__ void roi_warp_body_rmp(const WarpTables<7>& tb, int nrows, const char* __restrict__ fb, uint32_t swb,
                                                  float* __restrict__ my, bool& store_pending, const RowOp* rows_at = nullptr) {
  constexpr int P = 7, PP = 49;
  const uint32_t cs = CSW ? (uint32_t)CSW * 4u : swb;
  const RowOp* rows = rows_at ? rows_at : reinterpret_cast<const RowOp*>(tb.xs);  // (where build_row_program left it)
  uint32_t xo[P];
  float2 acc0[P], acc1[P], acc2[P];
  float xwr[XWR > 0 ? XWR : 1][NB];  // the folded x weights of bins 0 .. XWR-1 (the rest are re-read from shared memory per row)
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
    xo[pw] = tb.xoff[pw];
    acc0[pw] = acc1[pw] = acc2[pw] = make_float2(0.f, 0.f);
    if (pw < XWR) {
      const float4 w = tb.xw[pw];
      xwr[pw % (XWR > 0 ? XWR : 1)][0] = w.x;
      if (NB > 1) xwr[pw % (XWR > 0 ? XWR : 1)][1 % NB] = w.y;
      if (NB > 2) xwr[pw % (XWR > 0 ? XWR : 1)][2 % NB] = w.z;
      if (NB > 3) xwr[pw % (XWR > 0 ? XWR : 1)][3 % NB] = w.w;
    }
  }
  // (Not loading a bin's LAST column when it carries no weight — both samples on the first NB-1 columns, 20-60 % of the bins —
  // was tried for the LSU wavefronts it saves: the per-bin predicate costs registers the body does not have, 1.33 -> 1.37 ms.)
  const bool lower = (threadIdx.x & 16) == 0;  // conflict-free tile stores, see roi_warp_body
  float* const o_a = my + (lower ? 0 : PP);
  float* const o_b = my + (lower ? PP : 0);
  int base = 0;
  auto retire = [&](int upto) {  // bin rows base .. upto-1 are complete: write them out and rotate the accumulator sets
    if (base < upto && store_pending) {  // warp-uniform: the previous item's bulk store must have finished READING the tile
      if ((threadIdx.x & 31) == 0) bulk_wait_read_all();
      __syncwarp();
      store_pending = false;
    }
    while (base < upto) {
      const int o = base * P;
#pragma unroll
      for (int pw = 0; pw < P; ++pw) {
        const float e = acc0[pw].x, f = acc0[pw].y;
        o_a[o + pw] = lower ? e : f;
        o_b[o + pw] = lower ? f : e;
        acc0[pw] = acc1[pw];
        if (NACC > 2) {
          acc1[pw] = acc2[pw];
          acc2[pw] = make_float2(0.f, 0.f);
        } else {
          acc1[pw] = make_float2(0.f, 0.f);
        }
      }
      ++base;
    }
  };
  float2 v[P][NB];
  {
    const char* row = fb + (rows[0].off_pa & ~7u);
#pragma unroll
    for (int pw = 0; pw < P; ++pw) {
      const char* q = row + xo[pw];
#pragma unroll
      for (int j = 0; j < NB; ++j) v[pw][j] = ldg_f2b(q + j * cs);
    }
  }
#pragma unroll 1
  for (int i = 0; i < nrows; ++i) {
    const bool more = i + 1 < nrows;  // warp-uniform
    const RowOp op = rows[i];  // (read one row ahead it costs four registers: ptxas spills in the loop, 1.33 -> 1.43 ms)
    const char* nrow = fb + (rows[more ? i + 1 : i].off_pa & ~7u);
    retire((int)(op.off_pa & 7u));
    const float2 w0 = splat(op.w0), w1 = splat(op.w1), w2 = splat(op.w2);
#pragma unroll
    for (int pw = 0; pw < P; ++pw) {
      float xw[4];
      if (pw < XWR) {
#pragma unroll
        for (int j = 0; j < NB; ++j) xw[j] = xwr[pw % (XWR > 0 ? XWR : 1)][j];
      } else {
        const float4 xw4 = tb.xw[pw];
        xw[0] = xw4.x; xw[1] = xw4.y; xw[2] = xw4.z; xw[3] = xw4.w;
      }
      float2 t = __fmul2_rn(splat(xw[0]), v[pw][0]);
#pragma unroll
      for (int j = 1; j < NB; ++j) t = ffma2(splat(xw[j]), v[pw][j], t);
      if (more) {  // this bin's columns of the NEXT row, into the registers just consumed
        const char* q = nrow + xo[pw];
#pragma unroll
        for (int j = 0; j < NB; ++j) v[pw][j] = ldg_f2b(q + j * cs);
      }
      acc0[pw] = ffma2(w0, t, acc0[pw]);
      acc1[pw] = ffma2(w1, t, acc1[pw]);
      if (NACC > 2) acc2[pw] = ffma2(w2, t, acc2[pw]);
    }
  }
  retire(P);
}

template <int CSW>
__global__ void __launch_bounds__(128, 4)
    roi_fwd_rmp_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int flags, int ipw) {
  using WI = WarpItem<7, 7>;
  constexpr int WARPS = 4, PP = 49;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * WI::kTileFloats;
  WarpTables<7>* tbs = reinterpret_cast<WarpTables<7>*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats);
  __shared__ int s_run[WARPS], s_b[WARPS], s_lvl[WARPS], s_rows[WARPS], s_span[WARPS];
  float* my = tile + (size_t)(2 * lane) * PP;
  const uint64_t pol = l2_policy_evict_first();
  // (an L2 prefetch of the RoI rows of the CTA two waves further down changed nothing: 1.331 ms either way)
  // Table build, split over the CTA's four warps (ipw <= 2 RoIs per CTA): warp i builds RoI i's x / y tap tables and folded
  // bin weights in slot i, warp 2 + i builds RoI i's row program (from its own copy of the y taps) in slot 2 + i — the two
  // halves of what was one warp's dependent chain in front of the barrier (7.6 % of the warp samples waited there).
  const int k0 = blockIdx.x * ipw;
  {
    const int r = warp & 1;  // RoI of this warp
    if (r < ipw && k0 + r < p.K) {
      const RoiGeom g = roi_geom(p, k0 + r);
      const bool live = __any_sync(kAll, g.live);
      if (warp < 2) {
        int run = -1;
        if (live) run = build_tables_warp<7>(tbs[r], g, p.lv[g.lvl], lane, 7);
        if (lane == 0) {
          s_run[r] = run;
          s_b[r] = g.b;
          s_lvl[r] = live ? g.lvl : 0;
        }
      } else {
        int nrows = -1, span = 2;
        if (live) {
          const LvParam& lv = p.lv[g.lvl];
          if (!row_program_hopeless(g, lv)) {
            build_y_taps(tbs[2 + r], g, lv, lane);
            nrows = build_row_program(tbs[2 + r], (uint32_t)lv.sh * 4u, lane, &span);
          }
        }
        if (lane == 0) {
          s_rows[r] = nrows;
          s_span[r] = span;
        }
      }
    }
  }
  __syncthreads();
  bool store_pending = false;
  for (int slot = 0; slot < ipw && k0 + slot < p.K; ++slot) {
    const int k = k0 + slot;
    const int run = s_run[slot], nrows = s_rows[slot];
    const WarpTables<7>& tb = tbs[slot];
    const int c0 = warp * WI::kChannels;
    const LvParam& lv = p.lv[s_lvl[slot]];
    const char* fb = reinterpret_cast<const char*>(lv.data + (size_t)s_b[slot] * lv.sn + c0 + 2 * lane);
    const uint32_t swb = (uint32_t)lv.sw * 4u;
    const bool two = s_span[slot] <= 1;  // every window row feeds at most two bin rows (bins at least one pixel tall)
    const RowOp* rows = reinterpret_cast<const RowOp*>(tbs[2 + slot].xs);
    // (XWR = 0: the folded x weights are re-read from shared memory per bin and row — 7 of a row's ~51 LSU wavefronts, on the
    // unit that bounds the kernel.  The bodies sit exactly at 128 registers: holding even three bins' weights in registers
    // makes ptxas spill inside the row loop (XWR = 7: 1.45 against 1.39 ms; XWR = 3: 16 bytes spilled).)
    if (nrows > 0 && run >= 0 && run <= 3) {
      if (two) roi_warp_body_rmp<3, 2, CSW, 0>(tb, nrows, fb, swb, my, store_pending, rows);
      else roi_warp_body_rmp<3, 3, CSW, 0>(tb, nrows, fb, swb, my, store_pending, rows);
    } else if (nrows > 0 && run == 4 && two) {
      roi_warp_body_rmp<4, 2, CSW, 0>(tb, nrows, fb, swb, my, store_pending, rows);
    } else {
      if (store_pending) {
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        store_pending = false;
      }
      if (run < 0 || nrows == 0) {  // padding row, or every sample outside the map
        for (int j = lane; j < WI::kTileFloats; j += 32) tile[j] = 0.f;
      } else {
        // one sample-walk body for every RoI without a row program (per-sample taps: any geometry) — with the bin-dense
        // NB = 3 / 4 walks inlined as well, a list mixing small and large RoIs ran 9 % slower than through
        // roi_fwd_warp_kernel (instruction-cache misses: seven loop bodies side by side on an SM)
        roi_warp_body<7, 7, 0, CSW>(tb, 0, fb, swb, my);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      float* dst = out + ((size_t)k * p.C + c0) * PP;
      const uint32_t bytes = (uint32_t)(WI::kChannels * PP * sizeof(float));
      if (flags & 1) bulk_store_smem_to_global_hint(dst, tile, bytes, pol);
      else bulk_store_smem_to_global(dst, tile, bytes);
      bulk_commit();
    }
    store_pending = true;
  }
  if (lane == 0) bulk_wait_read_all();
}

// ---- backward -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load_global_to_smem(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

// Scatter one window row: for the row's distinct columns accumulate sum_t w * U[t>>1] and issue one
// vector atomic per (column, channel pair).
template <int P>
__device__ __forceinline__ void scatter_row(const RoiTables<P>& tb, float* __restrict__ row, const float2 (&U)[P]) {
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  int o0 = -1, o1 = -1;  // element offsets of the two live columns (-1: empty)
#pragma unroll
  for (int t = 0; t < 2 * P; ++t) {
    const uint32_t m = tb.xmode[t];
    if (m & kValid) {
      const uint32_t mode = m & kModeMask;
      const AxisSample s = tb.xs[t];
      if (mode == kShift) {
        if (o0 >= 0) atomicAdd(reinterpret_cast<float2*>(row + o0), a0);
        a0 = a1; o0 = o1;
        a1 = make_float2(0.f, 0.f); o1 = -1;
      } else if (mode == kNew) {
        if (o0 >= 0) atomicAdd(reinterpret_cast<float2*>(row + o0), a0);
        if (o1 >= 0) atomicAdd(reinterpret_cast<float2*>(row + o1), a1);
        a0 = a1 = make_float2(0.f, 0.f);
        o1 = -1;
      }
      o0 = s.off_lo;
      a0 = ffma2(splat(s.w_lo), U[t >> 1], a0);
      if (m & kBorder) {
        a0 = ffma2(splat(s.w_hi), U[t >> 1], a0);
      } else {
        o1 = s.off_hi;
        a1 = ffma2(splat(s.w_hi), U[t >> 1], a1);
      }
    }
  }
  if (o0 >= 0) atomicAdd(reinterpret_cast<float2*>(row + o0), a0);
  if (o1 >= 0) atomicAdd(reinterpret_cast<float2*>(row + o1), a1);
}

template <int P, int NT>
__global__ void __launch_bounds__(NT, 4) roi_bwd_nhwc_kernel(const __grid_constant__ RoiParams p, const float* __restrict__ gout,
                                                          int groups, int ipc) {
  constexpr int CPB = 2 * NT;
  constexpr int PP = P * P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);  // grad_out tile [CPB][PP]
  RoiTables<P>& tb = *reinterpret_cast<RoiTables<P>*>(smem_raw + sizeof(float) * CPB * PP);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * CPB * PP + sizeof(RoiTables<P>));
  const int tid = threadIdx.x;
  const long long items = (long long)p.K * groups;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t parity = 0;

  // CTA b owns the ipc consecutive items starting at b*ipc: CTAs launch in index order, so the RoIs in flight
  // are neighbours in the list and the gradient lines they hit stay L2-resident (see roi_fwd_warp_kernel)
  const long long last = min(items, ((long long)blockIdx.x + 1) * ipc);
  for (long long item = (long long)blockIdx.x * ipc; item < last; ++item) {
    const int k = (int)(item / groups);
    const int cg = (int)(item - (long long)k * groups);
    const int c0 = cg * CPB;
    const int nch = min(CPB, p.C - c0);
    const RoiGeom g = roi_geom(p, k);
    if (!g.live) continue;  // block-uniform
    const LvParam& lv = p.lv[g.lvl];
    fence_proxy_async_smem();  // generic-proxy reads of the old tile are ordered before the async-proxy refill
    __syncthreads();           // everyone is done with the previous tile and tables
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(nch * PP * sizeof(float));
      mbar_expect_tx(bar, bytes);
      bulk_load_global_to_smem(tile, gout + ((size_t)k * p.C + c0) * PP, bytes, bar);
    }
    build_tables<P>(tb, g, lv, tid);
    mbar_wait(bar, parity);
    parity ^= 1u;

    const int c = c0 + 2 * tid;
    if (2 * tid < nch) {
      const float* my = tile + (size_t)(2 * tid) * PP;
      float* fb = lv.data + (size_t)g.b * lv.sn + c;
      float2 U0[P], U1[P];
      int r0 = -1, r1 = -1;  // element offsets of the two live rows
#pragma unroll
      for (int pw = 0; pw < P; ++pw) U0[pw] = U1[pw] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int ph = 0; ph < P; ++ph) {
        float2 gq[P];  // grad_out[ph][:] / count  (count = 4, exact)
#pragma unroll
        for (int pw = 0; pw < P; ++pw) gq[pw] = make_float2(my[ph * P + pw] * 0.25f, my[PP + ph * P + pw] * 0.25f);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t m = tb.ymode[2 * ph + i];
          if (m & kValid) {
            const uint32_t mode = m & kModeMask;
            const AxisSample s = tb.ys[2 * ph + i];
            if (mode == kShift) {
              if (r0 >= 0) scatter_row<P>(tb, fb + r0, U0);
#pragma unroll
              for (int pw = 0; pw < P; ++pw) {
                U0[pw] = U1[pw];
                U1[pw] = make_float2(0.f, 0.f);
              }
              r0 = r1;
              r1 = -1;
            } else if (mode == kNew) {
              if (r0 >= 0) scatter_row<P>(tb, fb + r0, U0);
              if (r1 >= 0) scatter_row<P>(tb, fb + r1, U1);
#pragma unroll
              for (int pw = 0; pw < P; ++pw) U0[pw] = U1[pw] = make_float2(0.f, 0.f);
              r1 = -1;
            }
            r0 = s.off_lo;
            if (m & kBorder) {
#pragma unroll
              for (int pw = 0; pw < P; ++pw) {
                U0[pw] = ffma2(splat(s.w_lo), gq[pw], U0[pw]);
                U0[pw] = ffma2(splat(s.w_hi), gq[pw], U0[pw]);
              }
            } else {
              r1 = s.off_hi;
#pragma unroll
              for (int pw = 0; pw < P; ++pw) {
                U0[pw] = ffma2(splat(s.w_lo), gq[pw], U0[pw]);
                U1[pw] = ffma2(splat(s.w_hi), gq[pw], U1[pw]);
              }
            }
          }
        }
      }
      if (r0 >= 0) scatter_row<P>(tb, fb + r0, U0);
      if (r1 >= 0) scatter_row<P>(tb, fb + r1, U1);
    }
  }
}

// ---- backward, warp items (P = 7 and P = 14) ------------------------------------------------------
// The transpose of roi_fwd_warp_kernel (same item shapes): the warp's grad_out tile (12.5 / 25 KB) arrives
// with one bulk load on the warp's own mbarrier; the y direction accumulates, per distinct window row,
// U[pw] = sum of w * grad_out[ph][pw] / 4 with the same two-row cache; a row that leaves the cache is
// scattered ONCE: the folded bin weights are walked with a sliding window of NB column accumulators,
// so every distinct column of the row receives exactly one vector red.global.add (lanes = consecutive
// channel pairs: a warp's reduction covers two full 128-byte lines at L2).
// One call site for the row scatter (the r01 kernel inlined it five times and was bound by instruction
// cache misses: no_instruction stall 10 per issue, ncu r01b).
__device__ __forceinline__ void red_add_f2(char* p, float2 v, bool active) {
  if (active) atomicAdd(reinterpret_cast<float2*>(p), v);
}

template <int P, int XB, int NB, int CSW>
__device__ __forceinline__ void scatter_row_warp(const WarpTables<P>& tb, int xb0, const uint32_t (&xo)[XB], const int (&xstep)[XB],
                                                 char* __restrict__ row, uint32_t swb, const float2 (&U)[XB], bool active) {
  const uint32_t cs = CSW ? (uint32_t)CSW * 4u : swb;
  if (NB > 0) {
    float2 A[NB > 0 ? NB : 1];
#pragma unroll
    for (int j = 0; j < NB; ++j) A[j] = make_float2(0.f, 0.f);
    uint32_t base = xo[0];
#pragma unroll
    for (int pw = 0; pw < XB; ++pw) {
      if (pw > 0) {
        const int d = min(xstep[pw], NB);  // columns the window slides (uniform per half-warp)
#pragma unroll 1
        for (int q = 0; q < d; ++q) {
          red_add_f2(row + base + q * cs, A[0], active);
#pragma unroll
          for (int j = 0; j + 1 < NB; ++j) A[j] = A[j + 1];
          A[NB - 1] = make_float2(0.f, 0.f);
        }
        base = xo[pw];
      }
      const float4 w = tb.xw[xb0 + pw];
      A[0] = ffma2(splat(w.x), U[pw], A[0]);
      A[1] = ffma2(splat(w.y), U[pw], A[1]);
      if (NB > 2) A[2] = ffma2(splat(w.z), U[pw], A[2]);
      if (NB > 3) A[3] = ffma2(splat(w.w), U[pw], A[3]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) red_add_f2(row + base + j * cs, A[j], active);
  } else {
#pragma unroll
    for (int t = 0; t < 2 * XB; ++t) {
      const AxisTapB s = tb.xs[2 * xb0 + t];
      if (s.w_lo + s.w_hi != 0.f) {  // uniform per half-warp
        red_add_f2(row + s.off_lo, make_float2(s.w_lo * U[t >> 1].x, s.w_lo * U[t >> 1].y), active);
        red_add_f2(row + s.off_hi, make_float2(s.w_hi * U[t >> 1].x, s.w_hi * U[t >> 1].y), active);
      }
    }
  }
}

template <int P, int XB, int NB, int CSW>
__device__ __forceinline__ void roi_warp_body_bwd(const WarpTables<P>& tb, int xb0, char* __restrict__ fb, uint32_t swb,
                                                  const float* __restrict__ my, bool active) {
  constexpr int PP = P * P;
  constexpr uint32_t kAll = 0xffffffffu;
  float2 U0[XB], U1[XB], gq[XB];
  uint32_t xo[XB];
  int xstep[XB];
#pragma unroll
  for (int pw = 0; pw < XB; ++pw) {
    U0[pw] = U1[pw] = gq[pw] = make_float2(0.f, 0.f);
    xo[pw] = NB > 0 ? tb.xoff[xb0 + pw] : 0u;
    xstep[pw] = (NB > 0 && pw > 0) ? tb.xfirst[xb0 + pw] - tb.xfirst[xb0 + pw - 1] : 0;
  }
  // P = 7: conflict-free tile reads, see the forward kernel's store order
  const bool lower = P != 7 || (threadIdx.x & 16) == 0;
  const float* const i_a = my + (lower ? 0 : PP);
  const float* const i_b = my + (lower ? PP : 0);
  const int nmine = min(XB, P - xb0);
  uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu;  // byte offsets of the two cached rows (all-ones: empty)
#pragma unroll 1
  for (int t = 0; t <= 2 * P; ++t) {  // the extra iteration flushes both rows
    const uint32_t m = t < 2 * P ? tb.ymode[t] : (kValid | kNew);
    if (!(t & 1) && t < 2 * P) {  // entering bin row t/2: grad_out[ph][:] / count (count = 4, exact)
      const int o = (t >> 1) * P;
#pragma unroll
      for (int pw = 0; pw < XB; ++pw) {
        const float e = pw < nmine ? i_a[o + pw] * 0.25f : 0.f, f = pw < nmine ? i_b[o + pw] * 0.25f : 0.f;
        gq[pw] = lower ? make_float2(e, f) : make_float2(f, e);
      }
    }
    if (__any_sync(kAll, m & kValid)) {
      const uint32_t mode = m & kModeMask;
      const int nflush = __any_sync(kAll, mode == kShift) ? 1 : (__any_sync(kAll, mode == kNew) ? 2 : 0);
#pragma unroll 1
      for (int f = 0; f < nflush; ++f) {  // the single scatter site
        if (__any_sync(kAll, r0 != 0xffffffffu)) scatter_row_warp<P, XB, NB, CSW>(tb, xb0, xo, xstep, fb + r0, swb, U0, active);
#pragma unroll
        for (int pw = 0; pw < XB; ++pw) {
          U0[pw] = U1[pw];
          U1[pw] = make_float2(0.f, 0.f);
        }
        r0 = r1;
        r1 = 0xffffffffu;
      }
      if (t < 2 * P) {
        const AxisTapB s = tb.ys[t];
        r0 = s.off_lo;
        if (__any_sync(kAll, m & kBorder)) {
#pragma unroll
          for (int pw = 0; pw < XB; ++pw) {
            U0[pw] = ffma2(splat(s.w_lo), gq[pw], U0[pw]);
            U0[pw] = ffma2(splat(s.w_hi), gq[pw], U0[pw]);
          }
        } else {
          r1 = s.off_hi;
#pragma unroll
          for (int pw = 0; pw < XB; ++pw) {
            U0[pw] = ffma2(splat(s.w_lo), gq[pw], U0[pw]);
            U1[pw] = ffma2(splat(s.w_hi), gq[pw], U1[pw]);
          }
        }
      }
    }
  }
}

template <int P, int XB, int WARPS, int CSW>
__global__ void __launch_bounds__(WARPS * 32, P == 7 ? (XB == 7 ? 4 : 6) : 2)
    roi_bwd_warp_kernel(const __grid_constant__ RoiParams p, const float* __restrict__ gout, int groups, int flags, int ipw) {
  using WI = WarpItem<P, XB>;
  constexpr int PP = P * P;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int pair = lane % WI::kPairs;
  const int xb0 = (lane / WI::kPairs) * XB;
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * WI::kTileFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats + WARPS * sizeof(WarpTables<P>)) + warp;
  const float* my = tile + (size_t)(2 * pair) * PP + xb0;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;
  const long long items = (long long)p.K * groups;
  const long long first = (long long)blockIdx.x * WARPS * ipw + warp;
  const long long last = min(items, ((long long)blockIdx.x + 1) * WARPS * ipw);
  // shared tables: as in the forward kernel (one table build per RoI when a RoI has exactly WARPS channel groups)
  __shared__ int s_run[WARPS], s_b[WARPS], s_lvl[WARPS];
  WarpTables<P>* tbs = reinterpret_cast<WarpTables<P>*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats);
  const bool shared_tables = groups == WARPS && ipw >= 1 && ipw <= WARPS && !(flags & 2);
  const int k0 = blockIdx.x * ipw;
  if (shared_tables) {
    if (warp < ipw && k0 + warp < p.K) {
      const RoiGeom g = roi_geom(p, k0 + warp);
      const bool live = __any_sync(kAll, g.live);
      int run = -1;
      if (live) {
        run = build_tables_warp<P>(tbs[warp], g, p.lv[g.lvl], lane, WI::kHalves * XB);
        bool mono = true;  // the sliding window needs non-decreasing bin starts (always true for x2 >= x1)
        if (lane > 0 && lane < P) mono = tbs[warp].xfirst[lane] >= tbs[warp].xfirst[lane - 1];
        if (!__all_sync(kAll, mono)) run = 99;  // per-sample path
      }
      if (lane == 0) {
        s_run[warp] = run;
        s_b[warp] = g.b;
        s_lvl[warp] = live ? g.lvl : 0;
      }
    }
    __syncthreads();
  }

  for (long long item = first; item < last; item += WARPS) {
    int k, cg, run, gb, glvl;
    bool live;
    const WarpTables<P>* tbp;
    if (shared_tables) {  // block-uniform
      const int slot = (int)((item - first) / WARPS);
      k = k0 + slot;
      cg = warp;
      run = s_run[slot];
      live = __any_sync(kAll, run >= 0);
      gb = s_b[slot];
      glvl = s_lvl[slot];
      tbp = &tbs[slot];
    } else {
      k = (int)(item / groups);
      cg = (int)(item - (long long)k * groups);
      run = -1;
      tbp = &tbs[warp];
    }
    RoiGeom g{};
    if (!shared_tables) {
      g = roi_geom(p, k);
      live = __any_sync(kAll, g.live);
      gb = g.b;
      glvl = live ? g.lvl : 0;
    }
    if (!live) continue;  // warp-uniform
    const WarpTables<P>& tb = *tbp;
    const int c0 = cg * WI::kChannels;
    const int nch = min(WI::kChannels, p.C - c0);
    const LvParam& lv = p.lv[glvl];
    fence_proxy_async_smem();  // generic-proxy reads of the old tile are ordered before the async-proxy refill
    __syncwarp();              // every lane is done with the previous tile
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(nch * PP * sizeof(float));
      mbar_expect_tx(bar, bytes);
      bulk_load_global_to_smem(tile, gout + ((size_t)k * p.C + c0) * PP, bytes, bar);
    }
    if (!shared_tables) {  // per-warp tables, built while the tile is in flight
      run = build_tables_warp<P>(tbs[warp], g, lv, lane, WI::kHalves * XB);
      bool mono = true;
      if (lane > 0 && lane < P) mono = tbs[warp].xfirst[lane] >= tbs[warp].xfirst[lane - 1];
      if (!__all_sync(kAll, mono)) run = 99;
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    // every lane runs the (vote-synchronised) walk; lanes past a short channel group only skip the atomics
    const bool active = 2 * pair < nch;
    char* fb = reinterpret_cast<char*>(lv.data + (size_t)gb * lv.sn + c0 + min(2 * pair, nch - 2));
    const uint32_t swb = (uint32_t)lv.sw * 4u;
    if (__all_sync(kAll, run <= 3)) roi_warp_body_bwd<P, XB, 3, CSW>(tb, xb0, fb, swb, my, active);
    else if (__all_sync(kAll, run == 4)) roi_warp_body_bwd<P, XB, 4, CSW>(tb, xb0, fb, swb, my, active);
    else roi_warp_body_bwd<P, XB, 0, CSW>(tb, xb0, fb, swb, my, active);
  }
}

// ---- forward, staged rows (P = 7, dense NHWC, C in {64, 128, 192, 256}) ------------------------------
// The north-star design: the feature window of a RoI is STAGED through shared memory by the copy engine.
//
// Why: roi_fwd_warp_kernel keeps the gather in the pooling warps — per window row a warp issues 21-28 loads, waits
// one L2 round trip (~600 cycles under load), does its FMAs, and starts over: ~15 such round trips per item plus the
// table build and the tile store, all on ONE warp's critical path, 16 warps per SM.  The L2 -> SM path itself is
// nowhere near its limit (tools/l2_probe.cu: 15-19 TB/s from L2 with >= 64 KB in flight per SM, the gather of a bench
// step is 11 GB), the chain is.  Here the chain is cut:
//
//   * In dense NHWC memory the 256 channels of `ncols` neighbouring pixels of a map row are ONE contiguous run of
//     ncols KB, so a window row is a single 1-D bulk copy (cp.async.bulk.shared.global, SASS UBLKCP — the TMA engine
//     without a tensor map; a tensor map's box is fixed per map, window widths vary per RoI).
//   * One PRODUCER warp per CTA builds the tables of the CTA's RoIs one to two RoIs ahead and issues the row copies
//     into a 56 KB byte ring (rows are 4-28 KB; allocation wraps at row boundaries), each row on its own mbarrier:
//     ~45 KB in flight per CTA, two CTAs per SM.
//   * Four CONSUMER warps (warp = 64-channel group, lane = channel pair) wait for a row, pool it in x with the same
//     bin-dense folded weights as the warp kernel — LDS.64 at ~30 cycles instead of LDG at ~600 — release it, and
//     accumulate in y with the two-row cache.  Output tile and bulk store as before (one 12.5 KB store per warp).
//   * RoIs whose bins are wider than 4 feature pixels (samples more than 2 px apart: a contiguous span would be mostly
//     unused columns) or whose span exceeds 28 columns are gathered from global memory by the consumers, as before.
//
// Arithmetic is exactly that of roi_warp_body (same folded weights, same FMA order): results are bit-identical to
// roi_fwd_warp_kernel.
constexpr int kStRing = 56 * 1024;  // default byte ring of staged rows (two CTAs per SM)
constexpr int kStBars = 16;         // rows in flight (mbarrier slots)
constexpr int kStDescs = 4;         // RoI descriptors (tables) built ahead
constexpr int kStMaxCols = 28;      // widest staged row in pixels (28 KB at C = 256: two fit in the ring)

struct __align__(16) StagedDesc {
  WarpTables<7> tb;
  int run;             // widest bin run in columns; -1: dead RoI (zero tile)
  int staged;          // 1: rows arrive through the ring, 0: consumers gather from global memory
  int b, lvl;
  uint32_t col0_bytes;  // byte offset of the first staged column inside a map row
  uint32_t row_bytes;   // staged bytes per row
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// One staged window row -> T[0..XB) = this lane's x-bins (x-pooled), then the row's ring space is handed back to the
// producer.
template <int XB, int NB, int CSW>
__device__ __forceinline__ void pool_row_staged(const unsigned char* __restrict__ ring_lane, const uint32_t (&xo)[XB],
                                                const float (&xw)[XB][NB], const uint32_t* __restrict__ row_off, uint64_t* full,
                                                uint64_t* empty, uint32_t& q, uint32_t csb, float2 (&T)[XB]) {
  const uint32_t cs = CSW ? (uint32_t)CSW * 4u : csb;
  const uint32_t i = q % kStBars;
  mbar_wait(&full[i], (q / kStBars) & 1u);
  const unsigned char* row = ring_lane + row_off[i];
  float2 v[XB][NB];
#pragma unroll
  for (int pw = 0; pw < XB; ++pw) {
#pragma unroll
    for (int j = 0; j < NB; ++j) v[pw][j] = *reinterpret_cast<const float2*>(row + xo[pw] + j * cs);
  }
#pragma unroll
  for (int pw = 0; pw < XB; ++pw) {
    float2 t = __fmul2_rn(splat(xw[pw][0]), v[pw][0]);
#pragma unroll
    for (int j = 1; j < NB; ++j) t = ffma2(splat(xw[pw][j]), v[pw][j], t);
    T[pw] = t;
  }
  __syncwarp();  // every lane's loads of this row have returned (their FMAs consumed them)
  if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[i]);
  ++q;
}

// roi_warp_body with the rows coming from the ring (P = 7; XB = 7: one bin group per warp, XB = 4: two).
template <int XB, int NB, int CSW>
__device__ __forceinline__ void roi_warp_body_staged(const WarpTables<7>& tb, int xb0, const unsigned char* __restrict__ ring_lane,
                                                     uint32_t col0_bytes, uint32_t csb, const uint32_t* __restrict__ row_off,
                                                     uint64_t* full, uint64_t* empty, uint32_t& q, float* __restrict__ my,
                                                     bool& tile_free) {
  constexpr int P = 7, PP = 49;
  constexpr uint32_t kAll = 0xffffffffu;
  float2 T0[XB], T1[XB], acc[XB];
  uint32_t xo[XB];
  float xw[XB][NB];
#pragma unroll
  for (int pw = 0; pw < XB; ++pw) {
    T0[pw] = T1[pw] = acc[pw] = make_float2(0.f, 0.f);
    xo[pw] = tb.xoff[xb0 + pw] - col0_bytes;
    const float4 w = tb.xw[xb0 + pw];
    xw[pw][0] = w.x;
    xw[pw][1] = w.y;
    xw[pw][2] = w.z;
    if (NB > 3) xw[pw][3 % NB] = w.w;
  }
  const bool lower = XB != 7 || (threadIdx.x & 16) == 0;  // XB = 7: conflict-free tile stores, see roi_warp_body
  float* const o_a = my + (lower ? 0 : PP);
  float* const o_b = my + (lower ? PP : 0);
  const int nmine = min(XB, P - xb0);
  bool flip = false;
#pragma unroll 1
  for (int t = 0; t < 2 * P; ++t) {
    const uint32_t m = tb.ymode[t];
    if (__any_sync(kAll, m & kValid)) {
      const uint32_t mode = m & kModeMask;
      const AxisTapB s = tb.ys[t];
      const float2 wl = splat(s.w_lo * 0.25f), wh = splat(s.w_hi * 0.25f);
      const bool border = __any_sync(kAll, m & kBorder);
      const bool is_new = __any_sync(kAll, mode == kNew), is_shift = __any_sync(kAll, mode == kShift);
      auto y_step = [&](auto& LO, auto& HI) -> bool {
        if (is_new || (is_shift && !border)) pool_row_staged<XB, NB, CSW>(ring_lane, xo, xw, row_off, full, empty, q, csb, LO);
        if (is_shift && !border) {
#pragma unroll
          for (int pw = 0; pw < XB; ++pw) {
            acc[pw] = ffma2(wl, HI[pw], acc[pw]);
            acc[pw] = ffma2(wh, LO[pw], acc[pw]);
          }
          return true;
        }
        if (is_shift) {
#pragma unroll
          for (int pw = 0; pw < XB; ++pw) LO[pw] = HI[pw];
        } else if (is_new) {
          if (border) {
#pragma unroll
            for (int pw = 0; pw < XB; ++pw) HI[pw] = LO[pw];
          } else {
            pool_row_staged<XB, NB, CSW>(ring_lane, xo, xw, row_off, full, empty, q, csb, HI);
          }
        }
#pragma unroll
        for (int pw = 0; pw < XB; ++pw) {
          acc[pw] = ffma2(wl, LO[pw], acc[pw]);
          acc[pw] = ffma2(wh, HI[pw], acc[pw]);
        }
        return false;
      };
      if (!flip) flip = y_step(T0, T1);
      else flip = !y_step(T1, T0);
    }
    if (t & 1) {  // bin row complete
      if (!tile_free) {  // the previous RoI's bulk store must have finished READING the tile (first write only)
        if ((threadIdx.x & 31) == 0) bulk_wait_read_all();
        __syncwarp();
        tile_free = true;
      }
      const int o = (t >> 1) * P;
#pragma unroll
      for (int pw = 0; pw < XB; ++pw) {
        const float e = acc[pw].x, f = acc[pw].y;
        if (pw < nmine) {
          o_a[o + pw] = lower ? e : f;
          o_b[o + pw] = lower ? f : e;
        }
        acc[pw] = make_float2(0.f, 0.f);
      }
    }
  }
}

// XB = 7: 4 consumer warps of 64 channels; XB = 4: 8 consumer warps of 32 channels (half-warps take bins 0-3 / 4-6).  The
// output tile belongs to the RoI (C x 49 floats), not to a warp, so more pooling warps cost no shared memory — the
// warp kernel cannot do that (one 12.5 KB tile per warp caps it at 16 warps per SM).
template <int XB, int CSW>
__global__ void __launch_bounds__((256 / WarpItem<7, XB>::kChannels + 2) * 32, 2)
    roi_fwd_staged_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int rois_per_cta, int flags,
                          uint32_t ring_bytes) {
  using WI = WarpItem<7, XB>;
  constexpr int WARPS = 256 / WI::kChannels;  // consumer warps at C = 256
  constexpr int PP = 49;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;
  float* tile_roi = reinterpret_cast<float*>(smem_raw + ring_bytes);  // [C][49] of the RoI being pooled
  StagedDesc* descs = reinterpret_cast<StagedDesc*>(smem_raw + ring_bytes + sizeof(float) * 256 * PP);
  uint64_t* full = reinterpret_cast<uint64_t*>(descs + kStDescs);
  uint64_t* empty = full + kStBars;
  uint64_t* dfull = empty + kStBars;
  uint64_t* dempty = dfull + kStDescs;
  uint32_t* row_off = reinterpret_cast<uint32_t*>(dempty + kStDescs);
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int ncw = p.C / WI::kChannels;  // active consumer warps (C is a multiple of 64, at most 256)
  const uint32_t csb = (uint32_t)p.C * 4u;
  const int k_begin = blockIdx.x * rois_per_cta;
  const int k_end = min(p.K, k_begin + rois_per_cta);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStBars; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], (uint32_t)ncw);
    }
    for (int i = 0; i < kStDescs; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], (uint32_t)ncw + 1u);  // the pooling warps and the row issuer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == WARPS) {
    // ---------------- table builder: geometry, taps and folded weights of the CTA's RoIs, kStDescs - 1 RoIs ahead ----------
    for (int k = k_begin; k < k_end; ++k) {
      const int r = k - k_begin, d = r % kStDescs;
      if (r >= kStDescs) mbar_wait(&dempty[d], (uint32_t)((r / kStDescs) - 1) & 1u);  // consumers and issuer are done with the slot
      StagedDesc& ds = descs[d];
      const RoiGeom g = roi_geom(p, k);
      const bool live = __any_sync(kAll, g.live);
      const int lvl = live ? g.lvl : 0;
      const LvParam& lv = p.lv[lvl];
      int run = -1, staged = 0;
      uint32_t col0b = 0, rowb = 0;
      if (live) {
        run = build_tables_warp<7>(ds.tb, g, lv, lane, WI::kHalves * XB);
        if (run <= 4) {
          const int nb = run <= 3 ? 3 : 4;
          int f = lane < 7 ? ds.tb.xfirst[lane] : 0x7fffffff;
          int l = lane < 7 ? ds.tb.xfirst[lane] + nb - 1 : -1;
          f = __reduce_min_sync(kAll, f);
          l = __reduce_max_sync(kAll, l);
          const int ncols = l - f + 1;
          if (ncols <= kStMaxCols && (uint32_t)ncols * csb * 2u <= ring_bytes && !(flags & 4)) {
            staged = 1;
            col0b = (uint32_t)f * csb;
            rowb = (uint32_t)ncols * csb;
          }
        }
      }
      if (lane == 0) {
        ds.run = run;
        ds.staged = staged;
        ds.b = g.b;
        ds.lvl = lvl;
        ds.col0_bytes = col0b;
        ds.row_bytes = rowb;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dfull[d]);  // release: tables and header are visible to whoever observes the phase
    }
  } else if (warp == WARPS + 1) {
    // ---------------- row issuer (one thread): the RoIs' window rows into the ring, in consumption order ----------------
    if (lane == 0) {
      uint32_t q = 0, tail_q = 0, head = 0, tail_off = 0;
      int inflight = 0;
      bool wrapped = false;
      for (int k = k_begin; k < k_end; ++k) {
        const int r = k - k_begin, d = r % kStDescs;
        mbar_wait(&dfull[d], (uint32_t)(r / kStDescs) & 1u);
        const StagedDesc& ds = descs[d];
        if (ds.staged) {
          const LvParam& lv = p.lv[ds.lvl];
          const uint32_t rowb = ds.row_bytes;
          const char* fbase = reinterpret_cast<const char*>(lv.data + (size_t)ds.b * lv.sn) + ds.col0_bytes;
          auto issue = [&](uint32_t row_byte_off) {
            uint32_t place = 0;
            for (;;) {
              bool ok = false, wrap_now = false;
              if (inflight == 0) {
                head = 0;
                wrapped = false;
                place = 0;
                ok = true;
              } else if (!wrapped) {
                if (head + rowb <= ring_bytes) {
                  place = head;
                  ok = true;
                } else if (rowb <= tail_off) {
                  place = 0;
                  wrap_now = true;
                  ok = true;
                }
              } else if (head + rowb <= tail_off) {
                place = head;
                ok = true;
              }
              if (ok && inflight < kStBars) {
                if (wrap_now) wrapped = true;
                break;
              }
              // hand the oldest row's space back (rows are released in issue order)
              mbar_wait(&empty[tail_q % kStBars], (tail_q / kStBars) & 1u);
              ++tail_q;
              --inflight;
              if (inflight > 0) {
                const uint32_t nt = row_off[tail_q % kStBars];
                if (nt < tail_off) wrapped = false;  // the tail itself wrapped to the front
                tail_off = nt;
              }
            }
            const uint32_t i = q % kStBars;
            row_off[i] = place;
            if (inflight == 0) tail_off = place;
            mbar_expect_tx(&full[i], rowb);
            bulk_load_global_to_smem(ring + place, fbase + row_byte_off, rowb, &full[i]);
            head = place + rowb;
            ++q;
            ++inflight;
          };
#pragma unroll 1
          for (int t = 0; t < 14; ++t) {
            const uint32_t m = ds.tb.ymode[t];
            if (!(m & kValid)) continue;
            const uint32_t mode = m & kModeMask;
            const bool border = (m & kBorder) != 0;
            const AxisTapB s = ds.tb.ys[t];
            if (mode == kNew) {
              issue(s.off_lo);
              if (!border) issue(s.off_hi);
            } else if (mode == kShift && !border) {
              issue(s.off_hi);
            }
          }
        }
        mbar_arrive(&dempty[d]);
      }
    }
  } else if (warp < ncw) {
    // ---------------- consumers: warp = channel group of every RoI of the CTA ----------------
    const int pair = lane % WI::kPairs;
    const int xb0 = (lane / WI::kPairs) * XB;
    const int c0 = warp * WI::kChannels;
    float* tile = tile_roi + (size_t)c0 * PP;                     // this warp's slice of the RoI tile
    float* my = tile + (size_t)(2 * pair) * PP + xb0;
    const unsigned char* ring_lane = ring + (size_t)(c0 + 2 * pair) * 4;
    const uint64_t pol = l2_policy_evict_first();
    uint32_t q = 0;
    bool tile_free = true;
    for (int k = k_begin; k < k_end; ++k) {
      const int r = k - k_begin, d = r % kStDescs;
      mbar_wait(&dfull[d], (uint32_t)(r / kStDescs) & 1u);
      const StagedDesc& ds = descs[d];
      const int run = ds.run;
      if (run >= 0 && ds.staged) {
        if (run <= 3) roi_warp_body_staged<XB, 3, CSW>(ds.tb, xb0, ring_lane, ds.col0_bytes, csb, row_off, full, empty, q, my, tile_free);
        else roi_warp_body_staged<XB, 4, CSW>(ds.tb, xb0, ring_lane, ds.col0_bytes, csb, row_off, full, empty, q, my, tile_free);
      } else {
        if (!tile_free) {
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
          tile_free = true;
        }
        if (run < 0) {
          for (int j = lane; j < WI::kTileFloats; j += 32) tile[j] = 0.f;
        } else {
          const LvParam& lv = p.lv[ds.lvl];
          const char* fb = reinterpret_cast<const char*>(lv.data + (size_t)ds.b * lv.sn + c0 + 2 * pair);
          const uint32_t swb = (uint32_t)lv.sw * 4u;
          if (run <= 3) roi_warp_body<7, XB, 3, CSW>(ds.tb, xb0, fb, swb, my);
          else if (run == 4) roi_warp_body<7, XB, 4, CSW>(ds.tb, xb0, fb, swb, my);
          else roi_warp_body<7, XB, 0, CSW>(ds.tb, xb0, fb, swb, my);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();  // the warp's slice is complete and every lane is done with the descriptor
      if (lane == 0) {
        float* dst = out + ((size_t)k * p.C + c0) * PP;
        const uint32_t bytes = (uint32_t)(WI::kChannels * PP * sizeof(float));
        if (flags & 1) bulk_store_smem_to_global_hint(dst, tile, bytes, pol);
        else bulk_store_smem_to_global(dst, tile, bytes);
        bulk_commit();
        mbar_arrive(&dempty[d]);
      }
      tile_free = false;
    }
    if (lane == 0) bulk_wait_read_all();
  }
}

// ---- forward, persistent teams (P = 7, C == 256): LCR_ROI_FWD=team --------------------------------------------------
// ncu of the launch-ordered kernels (profiles/r02b_*): a CTA lives for two RoIs; its table build (RoI load from DRAM, taps,
// fold, row program: ~650 instructions on ONE warp's dependent chain) sits in front of a __syncthreads, so all four warps
// spend 7-20 % of the CTA's life in it or waiting for it, then the CTA drains and the slot waits for the next launch.
// Here ONE persistent CTA of 16 warps per SM (same 16 pooling warps, same 12.5 KB tile each; the per-CTA reserved shared
// memory of four CTAs becomes room for 16 table slots) is split into four TEAMS of four warps (warp = 64-channel group).
// A team pulls RoIs from a global counter IN ORDER (the RoIs in flight on the GPU stay ~600 consecutive list entries: the
// L2 footprint of the launch-ordered grid, halved) through a ring of four table slots: warp w builds every fourth RoI's
// tables TWO RoIs ahead of where it pools, hands them over on an mbarrier (full), and reuses the slot when all four warps
// have released it (empty).  No block barrier: a warp that is building does not stop the other three, and the build costs
// each warp a quarter of a build per item instead of a whole one per two items in front of a barrier.
// Marks a RoI the persistent kernel left to the second pass: an fp32 quiet-NaN bit pattern in the first word of the RoI's
// output block.  The second pass recomputes every RoI whose first word holds it, so a genuine output that happened to carry
// the same bits (a NaN with this payload) is merely pooled twice, with the same result.
constexpr unsigned int kRestSentinel = 0x7fd5c3a9u;

// Second pass of the two-pass forward: the RoIs the persistent kernel marked, through the sample-walk bodies of
// roi_fwd_warp_kernel.  CTA b scans the first output words of RoIs [b*R, (b+1)*R) (one per thread), compacts the marked ones
// and pools them four at a time (warp = 64-channel group, warp i builds the tables of the i-th RoI of the group).
template <int CSW>
__global__ void __launch_bounds__(128, 4)
    roi_fwd_rest_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int flags, int R,
                        const unsigned int* __restrict__ n_marked) {
  if (__ldcg(n_marked) == 0u) return;  // nothing was left over: no scan (block-uniform)
  using WI = WarpItem<7, 7>;
  constexpr int WARPS = 4, PP = 49;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * WI::kTileFloats;
  WarpTables<7>* tbs = reinterpret_cast<WarpTables<7>*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats);
  __shared__ int s_run[WARPS], s_b[WARPS], s_lvl[WARPS];
  __shared__ int s_cnt[WARPS + 1];
  __shared__ int s_list[128];
  const int k_mine = blockIdx.x * R + (int)threadIdx.x;
  bool marked = false;
  if ((int)threadIdx.x < R && k_mine < p.K)
    marked = __ldcg(reinterpret_cast<const unsigned int*>(out + (size_t)k_mine * p.C * PP)) == kRestSentinel;
  const uint32_t bal = __ballot_sync(kAll, marked);
  if (lane == 0) s_cnt[warp + 1] = __popc(bal);
  if (threadIdx.x == 0) s_cnt[0] = 0;
  __syncthreads();
  int before = 0;
  for (int q = 0; q <= warp; ++q) before += s_cnt[q];
  const int n = s_cnt[1] + s_cnt[2] + s_cnt[3] + s_cnt[4];
  if (n == 0) return;  // block-uniform
  if (marked) s_list[before + __popc(bal & ((1u << lane) - 1u))] = k_mine;
  float* my = tile + (size_t)(2 * lane) * PP;
  const uint64_t pol = l2_policy_evict_first();
  const int c0 = warp * WI::kChannels;
  constexpr int G = 2;  // RoIs per group
  for (int g0 = 0; g0 < n; g0 += G) {
    __syncthreads();  // the list is complete / the previous group's tables are no longer read
    if (warp < G && g0 + warp < n) {
      const RoiGeom g = roi_geom(p, s_list[g0 + warp]);
      const bool live = __any_sync(kAll, g.live);
      int run = -1;
      if (live) run = build_tables_warp<7>(tbs[warp], g, p.lv[g.lvl], lane, 7);
      if (lane == 0) {
        s_run[warp] = run;
        s_b[warp] = g.b;
        s_lvl[warp] = live ? g.lvl : 0;
      }
    }
    __syncthreads();
    for (int slot = 0; slot < G && g0 + slot < n; ++slot) {
      const int k = s_list[g0 + slot];
      const int run = s_run[slot];
      const WarpTables<7>& tb = tbs[slot];
      const LvParam& lv = p.lv[s_lvl[slot]];
      if (lane == 0) bulk_wait_read_all();  // the previous item's bulk store has finished reading the tile
      __syncwarp();
      if (run < 0) {
        for (int q = lane; q < WI::kTileFloats; q += 32) tile[q] = 0.f;
      } else {
        const char* fb = reinterpret_cast<const char*>(lv.data + (size_t)s_b[slot] * lv.sn + c0 + 2 * lane);
        const uint32_t swb = (uint32_t)lv.sw * 4u;
        if (run <= 3) roi_warp_body<7, 7, 3, CSW>(tb, 0, fb, swb, my);
        else if (run == 4) roi_warp_body<7, 7, 4, CSW>(tb, 0, fb, swb, my);
        else roi_warp_body<7, 7, 0, CSW>(tb, 0, fb, swb, my);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        float* dst = out + ((size_t)k * p.C + c0) * PP;
        const uint32_t bytes = (uint32_t)(WI::kChannels * PP * sizeof(float));
        if (flags & 1) bulk_store_smem_to_global_hint(dst, tile, bytes, pol);
        else bulk_store_smem_to_global(dst, tile, bytes);
        bulk_commit();
      }
    }
  }
  if (lane == 0) bulk_wait_read_all();
}

constexpr int kTeamSlots = 4;
constexpr int kClaimSlots = 1024;  // a CUDA graph keeps the slot it captured: two graphs replayed at the same time must not share one
__device__ unsigned int g_roi_claim[kClaimSlots][2];  // [work counter, RoIs left to the second pass]

struct TeamMeta {
  int k, run, b, lvl, nrows, span;
};

template <int CSW>
__global__ void __launch_bounds__(512, 1)
    roi_fwd_team_kernel(const __grid_constant__ RoiParams p, float* __restrict__ out, int flags, unsigned int* __restrict__ claim) {
  using WI = WarpItem<7, 7>;
  constexpr int TEAMS = 4, WARPS = 16, PP = 49;
  constexpr uint32_t kAll = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(kAll, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int team = warp >> 2, w = warp & 3;
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * WI::kTileFloats;
  WarpTables<7>* tbs = reinterpret_cast<WarpTables<7>*>(smem_raw + sizeof(float) * WARPS * WI::kTileFloats) + team * kTeamSlots;
  __shared__ TeamMeta s_meta[TEAMS][kTeamSlots];
  __shared__ __align__(8) uint64_t s_full[TEAMS][kTeamSlots], s_empty[TEAMS][kTeamSlots];
  if (threadIdx.x < TEAMS * kTeamSlots) {
    mbar_init(&s_full[0][0] + threadIdx.x, 1);
    mbar_init(&s_empty[0][0] + threadIdx.x, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float* my = tile + (size_t)(2 * lane) * PP;
  const uint64_t pol = l2_policy_evict_first();
  const int c0 = w * WI::kChannels;
  bool store_pending = false;
  int exhausted = 0;
#pragma unroll 1
  for (int i = -2;; ++i) {
    const int j = i + 2;
    if ((j & 3) == w) {  // this warp builds the team's RoI number j into slot w
      if (j >= 4) mbar_wait(&s_empty[team][w], (uint32_t)(((j >> 2) - 1) & 1));
      unsigned int kc = 0;
      if (lane == 0) kc = atomicAdd(claim, 1u);
      const int k = (int)min(__shfl_sync(kAll, kc, 0), (unsigned int)p.K);
      int run = -1, nrows = -1, span = 2, gb = 0, glvl = 0;
      if (k < p.K) {
        const RoiGeom g = roi_geom(p, k);
        const bool live = __any_sync(kAll, g.live);
        gb = g.b;
        if (live) {
          glvl = g.lvl;
          const LvParam& lv = p.lv[g.lvl];
          if (row_program_hopeless(g, lv)) {
            run = 99;  // no tables: the second pass builds its own
          } else {
            run = build_tables_warp<7>(tbs[w], g, lv, lane, 7);
            if (run <= 4) nrows = build_row_program(tbs[w], (uint32_t)lv.sh * 4u, lane, &span);
          }
        }
      }
      if (lane == 0) s_meta[team][w] = TeamMeta{k, run, gb, glvl, nrows, span};
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[team][w]);  // release: tables and header are visible to whoever observes the phase
    }
    if (i < 0) continue;
    const int s = i & 3;
    mbar_wait(&s_full[team][s], (uint32_t)((i >> 2) & 1));
    const TeamMeta m = s_meta[team][s];
    if (m.k >= p.K) {
      // Past the end of the list.  The four builders claim concurrently, so the team's claims are not ordered ACROSS warps
      // (RoI #i+1 may hold a smaller list index than RoI #i): one exhausted header does not end the team.  Each warp's own
      // claims do increase, so four exhausted headers in a row (one from every builder) do.  (Team-uniform: every warp
      // reads the same headers.)
      if (++exhausted == kTeamSlots) break;
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[team][s]);
      continue;
    }
    exhausted = 0;
    const WarpTables<7>& tb = tbs[s];
    const LvParam& lv = p.lv[m.lvl];
    const char* fb = reinterpret_cast<const char*>(lv.data + (size_t)m.b * lv.sn + c0 + 2 * lane);
    const uint32_t swb = (uint32_t)lv.sw * 4u;
    // Row-program RoIs only.  16 warps of one CTA run their bodies side by side: with the sample-walk bodies inlined here
    // too, a list that mixes small and large RoIs spent 56 % of its warp samples waiting for instructions (ncu
    // stall_no_inst; 0.75 ms against 0.46 ms for the same list sorted by size).  A RoI without a row program (bins wider
    // than four columns, windows taller than 32 rows, bins under ~0.7 px) is LEFT to the second pass
    // (roi_fwd_rest_kernel): the first word of its output block receives kRestSentinel.
    const bool two = m.span <= 1;  // every window row feeds at most two bin rows
    if (m.nrows > 0 && m.run <= 3 && two) {
      roi_warp_body_rmp<3, 2, CSW, 0>(tb, m.nrows, fb, swb, my, store_pending);
    } else if (m.nrows > 0 && m.run <= 3) {
      roi_warp_body_rmp<3, 3, CSW, 0>(tb, m.nrows, fb, swb, my, store_pending);
    } else if (m.nrows > 0 && two) {
      roi_warp_body_rmp<4, 2, CSW, 0>(tb, m.nrows, fb, swb, my, store_pending);
    } else if (m.run < 0 || m.nrows == 0) {  // padding row, or every sample outside the map: zeros
      if (store_pending) {
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        store_pending = false;
      }
      for (int q = lane; q < WI::kTileFloats; q += 32) tile[q] = 0.f;
    } else {
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[team][s]);
        if (w == 0) {
          *reinterpret_cast<unsigned int*>(out + (size_t)m.k * p.C * PP) = kRestSentinel;
          atomicAdd(claim + 1, 1u);
        }
      }
      continue;
    }
    fence_proxy_async_smem();
    __syncwarp();  // the tile is complete and every lane is done with the table slot
    if (lane == 0) {
      mbar_arrive(&s_empty[team][s]);
      float* dst = out + ((size_t)m.k * p.C + c0) * PP;
      const uint32_t bytes = (uint32_t)(WI::kChannels * PP * sizeof(float));
      if (flags & 1) bulk_store_smem_to_global_hint(dst, tile, bytes, pol);
      else bulk_store_smem_to_global(dst, tile, bytes);
      bulk_commit();
    }
    store_pending = true;
  }
  if (lane == 0) bulk_wait_read_all();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int fill_params(RoiParams& p, const LcrFeatLevel* lv, int L, int C, const float* rois, const int* roi_level, int K,
                       int PH, int PW, int sr, int aligned) {
  LCR_REQUIRE(lv && L > 0 && C > 0 && K >= 0 && PH > 0 && PW > 0 && sr >= 0, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(L <= LCR_MAX_LEVELS, LCR_ERR_CAPACITY);
  LCR_REQUIRE(K == 0 || rois, LCR_ERR_INVALID_ARG);
  for (int l = 0; l < L; ++l) {
    LCR_REQUIRE(lv[l].data && lv[l].N > 0 && lv[l].H > 0 && lv[l].W > 0, LCR_ERR_INVALID_ARG);
    p.lv[l] = LvParam{lv[l].data, lv[l].N, lv[l].H, lv[l].W, lv[l].sn, lv[l].sc, lv[l].sh, lv[l].sw, lv[l].spatial_scale};
  }
  p.rois = rois;
  p.roi_level = roi_level;
  p.L = L; p.C = C; p.K = K; p.PH = PH; p.PW = PW; p.sr = sr;
  p.aligned = aligned & 1;            // LCR_ROI_ALIGNED
  p.cpu_coords = (aligned >> 1) & 1;  // LCR_ROI_CPU_COORDS
  return LCR_OK;
}

// NHWC fast path eligibility: sampling_ratio 2, P in {7, 14}, channel stride 1, 8-byte aligned
// channel pairs, offsets that fit int32, C % 4 == 0 (16-byte bulk-copy granularity of the tile).
static bool fast_eligible(const RoiParams& p, const void* out) {
  if (p.sr != 2 || p.PH != p.PW || (p.PH != 7 && p.PH != 14)) return false;
  if (p.C % 4 != 0 || !aligned_to(out, 16)) return false;
  for (int l = 0; l < p.L; ++l) {
    const LvParam& v = p.lv[l];
    if (v.sc != 1 || v.sw % 2 || v.sh % 2 || v.sn % 2 || !aligned_to(v.data, 8)) return false;
    if ((long long)v.H * v.sh >= (1ll << 30) || (long long)v.W * v.sw >= (1ll << 30)) return false;  // 32-bit byte offsets
  }
  return true;
}

template <int P, int NT>
static size_t fast_smem_bytes() {
  return sizeof(float) * 2 * NT * P * P + sizeof(RoiTables<P>) + 16;
}

template <int P, int NT, bool BWD>
static int launch_fast(const RoiParams& p, float* out_or_gout, cudaStream_t st) {
  const int groups = (p.C + 2 * NT - 1) / (2 * NT);
  const size_t smem = fast_smem_bytes<P, NT>();
  const long long items = (long long)p.K * groups;
  const int per_sm = (int)((227 * 1024) / (smem + 1024));
  const long long max_blocks = (long long)sm_count() * (per_sm > 0 ? per_sm : 1);
  const int blocks = (int)(items < max_blocks ? items : max_blocks);
  static thread_local int configured_dev = -1;  // opt in to > 48 KB dynamic shared memory once per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (BWD) {
    auto kern = roi_bwd_nhwc_kernel<P, NT>;
    if (configured_dev != dev) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e);
      configured_dev = dev;
    }
    const int ipc = 2;
    const long long nb = (items + ipc - 1) / ipc;
    LCR_REQUIRE(nb < (1ll << 31), LCR_ERR_CAPACITY);
    kern<<<(unsigned)nb, NT, smem, st>>>(p, out_or_gout, groups, ipc);
  } else {
    auto kern = roi_fwd_nhwc_kernel<P, NT>;
    if (configured_dev != dev) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e);
      configured_dev = dev;
    }
    kern<<<blocks, NT, smem, st>>>(p, out_or_gout, groups);
  }
  return after_launch();
}

// Tuning switch read per call (A/B runs): LCR_ROI_FWD = "cta" selects the per-CTA kernel,
// LCR_ROI_STREAM_OUT = "0" drops the evict-first hint of the output stores.

template <int P, int XB, int CSW>
static int launch_fwd_warp(const RoiParams& p, float* out, cudaStream_t st) {
  using WI = WarpItem<P, XB>;
  constexpr int WARPS = 4;
  const int groups = (p.C + WI::kChannels - 1) / WI::kChannels;
  const size_t smem = sizeof(float) * WARPS * WI::kTileFloats + WARPS * sizeof(WarpTables<P>);
  const long long items = (long long)p.K * groups;
  int ipw = items >= (long long)WARPS * 2 * 4 * sm_count() ? 2 : 1;  // items per warp per CTA (small K: spread over all SMs)
  if (const char* v = tune_get("LCR_ROI_IPW")) ipw = atoi(v);
  long long want = (items + WARPS - 1) / WARPS;
  if (ipw > 0) {
    want = (items + (long long)WARPS * ipw - 1) / ((long long)WARPS * ipw);
  } else {
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    const long long max_blocks = (long long)sm_count() * (per_sm > 4 ? 4 : (per_sm > 0 ? per_sm : 1));
    want = want < max_blocks ? want : max_blocks;
  }
  LCR_REQUIRE(want < (1ll << 31), LCR_ERR_CAPACITY);
  const int blocks = (int)want;
  auto kern = roi_fwd_warp_kernel<P, XB, WARPS, CSW>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
  }
  // flags: bit 0 = evict-first output stores, bit 1 = per-warp tables even when a RoI has exactly WARPS channel groups
  const int flags = (tune_is("LCR_ROI_STREAM_OUT", "0") ? 0 : 1) | (tune_is("LCR_ROI_SHARED_TABLES", "0") ? 2 : 0);
  kern<<<blocks, WARPS * 32, smem, st>>>(p, out, groups, flags, ipw);
  return after_launch();
}

template <int P, int XB, int CSW>
static int launch_bwd_warp(const RoiParams& p, const float* gout, cudaStream_t st) {
  using WI = WarpItem<P, XB>;
  constexpr int WARPS = 4;
  const int groups = (p.C + WI::kChannels - 1) / WI::kChannels;
  const size_t smem = sizeof(float) * WARPS * WI::kTileFloats + WARPS * sizeof(WarpTables<P>) + WARPS * 8;
  const long long items = (long long)p.K * groups;
  const int ipw = items >= (long long)WARPS * 2 * 4 * sm_count() ? 2 : 1;
  const long long want = (items + (long long)WARPS * ipw - 1) / ((long long)WARPS * ipw);
  LCR_REQUIRE(want < (1ll << 31), LCR_ERR_CAPACITY);
  auto kern = roi_bwd_warp_kernel<P, XB, WARPS, CSW>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
  }
  kern<<<(unsigned)want, WARPS * 32, smem, st>>>(p, gout, groups, tune_is("LCR_ROI_SHARED_TABLES", "0") ? 2 : 0, ipw);
  return after_launch();
}

// warp-item kernel: additionally every map at least 4 columns wide
static bool warp_eligible(const RoiParams& p) {
  for (int l = 0; l < p.L; ++l)
    if (p.lv[l].W < 4) return false;
  return true;
}

static bool all_sw_equal(const RoiParams& p, long long v) {
  for (int l = 0; l < p.L; ++l)
    if (p.lv[l].sw != v) return false;
  return true;
}


// staged forward: P = 7, every level dense NHWC (sc == 1, sw == C), C a multiple of 64 up to 256, 16-byte aligned maps
static bool staged_eligible(const RoiParams& p) {
  if (p.PH != 7 || p.C % 64 != 0 || p.C > 256) return false;
  for (int l = 0; l < p.L; ++l) {
    const LvParam& v = p.lv[l];
    if (v.sc != 1 || v.sw != p.C || v.sh != (long long)v.W * p.C || v.sn % 4 != 0 || !aligned_to(v.data, 16)) return false;
    if (v.W < 4) return false;
  }
  return true;
}

template <int XB, int CSW>
static int launch_fwd_staged(const RoiParams& p, float* out, cudaStream_t st, int flags) {
  using WI = WarpItem<7, XB>;
  constexpr int WARPS = 256 / WI::kChannels;
  size_t ring = kStRing;
  if (const char* v = tune_get("LCR_ROI_RING_KB")) ring = (size_t)(atoi(v) > 0 ? atoi(v) : 56) * 1024;  // > 56: one CTA per SM
  const size_t fixed = sizeof(float) * 256 * 49 + kStDescs * sizeof(StagedDesc) + (2 * kStBars + 2 * kStDescs) * sizeof(uint64_t) +
                       kStBars * sizeof(uint32_t);
  if (ring + fixed > 227 * 1024) ring = (227 * 1024 - fixed) / 1024 * 1024;
  const size_t smem = ring + fixed;
  // RoIs per CTA: long enough to amortise the pipeline fill, short enough that small K still covers every SM twice
  int rpc = p.K / (2 * sm_count());
  rpc = rpc < 1 ? 1 : (rpc > 8 ? 8 : rpc);
  if (const char* v = tune_get("LCR_ROI_RPC")) rpc = atoi(v) > 0 ? atoi(v) : rpc;
  const long long want = ((long long)p.K + rpc - 1) / rpc;
  LCR_REQUIRE(want < (1ll << 31), LCR_ERR_CAPACITY);
  auto kern = roi_fwd_staged_kernel<XB, CSW>;
  static thread_local int configured_dev = -1;
  static thread_local size_t configured_smem = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev || configured_smem != smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
    configured_smem = smem;
  }
  kern<<<(unsigned)want, (WARPS + 2) * 32, smem, st>>>(p, out, rpc, flags, (uint32_t)ring);
  return after_launch();
}

// row-major forward: P = 7, C == 256 (a RoI = the CTA's four 64-channel warps), NHWC with 16-byte aligned rows
static bool rm_eligible(const RoiParams& p) {
  if (p.PH != 7 || p.C != 256) return false;
  for (int l = 0; l < p.L; ++l) {
    const LvParam& v = p.lv[l];
    if (v.W < 4 || v.sh % 4 != 0 || (long long)v.H * v.sh * 4 >= (1ll << 31)) return false;
  }
  return true;
}

template <int ROWS, int CSW>
static int launch_fwd_rm(const RoiParams& p, float* out, cudaStream_t st) {
  using WI = WarpItem<7, 7>;
  constexpr int WARPS = 4;
  const size_t smem = sizeof(float) * WARPS * WI::kTileFloats + WARPS * sizeof(WarpTables<7>);
  int ipw = p.K >= 2 * 4 * sm_count() ? 2 : 1;
  if (const char* v = tune_get("LCR_ROI_IPW")) ipw = atoi(v) >= 1 && atoi(v) <= WARPS ? atoi(v) : ipw;
  const long long want = ((long long)p.K + ipw - 1) / ipw;
  LCR_REQUIRE(want < (1ll << 31), LCR_ERR_CAPACITY);
  auto kern = roi_fwd_rm_kernel<ROWS, CSW>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
  }
  kern<<<(unsigned)want, WARPS * 32, smem, st>>>(p, out, tune_is("LCR_ROI_STREAM_OUT", "0") ? 0 : 1, ipw);
  return after_launch();
}

template <int CSW>
static int launch_fwd_rmp(const RoiParams& p, float* out, cudaStream_t st) {
  using WI = WarpItem<7, 7>;
  constexpr int WARPS = 4;
  const size_t smem = sizeof(float) * WARPS * WI::kTileFloats + WARPS * sizeof(WarpTables<7>);
  int ipw = p.K >= 2 * 4 * sm_count() ? 2 : 1;  // RoIs per CTA (1 or 2: two table slots + two row-program slots)
  if (const char* v = tune_get("LCR_ROI_IPW")) ipw = atoi(v) == 1 ? 1 : (atoi(v) == 2 ? 2 : ipw);
  const long long want = ((long long)p.K + ipw - 1) / ipw;
  LCR_REQUIRE(want < (1ll << 31), LCR_ERR_CAPACITY);
  auto kern = roi_fwd_rmp_kernel<CSW>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
  }
  const int flags = tune_is("LCR_ROI_STREAM_OUT", "0") ? 0 : 1;  // bit 0 = evict-first output stores
  kern<<<(unsigned)want, WARPS * 32, smem, st>>>(p, out, flags, ipw);
  return after_launch();
}

// persistent-team forward: rm_eligible + a list long enough to keep every team busy
static int launch_fwd_team_impl(const RoiParams& p, float* out, cudaStream_t st, bool csw256) {
  using WI = WarpItem<7, 7>;
  constexpr int WARPS = 16;
  const size_t smem = sizeof(float) * WARPS * WI::kTileFloats + 4 * kTeamSlots * sizeof(WarpTables<7>);
  static thread_local int configured_dev = -1;
  static thread_local unsigned int* claim_base = nullptr;
  static std::atomic<unsigned int> next_claim{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(roi_fwd_team_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(roi_fwd_team_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaGetSymbolAddress(reinterpret_cast<void**>(&claim_base), g_roi_claim);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
  }
  // one work counter per launch (launches in flight at the same time on different streams must not share one)
  unsigned int* claim = claim_base + 2 * (next_claim.fetch_add(1, std::memory_order_relaxed) % kClaimSlots);
  cudaError_t e = cudaMemsetAsync(claim, 0, 2 * sizeof(unsigned int), st);
  if (e != cudaSuccess) return cuda_status(e);
  const int flags = tune_is("LCR_ROI_STREAM_OUT", "0") ? 0 : 1;  // bit 0 = evict-first output stores
  const int grid = sm_count();
  if (csw256) roi_fwd_team_kernel<256><<<grid, WARPS * 32, smem, st>>>(p, out, flags, claim);
  else roi_fwd_team_kernel<0><<<grid, WARPS * 32, smem, st>>>(p, out, flags, claim);
  int rc = after_launch();
  if (rc != LCR_OK || tune_is("LCR_ROI_TEAM_PASS2", "0")) return rc;  // ("0": A/B timing of the first pass alone)
  // second pass: the RoIs the first left marked.  R RoIs per CTA: few enough CTAs that an all-row-program list costs one
  // short wave of scans, enough of them that a list of large RoIs still fills the GPU.
  constexpr int RW = 4;
  const size_t smem2 = sizeof(float) * RW * WI::kTileFloats + RW * sizeof(WarpTables<7>);
  static thread_local int configured2 = -1;
  if (configured2 != dev) {
    cudaError_t e2 = cudaFuncSetAttribute(roi_fwd_rest_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(roi_fwd_rest_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e2 != cudaSuccess) return cuda_status(e2);
    configured2 = dev;
  }
  int R = (p.K + 4 * 4 * grid - 1) / (4 * 4 * grid);
  R = R < 4 ? 4 : (R > 128 ? 128 : (R + 3) / 4 * 4);
  const int blocks = (p.K + R - 1) / R;
  if (csw256) roi_fwd_rest_kernel<256><<<blocks, RW * 32, smem2, st>>>(p, out, flags & 1, R, claim + 1);
  else roi_fwd_rest_kernel<0><<<blocks, RW * 32, smem2, st>>>(p, out, flags & 1, R, claim + 1);
  return after_launch();
}
}  // namespace lcr

using namespace lcr;

extern "C" int lcr_roi_align_fwd_f32(const LcrFeatLevel* levels_host, int L, int C, const float* rois, const int* roi_level,
                                     int K, int PH, int PW, int sampling_ratio, int aligned, float* out, void* stream) {
  RoiParams p{};
  int rc = fill_params(p, levels_host, L, C, rois, roi_level, K, PH, PW, sampling_ratio, aligned);
  if (rc != LCR_OK) return rc;
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(out, LCR_ERR_INVALID_ARG);
  cudaStream_t st = as_stream(stream);
  if (fast_eligible(p, out)) {
    // roi_fwd_staged_kernel (the north star's "window staged through TMA into shared memory") is complete, bit-identical and
    // covered by the GPU tests, but measured SLOWER than the warp kernel on B200 (bench list 1.60 vs 1.46 ms,
    // profiles/r02_roi_staged_ab.jsonl): staging moves every window byte through the SM's 128 B/clk shared-memory data path
    // twice (copy-engine fill + LDS) where a global load passes once.  It is therefore opt-in: LCR_ROI_FWD=staged.
    if (PH == 7 && staged_eligible(p) && (tune_is("LCR_ROI_FWD", "staged") || tune_is("LCR_ROI_FWD", "staged_direct"))) {
      // flags: bit 0 = evict-first output stores, bit 2 = stage nothing (every RoI gathered by the consumers: A/B)
      const int flags = (tune_is("LCR_ROI_STREAM_OUT", "0") ? 0 : 1) | (tune_is("LCR_ROI_FWD", "staged_direct") ? 4 : 0);
      if (tune_is("LCR_ROI_STAGED_WARPS", "4")) return p.C == 256 ? launch_fwd_staged<7, 256>(p, out, st, flags) : launch_fwd_staged<7, 0>(p, out, st, flags);
      return p.C == 256 ? launch_fwd_staged<4, 256>(p, out, st, flags) : launch_fwd_staged<4, 0>(p, out, st, flags);
    }
    if (rm_eligible(p) && tune_is("LCR_ROI_FWD", "team")) return launch_fwd_team_impl(p, out, st, all_sw_equal(p, 256));
    if (rm_eligible(p) && tune_is("LCR_ROI_FWD", "rmp")) return all_sw_equal(p, 256) ? launch_fwd_rmp<256>(p, out, st) : launch_fwd_rmp<0>(p, out, st);
    if (rm_eligible(p) && (tune_is("LCR_ROI_FWD", "rm") || tune_is("LCR_ROI_FWD", "rm1"))) {
      const bool two = tune_is("LCR_ROI_FWD", "rm");
      if (all_sw_equal(p, 256)) return two ? launch_fwd_rm<2, 256>(p, out, st) : launch_fwd_rm<1, 256>(p, out, st);
      return two ? launch_fwd_rm<2, 0>(p, out, st) : launch_fwd_rm<1, 0>(p, out, st);
    }
    // Default for the reference's pooler shape (7x7, 256 channels): the launch-ordered row-program kernel with pipelined rows
    // (bench list 1.40 against 1.47 ms for roi_fwd_warp_kernel, and +2 % on the streamed step, where the persistent team
    // kernel — 1.29 ms on its own — loses the overlap with the paste kernel because its CTAs hold the SM's whole shared
    // memory from start to end).  LCR_ROI_FWD=warp selects the sample-walk kernel.
    // Serving batches only: lists that fill the GPU with two-RoI CTAs (below that the kernel is latency-bound per CTA and the
    // shorter table build of the sample-walk kernel wins by 2-3 us — C3: one frame, 896 RoIs) over at least four frames (the
    // single-map sweeps of BASELINE config C5 mix in RoIs of 30-45 feature px, which take this kernel's one sample-walk body:
    // 2-5 % slower there than roi_fwd_warp_kernel, profiles/r02b_roi_ksweep.jsonl).
    if (rm_eligible(p) && !tune_get("LCR_ROI_FWD") && K >= 2 * 4 * sm_count() && p.lv[0].N >= 4)
      return all_sw_equal(p, 256) ? launch_fwd_rmp<256>(p, out, st) : launch_fwd_rmp<0>(p, out, st);
    if (warp_eligible(p) && !tune_is("LCR_ROI_FWD", "cta")) {
      // (P = 7 with 32-channel items — half-warps taking x-bins 0-3 / 4-6, XB = 4, 6 CTAs/SM instead of 4 — was measured
      // 10 % slower than the 64-channel items on B200: the extra occupancy does not pay for the 8-slots-for-7-bins padding.)
      if (PH == 7) return all_sw_equal(p, 256) ? launch_fwd_warp<7, 7, 256>(p, out, st) : launch_fwd_warp<7, 7, 0>(p, out, st);
      return all_sw_equal(p, 256) ? launch_fwd_warp<14, 7, 256>(p, out, st) : launch_fwd_warp<14, 7, 0>(p, out, st);
    }
    if (PH == 7) return launch_fast<7, 128, false>(p, out, st);
    return launch_fast<14, 32, false>(p, out, st);
  }
  if (!tune_is("LCR_ROI_FWD", "generic") && (long long)PH * PW <= (1 << 20)) {
    // per-RoI tap tables (roi_fwd_planes_kernel).  Block = a multiple of PH*PW threads where possible (one bin per thread, sr == 2);
    // channel chunk = a whole number of passes of the block's channel lanes, the most passes that still give `per_sm` CTAs per SM
    // (short lists are latency-bound: more, shorter CTAs).
    const int pp = PH * PW;
    const int threads = pp <= 256 ? (256 / pp) * pp : 256;
    const int unit = pp <= 256 ? 256 / pp : 1;  // channels a block covers per pass
    int per_sm = 8;
    if (const char* v = tune_get("LCR_ROI_PLANES_CTAS")) per_sm = atoi(v) > 0 ? atoi(v) : per_sm;
    int passes = (int)(((long long)K * C) / ((long long)unit * per_sm * sm_count()));
    passes = passes < 1 ? 1 : passes;
    int cchunk = unit * passes;
    cchunk = cchunk > C ? C : (cchunk > 256 ? 256 / unit * unit : cchunk);
    const long long blocks = (long long)K * ((C + cchunk - 1) / cchunk);
    LCR_REQUIRE(blocks < (1ll << 31) && (long long)cchunk * PH * PW < (1ll << 31), LCR_ERR_CAPACITY);
    if (p.sr == 2) roi_fwd_planes_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(p, out, cchunk);
    else roi_fwd_planes_kernel<0><<<(unsigned)blocks, threads, 0, st>>>(p, out, cchunk);
    return after_launch();
  }
  const size_t total = (size_t)K * C * PH * PW;
  const size_t want = (total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 32;
  roi_fwd_generic_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(p, out);
  return after_launch();
}

extern "C" int lcr_roi_align_bwd_f32(const float* grad_out, const LcrFeatLevel* grad_levels_host, int L, int C, const float* rois,
                                     const int* roi_level, int K, int PH, int PW, int sampling_ratio, int aligned,
                                     int zero_grad, void* stream) {
  RoiParams p{};
  int rc = fill_params(p, grad_levels_host, L, C, rois, roi_level, K, PH, PW, sampling_ratio, aligned);
  if (rc != LCR_OK) return rc;
  cudaStream_t st = as_stream(stream);
  if (zero_grad) {
    for (int l = 0; l < L; ++l) {  // dense levels only: the level must be exactly one run of N*C*H*W floats from `data`
      const LvParam& v = p.lv[l];
      const long long HW = (long long)v.H * v.W;
      const bool sn_ok = v.N == 1 || v.sn == HW * C;  // a size-1 batch dimension may carry any stride
      const bool nchw = v.sw == 1 && v.sh == v.W && v.sc == HW && sn_ok;
      const bool nhwc = v.sc == 1 && v.sw == C && v.sh == (long long)v.W * C && sn_ok;
      LCR_REQUIRE(nchw || nhwc, LCR_ERR_INVALID_ARG);
    }
    for (int l = 0; l < L; ++l) {
      const LvParam& v = p.lv[l];
      cudaError_t e = cudaMemsetAsync(v.data, 0, sizeof(float) * (size_t)v.N * C * v.H * v.W, st);
      if (e != cudaSuccess) return cuda_status(e);
    }
  }
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(grad_out, LCR_ERR_INVALID_ARG);
  if (fast_eligible(p, grad_out)) {
    if (warp_eligible(p) && !tune_is("LCR_ROI_BWD", "cta")) {
      if (PH == 7) return all_sw_equal(p, 256) ? launch_bwd_warp<7, 7, 256>(p, grad_out, st) : launch_bwd_warp<7, 7, 0>(p, grad_out, st);
      return all_sw_equal(p, 256) ? launch_bwd_warp<14, 7, 256>(p, grad_out, st) : launch_bwd_warp<14, 7, 0>(p, grad_out, st);
    }
    if (PH == 7) return launch_fast<7, 128, true>(p, const_cast<float*>(grad_out), st);
    return launch_fast<14, 32, true>(p, const_cast<float*>(grad_out), st);
  }
  const size_t total = (size_t)K * C * PH * PW;
  const size_t want = (total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 32;
  roi_bwd_generic_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(p, grad_out);
  return after_launch();
}
