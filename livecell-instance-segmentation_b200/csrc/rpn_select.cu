// rpn_select.cu — RPN objectness -> proposals (a2/a3/a12), ONE launch for all images and levels.
//
// Replaces, per image, sigmoid -> permute/reshape copy -> torch.topk -> bool-mask index (host sync)
// -> anchors[idx] gather -> clamp x2 -> bool-mask index (host sync) of
// generate_inference_proposals / generate_training_proposals (src/utils/proposal_utils.py:16-29,
// :38-52, src/utils/box_utils.py:32-44), and torchvision's per-level top-k + decode + filter
// (TV:models/detection/rpn.py:231-297) when deltas are given.
//
// One thread-block CLUSTER of 8 CTAs owns one (image, level) segment:
//   * each CTA reads its 1/8 position-slice of the [A,h,w] logit map once from HBM with coalesced
//     loads, turns every logit into an order-preserving 32-bit key (sigmoid included) and keeps the
//     keys in shared memory (103 KB per CTA at 704x520) — later passes never touch HBM again;
//   * exact k-th key by 3-pass radix select (11+11+10 bits): per-CTA shared-memory histograms with
//     warp-aggregated atomics (__match_any_sync), merged across the cluster through distributed
//     shared memory (cluster.map_shared_rank), suffix-scanned redundantly by every CTA;
//   * candidates > k-th key (plus the ties with the LOWEST flat indices) are appended with
//     warp-aggregated atomics, rank-sorted on 64-bit (key, ~index) keys by all 8 CTAs, and CTA 0
//     does the ordered epilogue: score threshold, anchor generate/gather, optional delta decode,
//     clip, min-size filter, ordered compaction, device-side count.
// No host sync, no intermediate tensors: 4*A*h*w bytes in, k*(16+4+8) bytes out per segment.
//
// Tie rule (torch.topk leaves it unspecified): (key desc, flat index asc); NaN ranks highest.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lcr {

constexpr int kCluster = 8;
constexpr int kSelThreads = 512;
constexpr int kBins = 2048;

struct SelLevel {
  const float* obj;
  const float* deltas;
  const float* anchors;
  int h, w, stride, pad;
};

struct SelParams {
  SelLevel lv[LCR_MAX_LEVELS];
  float base[LCR_MAX_LEVELS][LCR_MAX_ANCHORS * 4];
  int L, B, A, k;
  float score_thresh, min_size, img_h, img_w;
  int score_strict, topk_on_sigmoid;
  DecodeCfg dec;
  float* boxes;
  float* scores;
  long long* index;
  int* counts;
  unsigned long long* cand;    // [S][k]
  unsigned long long* sorted;  // [S][k]
  unsigned int* counters;      // [S]
  unsigned long long* pre_cand;  // [S][kPreCap] threshold-first survivors (NULL: fast path off)
  unsigned int* pre_count;     // [S][2]: survivors, NaNs
  int region_bytes;            // shared-memory bytes after the histogram (key cache / candidate copy)
};

__device__ __forceinline__ uint32_t make_key(float logit, int on_sigmoid) {
  return order_key(on_sigmoid ? sigmoid_f32(logit) : logit);
}

// Warp-aggregated shared-memory histogram update: lanes with equal bins elect one atomicAdd.
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool active) {
  const unsigned am = __ballot_sync(0xFFFFFFFFu, active);
  if (active) {
    const unsigned peers = __match_any_sync(am, bin);
    if (lane_id() == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
  }
}

// Block-wide exclusive prefix sum (blockDim.x == NT); returns the block total in `total`.
template <int NT = 512>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) woff += t;
    tot += t;
  }
  total = tot;
  return woff + incl - v;
}

struct SliceView {
  const float* obj;      // this segment's [A][P] logits
  const uint32_t* keys;  // shared-memory key cache [A][chunk] (valid iff cached)
  int A, P, p0, np, chunk;
  bool cached;
  int on_sigmoid;
  __device__ __forceinline__ uint32_t key(int a, int pp) const {
    return cached ? keys[a * chunk + pp] : make_key(__ldg(obj + (size_t)a * P + p0 + pp), on_sigmoid);
  }
};

// Cluster-wide digit selection.  Every CTA merges the 8 per-CTA histograms over DSMEM and finds the
// bin holding the `need`-th largest element: returns bin, #elements in higher bins, #elements in it.
__device__ __forceinline__ void cluster_select_bin(cg::cluster_group& cluster, uint32_t* hist, int nbins, uint32_t need,
                                                   uint32_t* s_warp, uint32_t* s_sel, uint32_t& bin, uint32_t& above,
                                                   uint32_t& in_bin) {
  cluster.sync();  // all per-CTA histograms of this pass are complete
  const int bpt = nbins / kSelThreads;  // 4 (2048 bins) or 2 (1024 bins)
  uint32_t c[4] = {0u, 0u, 0u, 0u};
  for (int r = 0; r < kCluster; ++r) {
    const uint32_t* rh = cluster.map_shared_rank(hist, r);
    for (int j = 0; j < bpt; ++j) c[j] += rh[threadIdx.x * bpt + j];
  }
  uint32_t tot = 0;
  for (int j = 0; j < bpt; ++j) tot += c[j];
  uint32_t all;
  const uint32_t excl = block_exclusive_scan(tot, s_warp, all);
  uint32_t run = all - excl - tot;  // elements in bins owned by higher threads
  for (int j = bpt - 1; j >= 0; --j) {
    if (run < need && run + c[j] >= need) {
      s_sel[0] = threadIdx.x * bpt + j;
      s_sel[1] = run;
      s_sel[2] = c[j];
    }
    run += c[j];
  }
  __syncthreads();
  bin = s_sel[0];
  above = s_sel[1];
  in_bin = s_sel[2];
  cluster.sync();  // remote reads done: histograms may be reused
}

// Ordered epilogue over `take` candidates sorted by (key desc, flat index asc): score threshold,
// anchor generate/gather, optional delta decode, clip, min-size filter, ordered compaction,
// device-side count (src/utils/proposal_utils.py:21-29,43-52; TV:models/detection/rpn.py:277-287).
template <int NT, bool SORTED_IN_SMEM>
__device__ __forceinline__ void ordered_epilogue(const SelParams& p, int seg, int b, int l, const unsigned long long* sorted,
                                                 int take, uint32_t* s_warp) {
  const SelLevel& lv = p.lv[l];
  const int A = p.A, P = lv.h * lv.w, n = A * P;
  const float* obj = lv.obj + (size_t)b * n;
  const int tid = threadIdx.x;
  const int per = (take + NT - 1) / NT;
  const int jb = min(take, tid * per), je = min(take, jb + per);
  const int W = lv.w;

  auto eval = [&](int j, float4& box, float& score, uint32_t& flat) -> bool {
    const unsigned long long v = SORTED_IN_SMEM ? sorted[j] : __ldcg(sorted + j);
    flat = 0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFull);
    const int a = (int)(flat % (uint32_t)A), pos = (int)(flat / (uint32_t)A);
    score = sigmoid_f32(__ldg(obj + (size_t)a * P + pos));
    const bool pass = p.score_strict ? (score > p.score_thresh) : (score >= p.score_thresh);
    if (!pass) return false;
    if (lv.anchors) {
      box = __ldg(reinterpret_cast<const float4*>(lv.anchors) + flat);
    } else {
      const int y = pos / W, x = pos - y * W;
      const float sx = __fmul_rn((float)x, (float)lv.stride), sy = __fmul_rn((float)y, (float)lv.stride);
      const float* ba = p.base[l] + a * 4;
      box = make_float4(__fadd_rn(sx, ba[0]), __fadd_rn(sy, ba[1]), __fadd_rn(sx, ba[2]), __fadd_rn(sy, ba[3]));
    }
    if (lv.deltas) {
      const float* d = lv.deltas + (size_t)b * 4 * n + (size_t)(a * 4) * P + pos;
      box = decode_one(make_float4(__ldg(d), __ldg(d + P), __ldg(d + 2 * (size_t)P), __ldg(d + 3 * (size_t)P)), box, p.dec);
    }
    box.x = clampf(box.x, 0.f, p.img_w);
    box.z = clampf(box.z, 0.f, p.img_w);
    box.y = clampf(box.y, 0.f, p.img_h);
    box.w = clampf(box.w, 0.f, p.img_h);
    return (__fsub_rn(box.z, box.x) >= p.min_size) && (__fsub_rn(box.w, box.y) >= p.min_size);
  };

  uint32_t cnt = 0;
  for (int j = jb; j < je; ++j) {
    float4 box;
    float score;
    uint32_t flat;
    cnt += eval(j, box, score, flat) ? 1u : 0u;
  }
  uint32_t total;
  uint32_t o = block_exclusive_scan<NT>(cnt, s_warp, total);
  float4* out_boxes = reinterpret_cast<float4*>(p.boxes) + (size_t)seg * p.k;
  float* out_scores = p.scores + (size_t)seg * p.k;
  long long* out_index = p.index + (size_t)seg * p.k;
  for (int j = jb; j < je; ++j) {
    float4 box;
    float score;
    uint32_t flat;
    if (eval(j, box, score, flat)) {
      out_boxes[o] = box;
      out_scores[o] = score;
      out_index[o] = (long long)flat;
      ++o;
    }
  }
  if (tid == 0) p.counts[seg] = (int)total;
}

// ------------------------------------------------------------------------------------------------
// Threshold-first fast path.
// top-k followed by "score > thr" keeps a PREFIX of the sorted top-k, so it equals: drop everything
// that fails the threshold, then take the top-k of the survivors (NaN scores, which rank first in
// torch.topk but fail the threshold, are the one exception and send the segment to the general
// kernel).  On a real objectness map only a few thousand of the 205 920 anchors pass 0.3, so:
//   1. rpn_prefilter_kernel — one streaming pass over the logits at HBM speed (8 independent loads
//      per thread, one compare per element; the exact sigmoid is evaluated only inside a narrow band
//      around logit(thr)); survivors are appended as 64-bit (key, ~index) with warp-aggregated atomics;
//   2. rpn_sortfilter_kernel — one CTA per segment: survivors (<= kPreCap) into shared memory,
//      bitonic sort on (key desc, index asc), first min(k, M) through the ordered epilogue.
// Segments with more than kPreCap survivors (low thresholds, e.g. the training path's 0.01) or a NaN
// are left to the general cluster kernel, which returns immediately for the others.
// ------------------------------------------------------------------------------------------------
constexpr int kPreCap = 8192;
constexpr int kPreThreads = 256;
constexpr int kPreEPT = 8;
constexpr int kSortThreads = 1024;

__global__ void __launch_bounds__(kPreThreads) rpn_prefilter_kernel(const __grid_constant__ SelParams p, float band_lo, float band_hi) {
  __shared__ uint32_t s_warp[kPreThreads / 32];
  __shared__ uint32_t s_base;
  const int seg = blockIdx.y;
  const int b = seg / p.L, l = seg - b * p.L;
  const SelLevel& lv = p.lv[l];
  const int A = p.A, P = lv.h * lv.w, n = A * P;
  const int m0 = blockIdx.x * (kPreThreads * kPreEPT) + threadIdx.x;
  if (blockIdx.x * (kPreThreads * kPreEPT) >= n) return;
  const float* obj = lv.obj + (size_t)b * n;
  unsigned long long* cand = p.pre_cand + (size_t)seg * kPreCap;
  float v[kPreEPT];
#pragma unroll
  for (int j = 0; j < kPreEPT; ++j) {
    const int m = m0 + j * kPreThreads;
    v[j] = m < n ? __ldg(obj + m) : -INFINITY;   // -inf never passes and is not NaN
  }
  uint32_t bits = 0u;   // bit j: element j of this thread survives the threshold
  bool any_nan = false;
#pragma unroll
  for (int j = 0; j < kPreEPT; ++j) {
    const int m = m0 + j * kPreThreads;
    const float x = v[j];
    bool pass = x > band_hi;   // (padding lanes hold -inf: never above the band)
    if (!pass && x >= band_lo && m < n) {
      const float s = sigmoid_f32(x);
      pass = p.score_strict ? (s > p.score_thresh) : (s >= p.score_thresh);
    }
    bits |= pass ? (1u << j) : 0u;
    any_nan |= (x != x);
  }
  // ONE atomic per CTA (2048 logits): same-address atomics with a return value serialise at L2, and a per-warp
  // append made them the whole cost of this kernel (ncu r01c: 67 % of the stall samples on the broadcast after it).
  const uint32_t cnt = __popc(bits);
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  const unsigned nm = __ballot_sync(0xFFFFFFFFu, any_nan);
  if (nm && lane == 0) atomicAdd(&p.pre_count[2 * seg + 1], (uint32_t)__popc(nm));   // rare: sends the segment to the general kernel
  __syncthreads();
  uint32_t woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kPreThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) woff += t;
    total += t;
  }
  if (threadIdx.x == 0) s_base = total ? atomicAdd(&p.pre_count[2 * seg], total) : 0u;
  __syncthreads();
  uint32_t pos = s_base + woff + incl - cnt;
#pragma unroll
  for (int j = 0; j < kPreEPT; ++j) {
    if ((bits >> j) & 1u) {
      if (pos < (uint32_t)kPreCap) {
        const int m = m0 + j * kPreThreads;
        const int a = m / P, pp = m - a * P;
        const uint32_t flat = (uint32_t)(pp * A + a);
        cand[pos] = ((unsigned long long)make_key(v[j], p.topk_on_sigmoid) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
      }
      ++pos;
    }
  }
}

__device__ __forceinline__ bool fast_path_done(const SelParams& p, int seg) {
  return p.pre_count && p.pre_count[2 * seg] <= (uint32_t)kPreCap && p.pre_count[2 * seg + 1] == 0u;
}

// Block-wide bitonic sort (descending) of SZ 64-bit keys in shared memory, SZ a power of two in [32, 8 * kSortThreads].
// Thread t owns the E consecutive keys t*E .. t*E+E-1 in REGISTERS: compare-exchange passes with stride < E are register
// swaps, strides < 32*E are warp shuffles, and only the strides that cross warps go through shared memory behind a
// __syncthreads (15 of the 78 passes of a 4096-key sort).  The all-shared-memory version spent 3/4 of the kernel waiting
// on LDS -> compare -> STS chains (ncu r01d source page).
template <int E>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* s, int SZ, int tid) {
  const int lane = tid & 31;
  const int base = tid * E;
  const bool active = base < SZ;
  unsigned long long k[E];
#pragma unroll
  for (int r = 0; r < E; ++r) k[r] = active ? s[base + r] : 0ull;
  bool in_smem = false;  // uniform: where the current keys live
  for (int size = 2; size <= SZ; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32 * E) {  // pairs in different warps
        if (!in_smem) {
          if (active) {
#pragma unroll
            for (int r = 0; r < E; ++r) s[base + r] = k[r];
          }
          in_smem = true;
        }
        __syncthreads();
        for (int t = tid; t < (SZ >> 1); t += kSortThreads) {
          const int i = 2 * t - (t & (stride - 1));
          const int j = i + stride;
          const unsigned long long x = s[i], y = s[j];
          if ((x < y) == ((i & size) == 0)) {
            s[i] = y;
            s[j] = x;
          }
        }
      } else {
        if (in_smem) {
          __syncthreads();
#pragma unroll
          for (int r = 0; r < E; ++r) k[r] = active ? s[base + r] : 0ull;
          in_smem = false;
        }
        if (stride >= E) {  // partner key sits in another lane of this warp
          const int lm = stride / E;
          const bool lower = (lane & lm) == 0;
#pragma unroll
          for (int r = 0; r < E; ++r) {
            const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, k[r], lm);
            const bool desc = ((base + r) & size) == 0;
            const unsigned long long hi = k[r] > o ? k[r] : o, lo = k[r] > o ? o : k[r];
            k[r] = (lower == desc) ? hi : lo;  // descending block: the lower index keeps the larger key
          }
        } else {  // both keys in this thread (static register indices: one unrolled body per possible stride)
#pragma unroll
          for (int ls = E >> 1; ls > 0; ls >>= 1) {
            if (stride == ls) {
#pragma unroll
              for (int r = 0; r < E; ++r) {
                if ((r & ls) == 0) {
                  const bool desc = ((base + r) & size) == 0;
                  const unsigned long long x = k[r], y = k[r + ls];
                  if ((x < y) == desc) {
                    k[r] = y;
                    k[r + ls] = x;
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  if (!in_smem && active) {
#pragma unroll
    for (int r = 0; r < E; ++r) s[base + r] = k[r];
  }
}

__global__ void __launch_bounds__(kSortThreads) rpn_sortfilter_kernel(const __grid_constant__ SelParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* s = reinterpret_cast<unsigned long long*>(smem);
  __shared__ uint32_t s_warp[kSortThreads / 32];
  const int seg = blockIdx.x;
  if (!fast_path_done(p, seg)) return;  // block-uniform: the general kernel owns this segment
  const int b = seg / p.L, l = seg - b * p.L;
  const SelLevel& lv = p.lv[l];
  const int n = p.A * lv.h * lv.w;
  const int M = (int)p.pre_count[2 * seg];
  const int tid = threadIdx.x;
  int SZ = 32;
  while (SZ < M) SZ <<= 1;
  const unsigned long long* cand = p.pre_cand + (size_t)seg * kPreCap;
  for (int i = tid; i < SZ; i += kSortThreads) s[i] = i < M ? __ldcg(cand + i) : 0ull;  // 0 sorts last
  __syncthreads();
  const int epl = SZ / kSortThreads;  // keys per thread in the register phases (1 when the array is shorter than the CTA)
  if (epl <= 1) bitonic_sort_desc<1>(s, SZ, tid);
  else if (epl == 2) bitonic_sort_desc<2>(s, SZ, tid);
  else if (epl == 4) bitonic_sort_desc<4>(s, SZ, tid);
  else bitonic_sort_desc<8>(s, SZ, tid);
  __syncthreads();
  ordered_epilogue<kSortThreads, true>(p, seg, b, l, s, min(M, min(p.k, n)), s_warp);
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kSelThreads, 1)
    rpn_select_kernel(const __grid_constant__ SelParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int seg = blockIdx.x / kCluster;
  if (p.pre_count && p.pre_count[2 * seg] <= 8192u && p.pre_count[2 * seg + 1] == 0u) return;  // cluster-uniform: fast path did it
  const int b = seg / p.L, l = seg - b * p.L;
  const SelLevel& lv = p.lv[l];
  const int A = p.A, P = lv.h * lv.w, n = A * P;
  const int kk = min(p.k, n);
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem);
  uint32_t* keys = hist + kBins;
  __shared__ uint32_t s_warp[kSelThreads / 32];
  __shared__ uint32_t s_sel[4];
  __shared__ uint32_t s_ties;

  // position slice of this CTA
  const int chunk = (((P + kCluster - 1) / kCluster) + 3) & ~3;
  const int p0 = min(P, rank * chunk), p1 = min(P, p0 + chunk);
  SliceView sv;
  sv.obj = lv.obj + (size_t)b * n;
  sv.keys = keys;
  sv.A = A; sv.P = P; sv.p0 = p0; sv.np = p1 - p0; sv.chunk = chunk;
  sv.cached = (size_t)A * chunk * sizeof(uint32_t) <= (size_t)p.region_bytes;
  sv.on_sigmoid = p.topk_on_sigmoid;
  const int np = sv.np;

  if (rank == 0 && tid == 0) p.counters[seg] = 0u;
  for (int i = tid; i < kBins; i += kSelThreads) hist[i] = 0u;
  __syncthreads();

  // ---- pass 1: read HBM once, build keys, histogram of the top 11 bits -------------------------
  for (int a = 0; a < A; ++a) {
    const float* row = sv.obj + (size_t)a * P + p0;
    for (int base = 0; base < np; base += kSelThreads) {
      const int pp = base + tid;
      const bool act = pp < np;
      uint32_t key = 0u;
      if (act) {
        key = make_key(__ldg(row + pp), sv.on_sigmoid);
        if (sv.cached) keys[a * chunk + pp] = key;
      }
      hist_add(hist, key >> 21, act);
    }
  }
  uint32_t b1, above1, in1;
  cluster_select_bin(cluster, hist, kBins, (uint32_t)kk, s_warp, s_sel, b1, above1, in1);

  // ---- pass 2: next 11 bits among keys with the selected top digit -----------------------------
  for (int i = tid; i < kBins; i += kSelThreads) hist[i] = 0u;
  __syncthreads();
  for (int a = 0; a < A; ++a)
    for (int base = 0; base < np; base += kSelThreads) {
      const int pp = base + tid;
      uint32_t key = 0u;
      bool act = pp < np;
      if (act) {
        key = sv.key(a, pp);
        act = (key >> 21) == b1;
      }
      hist_add(hist, (key >> 10) & 0x7FFu, act);
    }
  uint32_t b2, above2, in2;
  cluster_select_bin(cluster, hist, kBins, (uint32_t)kk - above1, s_warp, s_sel, b2, above2, in2);
  const uint32_t prefix22 = (b1 << 11) | b2;

  // ---- pass 3: last 10 bits --------------------------------------------------------------------
  for (int i = tid; i < kBins; i += kSelThreads) hist[i] = 0u;
  __syncthreads();
  for (int a = 0; a < A; ++a)
    for (int base = 0; base < np; base += kSelThreads) {
      const int pp = base + tid;
      uint32_t key = 0u;
      bool act = pp < np;
      if (act) {
        key = sv.key(a, pp);
        act = (key >> 10) == prefix22;
      }
      hist_add(hist, key & 0x3FFu, act);
    }
  uint32_t b3, above3, n_eq;
  cluster_select_bin(cluster, hist, 1024, (uint32_t)kk - above1 - above2, s_warp, s_sel, b3, above3, n_eq);
  const uint32_t kstar = (prefix22 << 10) | b3;         // the k-th largest key
  const uint32_t n_gt = above1 + above2 + above3;       // keys strictly greater
  const uint32_t need_eq = (uint32_t)kk - n_gt;         // ties to take (1 <= need_eq <= n_eq)
  const bool take_all_eq = need_eq == n_eq;

  // ---- collect: unordered append of everything above the k-th key (and all ties if they all fit)
  unsigned long long* cand = p.cand + (size_t)seg * p.k;
  for (int a = 0; a < A; ++a)
    for (int base = 0; base < np; base += kSelThreads) {
      const int pp = base + tid;
      uint32_t key = 0u;
      bool take = false;
      if (pp < np) {
        key = sv.key(a, pp);
        take = key > kstar || (take_all_eq && key == kstar);
      }
      const unsigned tm = __ballot_sync(0xFFFFFFFFu, take);
      if (tm) {
        uint32_t pos0 = 0;
        if (lane_id() == 0) pos0 = atomicAdd(&p.counters[seg], (uint32_t)__popc(tm));
        pos0 = __shfl_sync(0xFFFFFFFFu, pos0, 0);
        if (take) {
          const uint32_t flat = (uint32_t)((p0 + pp) * A + a);
          cand[pos0 + __popc(tm & ((1u << lane_id()) - 1u))] =
              ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
        }
      }
    }
  if (!take_all_eq) {
    // Rare path (ties straddle the k-th place): take the need_eq ties with the lowest flat index.
    // Flat order == (position, anchor) order and the CTA slices are position-contiguous, so: per
    // thread contiguous position chunks, block scan, cluster offset by rank.
    const int per = (np + kSelThreads - 1) / kSelThreads;
    const int pb = min(np, tid * per), pe = min(np, pb + per);
    uint32_t cnt = 0;
    for (int pp = pb; pp < pe; ++pp)
      for (int a = 0; a < A; ++a) cnt += sv.key(a, pp) == kstar ? 1u : 0u;
    uint32_t tot;
    const uint32_t excl = block_exclusive_scan(cnt, s_warp, tot);
    if (tid == 0) s_ties = tot;
    cluster.sync();
    uint32_t g = excl;
    for (int r = 0; r < rank; ++r) g += *cluster.map_shared_rank(&s_ties, r);
    for (int pp = pb; pp < pe && g < need_eq; ++pp)
      for (int a = 0; a < A && g < need_eq; ++a)
        if (sv.key(a, pp) == kstar) {
          const uint32_t flat = (uint32_t)((p0 + pp) * A + a);
          cand[n_gt + g] = ((unsigned long long)kstar << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
          ++g;
        }
  }
  __threadfence();
  cluster.sync();  // candidate list complete and visible cluster-wide; key cache no longer needed

  // ---- rank sort of the kk candidates on (key desc, index asc), all 8 CTAs ---------------------
  unsigned long long* sc = reinterpret_cast<unsigned long long*>(keys);
  for (int i = tid; i < kk; i += kSelThreads) sc[i] = __ldcg(cand + i);
  __syncthreads();
  unsigned long long* sorted = p.sorted + (size_t)seg * p.k;
  for (int j = rank * kSelThreads + tid; j < kk; j += kCluster * kSelThreads) {
    const unsigned long long mine = sc[j];
    int r = 0;
#pragma unroll 8
    for (int i = 0; i < kk; ++i) r += (sc[i] > mine) ? 1 : 0;
    sorted[r] = mine;
  }
  __threadfence();
  cluster.sync();
  if (rank != 0) return;  // no DSMEM access after this point

  // ---- ordered epilogue (CTA 0): threshold, box, clip, min-size, compaction ---------------------
  ordered_epilogue<kSelThreads, false>(p, seg, b, l, p.sorted + (size_t)seg * p.k, kk, s_warp);
}

static size_t select_ws_layout(int S, int k, void* base, SelParams* p) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = round_up(off + bytes, 256);
    return o;
  };
  const size_t o_cand = take((size_t)S * k * 8);
  const size_t o_sorted = take((size_t)S * k * 8);
  const size_t o_cnt = take((size_t)S * 4);
  const size_t o_pcnt = take((size_t)S * 8);
  const size_t o_pcand = take((size_t)S * 8192 * 8);
  if (p) {
    char* bp = static_cast<char*>(base);
    p->cand = reinterpret_cast<unsigned long long*>(bp + o_cand);
    p->sorted = reinterpret_cast<unsigned long long*>(bp + o_sorted);
    p->counters = reinterpret_cast<unsigned int*>(bp + o_cnt);
    p->pre_count = reinterpret_cast<unsigned int*>(bp + o_pcnt);
    p->pre_cand = reinterpret_cast<unsigned long long*>(bp + o_pcand);
  }
  return off;
}

}  // namespace lcr

using namespace lcr;

extern "C" size_t lcr_rpn_select_workspace_bytes(int B, int L, int k) {
  if (B <= 0 || L <= 0 || k <= 0) return 256;
  return select_ws_layout(B * L, k, nullptr, nullptr);
}

extern "C" int lcr_rpn_select_f32(const LcrRpnLevel* levels_host, int L, int B, const LcrRpnCfg* cfg, float* boxes,
                                  float* scores, int64_t* index, int* counts, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  LCR_REQUIRE(levels_host && cfg && L > 0 && B >= 0, LCR_ERR_INVALID_ARG);
  if (B == 0) return LCR_OK;
  LCR_REQUIRE(boxes && scores && index && counts, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(L <= LCR_MAX_LEVELS && cfg->num_anchors > 0 && cfg->num_anchors <= LCR_MAX_ANCHORS, LCR_ERR_CAPACITY);
  LCR_REQUIRE(cfg->pre_nms_top_n > 0 && cfg->pre_nms_top_n <= LCR_MAX_TOPK, LCR_ERR_CAPACITY);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  const int S = B * L, k = cfg->pre_nms_top_n;
  LCR_REQUIRE((long long)S * kCluster < (1ll << 31), LCR_ERR_CAPACITY);
  LCR_REQUIRE(workspace && aligned_to(workspace, 256) && workspace_bytes >= lcr_rpn_select_workspace_bytes(B, L, k),
              LCR_ERR_WORKSPACE);

  SelParams p{};
  select_ws_layout(S, k, workspace, &p);
  const size_t smem_cap = 200 * 1024;  // key-cache budget per CTA (B200: 227 KB per CTA opt-in)
  size_t region = (size_t)k * 8;
  for (int l = 0; l < L; ++l) {
    const LcrRpnLevel& lv = levels_host[l];
    LCR_REQUIRE(lv.objectness && lv.h > 0 && lv.w > 0, LCR_ERR_INVALID_ARG);
    LCR_REQUIRE(lv.anchors || lv.stride > 0, LCR_ERR_INVALID_ARG);
    LCR_REQUIRE((long long)lv.h * lv.w * cfg->num_anchors < (1ll << 31), LCR_ERR_CAPACITY);
    LCR_REQUIRE(!lv.anchors || aligned_to(lv.anchors, 16), LCR_ERR_ALIGNMENT);
    p.lv[l] = SelLevel{lv.objectness, lv.deltas, lv.anchors, lv.h, lv.w, lv.stride, 0};
    for (int i = 0; i < cfg->num_anchors * 4; ++i) p.base[l][i] = lv.base_anchors[i];
    const int P = lv.h * lv.w;
    const size_t chunk = (size_t)((((P + kCluster - 1) / kCluster) + 3) & ~3);
    const size_t need = chunk * cfg->num_anchors * sizeof(uint32_t);
    if (need <= smem_cap && need > region) region = need;  // larger maps fall back to re-reading L2
  }
  p.L = L; p.B = B; p.A = cfg->num_anchors; p.k = k;
  p.score_thresh = cfg->score_thresh; p.min_size = cfg->min_size;
  p.img_h = (float)cfg->img_h; p.img_w = (float)cfg->img_w;
  p.score_strict = cfg->score_strict; p.topk_on_sigmoid = cfg->topk_on_sigmoid;
  p.dec = DecodeCfg{cfg->decode_weights[0], cfg->decode_weights[1], cfg->decode_weights[2], cfg->decode_weights[3],
                    cfg->xform_clip, 0.f, 0.f};
  p.boxes = boxes; p.scores = scores; p.index = reinterpret_cast<long long*>(index); p.counts = counts;
  p.region_bytes = (int)region;

  const size_t smem = kBins * sizeof(uint32_t) + region;
  static thread_local size_t configured_smem = 0;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev || configured_smem < smem) {
    cudaError_t e = cudaFuncSetAttribute(rpn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    configured_dev = dev;
    configured_smem = smem;
  }
  cudaStream_t st = as_stream(stream);
  const char* mode = tune_get("LCR_SELECT");  // tuning switch for A/B runs: "general" disables the threshold-first path
  const bool fast = !(mode && strcmp(mode, "general") == 0);
  if (fast) {
    // logit band outside which the threshold decision needs no sigmoid: logit(thr) -+ 0.01 (thr away from 0 and 1)
    float lo = -INFINITY, hi = INFINITY;
    const double t = (double)cfg->score_thresh;
    if (t > 1e-4 && t < 1.0 - 1e-4) {
      const double x = log(t / (1.0 - t));
      lo = (float)(x - 0.01);
      hi = (float)(x + 0.01);
    }
    cudaError_t e = cudaMemsetAsync(p.pre_count, 0, (size_t)S * 8, st);
    if (e != cudaSuccess) return cuda_status(e);
    int max_n = 0;
    for (int l = 0; l < L; ++l) max_n = max_n > levels_host[l].h * levels_host[l].w * cfg->num_anchors ? max_n : levels_host[l].h * levels_host[l].w * cfg->num_anchors;
    dim3 g1((max_n + kPreThreads * kPreEPT - 1) / (kPreThreads * kPreEPT), S);
    rpn_prefilter_kernel<<<g1, kPreThreads, 0, st>>>(p, lo, hi);
    int rc = after_launch();
    if (rc != LCR_OK) return rc;
    static thread_local int sort_dev = -1;
    if (sort_dev != dev) {
      e = cudaFuncSetAttribute(rpn_sortfilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPreCap * 8);
      if (e != cudaSuccess) return cuda_status(e);
      sort_dev = dev;
    }
    rpn_sortfilter_kernel<<<S, kSortThreads, kPreCap * 8, st>>>(p);
    rc = after_launch();
    if (rc != LCR_OK) return rc;
  } else {
    p.pre_count = nullptr;
    p.pre_cand = nullptr;
  }
  rpn_select_kernel<<<S * kCluster, kSelThreads, smem, st>>>(p);
  return after_launch();
}
