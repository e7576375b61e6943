// nms.cu — batched greedy NMS (a8): stable rank sort -> bitmask IoU build -> warp resolve.
//
// Replaces torchvision.ops.nms at src/utils/proposal_utils.py:55 (thr 0.4) and
// src/custom_maskrcnn.py:192 (thr 0.5).  torchvision's CUDA op is sort + index_select + a 64x64
// tiled mask kernel + a single-block gather + nonzero (a host sync), per image, in a Python loop;
// here S segments (images x levels) are processed by three stream-ordered launches with no sync:
//
//   1. nms_rank_kernel     stable descending order of the scores by rank counting on 64-bit keys
//                          (score key << 32 | ~index): exact for any input, no sort network.
//                          Skipped when the input is already sorted (the rpn_select output).
//   2. nms_mask_kernel     upper-triangular suppression bitmask.  A CTA owns a 32-row block and a
//                          span of 32 column words; the column boxes sit in registers, the row
//                          boxes in shared memory; one __ballot_sync yields a 32-bit mask word, and
//                          the 32x32 words are staged in shared memory so rows go out as 128-byte
//                          coalesced stores.  IoU arithmetic is torchvision's: fp32
//                          inter/(a_i+a_j-inter), compared in double against the threshold.
//   3. nms_resolve_kernel  one warp per segment walks the boxes in chunks of 32: the chunk's
//                          32x32 diagonal block is resolved with register shuffles, and the rows of
//                          the kept boxes are OR-ed into the `removed` bit-vector (shared memory),
//                          loads for the next chunk being issued before the serial part.
//                          Stops as soon as post_n boxes are kept.
//   3'. nms_jacobi_kernel  segments of <= 2048 boxes: the greedy keep vector is the unique solution of the
//                          triangular system kept[i] = !OR_{j<i}(M[i][j] & kept[j]); one thread per box holds its
//                          (lower-triangle) mask row in registers and the whole vector is re-evaluated in parallel
//                          rounds (a cluster of 8 CTAs exchanges the 64 kept words through distributed shared
//                          memory) until it stops changing — 5-7 rounds on crowded scenes instead of a
//                          63-chunk serial chain, exact by induction over the box index.
//
// Bound: latency (N <= 2000 boxes, 20 B each) — reported in microseconds, not GB/s.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lcr {

constexpr int kRankThreads = 256;

struct NmsWorkspace {
  float4* sorted_boxes;  // [S][stride]
  int* order;            // [S][stride]  sorted position -> original index
  int* sorted_cat;       // [S][stride]  (only when category given)
  int* n_valid;          // [S]
  uint32_t* mask;        // [S][stride][nw]
  int nw;                // mask words per row
};

static size_t nms_ws_layout(int S, int stride, void* base, NmsWorkspace* ws) {
  const int nw = ((stride + 31) / 32 + 3) & ~3;  // row pitch in words: 16-byte aligned rows (uint4 row loads)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = round_up(off + bytes, 256);
    return o;
  };
  const size_t o_boxes = take((size_t)S * stride * sizeof(float4));
  const size_t o_order = take((size_t)S * stride * sizeof(int));
  const size_t o_cat = take((size_t)S * stride * sizeof(int));
  const size_t o_nv = take((size_t)S * sizeof(int));
  const size_t o_mask = take((size_t)S * stride * nw * sizeof(uint32_t));
  if (ws) {
    char* b = static_cast<char*>(base);
    ws->sorted_boxes = reinterpret_cast<float4*>(b + o_boxes);
    ws->order = reinterpret_cast<int*>(b + o_order);
    ws->sorted_cat = reinterpret_cast<int*>(b + o_cat);
    ws->n_valid = reinterpret_cast<int*>(b + o_nv);
    ws->mask = reinterpret_cast<uint32_t*>(b + o_mask);
    ws->nw = nw;
  }
  return off;
}

// ---------------------------------------------------------------------------------------------
// 1. rank sort.  grid (ceil(stride/256), S).  key = order_key(score) << 32 | (0xFFFFFFFF - i):
// descending key order == (score desc, index asc) == torch's stable descending sort.
// scores == NULL: identity order (already sorted), only n_valid / copies are produced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRankThreads) nms_rank_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                                                 const int* __restrict__ category,
                                                                 const int* __restrict__ counts, int stride,
                                                                 float score_thresh, int use_thresh, NmsWorkspace ws) {
  __shared__ unsigned long long tile[kRankThreads];
  const int s = blockIdx.y;
  const int n = counts ? min(max(counts[s], 0), stride) : stride;
  const int i = blockIdx.x * kRankThreads + threadIdx.x;
  const float* sc = scores ? scores + (size_t)s * stride : nullptr;

  auto key_of = [&](int j) -> unsigned long long {
    if (j >= n) return 0ull;
    if (!sc) return ((unsigned long long)0xFFFFFFFFu << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
    const float v = __ldg(sc + j);
    if (use_thresh && !(v > score_thresh)) return 0ull;  // dropped (src/custom_maskrcnn.py:185)
    return ((unsigned long long)order_key(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
  };

  const unsigned long long mine = key_of(i);
  int rank = 0, nvalid = 0;
  if (!sc) {  // already sorted: identity order, no counting
    rank = i;
    nvalid = n;
  }
  for (int t0 = 0; sc && t0 < n; t0 += kRankThreads) {
    __syncthreads();
    tile[threadIdx.x] = key_of(t0 + threadIdx.x);
    __syncthreads();
    const int lim = min(kRankThreads, n - t0);
    for (int j = 0; j < lim; ++j) {
      const unsigned long long k = tile[j];
      rank += (k > mine) ? 1 : 0;
      nvalid += (k != 0ull) ? 1 : 0;
    }
  }
  if (mine != 0ull) {
    const size_t o = (size_t)s * stride + rank;
    ws.sorted_boxes[o] = __ldg(boxes + (size_t)s * stride + i);
    ws.order[o] = i;
    if (category) ws.sorted_cat[o] = __ldg(category + (size_t)s * stride + i);
  }
  if (i == 0) ws.n_valid[s] = nvalid;
}

// ---------------------------------------------------------------------------------------------
// 2. mask build.  grid (spans, row_blocks, S), 256 threads = 8 warps.
// CTA (span sp, row block rb): rows [32rb, 32rb+32), column words [32sp, 32sp+32).
// ---------------------------------------------------------------------------------------------
// torchvision's CPU op compares the fp32 IoU with the threshold in DOUBLE (its CUDA op rounds the threshold to fp32 first:
// callers wanting that rule pass (double)(float)thr — see include/lcr.h).  For a float x and a double t,
// (double)x > t  <=>  x > tf  with tf = the largest float whose double value is <= t (host side,
// float_threshold()), so the kernel stays in fp32 — same decisions, no F2F.F64/DSETP per pair.
// Disjoint pairs (inter == 0: IoU is 0 or NaN) can only "suppress" when tf < 0, which the
// NEG_THR = false instantiation rules out: they skip the IEEE division (the common case by far).
template <bool NEG_THR>
__device__ __forceinline__ bool iou_gt(const float4 a, float area_a, const float4 b, float area_b, float tf) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(width, height);
  if (!NEG_THR && !(inter > 0.f)) return false;
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return iou > tf;  // NaN (0/0) compares false: never suppresses
}

static float float_threshold(double thr) {
  if (thr != thr) return nanf("");                       // x > NaN is false for every x
  if (thr >= 3.4028234663852886e38) return INFINITY;     // nothing exceeds it
  if (thr < -3.4028234663852886e38) return -INFINITY;    // every non-NaN IoU exceeds it
  float tf = (float)thr;                                  // round to nearest
  if ((double)tf > thr) tf = nextafterf(tf, -INFINITY);   // largest float <= thr
  return tf;
}

template <bool NEG_THR>
__global__ void __launch_bounds__(256) nms_mask_kernel(int stride, float thr, int use_cat, int span, int lower, NmsWorkspace ws) {
  __shared__ float4 row_box[32];
  __shared__ float row_area[32];
  __shared__ int row_cat[32];
  __shared__ uint32_t words[32][33];

  const int s = blockIdx.z, rb = blockIdx.y, sp = blockIdx.x;
  const int n = ws.n_valid[s];
  const int row0 = rb * 32;
  const int w_lo = sp * span;  // span = mask words per CTA: 32 when many segments fill the GPU, 8 for one segment (latency)
  // nothing to do for empty row blocks or spans entirely below the diagonal / beyond n
  if (row0 >= n || w_lo + span - 1 < rb || w_lo * 32 >= n) return;

  const float4* boxes = ws.sorted_boxes + (size_t)s * stride;
  const int* cats = ws.sorted_cat + (size_t)s * stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x < 32) {
    const int r = row0 + threadIdx.x;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n) b = boxes[r];
    row_box[threadIdx.x] = b;
    row_area[threadIdx.x] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    row_cat[threadIdx.x] = (use_cat && r < n) ? cats[r] : 0;
  }
  __syncthreads();

  for (int wl = warp; wl < span; wl += 8) {
    const int w = w_lo + wl;
    uint32_t my_word = 0u;  // lane r ends up holding the word of row r
    if (w >= rb && w * 32 < n) {
      const int col = w * 32 + lane;
      const bool col_ok = col < n;
      const bool diag_word = (w == rb);
      float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok) cb = boxes[col];
      const float carea = __fmul_rn(__fsub_rn(cb.z, cb.x), __fsub_rn(cb.w, cb.y));
      const int ccat = (use_cat && col_ok) ? cats[col] : 0;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        // upper triangle only, except the 32x32 diagonal block, which is stored symmetric (minus the diagonal): the
        // resolve kernel then reads, for box b, the EARLIER boxes of its chunk that suppress it from b's own word
        const bool wanted = diag_word ? (col != row0 + r) : (col > row0 + r);
        const bool hit = col_ok && wanted && (!use_cat || ccat == row_cat[r]) &&
                         iou_gt<NEG_THR>(row_box[r], row_area[r], cb, carea, thr);
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, hit);
        if (lane == r) my_word = word;
      }
      if (lower && !diag_word) {
        // the mirror image for the parallel resolve: row (32w + c) gets, as its word rb, the rows of this block that
        // suppress column c — a 32x32 bit transpose by ballots, lane c keeping the ballot of bit c
        // (most blocks are empty and the rest hold a few bits: only the columns that occur are visited)
        uint32_t t_word = 0u;
        uint32_t cols = __reduce_or_sync(0xFFFFFFFFu, my_word);
        while (cols) {
          const int c = __ffs(cols) - 1;
          cols &= cols - 1u;
          const uint32_t word = __ballot_sync(0xFFFFFFFFu, (my_word >> c) & 1u);
          if (lane == c) t_word = word;
        }
        if (col_ok) ws.mask[((size_t)s * stride + col) * ws.nw + rb] = t_word;
      }
    }
    words[lane][wl] = my_word;
  }
  __syncthreads();
  // coalesced write-out: warp q writes rows q, q+8, ...; lanes cover 32 consecutive words (128 B)
  const int nw = ws.nw;
  for (int r = warp; r < 32; r += 8) {
    const int row = row0 + r;
    const int w = w_lo + lane;
    if (row < n && w < nw && lane < span && w >= rb) ws.mask[((size_t)s * stride + row) * nw + w] = words[r][lane];
  }
}

// ---------------------------------------------------------------------------------------------
// 3. resolve.  grid S, one CTA of kResolveThreads per segment.
// Warp 0 owns the serial part (greedy pass over the chunk's 32x32 diagonal block, evaluated as a
// ballot fixpoint on the symmetric block: 2-4 rounds instead of a 32-step chain); all warps
// then OR the rows of the kept boxes into the `removed` bit-vector (warp = row pair, lane = word,
// coalesced row segments + shared-memory atomicOr).  The row words are loaded one chunk ahead of the greedy pass, speculatively for all 32
// rows, so no L2 round trip sits on the per-chunk critical path.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxWords = LCR_MAX_NMS_BOXES / 32;
constexpr int kResolveThreads = 512;
constexpr int kResolveWarps = kResolveThreads / 32;
constexpr int kResolveWpt = 2 * (32 / kResolveWarps);  // mask words per thread held in registers: its rows x 2 word groups (2048 boxes per pass)

__global__ void __launch_bounds__(kResolveThreads) nms_resolve_kernel(int stride, int post_n, NmsWorkspace ws,
                                                                      int64_t* __restrict__ keep, int* __restrict__ keep_counts) {
  __shared__ uint32_t removed[kMaxWords];
  __shared__ uint32_t s_kept;
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = ws.n_valid[s];
  const int nw = ws.nw;
  const int nchunks = (n + 31) >> 5;
  const uint32_t* mask = ws.mask + (size_t)s * stride * nw;
  const int* order = ws.order + (size_t)s * stride;
  int64_t* out = keep + (size_t)s * post_n;

  for (int w = tid; w < nchunks; w += kResolveThreads) removed[w] = 0u;
  int count = 0;
  // diagonal words (symmetric 32x32 block of the chunk), two chunks ahead
  auto load_diag = [&](int c) -> uint32_t {
    const int row = (c << 5) + lane;
    return (warp == 0 && c < nchunks && row < n) ? __ldg(mask + (size_t)row * nw + c) : 0u;
  };
  uint32_t diag0 = load_diag(0), diag1 = load_diag(1);
  // original indices of the chunk's boxes, also two chunks ahead (a load -> store dependency inside the chunk loop
  // would put an L2 round trip on warp 0's critical path)
  auto load_order = [&](int c) -> int {
    const int row = (c << 5) + lane;
    return (warp == 0 && row < n) ? __ldg(order + row) : 0;
  };
  int ord0 = load_order(0), ord1 = load_order(1);

  // Row words of chunk c, issued TWO chunks ahead and speculatively for all 32 rows, so no L2 round trip sits on the
  // per-chunk critical path.  Layout: warp w owns rows 2w and 2w+1 of the chunk, lane l owns words c+1+l and c+33+l of
  // those rows — every load is a coalesced 128-byte row segment (a lane = row layout costs 32 L1 tag cycles per
  // load and made the L1 the bottleneck: 2 000 cycles per chunk).
  constexpr int kRowsPerWarp = 32 / kResolveWarps;  // 2
  auto prefetch = [&](int c, uint32_t (&v)[kResolveWpt]) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int row = min((c << 5) + warp * kRowsPerWarp + r, max(n - 1, 0));  // rows past n: clamped, their kept bit is 0
      const uint32_t* mrow = mask + (size_t)row * nw;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int w = c + 1 + lane + 32 * j;
        v[r * 2 + j] = (c < nchunks && w < nchunks) ? __ldg(mrow + w) : 0u;
      }
    }
  };
  uint32_t v0[kResolveWpt], v1[kResolveWpt];
  prefetch(0, v0);
  prefetch(1, v1);
  __syncthreads();

  for (int c = 0; c < nchunks && count < post_n; ++c) {
    const int row0 = c << 5;
    if (warp == 0) {
      const uint32_t diag2 = load_diag(c + 2);
      const int ord2 = load_order(c + 2);
      const int left = n - row0;
      const uint32_t in_range = left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
      const uint32_t alive = ~removed[c] & in_range;
      // Greedy pass over the chunk as a fixpoint: box b is kept iff it is alive and no KEPT earlier box of the chunk
      // suppresses it.  Starting from kept = alive, box b is final after b+1 rounds at the latest (it only depends on
      // earlier boxes); in practice the dependency chains are 2-3 boxes deep, so the loop runs 2-4 ballots instead
      // of a 32-step serial chain.
      const uint32_t earlier = diag0 & ((1u << lane) - 1u);
      const bool me = (alive >> lane) & 1u;
      uint32_t kept = alive;
      for (int it = 0; it < 32; ++it) {
        const uint32_t nk = __ballot_sync(0xFFFFFFFFu, me && !(earlier & kept));
        if (nk == kept) break;
        kept = nk;
      }
      if ((kept >> lane) & 1u) {  // emit the kept boxes of this chunk in order
        const int pos = count + __popc(kept & ((1u << lane) - 1u));
        if (pos < post_n) out[pos] = (int64_t)ord0;
      }
      if (lane == 0) s_kept = kept;
      diag0 = diag1;
      diag1 = diag2;
      ord0 = ord1;
      ord1 = ord2;
    }
    __syncthreads();
    const uint32_t kept = s_kept;
    count += __popc(kept);
    uint32_t v2[kResolveWpt];
    prefetch(c + 2, v2);  // in flight during this fold and the next chunk
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      if ((kept >> (warp * kRowsPerWarp + r)) & 1u) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int w = c + 1 + lane + 32 * j;
          if (w < nchunks && v0[r * 2 + j]) atomicOr(&removed[w], v0[r * 2 + j]);
        }
        // segments longer than 2048 boxes: the remaining words of the kept rows, loaded on demand
        for (int w = c + 1 + lane + 64; w < nchunks; w += 32) {
          const uint32_t x = __ldg(mask + (size_t)(row0 + warp * kRowsPerWarp + r) * nw + w);
          if (x) atomicOr(&removed[w], x);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kResolveWpt; ++j) {
      v0[j] = v1[j];
      v1[j] = v2[j];
    }
    __syncthreads();
  }
  if (tid == 0) keep_counts[s] = min(count, post_n);
}

// ---------------------------------------------------------------------------------------------
// 3'. parallel resolve.  grid (CL, S), cluster (CL, 1, 1), 256 threads; segments of <= 256 * CL boxes.
// Box i lives in 32-box block b = i / 32; block b belongs to CTA (b % CL), warp (b / CL), lane (i % 32).
// The thread keeps words 0..b of its mask row in registers (the diagonal word cut to the earlier boxes).  A round
// evaluates kept[i] = valid[i] && !OR_w(row[w] & K[w]) for every box at once against the previous round's vector K
// (the block's own word is refined on the spot by the ballot fixpoint of the serial kernel, so after round t the
// blocks 0..t-1 are final and the loop ends after at most nblocks + 1 rounds; crowded scenes take 5-7).  Every warp
// stores its new word into the K buffers of all CTAs of the cluster (distributed shared memory) and raises a
// cluster-wide `changed` flag; one cluster barrier per round.
// ---------------------------------------------------------------------------------------------
template <int NW4, int CL>
__global__ void __launch_bounds__(256) nms_jacobi_kernel(int stride, int post_n, NmsWorkspace ws, int64_t* __restrict__ keep,
                                                         int* __restrict__ keep_counts) {
  constexpr int NW = NW4 * 4;
  constexpr uint32_t kAll = 0xFFFFFFFFu;
  __shared__ __align__(16) uint32_t K[2][NW];
  __shared__ uint32_t chg[4];
  cg::cluster_group cluster = cg::this_cluster();
  const int q = CL > 1 ? (int)cluster.block_rank() : 0;
  const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(ws.n_valid[s], 32 * NW);
  const int nw = ws.nw;
  const int b = warp * CL + q;       // this warp's block
  const int i = b * 32 + lane;       // this thread's box (sorted position)
  const bool valid = i < n;
  const uint32_t* mrow = ws.mask + ((size_t)s * stride + (valid ? i : 0)) * nw;
  const int ord = valid ? __ldg(ws.order + (size_t)s * stride + i) : 0;

  uint4 row[NW4];
#pragma unroll
  for (int j = 0; j < NW4; ++j) {
    row[j] = make_uint4(0u, 0u, 0u, 0u);
    if (valid && 4 * j <= b) row[j] = __ldg(reinterpret_cast<const uint4*>(mrow) + j);
  }
  // words past the block are not part of the lower triangle (they hold the upper one): clear them; the diagonal
  // word splits into the earlier boxes of the block (in-warp fixpoint) and nothing else
  uint32_t earlier = 0u;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < NW4; ++j) {
    uint32_t* wv = reinterpret_cast<uint32_t*>(&row[j]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int w = 4 * j + e;
      if (w == b) earlier = wv[e] & lt;
      if (w >= b) wv[e] = 0u;
    }
  }
  for (int w = tid; w < NW; w += 256) {
    const int left = n - 32 * w;
    const uint32_t v = left >= 32 ? kAll : (left > 0 ? (1u << left) - 1u : 0u);
    K[0][w] = v;
    K[1][w] = v;
  }
  if (tid < 4) chg[tid] = 0u;
  if (CL > 1) cluster.sync();  // nobody stores into a peer's K / chg before the peer initialised them
  else __syncthreads();

  int cur = 0;
  for (int r = 0;; ++r) {
    uint32_t acc = 0u;
    const uint4* kv = reinterpret_cast<const uint4*>(K[cur]);
#pragma unroll
    for (int j = 0; j < NW4; ++j) {
      const uint4 k = kv[j];
      acc |= (row[j].x & k.x) | (row[j].y & k.y) | (row[j].z & k.z) | (row[j].w & k.w);
    }
    const bool me = valid && acc == 0u;
    uint32_t word = __ballot_sync(kAll, me);
    for (int it = 0; it < 32; ++it) {  // greedy pass inside the block, as a fixpoint (2-4 ballots)
      const uint32_t nk = __ballot_sync(kAll, me && !(earlier & word));
      if (nk == word) break;
      word = nk;
    }
    const int slot = r % 3;
    if (b < NW) {
      const bool changed = word != K[cur][b];
      if (CL > 1) {
        if (lane < CL) {
          cluster.map_shared_rank(&K[cur ^ 1][0], lane)[b] = word;
          if (changed) *cluster.map_shared_rank(&chg[slot], lane) = 1u;
        }
      } else if (lane == 0) {
        K[cur ^ 1][b] = word;
        if (changed) chg[slot] = 1u;
      }
    }
    if (tid == 0) chg[(r + 1) % 3] = 0u;  // next round's flag: last read two barriers ago, first written after this one
    if (CL > 1) cluster.sync();
    else __syncthreads();
    cur ^= 1;
    if (chg[slot] == 0u) break;  // uniform over the cluster: every CTA received the same flags
  }

  // K[cur] is the keep vector.  Kept boxes go out in score order: position = kept boxes in earlier blocks + earlier
  // kept boxes of the block.
  const uint32_t* kf = K[cur];
  int before = 0, total = 0;
  for (int w = lane; w < NW; w += 32) {
    const int c = __popc(kf[w]);
    total += c;
    if (w < b) before += c;
  }
  before = __reduce_add_sync(kAll, before);
  total = __reduce_add_sync(kAll, total);
  if (b < NW) {
    const uint32_t word = kf[b];
    if ((word >> lane) & 1u) {
      const int pos = before + __popc(word & lt);
      if (pos < post_n) keep[(size_t)s * post_n + pos] = (int64_t)ord;
    }
  }
  if (q == 0 && tid == 0) keep_counts[s] = min(total, post_n);
}

template <int NW4, int CL>
static int launch_jacobi(int S, int stride, int post_n, const NmsWorkspace& ws, int64_t* keep, int* keep_counts, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, S, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, nms_jacobi_kernel<NW4, CL>, stride, post_n, ws, keep, keep_counts);
  if (e != cudaSuccess) return cuda_status(e);
  return after_launch();
}

}  // namespace lcr

using namespace lcr;

extern "C" size_t lcr_nms_workspace_bytes(int S, int stride) {
  if (S <= 0 || stride <= 0) return 256;
  return nms_ws_layout(S, stride, nullptr, nullptr);
}

extern "C" int lcr_nms_f32(const float* boxes, const float* scores, const int* category, const int* counts, int S, int stride,
                           double iou_threshold, float score_thresh, int use_score_thresh, int post_n, int64_t* keep,
                           int* keep_counts, void* workspace, size_t workspace_bytes, void* stream) {
  LCR_REQUIRE(S >= 0 && stride > 0 && post_n > 0, LCR_ERR_INVALID_ARG);
  if (S == 0) return LCR_OK;
  LCR_REQUIRE(boxes && keep && keep_counts, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(stride <= LCR_MAX_NMS_BOXES && S <= 65535, LCR_ERR_CAPACITY);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  LCR_REQUIRE(workspace && aligned_to(workspace, 256) && workspace_bytes >= lcr_nms_workspace_bytes(S, stride),
              LCR_ERR_WORKSPACE);
  cudaStream_t st = as_stream(stream);
  NmsWorkspace ws;
  nms_ws_layout(S, stride, workspace, &ws);

  dim3 g1((stride + kRankThreads - 1) / kRankThreads, S);
  nms_rank_kernel<<<g1, kRankThreads, 0, st>>>(reinterpret_cast<const float4*>(boxes), scores, category, counts, stride,
                                               score_thresh, use_score_thresh, ws);
  int rc = after_launch();
  if (rc != LCR_OK) return rc;

  const int row_blocks = (stride + 31) / 32;
  int span = ((long long)S * row_blocks * ((ws.nw + 31) / 32) >= 4ll * sm_count()) ? 32 : 8;
  if (const char* v = tune_get("LCR_NMS_SPAN")) {  // tuning switch for A/B runs (tools/bench_kernels.py)
    const int e = atoi(v);
    if (e == 8 || e == 16 || e == 32) span = e;
  }
  const int spans = (ws.nw + span - 1) / span;
  dim3 g2(spans, row_blocks, S);
  const float tf = float_threshold(iou_threshold);
  // segments of <= 2048 boxes: parallel (Jacobi) resolve on the full symmetric mask; longer ones: serial chunk walk
  bool parallel = stride <= 2048;
  if (const char* v = tune_get("LCR_NMS_RESOLVE")) parallel = parallel && strcmp(v, "serial") != 0;  // tuning switch (A/B, tests)
  if (tf < 0.f) nms_mask_kernel<true><<<g2, 256, 0, st>>>(stride, tf, category != nullptr, span, parallel ? 1 : 0, ws);
  else nms_mask_kernel<false><<<g2, 256, 0, st>>>(stride, tf, category != nullptr, span, parallel ? 1 : 0, ws);
  rc = after_launch();
  if (rc != LCR_OK) return rc;

  if (parallel) {
    if (stride <= 256) return launch_jacobi<2, 1>(S, stride, post_n, ws, keep, keep_counts, st);
    if (stride <= 512) return launch_jacobi<4, 8>(S, stride, post_n, ws, keep, keep_counts, st);
    if (stride <= 1024) return launch_jacobi<8, 8>(S, stride, post_n, ws, keep, keep_counts, st);
    return launch_jacobi<16, 8>(S, stride, post_n, ws, keep, keep_counts, st);
  }
  nms_resolve_kernel<<<S, kResolveThreads, 0, st>>>(stride, post_n, ws, keep, keep_counts);
  return after_launch();
}
