// paste.cu — batched mask paste / threshold to full-resolution uint8 frames (a13).
//
// Replaces the per-detection Python loop of CustomMaskRCNN._generate_masks
// (src/custom_maskrcnn.py:276-295) / paste_masks_in_image (src/utils/mask_utils.py:149-171):
// ~6 host syncs + 4 launches + an fp32 [N,H,W] temporary per detection become ONE store-bound
// kernel that writes every output byte exactly once (1 B/px instead of ~17 B/px of traffic).
//
// Roofline: pure HBM write stream.  Algorithmic bytes per detection = H*W (u8 out) + M*M*4 + 16.
// Layout: out [N, H, W] u8; a thread owns one 16-byte column segment of the frame and walks down
// the rows of its band, so no per-vector integer division is needed; stores are st.global.cs
// (streaming: the 183 MB/image of masks must not evict the L2-resident feature maps).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace lcr {

struct PasteBox {
  int x1, y1, x2, y2;   // clamped integer box, half-open
  float sh, sw;         // M / box_h, M / box_w   (area_pixel_compute_scale, align_corners=False)
  bool live;
};

// box.int() truncation + clamp to the frame (src/custom_maskrcnn.py:279-283).
__device__ __forceinline__ PasteBox make_paste_box(const float* __restrict__ boxes, int i, int M, int H, int W) {
  const float4 b = __ldg(reinterpret_cast<const float4*>(boxes) + i);
  PasteBox p;
  p.x1 = max(0, __float2int_rz(b.x));
  p.y1 = max(0, __float2int_rz(b.y));
  p.x2 = min(W, __float2int_rz(b.z));
  p.y2 = min(H, __float2int_rz(b.w));
  p.live = (p.x2 > p.x1) && (p.y2 > p.y1);
  p.sh = p.live ? __fdiv_rn((float)M, (float)(p.y2 - p.y1)) : 0.f;
  p.sw = p.live ? __fdiv_rn((float)M, (float)(p.x2 - p.x1)) : 0.f;
  return p;
}

// ATen upsample_bilinear2d source index (align_corners=False): src = max(scale*(dst+0.5)-0.5, 0),
// the multiply-add contracted to one FMA as nvcc does for ATen's own kernel.
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  i0 = min(__float2int_rz(src), in_size - 1);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = __fsub_rn(src, (float)i0);
  l0 = __fsub_rn(1.0f, l1);
}

// One thresholded pixel: val = h0*(w0*a + w1*b) + h1*(w0*c + w1*d) with the FMA placement
// fma(w0,a,w1*b), fma(h0,top,h1*bot) (same tree as the oracle's restatement).
__device__ __forceinline__ bool paste_pixel(const float* __restrict__ prob, int M, int h0, int h1, float wy0, float wy1,
                                            float sw, int dx, float thr) {
  int w0, w1;
  float wx0, wx1;
  src_index(sw, dx, M, w0, w1, wx0, wx1);
  const float a = __ldg(prob + h0 * M + w0), b = __ldg(prob + h0 * M + w1);
  const float c = __ldg(prob + h1 * M + w0), d = __ldg(prob + h1 * M + w1);
  const float top = __fmaf_rn(a, wx0, __fmul_rn(b, wx1));
  const float bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
  const float v = __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
  return v > thr;
}

// Fast path: W % 16 == 0 and 16-byte aligned frames.  blockDim.x = VPR * RPP where VPR = W/16
// vectors per row and RPP rows per pass; a work item is (detection, band of `band_rows` rows).
__global__ void __launch_bounds__(512) paste_rows16_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                           const uint8_t* __restrict__ valid, int N, int M, int H, int W,
                                                           int vpr, int rpp, int band_rows, int bands, float thr,
                                                           uint32_t on_value, uint8_t* __restrict__ out) {
  const int xv = (int)threadIdx.x % vpr;   // one division per thread per kernel
  const int ry = (int)threadIdx.x / vpr;
  const int x0 = xv * 16;
  const long long items = (long long)N * bands;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int det = (int)(item / bands);
    const int band = (int)(item - (long long)det * bands);
    if (valid && !valid[det]) continue;
    const PasteBox pb = make_paste_box(boxes, det, M, H, W);
    const float* prob = probs + (size_t)det * M * M;
    uint8_t* frame = out + (size_t)det * H * W;
    const int y_end = min(H, (band + 1) * band_rows);
    const bool col_hit = pb.live && (x0 < pb.x2) && (x0 + 16 > pb.x1);
    for (int y = band * band_rows + ry; y < y_end; y += rpp) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (col_hit && y >= pb.y1 && y < pb.y2) {
        int h0, h1;
        float wy0, wy1;
        src_index(pb.sh, y - pb.y1, M, h0, h1, wy0, wy1);
        // rare path (a few % of the vectors): kept rolled so the store loop stays at low register count
        uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          uint32_t word = 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = x0 + q * 4 + j;
            if (x >= pb.x1 && x < pb.x2 && paste_pixel(prob, M, h0, h1, wy0, wy1, pb.sw, x - pb.x1, thr))
              word |= on_value << (8 * j);
          }
          w0 = q == 0 ? word : w0;
          w1 = q == 1 ? word : w1;
          w2 = q == 2 ? word : w2;
          w3 = q == 3 ? word : w3;
        }
        v = make_uint4(w0, w1, w2, w3);
      }
      __stcs(reinterpret_cast<uint4*>(frame + (size_t)y * W + x0), v);
    }
  }
}

// ---- TMA path ------------------------------------------------------------------------------------
// A frame is >99 % zeros: only the rows y1..y2 of the box carry data.  The SM therefore does not
// execute one store instruction per 16 bytes; it hands the copy engine a few large bulk stores
// (cp.async.bulk.global.shared::cta, SASS UBLKCP):
//   * rows above / below the box: straight from a zero-filled shared buffer (ZB bytes per store),
//   * box rows: composed (bilinear + threshold) into a double-buffered row chunk in shared memory,
//     one bulk store per chunk of RB / W rows.
// One CTA per detection (persistent), thread 0 issues and tracks the bulk groups; instruction issue
// drops from ~100 per 512 bytes to a handful per frame, leaving the kernel on the HBM write roofline.
__device__ __forceinline__ void paste_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void paste_bulk_store(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void paste_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void paste_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void paste_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// bilinear + threshold with the probabilities staged in shared memory (same expression tree as paste_pixel)
__device__ __forceinline__ bool paste_pixel_smem(const float* prob, int M, int h0, int h1, float wy0, float wy1, float sw, int dx,
                                                 float thr) {
  int w0, w1;
  float wx0, wx1;
  src_index(sw, dx, M, w0, w1, wx0, wx1);
  const float a = prob[h0 * M + w0], b = prob[h0 * M + w1];
  const float c = prob[h1 * M + w0], d = prob[h1 * M + w1];
  const float top = __fmaf_rn(a, wx0, __fmul_rn(b, wx1));
  const float bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
  const float v = __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
  return v > thr;
}

constexpr int kPasteThreads = 128;
constexpr int kPasteZB = 16 * 1024;   // zero buffer
constexpr int kPasteRB = 11 * 1024;   // one row chunk (16 rows of a 704-px frame)

__global__ void __launch_bounds__(kPasteThreads) paste_bulk_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                                   const uint8_t* __restrict__ valid, int N, int M, int H, int W,
                                                                   float thr, uint32_t on_value, uint8_t* __restrict__ out, int zb_bytes, int interleave) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint8_t* zb = smem;
  uint8_t* rb = smem + zb_bytes;                                               // two chunks of kPasteRB bytes
  float* sprob = reinterpret_cast<float*>(smem + zb_bytes + 2 * kPasteRB);    // [M*M]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < zb_bytes / 16; i += kPasteThreads) reinterpret_cast<uint4*>(zb)[i] = make_uint4(0u, 0u, 0u, 0u);
  paste_fence_async();
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int vpr = W / 16;
  const int chunk_rows = max(1, min(kPasteRB / W, H));
  const int MM = M * M;
  int buf = 0;
  // Bulk-group bookkeeping (thread 0): groups complete in commit order; a row chunk may be recomposed once the group
  // that stored it has finished READING shared memory.  The SM's copy engine drains its queue in order, so the wait
  // for a chunk only ends when everything queued before it has gone out — and the queue must not run dry meanwhile:
  // the frame's zero rows (95 % of its bytes) are therefore not issued in one piece after the box rows, but as one
  // slice (its own group) behind every chunk store, so there is always zero-row work queued BEHIND the chunk the
  // CTA waits for next (r01c issued them last: the queue ran dry during every chunk wait, 11 % of the kernel).
  int seq = 0;                      // groups committed so far
  int last_use[2] = {-1, -1};       // group id of each chunk buffer's latest store

  for (int det = blockIdx.x; det < N; det += gridDim.x) {
    if (valid && !valid[det]) continue;  // block-uniform
    const PasteBox pb = make_paste_box(boxes, det, M, H, W);
    uint8_t* frame = out + (size_t)det * H * W;
    const int y1 = pb.live ? pb.y1 : 0, y2 = pb.live ? pb.y2 : 0;
    // zero rows [0, y1) and [y2, H) as one virtual byte range walked by a cursor (thread 0)
    const size_t z_top = (size_t)y1 * W, z_all = z_top + (size_t)(H - y2) * W;
    size_t z_cur = 0;
    auto emit_zeros = [&](size_t upto) {  // thread 0: bulk stores from the zero buffer for virtual bytes [z_cur, upto), one group
      upto = upto < z_all ? upto : z_all;
      if (z_cur >= upto) return;
      while (z_cur < upto) {
        const bool top = z_cur < z_top;
        const size_t seg_end = top ? (z_top < upto ? z_top : upto) : upto;
        const size_t gaddr = top ? z_cur : (size_t)y2 * W + (z_cur - z_top);
        const uint32_t n = (uint32_t)min((size_t)zb_bytes, seg_end - z_cur);
        paste_bulk_store(frame + gaddr, zb, n, pol);
        z_cur += n;
      }
      paste_bulk_commit();
      ++seq;
    };
    if (pb.live) {
      // the previous frame's readers of sprob passed the barrier that followed their last chunk
      const float* prob = probs + (size_t)det * MM;
      for (int i = tid; i < MM; i += kPasteThreads) sprob[i] = __ldg(prob + i);
      const int nchunks = (y2 - y1 + chunk_rows - 1) / chunk_rows;
      // slice of the zero range issued behind each chunk (multiple of 16 B: W % 16 == 0 keeps every store aligned)
      const size_t slice = ((z_all / (size_t)nchunks) + 15) & ~(size_t)15;
      int ci = 0;
      for (int yc = y1; yc < y2; yc += chunk_rows, ++ci) {
        const int rows = min(chunk_rows, y2 - yc);
        uint8_t* chunk = rb + buf * kPasteRB;
        if (tid == 0) {  // groups younger than this buffer's last store may stay pending
          const int pending_ok = seq - 1 - last_use[buf];
          if (pending_ok <= 0) paste_bulk_wait_read<0>();
          else if (pending_ok == 1) paste_bulk_wait_read<1>();
          else if (pending_ok == 2) paste_bulk_wait_read<2>();
          else paste_bulk_wait_read<3>();
        }
        __syncthreads();                         // (also publishes sprob)
        // Compose the chunk.  The box covers a few 16-pixel vectors of each row: the vectors outside it are zero-filled,
        // and the 4-pixel words inside it are dealt round-robin to ALL threads (a lane = column mapping left the
        // bilinear work to the handful of lanes whose columns fall into the box: 5-10 us per chunk, on the CTA's
        // critical path between two bursts of stores).
        const int vb0 = pb.x1 >> 4, vb1 = (pb.x2 + 15) >> 4;  // vectors [vb0, vb1) intersect the box
        for (int i = tid; i < rows * vpr; i += kPasteThreads) {
          const int r = i / vpr, xv = i - r * vpr;
          if (xv < vb0 || xv >= vb1) *reinterpret_cast<uint4*>(chunk + (size_t)r * W + xv * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        const int nwb = (vb1 - vb0) * 4;  // words per row inside those vectors
        for (int j = tid; j < rows * nwb; j += kPasteThreads) {
          const int r = j / nwb, x0 = (vb0 << 4) + (j - r * nwb) * 4;
          int h0, h1;
          float wy0, wy1;
          src_index(pb.sh, yc + r - pb.y1, M, h0, h1, wy0, wy1);
          uint32_t word = 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int x = x0 + q;
            if (x >= pb.x1 && x < pb.x2 && paste_pixel_smem(sprob, M, h0, h1, wy0, wy1, pb.sw, x - pb.x1, thr)) word |= on_value << (8 * q);
          }
          *reinterpret_cast<uint32_t*>(chunk + (size_t)r * W + x0) = word;
        }
        paste_fence_async();
        __syncthreads();
        if (tid == 0) {
          paste_bulk_store(frame + (size_t)yc * W, chunk, (uint32_t)(rows * W), pol);
          paste_bulk_commit();
          last_use[buf] = seq++;
          if (interleave) emit_zeros(ci + 1 == nchunks ? z_all : z_cur + slice);
        }
        buf ^= 1;
      }
    }
    if (tid == 0) emit_zeros(z_all);  // whatever is left (all of it for an empty box)
  }
  if (tid == 0) paste_bulk_wait_all();
}

// ---- TMA path, warp-specialised (default) -----------------------------------------------------------
// Same stores as paste_bulk_kernel, but issued by two independent roles of a persistent CTA (160 threads):
//   * warp 0, lane 0 — the ZERO ISSUER: walks the CTA's detections and hands the copy engine the zero rows of every
//     frame (95 % of the bytes), never waiting for anything but the engine's own back-pressure;
//   * warps 1-4 — the COMPOSERS: stage the detection's probabilities, compose the box rows (bilinear + threshold,
//     box words dealt round-robin to the 128 threads) into a ring of kSlots row chunks and store them; their bulk
//     groups belong to composer thread 0, which only waits when a ring slot comes round again (kSlots chunks later).
// The two roles write disjoint bytes and never synchronise with each other (composers use a named barrier), so the
// HBM write stream no longer pauses while a CTA composes — the r01d single-role kernel left ~7 % on the table
// against its all-empty-boxes floor.
constexpr int kSplitThreads = 160;
constexpr int kComposers = 128;
constexpr int kSlots = 8;   // default ring depth; 4 / 2 leave more shared memory to a kernel co-running on another stream

__device__ __forceinline__ void composer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kComposers) : "memory"); }

template <int SLOTS>
__global__ void __launch_bounds__(kSplitThreads) paste_split_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                                    const uint8_t* __restrict__ valid, int N, int M, int H, int W,
                                                                    float thr, uint32_t on_value, uint8_t* __restrict__ out, int zb_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint8_t* zb = smem;
  uint8_t* rb = smem + zb_bytes;                                                   // SLOTS chunks of kPasteRB bytes
  float* sprob = reinterpret_cast<float*>(smem + zb_bytes + SLOTS * kPasteRB);   // [M*M]
  const int tid = threadIdx.x;
  for (int i = tid; i < zb_bytes / 16; i += kSplitThreads) reinterpret_cast<uint4*>(zb)[i] = make_uint4(0u, 0u, 0u, 0u);
  paste_fence_async();
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

  if (tid < 32) {
    // ---- zero issuer --------------------------------------------------------------------------
    if (tid != 0) return;
    for (int det = blockIdx.x; det < N; det += gridDim.x) {
      if (valid && !valid[det]) continue;
      const PasteBox pb = make_paste_box(boxes, det, M, H, W);
      uint8_t* frame = out + (size_t)det * H * W;
      const int y1 = pb.live ? pb.y1 : 0, y2 = pb.live ? pb.y2 : 0;
      for (int part = 0; part < 2; ++part) {
        size_t a = part == 0 ? 0 : (size_t)y2 * W;
        const size_t b = part == 0 ? (size_t)y1 * W : (size_t)H * W;
        while (a < b) {
          const uint32_t n = (uint32_t)min((size_t)zb_bytes, b - a);
          paste_bulk_store(frame + a, zb, n, pol);
          a += n;
        }
      }
      paste_bulk_commit();
    }
    paste_bulk_wait_read<0>();  // the zero buffer must outlive its readers
    return;
  }

  // ---- composers ----------------------------------------------------------------------------------
  const int ct = tid - 32;  // 0..127
  const int vpr = W / 16;
  const int chunk_rows = max(1, min(kPasteRB / W, H));
  const int MM = M * M;
  int slot = 0;
  for (int det = blockIdx.x; det < N; det += gridDim.x) {
    if (valid && !valid[det]) continue;  // uniform over the composers
    const PasteBox pb = make_paste_box(boxes, det, M, H, W);
    if (!pb.live) continue;
    uint8_t* frame = out + (size_t)det * H * W;
    const float* prob = probs + (size_t)det * MM;
    composer_barrier();  // the previous detection's readers of sprob are done
    for (int i = ct; i < MM; i += kComposers) sprob[i] = __ldg(prob + i);
    const int vb0 = pb.x1 >> 4, vb1 = (pb.x2 + 15) >> 4;  // vectors [vb0, vb1) intersect the box
    const int nwb = (vb1 - vb0) * 4;                       // words per row inside those vectors
    for (int yc = pb.y1; yc < pb.y2; yc += chunk_rows) {
      const int rows = min(chunk_rows, pb.y2 - yc);
      uint8_t* chunk = rb + slot * kPasteRB;
      // the slot's previous store (kSlots groups ago) must have finished reading it: at most kSlots - 1 younger groups pending
      if (ct == 0) paste_bulk_wait_read<SLOTS - 1>();
      composer_barrier();  // (also publishes sprob)
      for (int i = ct; i < rows * vpr; i += kComposers) {
        const int r = i / vpr, xv = i - r * vpr;
        if (xv < vb0 || xv >= vb1) *reinterpret_cast<uint4*>(chunk + (size_t)r * W + xv * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      for (int j = ct; j < rows * nwb; j += kComposers) {
        const int r = j / nwb, x0 = (vb0 << 4) + (j - r * nwb) * 4;
        int h0, h1;
        float wy0, wy1;
        src_index(pb.sh, yc + r - pb.y1, M, h0, h1, wy0, wy1);
        uint32_t word = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int x = x0 + q;
          if (x >= pb.x1 && x < pb.x2 && paste_pixel_smem(sprob, M, h0, h1, wy0, wy1, pb.sw, x - pb.x1, thr)) word |= on_value << (8 * q);
        }
        *reinterpret_cast<uint32_t*>(chunk + (size_t)r * W + x0) = word;
      }
      paste_fence_async();
      composer_barrier();
      if (ct == 0) {
        paste_bulk_store(frame + (size_t)yc * W, chunk, (uint32_t)(rows * W), pol);
        paste_bulk_commit();
      }
      slot = slot + 1 == SLOTS ? 0 : slot + 1;
    }
  }
  if (ct == 0) paste_bulk_wait_read<0>();
}

// Generic path (any W / alignment): one thread per pixel, byte stores.  Correctness path for odd
// frame widths (e.g. the reference's 300x222 tiles are fine: 300 % 4 == 0 but 300 % 16 != 0).
__global__ void __launch_bounds__(256) paste_generic_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                            const uint8_t* __restrict__ valid, int N, int M, int H, int W,
                                                            float thr, uint8_t on_value, uint8_t* __restrict__ out) {
  const int det = blockIdx.y;
  if (valid && !valid[det]) return;
  const PasteBox pb = make_paste_box(boxes, det, M, H, W);
  const float* prob = probs + (size_t)det * M * M;
  uint8_t* frame = out + (size_t)det * H * W;
  const int hw = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    uint8_t v = 0;
    if (pb.live && y >= pb.y1 && y < pb.y2 && x >= pb.x1 && x < pb.x2) {
      int h0, h1;
      float wy0, wy1;
      src_index(pb.sh, y - pb.y1, M, h0, h1, wy0, wy1);
      v = paste_pixel(prob, M, h0, h1, wy0, wy1, pb.sw, x - pb.x1, thr) ? on_value : 0;
    }
    frame[i] = v;
  }
}

// ---- torchvision-style paste (a13, P2 variant) --------------------------------------------------------------------------
// torchvision.models.detection.roi_heads.paste_masks_in_image (TV:models/detection/roi_heads.py:405-501), what the transfer
// model's postprocess runs: the M x M PROBABILITY map is zero-padded by `padding` pixels, the box is expanded by
// (M + 2*padding) / M around its centre and truncated to integers, the padded map is resized (bilinear, align_corners=False)
// to (y2 - y1 + 1, x2 - x1 + 1) and the part inside the frame is copied: float32 frames, no threshold.
// One CTA per (detection, band of rows); a thread owns four consecutive pixels of a row and stores one float4, so the frame
// (4 B/pixel: 1.46 MB per detection at 704x520) is written as full lines.  HBM-write-bound.
struct TvBox {
  int x0, y0, w, h;   // integer expanded box origin and resized extents (>= 1)
  int xe, ye;         // box[2] + 1, box[3] + 1: exclusive end of the pasted region (<= origin for an inverted box: nothing pasted)
};

__device__ __forceinline__ TvBox tv_box(const float* __restrict__ b, float scale) {
  // expand_boxes: every torch op is one fp32 rounding
  const float x1 = __ldg(b), y1 = __ldg(b + 1), x2 = __ldg(b + 2), y2 = __ldg(b + 3);
  const float wh = __fmul_rn(__fmul_rn(__fsub_rn(x2, x1), 0.5f), scale), hh = __fmul_rn(__fmul_rn(__fsub_rn(y2, y1), 0.5f), scale);
  const float xc = __fmul_rn(__fadd_rn(x2, x1), 0.5f), yc = __fmul_rn(__fadd_rn(y2, y1), 0.5f);
  const long long ex0 = (long long)__fsub_rn(xc, wh), ex1 = (long long)__fadd_rn(xc, wh);   // .to(torch.int64): truncation
  const long long ey0 = (long long)__fsub_rn(yc, hh), ey1 = (long long)__fadd_rn(yc, hh);
  TvBox t;
  t.x0 = (int)max(min(ex0, (long long)(1 << 29)), -(long long)(1 << 29));
  t.y0 = (int)max(min(ey0, (long long)(1 << 29)), -(long long)(1 << 29));
  const long long w = ex1 - ex0 + 1, h = ey1 - ey0 + 1;
  t.w = (int)max(min(w, (long long)(1 << 29)), 1ll);
  t.h = (int)max(min(h, (long long)(1 << 29)), 1ll);
  t.xe = (int)max(min(ex1 + 1, (long long)(1 << 29)), -(long long)(1 << 29));
  t.ye = (int)max(min(ey1 + 1, (long long)(1 << 29)), -(long long)(1 << 29));
  return t;
}

__global__ void __launch_bounds__(256) paste_tv_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                       const uint8_t* __restrict__ valid, int M, int H, int W, int pad, float scale,
                                                       int band_rows, float* __restrict__ out) {
  const int i = blockIdx.y;
  if (valid && !valid[i]) return;
  extern __shared__ float sprob[];   // the padded map, (M + 2 pad)^2
  const int Mp = M + 2 * pad;
  for (int j = threadIdx.x; j < Mp * Mp; j += blockDim.x) {
    const int y = j / Mp - pad, x = j % Mp - pad;
    sprob[j] = (y >= 0 && y < M && x >= 0 && x < M) ? __ldg(probs + (size_t)i * M * M + y * M + x) : 0.f;
  }
  __syncthreads();
  const TvBox tb = tv_box(boxes + (size_t)i * 4, scale);
  const float sh = __fdiv_rn((float)Mp, (float)tb.h), sw = __fdiv_rn((float)Mp, (float)tb.w);
  // region copied into the frame: [max(y0, 0), min(box[3] + 1, H)) x [max(x0, 0), min(box[2] + 1, W))
  const int ya = max(tb.y0, 0), yb = min(tb.ye, H), xa = max(tb.x0, 0), xb = min(tb.xe, W);
  float* frame = out + (size_t)i * H * W;
  const int row0 = blockIdx.x * band_rows, row1 = min(H, row0 + band_rows);
  const int vpr = (W + 3) / 4;
  for (int idx = threadIdx.x; idx < (row1 - row0) * vpr; idx += blockDim.x) {
    const int y = row0 + idx / vpr, xq = (idx % vpr) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= ya && y < yb && xq + 3 >= xa && xq < xb) {
      int h0, h1;
      float wy0, wy1;
      src_index(sh, y - tb.y0, Mp, h0, h1, wy0, wy1);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int x = xq + q;
        if (x >= xa && x < xb) {
          int w0, w1;
          float wx0, wx1;
          src_index(sw, x - tb.x0, Mp, w0, w1, wx0, wx1);
          const float top = __fmaf_rn(sprob[h0 * Mp + w0], wx0, __fmul_rn(sprob[h0 * Mp + w1], wx1));
          const float bot = __fmaf_rn(sprob[h1 * Mp + w0], wx0, __fmul_rn(sprob[h1 * Mp + w1], wx1));
          v[q] = __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
        }
      }
    }
    float* dst = frame + (size_t)y * W + xq;
    if (xq + 3 < W && (W % 4 == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int q = 0; q < 4 && xq + q < W; ++q) dst[q] = v[q];
    }
  }
}

static int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

}  // namespace lcr

using namespace lcr;

extern "C" int lcr_paste_masks_u8(const float* probs, const float* boxes, const uint8_t* valid, int N, int M, int H, int W,
                                  float threshold, uint8_t on_value, uint8_t* out, void* stream) {
  LCR_REQUIRE(N >= 0 && M > 0 && H > 0 && W > 0, LCR_ERR_INVALID_ARG);
  if (N == 0) return LCR_OK;
  LCR_REQUIRE(probs && boxes && out, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  LCR_REQUIRE((int64_t)H * W < (1ll << 31), LCR_ERR_CAPACITY);
  const int vpr = W / 16;
  const bool fast = (W % 16 == 0) && aligned_to(out, 16) && vpr <= 512;
  const char* mode = tune_get("LCR_PASTE");  // tuning switch for A/B runs: "rows16" selects the per-thread store kernel
  int zb_bytes = kPasteZB;
  if (const char* v = tune_get("LCR_PASTE_ZB_KB")) zb_bytes = atoi(v) * 1024;   // tuning switch
  if (zb_bytes < 1024 || zb_bytes > 128 * 1024 || zb_bytes % 1024) zb_bytes = kPasteZB;
  const size_t prob_bytes = round_up(sizeof(float) * (size_t)M * M, 16);
  const bool want_rows16 = mode && strcmp(mode, "rows16") == 0;
  const bool want_single = mode && (strcmp(mode, "single") == 0 || strcmp(mode, "zeros_last") == 0);  // the r01d single-role kernel
  int slots = kSlots;
  if (const char* v = tune_get("LCR_PASTE_SLOTS")) slots = atoi(v);  // tuning switch: 8 (default), 4 or 2 ring slots
  if (slots != 8 && slots != 4 && slots != 2) slots = kSlots;
  const size_t split_smem = (size_t)zb_bytes + (size_t)slots * kPasteRB + prob_bytes;
  if (fast && W <= kPasteRB && split_smem <= 200 * 1024 && !want_rows16 && !want_single) {
    auto kern = slots == 8 ? paste_split_kernel<8> : (slots == 4 ? paste_split_kernel<4> : paste_split_kernel<2>);
    static thread_local int configured_dev = -1;
    static thread_local size_t configured_smem[3] = {0, 0, 0};
    const int ki = slots == 8 ? 0 : (slots == 4 ? 1 : 2);
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev || configured_smem[ki] < split_smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split_smem);
      if (e != cudaSuccess) return cuda_status(e);
      if (configured_dev != dev) configured_smem[0] = configured_smem[1] = configured_smem[2] = 0;
      configured_dev = dev;
      configured_smem[ki] = split_smem;
    }
    // One persistent CTA per SM is enough for the zero issuer to keep the copy engine saturated (1.945 ms; 2 per SM: 1.959)
    // and leaves 120 KB of shared memory per SM to a kernel running beside paste on another stream.
    int per_sm = (int)((227 * 1024) / (split_smem + 1024));
    int cap = 1;
    if (const char* v = tune_get("LCR_PASTE_CTAS")) cap = atoi(v);  // tuning switch
    if (cap >= 1 && cap < per_sm) per_sm = cap;
    long long max_blocks = (long long)sm_count() * (per_sm > 0 ? per_sm : 1);
    if (const char* v = tune_get("LCR_PASTE_GRID")) max_blocks = atoi(v) > 0 ? atoi(v) : max_blocks;  // tuning switch: persistent CTAs in all
    const int blocks = (int)((long long)N < max_blocks ? N : max_blocks);
    kern<<<blocks, kSplitThreads, split_smem, as_stream(stream)>>>(probs, boxes, valid, N, M, H, W, threshold, (uint32_t)on_value, out,
                                                                  zb_bytes);
    return after_launch();
  }
  const size_t bulk_smem = (size_t)zb_bytes + 2 * kPasteRB + prob_bytes;
  if (fast && W <= kPasteRB && bulk_smem <= 200 * 1024 && !want_rows16) {
    static thread_local int configured_dev = -1;
    static thread_local size_t configured_smem = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev || configured_smem < bulk_smem) {
      cudaError_t e = cudaFuncSetAttribute(paste_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem);
      if (e != cudaSuccess) return cuda_status(e);
      configured_dev = dev;
      configured_smem = bulk_smem;
    }
    // Two persistent CTAs per SM already keep the copy engines saturated (measured: 1.98 ms at 2/SM, 1.99 at 3, 2.005 at
    // 5, 2.25 at 1) and leave 140 KB of shared memory per SM to a kernel running beside paste on another stream.
    int per_sm = (int)((227 * 1024) / (bulk_smem + 1024));
    int cap = 2;
    if (const char* v = tune_get("LCR_PASTE_CTAS")) cap = atoi(v);  // tuning switch
    if (cap >= 1 && cap < per_sm) per_sm = cap;
    const long long max_blocks = (long long)sm_count() * (per_sm > 0 ? per_sm : 1);
    const int blocks = (int)((long long)N < max_blocks ? N : max_blocks);
    paste_bulk_kernel<<<blocks, kPasteThreads, bulk_smem, as_stream(stream)>>>(probs, boxes, valid, N, M, H, W, threshold,
                                                                              (uint32_t)on_value, out, zb_bytes,
                                                                              (mode && strcmp(mode, "zeros_last") == 0) ? 0 : 1);
    return after_launch();
  }
  if (fast) {
    // threads = vpr * rpp, a multiple of 32 in [256, 512] where possible
    int rpp = 32 / gcd_int(vpr, 32);
    while (vpr * rpp < 256 && vpr * rpp * 2 <= 512) rpp *= 2;
    if (vpr * rpp > 512) rpp = 32 / gcd_int(vpr, 32);
    if (vpr * rpp <= 512) {
      const int threads = vpr * rpp;
      const int band_rows = rpp * 8;  // 8 stores per thread per work item
      const int bands = (H + band_rows - 1) / band_rows;
      const long long items = (long long)N * bands;
      const long long max_blocks = (long long)sm_count() * 16;
      const int blocks = (int)(items < max_blocks ? items : max_blocks);
      paste_rows16_kernel<<<blocks, threads, 0, as_stream(stream)>>>(probs, boxes, valid, N, M, H, W, vpr, rpp, band_rows,
                                                                    bands, threshold, (uint32_t)on_value, out);
      return after_launch();
    }
  }
  LCR_REQUIRE(N <= 65535, LCR_ERR_CAPACITY);
  dim3 grid((unsigned)((H * W + 255) / 256 < 64 ? (H * W + 255) / 256 : 64), (unsigned)N);
  paste_generic_kernel<<<grid, 256, 0, as_stream(stream)>>>(probs, boxes, valid, N, M, H, W, threshold, on_value, out);
  return after_launch();
}

extern "C" int lcr_paste_masks_tv_f32(const float* probs, const float* boxes, const uint8_t* valid, int N, int M, int H, int W,
                                      int padding, float* out, void* stream) {
  LCR_REQUIRE(N >= 0 && M > 0 && H > 0 && W > 0 && padding >= 0 && padding <= 8, LCR_ERR_INVALID_ARG);
  if (N == 0) return LCR_OK;
  LCR_REQUIRE(probs && boxes && out, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(out, 16), LCR_ERR_ALIGNMENT);
  LCR_REQUIRE(N <= 65535 && (int64_t)H * W < (1ll << 31), LCR_ERR_CAPACITY);
  const int Mp = M + 2 * padding;
  const size_t smem = sizeof(float) * (size_t)Mp * Mp;
  LCR_REQUIRE(smem <= 48 * 1024, LCR_ERR_CAPACITY);
  // scale = float(M + 2 * padding) / M as a Python float, applied to a float32 tensor: rounded to fp32 first
  const float scale = (float)((double)Mp / (double)M);
  const int band_rows = 32;
  dim3 grid((unsigned)((H + band_rows - 1) / band_rows), (unsigned)N);
  paste_tv_kernel<<<grid, 256, smem, as_stream(stream)>>>(probs, boxes, valid, M, H, W, padding, scale, band_rows, out);
  return after_launch();
}
