// common.cuh — shared host/device helpers of liblcr (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>

#include "lcr.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblcr is written for sm_100a (B200) only"
#endif

namespace lcr {

extern thread_local int g_last_cuda_error;
extern std::atomic<uint64_t> g_launch_count;

// Called after every kernel launch: records launch errors without synchronising.
inline int after_launch(int n_launches = 1) {
  g_launch_count.fetch_add((uint64_t)n_launches, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return LCR_ERR_CUDA;
  }
  return LCR_OK;
}

inline int cuda_status(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    (void)cudaGetLastError();
    return LCR_ERR_CUDA;
  }
  return LCR_OK;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached per device id).
int sm_count();

// Value of tuning switch `key` ("LCR_..."), or nullptr when unset.  The environment is read once per process; see api.cu.
const char* tune_get(const char* key);
inline bool tune_is(const char* key, const char* value) {
  const char* v = tune_get(key);
  return v && strcmp(v, value) == 0;
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define LCR_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31); }

// torch's CUDA sigmoid for float: 1 / (1 + exp(-x)) with libdevice expf and IEEE division
// (no fast-math in this build), the formula behind torch.sigmoid at proposal_utils.py:16,38.
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

// Order-preserving map float -> uint32 (ascending); -0 == +0; every NaN maps to the top
// (torch.topk / sort rank NaN highest).
__device__ __forceinline__ uint32_t order_key(float v) {
  if (v != v) return 0xFFFFFFFFu;
  v = v + 0.0f;
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// torch.clamp semantics (NaN propagates), as used by clip_boxes_to_image (src/utils/box_utils.py:35-36).
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct DecodeCfg {
  float wx, wy, ww, wh, clip, img_h, img_w;
};

// a7 — BoxCoder.decode_single (TV:models/detection/_utils.py:183-224).  Each torch op is one fp32
// rounding, so nothing here may be contracted into an FMA.
__device__ __forceinline__ float4 decode_one(float4 d, float4 a, const DecodeCfg& c) {
  const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);
  const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(d.x, c.wx), dy = __fdiv_rn(d.y, c.wy);
  float dw = __fdiv_rn(d.z, c.ww), dh = __fdiv_rn(d.w, c.wh);
  dw = dw > c.clip ? c.clip : dw;  // torch.clamp(max=)
  dh = dh > c.clip ? c.clip : dh;
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
  float4 o = make_float4(__fsub_rn(pcx, hw), __fsub_rn(pcy, hh), __fadd_rn(pcx, hw), __fadd_rn(pcy, hh));
  if (c.img_h > 0.f) {
    o.x = clampf(o.x, 0.f, c.img_w);
    o.z = clampf(o.z, 0.f, c.img_w);
    o.y = clampf(o.y, 0.f, c.img_h);
    o.w = clampf(o.w, 0.f, c.img_h);
  }
  return o;
}

#endif  // __CUDACC__

}  // namespace lcr
