// layout.cu — NCHW <-> NHWC tiled transposes through padded shared memory (64x64 tiles with 16-byte accesses when the
// shape allows, 32x32 tiles with 4-byte accesses otherwise; loads and stores are coalesced either way).  Used by the Python shim to feed the reference's
// NCHW FPN maps (src/components/fpn.py:38-55 output) to the NHWC RoIAlign fast path.
#include "common.cuh"

namespace lcr {

// in: [batch][rows][cols] -> out: [batch][cols][rows]
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ty + j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + j][tx] = __ldg(in + base + (size_t)r * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, r = r0 + tx;
    if (r < rows && c < cols) out[base + (size_t)c * rows + r] = tile[tx][ty + j];
  }
}

// Same transpose with 16-byte global accesses on both sides: 64x64 tiles, 256 threads, a thread loads four float4 along
// a row and stores four float4 along a column (rows, cols multiples of 4, 16-byte aligned bases) - four times fewer
// memory instructions and 16 KB in flight per CTA.
__global__ void __launch_bounds__(256) transpose_v4_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[64][65];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 float4 columns x 16 rows per pass
#pragma unroll
  for (int j = 0; j < 64; j += 16) {
    const int r = r0 + ty + j, c = c0 + tx * 4;
    if (r < rows && c < cols) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in + base + (size_t)r * cols + c));
      tile[ty + j][tx * 4 + 0] = v.x;
      tile[ty + j][tx * 4 + 1] = v.y;
      tile[ty + j][tx * 4 + 2] = v.z;
      tile[ty + j][tx * 4 + 3] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 64; j += 16) {
    const int c = c0 + ty + j, r = r0 + tx * 4;  // output row c, output columns r .. r+3
    if (c < cols && r < rows) {
      const float4 v = make_float4(tile[tx * 4 + 0][ty + j], tile[tx * 4 + 1][ty + j], tile[tx * 4 + 2][ty + j], tile[tx * 4 + 3][ty + j]);
      *reinterpret_cast<float4*>(out + base + (size_t)c * rows + r) = v;
    }
  }
}

static int launch_transpose(const float* in, float* out, int batch, int rows, int cols, cudaStream_t st) {
  if (rows % 4 == 0 && cols % 4 == 0 && aligned_to(in, 16) && aligned_to(out, 16)) {
    dim3 grid4((cols + 63) / 64, (rows + 63) / 64, batch);
    if (grid4.y <= 65535 && grid4.z <= 65535) {
      transpose_v4_kernel<<<grid4, 256, 0, st>>>(in, out, rows, cols);
      return after_launch();
    }
  }
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  if (grid.y > 65535 || grid.z > 65535) return LCR_ERR_CAPACITY;
  transpose_kernel<<<grid, block, 0, st>>>(in, out, rows, cols);
  return after_launch();
}

}  // namespace lcr

using namespace lcr;

extern "C" int lcr_nchw_to_nhwc_f32(const float* in, float* out, int N, int C, int H, int W, void* stream) {
  LCR_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0, LCR_ERR_INVALID_ARG);
  return launch_transpose(in, out, N, C, H * W, as_stream(stream));
}

extern "C" int lcr_nhwc_to_nchw_f32(const float* in, float* out, int N, int C, int H, int W, void* stream) {
  LCR_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0, LCR_ERR_INVALID_ARG);
  return launch_transpose(in, out, N, H * W, C, as_stream(stream));
}
