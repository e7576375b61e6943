// layout.cu — NCHW <-> NHWC tiled transposes (32x32 tiles through padded shared memory; both the
// loads and the stores are 128-byte coalesced).  Used by the Python shim to feed the reference's
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

static int launch_transpose(const float* in, float* out, int batch, int rows, int cols, cudaStream_t st) {
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
