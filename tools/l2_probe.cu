// l2_probe.cu — what can the SMs pull out of L2?  (measurement tool, not product code)
//
// RoIAlign forward gathers ~11 GB of overlapping feature windows per bench step out of a 23 MB/frame map that is
// L2-resident (DRAM reads are only 1.5 GB): its bound is the L2 -> SM path, not HBM.  This probe measures that path on
// the box it runs on, so that the kernel's gather rate can be quoted against a measured ceiling:
//   ldg   : every warp streams 256-byte rows (LDG.64 per lane, UNROLL loads in flight) over an L2-resident buffer
//   bulk  : one thread per CTA pulls CHUNK-byte pieces with cp.async.bulk (TMA engine, UBLKCP) into a shared-memory ring
// Output: one JSON line per configuration.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_probe tools/l2_probe.cu && tools/l2_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));               \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

template <int UNROLL>
__global__ void ldg_kernel(const float2* __restrict__ buf, size_t n_f2, int reps, float* sink) {
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    for (size_t i = tid; i + (UNROLL - 1) * nthreads < n_f2; i += UNROLL * nthreads) {
      float2 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = __ldcg(buf + i + u * nthreads);   // L2 only: no L1 allocation
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y;
    }
  }
  if (acc == 123.456f) *sink = acc;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
          (uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

// one issuing thread per CTA, SLOTS chunks in flight, nobody reads the data (pure transfer rate)
template <int SLOTS>
__global__ void bulk_kernel(const char* __restrict__ buf, size_t bytes, int chunk, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[SLOTS];
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t nchunks = bytes / chunk;
    size_t issued = 0;
    uint32_t parity[SLOTS];
    for (int s = 0; s < SLOTS; ++s) parity[s] = 0;
    for (int r = 0; r < reps; ++r)
      for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++issued) {
        const int s = (int)(issued % SLOTS);
        if (issued >= SLOTS) {
          mbar_wait(&bars[s], parity[s]);
          parity[s] ^= 1u;
        }
        mbar_expect_tx(&bars[s], (uint32_t)chunk);
        bulk_load(smem + (size_t)s * chunk, buf + c * (size_t)chunk, (uint32_t)chunk, &bars[s]);
      }
    const size_t pending = issued < SLOTS ? issued : SLOTS;
    for (size_t q = 0; q < pending; ++q) {
      const int s = (int)((issued - pending + q) % SLOTS);
      mbar_wait(&bars[s], parity[s]);
      parity[s] ^= 1u;
    }
  }
}

template <typename F>
static float timed(F launch, int iters) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / iters;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float* sink;
  CK(cudaMalloc(&sink, 4));
  const int reps = 20;
  for (size_t mb : {24, 48, 96, 512}) {   // 24/48 MB: L2-resident; 96: partly; 512: HBM
    const size_t bytes = mb << 20;
    char* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    for (int tpb : {128, 256}) {
      for (int cps : {4, 8}) {
        const float ms = timed([&] { ldg_kernel<8><<<sms * cps, tpb>>>((const float2*)buf, bytes / 8, reps, sink); }, 3);
        printf("{\"probe\": \"ldg.cg f2 x8\", \"buffer_MB\": %zu, \"threads_per_sm\": %d, \"TBps\": %.2f}\n", mb, tpb * cps,
               (double)bytes * reps / (ms * 1e-3) / 1e12);
      }
    }
    for (int chunk : {4096, 16384}) {
      for (int cps : {2, 4}) {
        const int slots = 4;
        auto kern = bulk_kernel<4>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, slots * chunk));
        const float ms = timed([&] { kern<<<sms * cps, 32, slots * chunk>>>(buf, bytes, chunk, reps); }, 3);
        printf("{\"probe\": \"cp.async.bulk\", \"buffer_MB\": %zu, \"chunk\": %d, \"ctas_per_sm\": %d, \"in_flight_per_sm_KB\": %d, \"TBps\": %.2f}\n",
               mb, chunk, cps, cps * slots * chunk / 1024, (double)bytes * reps / (ms * 1e-3) / 1e12);
      }
    }
    CK(cudaFree(buf));
  }
  return 0;
}
