// Microbenchmark: per-SM L2 -> shared-memory ingest rate of TMA bulk copies as a function of how many CTAs pull at once.
// mode 0: every CTA streams the SAME 384 KB region (activation-like, L2 resident); mode 1: every CTA streams its own region
// (weight-like, L2 resident after the warm-up pass).  4-deep ring of 24 KB chunks, one producer thread, one consumer thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
constexpr int CHUNK = 24 * 1024, STAGES = 4;
__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, long long stride, int n_chunks, int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* base = src + blockIdx.x * stride;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it)
      for (int c = 0; c < n_chunks; ++c) {
        while (!mbar_try(&empty[s], ph ^ 1)) {}
        mbar_expect(&full[s], CHUNK);
        bulk_load(smem + s * CHUNK, base + (size_t)c * CHUNK, CHUNK, &full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it)
      for (int c = 0; c < n_chunks; ++c) {
        while (!mbar_try(&full[s], ph)) {}
        mbar_arrive(&empty[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
  const int n_chunks = 16, iters = 20;   // 384 KB per pass
  uint8_t* buf; long long* out;
  const size_t region = (size_t)n_chunks * CHUNK;
  cudaMalloc(&buf, region * 148); cudaMalloc(&out, 148 * 8);
  cudaMemset(buf, 1, region * 148);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
  for (int mode = 0; mode < 2; ++mode)
    for (int g : {1, 8, 24, 32, 48, 64, 96, 148}) {
      for (int rep = 0; rep < 2; ++rep) k<<<g, 64, STAGES * CHUNK>>>(buf, mode ? (long long)region : 0ll, n_chunks, iters, out);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, out, g * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      const double us = mx / 1.9e3, bytes = (double)region * iters;
      printf("mode %d (%s) CTAs %3d: %7.1f us per %d x 384 KB -> %6.1f GB/s per SM, %6.2f TB/s aggregate  (%s)\n", mode, mode ? "own region " : "same region", g, us,
             iters, bytes / us / 1e3, bytes * g / us / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
