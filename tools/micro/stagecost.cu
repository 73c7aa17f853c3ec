// Microbenchmark: one elementwise "stage" of the persistent kernel (GELU-like): per thread NB float4 ld.cg from an L2-resident
// buffer written by other CTAs, compute, bf16 store + zero store, proxy fence, grid barrier.  Prints cycles per stage.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

template <int NB>
__global__ void __launch_bounds__(288, 1) k(unsigned* counter, float* acc, uint2* out, long long* tout, int iters, int mode, int total4) {
  const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
  if (tid < 32) return;                      // warp 0 = weight producer in the real kernel
  const unsigned stride = G * 256;
  long long t_ld = 0, t_st = 0, t_fence = 0, t_bar = 0;
  asm volatile("bar.sync 1, 256;");
  const long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    long long c0 = clock64();
    float4 a[NB];
    const unsigned i0 = cta * 256 + (tid - 32);
#pragma unroll
    for (int k2 = 0; k2 < NB; ++k2) { const unsigned i = i0 + k2 * stride; if (i < total4) a[k2] = __ldcg(reinterpret_cast<const float4*>(acc) + ((i + it * 977u) % total4)); }
    float s = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < NB; ++k2) { const unsigned i = i0 + k2 * stride; if (i < total4) s += a[k2].x + a[k2].y + a[k2].z + a[k2].w; }
    long long c1 = clock64();
#pragma unroll
    for (int k2 = 0; k2 < NB; ++k2) {
      const unsigned i = i0 + k2 * stride;
      if (i < total4) {
        if (mode & 1) out[i] = make_uint2(__float_as_uint(s), it);
        if (mode & 2) reinterpret_cast<float4*>(acc)[i] = make_float4(s * 1e-30f, 0.f, 0.f, 0.f);
      }
    }
    long long c2 = clock64();
    if (mode & 4) asm volatile("fence.proxy.async.global;" ::: "memory");
    long long c3 = clock64();
    asm volatile("bar.sync 1, 256;");
    if (tid == 32) { red_release(counter); while (ld_acquire(counter) < (unsigned)it * G) {} }
    asm volatile("bar.sync 1, 256;");
    long long c4 = clock64();
    t_ld += c1 - c0; t_st += c2 - c1; t_fence += c3 - c2; t_bar += c4 - c3;
  }
  const long long t1 = clock64();
  if (tid == 32 && cta == 0) { tout[0] = t1 - t0; tout[1] = t_ld; tout[2] = t_st; tout[3] = t_fence; tout[4] = t_bar; }
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int G = prop.multiProcessorCount, iters = 1000, total4 = 180 * 4096 / 4;
  unsigned* counter; float* acc; uint2* out; long long* tout;
  cudaMalloc(&counter, 4); cudaMalloc(&acc, (size_t)total4 * 16); cudaMalloc(&out, (size_t)total4 * 8); cudaMalloc(&tout, 64);
  cudaMemset(acc, 0, (size_t)total4 * 16);
  for (int mode : {0, 1, 3, 7}) {
    cudaMemset(counter, 0, 4);
    int it = iters, t4 = total4;
    void* args[] = {&counter, &acc, &out, &tout, &it, &mode, &t4};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k<5>, dim3(G), dim3(288), args, 0, 0);
    cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, tout, 40, cudaMemcpyDeviceToHost);
    printf("mode %d: %7.0f cyc/stage  ld %6.0f  st %6.0f  fence %6.0f  bar %6.0f   (%s)\n", mode, (double)h[0] / iters, (double)h[1] / iters,
           (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, cudaGetErrorString(e));
  }
  return 0;
}
