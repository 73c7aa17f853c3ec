// Microbenchmark: cost of one "stage" of a persistent kernel on B200 = stores + fence + grid barrier + dependent loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gridbar gridbar.cu && ./gridbar
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned* p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

// mode bits: 1 = each thread stores a float4 before the barrier; 2 = fence.proxy.async (all); 4 = fence.proxy.async.global;
// 8 = __threadfence by all threads; 16 = after the barrier each thread loads (ld.cg) what a thread of the next CTA stored;
// 32 = second dependent load; 64 = relaxed red + explicit fence by thread 0 instead of release
__global__ void __launch_bounds__(256, 1) k(unsigned* counter, float4* buf, long long* out, int iters, int mode) {
  const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
  float4 acc = make_float4(0, 0, 0, 0);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    if (mode & 1) buf[(size_t)cta * 256 + tid] = make_float4(it, tid, cta, acc.x);
    if (mode & 2) asm volatile("fence.proxy.async;" ::: "memory");
    if (mode & 4) asm volatile("fence.proxy.async.global;" ::: "memory");
    if (mode & 8) __threadfence();
    __syncthreads();
    if (mode & 1024) {
      // all-to-all flags: every CTA publishes its epoch in its own word; G threads of every CTA poll one flag each
      const int stride = (mode & 2048) ? 32 : 1;
      unsigned* flags = counter + 256;
      if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + cta * stride), "r"((unsigned)it) : "memory");
      if (tid < G) { while (ld_relaxed(flags + tid * stride) < (unsigned)it) {} }
      __threadfence();
    } else if (tid == 0 && (mode & 128)) {
      // arrive with a returning atomic; the LAST arriver publishes the epoch in a separate flag line that everyone polls
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
      if (old == (unsigned)it * G - 1) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(counter + 32), "r"((unsigned)it) : "memory"); }
      else {
        while (ld_acquire(counter + 32) < (unsigned)it) { if (mode & 512) __nanosleep(32); }
      }
    } else if (tid == 0 && (mode & 256)) {
      // two-level: 8 group counters (CTA % 8), last of a group bumps the top counter, last of all publishes the flag
      unsigned old; unsigned* gc = counter + 64 + 32 * (cta & 7);
      const unsigned gsize = (G - (cta & 7) + 7) / 8;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(gc) : "memory");
      bool pub = false;
      if (old == (unsigned)it * gsize - 1) {
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        pub = (old == (unsigned)it * 8 - 1);
      }
      if (pub) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(counter + 32), "r"((unsigned)it) : "memory"); }
      else { while (ld_acquire(counter + 32) < (unsigned)it) {} }
    } else if (tid == 0) {
      if (mode & 64) { __threadfence(); red_relaxed(counter); while (ld_relaxed(counter) < (unsigned)it * G) {} __threadfence(); }
      else { red_release(counter); while (ld_acquire(counter) < (unsigned)it * G) {} }
    }
    __syncthreads();
    if (mode & 16) {
      float4 v = __ldcg(&buf[(size_t)((cta + 1) % G) * 256 + tid]);
      acc.x += v.x;
      if (mode & 32) { float4 w = __ldcg(&buf[(size_t)((cta + 7 + ((int)v.y & 1)) % G) * 256 + tid]); acc.y += w.x; }
    }
  }
  const long long t1 = clock64();
  if (tid == 0) { out[cta * 2] = t1 - t0; out[cta * 2 + 1] = (long long)(acc.x + acc.y); }
}

int main() {
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  const int G = prop.multiProcessorCount, iters = 2000;
  unsigned* counter; float4* buf; long long* out;
  cudaMalloc(&counter, 65536); cudaMalloc(&buf, (size_t)G * 256 * 16); cudaMalloc(&out, G * 16);
  cudaMemset(buf, 0, (size_t)G * 256 * 16);
  int modes[] = {0, 1024, 1024 | 2048, 1 | 4 | 16, 1024 | 1 | 4 | 16, 1024 | 2048 | 1 | 4 | 16};
  for (int m : modes) {
    cudaMemset(counter, 0, 65536);
    void* args[] = {&counter, &buf, &out, (void*)&iters, &m};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(G), dim3(256), args, 0, 0);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("mode %3d: %8.1f cycles/iter (%.2f us @1.9GHz)  err=%s\n", m, (double)h[0] / iters, (double)h[0] / iters / 1900.0, cudaGetErrorString(e));
  }
  return 0;
}
