// Microbenchmark: how fast one SM pushes a partial-sum tile out to L2, by path, with all SMs doing it at once.
//   mode 0: cp.reduce.async.bulk (TMA reduce-add f32, 1D, 16 KB pieces) from shared memory, wait for completion
//   mode 1: cp.async.bulk store (TMA plain store), same pieces
//   mode 2: st.global.v4.f32 from registers, each warp instruction writes 512 contiguous bytes, then membar.gl
//   mode 3: st.global.v8.f32 (256-bit), 1 KB per warp instruction
//   mode 4: red.global.add.v4.f32 (LSU vector reds), 512 B per warp instruction
//   mode 5: st.global.f32 scalar coalesced (128 B per warp instruction)
//   mode 6: cp.reduce.async.bulk .add.s32 (the deterministic fixed-point split-K)
// `share` CTAs target the same region (split-K partials of one tile meet there); share = 1: every CTA has its own slot.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_reduce_add(float* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_s32(float* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.s32 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store(float* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(512, 1) k(float* dst, int bytes, int share, int mode, int iters, int nthreads, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  for (int i = tid; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1e-3f * (i & 15);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  float* base = dst + (size_t)(blockIdx.x / share) * (bytes / 4);
  long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    const long long t0 = clock64();
    if (mode <= 1 || mode == 6) {
      if (tid == 0) {
        for (int o = 0; o < bytes; o += 16384) {
          const int n = min(16384, bytes - o);
          if (mode == 0) bulk_reduce_add(base + o / 4, smem + o, n); else if (mode == 6) bulk_reduce_add_s32(base + o / 4, smem + o, n); else bulk_store(base + o / 4, smem + o, n);
          bulk_commit();
        }
        bulk_wait0();
      }
    } else if (tid < nthreads) {
      const float4 v = make_float4(1e-3f, 2e-3f, 3e-3f, 4e-3f);
      if (mode == 2) {
        for (int o = tid * 16; o < bytes; o += nthreads * 16) *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(base) + o) = v;
      } else if (mode == 3) {
        for (int o = tid * 32; o < bytes; o += nthreads * 32)
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(reinterpret_cast<uint8_t*>(base) + o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      } else if (mode == 4) {
        for (int o = tid * 16; o < bytes; o += nthreads * 16)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<uint8_t*>(base) + o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      } else {
        for (int o = tid * 4; o < bytes; o += nthreads * 4) *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(base) + o) = v.x;
      }
      __threadfence();
    }
    __syncthreads();
    acc += clock64() - t0;
  }
  if (tid == 0) out[blockIdx.x] = acc / iters;
}
int main(int argc, char** argv) {
  const int iters = 50;
  float* dst; long long* out;
  cudaMalloc(&dst, (size_t)148 * 256 * 1024); cudaMalloc(&out, 148 * 8);
  cudaMemset(dst, 0, (size_t)148 * 256 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const char* names[] = {"TMA reduce-add", "TMA store", "st.v4", "st.v8", "red.v4", "st.f32", "TMA red s32"};
  printf("%-15s %6s %5s %5s %4s | %9s %9s %9s\n", "path", "KB/SM", "share", "thr", "CTAs", "us median", "us max", "GB/s/SM");
  for (int bytes : {48 * 1024, 96 * 1024})
    for (int mode : {0, 6, 1})
      for (int share : {1, 6})
        for (int nthreads : {128, 256, 512}) {
          if ((mode <= 1 || mode == 6) && nthreads != 128) continue;
          if ((mode == 2 || mode == 3 || mode == 5) && share != 1) continue;
          for (int grid : {148, 32}) {
            for (int rep = 0; rep < 2; ++rep) k<<<grid, 512, 128 * 1024>>>(dst, bytes, share, mode, iters, nthreads, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            std::vector<long long> h(grid);
            cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost);
            std::sort(h.begin(), h.end());
            const double clk = 1.9e3;  // clocks per us (approximate)
            printf("%-15s %6d %5d %5d %4d | %9.2f %9.2f %9.1f\n", names[mode], bytes / 1024, share, nthreads, grid, h[grid / 2] / clk, h[grid - 1] / clk,
                   bytes / (h[grid / 2] / clk) / 1e3);
          }
        }
  return 0;
}
