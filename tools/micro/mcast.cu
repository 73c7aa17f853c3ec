// Microbenchmark: does TMA multicast inside a thread-block cluster shorten the fetch of a FRESH activation tile?
// Every iteration: (1) CTA i overwrites the 32 KB tile that another cluster will read (so the lines are fresh, written by a remote
// SM a moment ago - the situation of the window kernels), (2) grid-wide rendezvous, (3) every cluster fetches ITS tile into the
// shared memory of all its CTAs - mode 0: each CTA copies the whole tile itself (what the dataflow kernel does today), mode 1: each
// CTA copies 1/C of it with .multicast::cluster to all C CTAs.  Reported: time from the rendezvous to the tile being complete in a
// CTA's shared memory (median / max over CTAs, mean over iterations).  Cluster sizes 1, 2, 4; 1, 8 or 32 CTAs read the same tile (a
// K slice of the activations is read by every feature tile of the GEMM); cooperative launch, one CTA per SM.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

constexpr int TILE = 32 * 1024;
__global__ void __launch_bounds__(128, 1) k(float* tiles, unsigned* counter, int mode, int iters, long long* out, int* err, int share) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  cg::cluster_group cl = cg::this_cluster();
  const int C = cl.num_blocks(), rank = cl.block_rank(), cluster = blockIdx.x / C, n_clusters = gridDim.x / C;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  cl.sync();
  long long acc = 0;
  uint32_t phase = 0;
  for (int it = 1; it <= iters; ++it) {
    // (1) freshen the tile of the cluster "opposite" to this one: each of the C CTAs writes its 1/C share
    {
      const int victim = (cluster + n_clusters / 2 + it) % n_clusters;
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(tiles) + (size_t)victim * TILE + (size_t)rank * (TILE / C));
      for (int i = threadIdx.x; i < TILE / C / 16; i += blockDim.x) dst[i] = make_float4(it, i, rank, cluster);
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncthreads();
    // (2) grid-wide rendezvous
    if (threadIdx.x == 0) {
      red_rel(counter);
      const long long t0 = clock64();
      while (ld_acq(counter) < (unsigned)it * gridDim.x) { if (clock64() - t0 > 2000000000ll) { *err = 1; break; } }
      asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    __syncthreads();
    if (C > 1) cl.sync();                                    // every CTA of the cluster has armed nothing yet; keep them in step
    // (3) fetch
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      mbar_expect(&bar, TILE);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(tiles) + (size_t)(cluster / share) * TILE;    // `share` clusters read the same tile
      if (mode == 0 || C == 1) {
        bulk_load(smem, src, TILE, &bar);
      } else {
        const int share = TILE / C;
        bulk_load_mc(smem + rank * share, src + rank * share, share, &bar, (uint16_t)((1u << C) - 1));
      }
      const long long t0 = clock64();
      while (!mbar_try(&bar, phase)) { if (clock64() - t0 > 2000000000ll) { *err = 2; break; } }
      acc += clock64() - t1;
    }
    phase ^= 1;
    __syncthreads();
    if (C > 1) cl.sync();                                    // nobody re-arms or overwrites while a peer's multicast may still land
  }
  if (threadIdx.x == 0) out[blockIdx.x] = acc / iters;
}
int main() {
  const int iters = 200, grid = 148;
  float* tiles; unsigned* counter; long long* out; int* err;
  cudaMalloc(&tiles, (size_t)grid * TILE); cudaMalloc(&counter, 4); cudaMalloc(&out, grid * 8); cudaMalloc(&err, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("%-8s %-10s | %9s %9s   (us from the rendezvous until the 32 KB tile is complete in shared memory; 148 CTAs)\n", "cluster", "mode", "median", "max");
  for (int share_ctas : {1, 8, 32})
  for (int C : {1, 2, 4})
    for (int mode : {0, 1}) {
      if (C == 1 && mode == 1) continue;
      const int share = share_ctas / C > 0 ? share_ctas / C : 1;      // clusters per tile, so that share_ctas CTAs read the same tile
      cudaMemset(counter, 0, 4); cudaMemset(err, 0, 4);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = TILE;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
      at[1].id = cudaLaunchAttributeClusterDimension; at[1].val.clusterDim.x = C; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 2;
      cudaError_t e = cudaLaunchKernelEx(&cfg, k, tiles, counter, mode, iters, out, err, share);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%-8d %-10s | launch failed: %s\n", C, mode ? "multicast" : "unicast", cudaGetErrorString(e)); cudaGetLastError(); continue; }
      int herr = 0; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("%-8d %-10s | %9.2f %9.2f   %2d CTAs read the same tile %s\n", C, mode ? "multicast" : "unicast", h[grid / 2] / 1.9e3, h[grid - 1] / 1.9e3, share * C, herr ? "(a bounded spin tripped)" : "");
    }
  return 0;
}
