// Skinny-M GEMM for the weight-streaming regime (M = nb*B*60 <= 256 rows, i.e. one clip per GPU):
//     out[M, N] = epilogue( A[M, K] . W[N, K]^T )
// computed TRANSPOSED on the tensor core: the weight tile is the UMMA "A" operand (M_umma = 128 output features), the
// whole activation matrix is the UMMA "B" operand (N_umma = Mpad rows), so one CTA owns 128 features x ALL rows and the
// weights - the only HBM traffic that matters at this size - are read exactly once by exactly one SM.
//   * grid = (N/128 feature tiles) x (n_ks K-splits); each CTA streams its 128 x (K/n_ks) weight slab through a TMA ring
//   * clusters of CL CTAs (consecutive feature tiles, same K-split) share the activation K-blocks: CTA r loads rows
//     [r*Mpad/CL, (r+1)*Mpad/CL) and TMA-multicasts them to every CTA of the cluster -> activation L2 traffic / CL
//   * K-split partials are reduced with fp32 RED atomics in L2: directly into the residual stream for the gate+residual
//     epilogue (x += gate*(.)), or into a zeroed scratch tile that the last-arriving CTA finalises (bias / GELU / cast)
//   * PDL: the weight prefetch of the first ring stages is issued BEFORE griddepcontrol.wait, so it overlaps the tail
//     of the previous kernel in the captured graph
// TMEM lanes = features, TMEM columns = rows: an epilogue warp reads 32 rows x 32 features and every global access is
// coalesced across the warp's 32 consecutive features.
#pragma once
#include "gemm.cuh"

namespace fmt {

struct SkinnyParams {
  EpiParams ep;        // ep.M = valid rows, ep.N = features
  int K;
  int Mpad;            // UMMA N: rows padded to a multiple of max(16, 8*CL), <= 256
  int n_ft;            // N / 128
  int n_ks;            // K splits
  int kb_per_split;    // 64-wide K blocks per split
  float* scratch;      // (M, N) fp32, all zero between launches (n_ks > 1 and kind != EPI_GATE_RES)
  int* counters;       // (n_ft) tickets, all zero between launches
};

constexpr int SK_STAGES = 4;
constexpr int SK_THREADS = 192;
constexpr int SK_W_BYTES = 128 * 64 * 2;
__host__ __device__ constexpr int sk_smem_bytes(int Mpad) { return SK_STAGES * (SK_W_BYTES + Mpad * 128) + 256 + 1024; }

// single-element epilogue (feature n of row m), value v already includes the bias where applicable
template <typename TT>
__device__ __forceinline__ void sk_store(const EpiParams& p, int m, int n, float v) {
  if (p.kind == EPI_GELU) v = gelu_tanh(v);
  else if (p.kind == EPI_POS) v += p.pos[static_cast<size_t>(m % p.frames) * p.N + n];
  if (p.out_f32) reinterpret_cast<float*>(p.out)[static_cast<size_t>(m) * p.ldo + n] = v;
  else reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<size_t>(m) * p.ldo + n] = __float2bfloat16_rn(v);
}
template <typename TT>
__device__ __forceinline__ float sk_gate(const EpiParams& p, int m, int n) {
  const int ur = p.urow ? p.urow[m] : m;
  return to_f32<TT>(reinterpret_cast<const TT*>(p.gate)[static_cast<size_t>(ur) * p.ldg + p.gate_off + n]);
}

template <typename TT>
__global__ void __launch_bounds__(SK_THREADS, 1)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const SkinnyParams sp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int a_bytes = sp.Mpad * 128;
  const int stage_bytes = SK_W_BYTES + a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SK_STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + SK_STAGES;
  uint64_t* tfull_bar = empty_bar + SK_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t CL = cluster_nctarank(), cr = cluster_ctarank();
  const uint16_t mc_mask = static_cast<uint16_t>((1u << CL) - 1);
  const int ft = blockIdx.x % sp.n_ft, ks = blockIdx.x / sp.n_ft;
  const int n0 = ft * 128;
  const int kb0 = ks * sp.kb_per_split, nkb = sp.kb_per_split;
  const EpiParams& ep = sp.ep;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(sp.Mpad)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SK_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();                 // let the next graph node start its own prologue / weight prefetch

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int slice_rows = sp.Mpad / static_cast<int>(CL);
      const int npre = nkb < SK_STAGES ? nkb : SK_STAGES;
      // weights do not depend on the previous kernel: prefetch them before the grid dependency resolves
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(&full_bar[i], stage_bytes);
        tma_load_2d(&tmW, &full_bar[i], smem + i * stage_bytes, (kb0 + i) * 64, n0, kEvictFirst);
      }
      pdl_wait_prior_grid();
      for (int i = 0; i < npre; ++i) {
        uint8_t* sa = smem + i * stage_bytes + SK_W_BYTES + cr * slice_rows * 128;
        if (CL > 1) tma_load_2d_mc(&tmA, &full_bar[i], sa, (kb0 + i) * 64, cr * slice_rows, mc_mask, kEvictLast);
        else tma_load_2d(&tmA, &full_bar[i], sa, (kb0 + i) * 64, 0, kEvictLast);
      }
      int stage = npre % SK_STAGES; uint32_t phase = (npre == SK_STAGES) ? 1 : 0;
      for (int kb = npre; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);      // every CTA of the cluster has consumed this stage
        uint8_t* sw = smem + stage * stage_bytes;
        mbar_expect_tx(&full_bar[stage], stage_bytes);
        tma_load_2d(&tmW, &full_bar[stage], sw, (kb0 + kb) * 64, n0, kEvictFirst);
        uint8_t* sa = sw + SK_W_BYTES + cr * slice_rows * 128;
        if (CL > 1) tma_load_2d_mc(&tmA, &full_bar[stage], sa, (kb0 + kb) * 64, cr * slice_rows, mc_mask, kEvictLast);
        else tma_load_2d(&tmA, &full_bar[stage], sa, (kb0 + kb) * 64, 0, kEvictLast);
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(128, static_cast<uint32_t>(sp.Mpad));
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sw = smem_u32(smem + stage * stage_bytes);
        const uint64_t dw = make_sw128_kmajor_desc(sw);
        const uint64_t da = make_sw128_kmajor_desc(sw + SK_W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dw + 2 * k, da + 2 * k, idesc, (kb | k) != 0);
        if (CL > 1) umma_commit_mc(&empty_bar[stage], mc_mask);
        else umma_commit(&empty_bar[stage]);
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM lanes = features, columns = rows =====================
    pdl_wait_prior_grid();                             // x / gate / scratch are produced by earlier kernels
    const int quarter = warp & 3;
    const int n = n0 + quarter * 32 + lane;
    const bool split = sp.n_ks > 1;
    const float bias = (ep.bias != nullptr && (!split || (ep.kind == EPI_GATE_RES && ks == 0))) ? ep.bias[n] : 0.f;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    for (int c = 0; c < (sp.Mpad + 31) / 32; ++c) {
      if (c * 32 >= ep.M) break;
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (ep.kind == EPI_GATE_RES) {
        float* x = reinterpret_cast<float*>(ep.out);
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int m = c * 32 + j;
          if (m < ep.M) {
            const float g = sk_gate<TT>(ep, m, n);
            float* xp = x + static_cast<size_t>(m) * ep.ldo + n;
            if (split) atomicAdd(xp, g * (v[j] + bias));        // RED.ADD.F32 in L2; summation order is not fixed
            else *xp = fmaf(g, v[j] + bias, *xp);
          }
        }
      } else if (split) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int m = c * 32 + j;
          if (m < ep.M) atomicAdd(sp.scratch + static_cast<size_t>(m) * ep.N + n, v[j]);
        }
      } else {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int m = c * 32 + j;
          if (m < ep.M) sk_store<TT>(ep, m, n, v[j] + bias);
        }
      }
    }
    if (split && ep.kind != EPI_GATE_RES) {
      // split-K fix-up: the last CTA of this feature tile reads the reduced tile back, applies the epilogue and
      // leaves scratch / ticket zeroed for the next launch
      __threadfence();
      named_bar_sync(1, 128);
      if (threadIdx.x == 64) *s_last = (atomicAdd(sp.counters + ft, 1) == sp.n_ks - 1);
      named_bar_sync(1, 128);
      if (*s_last) {
        __threadfence();
        const float b = ep.bias != nullptr ? ep.bias[n] : 0.f;
#pragma unroll 4
        for (int m = 0; m < ep.M; ++m) {
          float* sp_ = sp.scratch + static_cast<size_t>(m) * ep.N + n;
          const float s = __ldcg(sp_);
          *sp_ = 0.f;
          sk_store<TT>(ep, m, n, s + b);
        }
        if (threadIdx.x == 64) sp.counters[ft] = 0;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // nobody exits while a peer may still multicast / arrive into this CTA
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace fmt
