// GEMMs of the FMT step:  out[M,N] = epilogue( A[M,K] . W[N,K]^T )
//   * gemm_tc_kernel   - bf16 operands, tcgen05.mma (cta_group::1, M=128 x N=BN x K=16) with fp32 accumulators in
//                        TMEM, operands staged by TMA (128B swizzle) through an mbarrier ring; persistent over
//                        output tiles with a double-buffered accumulator so the epilogue of tile i overlaps the
//                        MMAs of tile i+1.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue.
//   * gemm_simt_kernel - fp32 operands, plain FFMA tiles; FMT_MODE_FP32_VALIDATE only (the 1e-4 parity mode).
// Both share epi_apply(), which carries the fused epilogues of the reference's block
// (FMT.py:171-176): bias, GELU(tanh), +pos_embed, and x += gate * (.).
#pragma once
#include "ptx.cuh"

namespace fmt {

enum EpiKind : int { EPI_STORE = 0, EPI_GELU = 1, EPI_POS = 2, EPI_GATE_RES = 3 };

struct EpiParams {
  int kind;
  int M, N;
  const float* bias;   // [N] or nullptr
  void* out;           // [M, ldo]: float if out_f32 else the activation type AT
  int ldo;
  int out_f32;
  const float* pos;    // EPI_POS: out += pos[(m % frames) * N + n]
  int frames;
  const void* gate;    // EPI_GATE_RES: table (type TT), element (urow[m] * ldg + gate_off + n)
  const int* urow;     // row indirection into the table (nullptr = identity)
  long long ldg;
  long long gate_off;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // nn.GELU(approximate="tanh"), FMT.py:161
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, int NC> struct VecIO;
template <int NC> struct VecIO<float, NC> {
  static_assert(NC % 4 == 0, "");
  static __device__ __forceinline__ void load(const float* p, float (&v)[NC]) {
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[NC]) {
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
};
template <int NC> struct VecIO<__nv_bfloat16, NC> {
  static_assert(NC % 4 == 0, "");
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[NC]) {
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      uint2 t = reinterpret_cast<const uint2*>(p)[i];
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
      v[4 * i] = __low2float(a); v[4 * i + 1] = __high2float(a); v[4 * i + 2] = __low2float(b); v[4 * i + 3] = __high2float(b);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[NC]) {
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v[4 * i], v[4 * i + 1]), b = __floats2bfloat162_rn(v[4 * i + 2], v[4 * i + 3]);
      uint2 t;
      t.x = *reinterpret_cast<uint32_t*>(&a);
      t.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(p)[i] = t;
    }
  }
};

// One thread owns NC consecutive columns [n0, n0+NC) of row m (n0 % NC == 0, N % NC == 0).
template <typename AT, typename TT, int NC>
__device__ __forceinline__ void epi_apply(const EpiParams& p, int m, int n0, float (&v)[NC]) {
  if (p.bias != nullptr) {
    float b[NC];
    VecIO<float, NC>::load(p.bias + n0, b);
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] += b[i];
  }
  if (p.kind == EPI_GELU) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = gelu_tanh(v[i]);
  } else if (p.kind == EPI_POS) {
    float b[NC];
    VecIO<float, NC>::load(p.pos + static_cast<size_t>(m % p.frames) * p.N + n0, b);
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] += b[i];
  } else if (p.kind == EPI_GATE_RES) {
    const int ur = p.urow ? p.urow[m] : m;
    float g[NC], x[NC];
    VecIO<TT, NC>::load(reinterpret_cast<const TT*>(p.gate) + static_cast<size_t>(ur) * p.ldg + p.gate_off + n0, g);
    float* xp = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.ldo + n0;
    VecIO<float, NC>::load(xp, x);
#pragma unroll
    for (int i = 0; i < NC; ++i) x[i] = fmaf(g[i], v[i], x[i]);     // x = x + gate * branch  (FMT.py:174-175)
    VecIO<float, NC>::store(xp, x);
    return;
  }
  if (p.out_f32) {
    VecIO<float, NC>::store(reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.ldo + n0, v);
  } else {
    VecIO<AT, NC>::store(reinterpret_cast<AT*>(p.out) + static_cast<size_t>(m) * p.ldo + n0, v);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMA GEMM
// ------------------------------------------------------------------------------------------------
template <int BN> struct TcCfg {
  static constexpr int BM = 128, BK = 64, UMMA_K = 16;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;     // 128 / 256 / 512: powers of two >= 32
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // +1024: manual 1 KB alignment for SWIZZLE_128B
  static constexpr int THREADS = 192;
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
};

template <int BN, typename TT>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams ep, const int K) {
  using C = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + C::ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + C::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + C::BK - 1) / C::BK;
  const int m_tiles = (ep.M + C::BM - 1) / C::BM;
  const int n_tiles = (ep.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < C::ACC_STAGES; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      pdl_wait_prior_grid();
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t % m_tiles) * C::BM, n0 = (t / m_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_2d(&tmA, &full_bar[stage], sa, kb * C::BK, m0, kEvictNormal);
          tma_load_2d(&tmB, &full_bar[stage], sa + C::A_BYTES, kb * C::BK, n0, kEvictNormal);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(C::BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);        // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);              // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da = make_sw128_kmajor_desc(sa);
          const uint64_t db = make_sw128_kmajor_desc(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < C::BK / C::UMMA_K; ++k) {
            // advance along K inside the 128B swizzle atom: +32 bytes (>>4 = 2) per UMMA_K
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);                      // accumulator complete -> epilogue
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    const int quarter = warp & 3;                          // a warp may only touch TMEM lanes [32*(warp%4), +32)
    pdl_wait_prior_grid();
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = (t % m_tiles) * C::BM, n0 = (t / m_tiles) * BN;
      const int m = m0 + quarter * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int n = n0 + c * 32;
        if (m < ep.M && n < ep.N) epi_apply<__nv_bfloat16, TT, 32>(ep, m, n, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// fp32 SIMT GEMM (validation mode).  64x64 tile, K step 16, 256 threads x (4x4) outputs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                        const EpiParams ep, const int K) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sW[16][64 + 4];
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    // each thread loads one float4 of A and one of W: row = tid/4 (0..63), k-quad = tid%4
    const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
    if (m0 + r < ep.M) a = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + r) * lda + k0 + kq);
    if (n0 + r < ep.N) w = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + r) * ldw + k0 + kq);
    sA[kq][r] = a.x; sA[kq + 1][r] = a.y; sA[kq + 2][r] = a.z; sA[kq + 3][r] = a.w;
    sW[kq][r] = w.x; sW[kq + 1][r] = w.y; sW[kq + 2][r] = w.z; sW[kq + 3][r] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sA[k][ty * 4 + i]; wv[i] = sW[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i, n = n0 + tx * 4;
    if (m < ep.M && n < ep.N) epi_apply<float, float, 4>(ep, m, n, acc[i]);
  }
}

}  // namespace fmt
