// libfmt_b200.so - host side of the C ABI declared in include/fmt_b200.h.
//
// Owns: the packed weights (bf16 for the tcgen05 path + fp32 for the validation mode), the workspace, the per-plan
// CUDA graph of one sampling window, and the window loop of _perform_ode_sampling_loop (nodes_adv.py:545-694),
// which runs entirely on the device: windows are chained through prev_x / prev_wa / prev_we without host syncs.
#include "../../include/fmt_b200.h"
#include "kernels.cuh"
#include "window.cuh"
#include "flow.cuh"

#include <algorithm>
#include <map>
#include <tuple>
#include <cstdlib>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

using namespace fmt;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int set_err(int code, const char* fmtstr, ...) {
  va_list ap;
  va_start(ap, fmtstr);
  vsnprintf(g_err, sizeof(g_err), fmtstr, ap);
  va_end(ap);
  return code;
}
#define CUDA_OK(expr)                                                                                         \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess) return set_err(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define FMT_OK(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)
#define REQUIRE(cond, ...)                          \
  do {                                              \
    if (!(cond)) return set_err(-1, __VA_ARGS__);   \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ------------------------------------------------------------------------------------------------ handle
struct Linear {
  bf16* w16 = nullptr;    // (N, K) bf16, K possibly padded
  float* w32 = nullptr;   // (N, K) fp32 (validation mode)
  float* b = nullptr;     // (N)
  int N = 0, K = 0;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct FmtHandle {
  int device = 0;
  int num_sms = 148;
  FmtDims d{};
  int N = 0, Kc = 0, NT = 0;
  PFN_encodeTiled encode = nullptr;

  Linear x_emb, c_emb, ada, dec;
  std::vector<Linear> qkv, proj, fc1, fc2;
  float *t0_w = nullptr, *t0_b = nullptr, *t2_w = nullptr, *t2_b = nullptr, *pos = nullptr;
  std::vector<void*> owned;

  // plan
  bool configured = false;
  FmtPlan plan{};
  std::vector<float> t_eval, dt, rk_a, rk_b;
  ModelShape shape{};
  int R = 0, U = 0, n_eval = 0, table_chunk = 0;
  size_t tsize = 2;   // sizeof(activation type)

  // workspace
  DevBuf cond, cemb, temb, tfreq, th, silu, table, xstate, ystage, kbuf, prevx, ax, X, A1, QKV, A2, Hm, V, ddt, dteval, wargs;
  DevBuf urow_buf, uidx_buf;                     // condition-row deduplication: table row of every token row / token row each table row stands for
  bool dedup = false;                            // U < R: the AdaLN tables hold one row per DISTINCT condition row (FMT_DEDUP=0 switches it off)
  int use_dedup = 1;
  DevBuf st_rs, st_wa, st_we, st_noise, st_rd;   // staging for host-located clips
  bool use_splitk = false;                       // 2-way K slicing of gate+residual GEMMs with a ragged last wave (FMT_SPLITK=1; measured 3 % slower at 32 clips)
  int raster_gm = 8;                             // FMT_RASTER_GM
  bool use_pair = true;                          // CTA-pair (cta_group::2) GEMM for M >= 512 (FMT_PAIR=0 disables)
  bool use_pdl = true;                           // programmatic dependent launch between graph nodes (FMT_PDL=0 disables)
  size_t ws_bytes = 0;

  // persistent window kernel (window.cuh): used when the plan has <= 256 token rows in bf16 mode
  int use_window = 3;                            // FMT_WINDOW: 0 = one kernel per op, 1 = barrier-stepped split-K window kernel, 2 = grouped one, 3 = dataflow kernel (flow.cuh)
  bool window_active = false;
  // dataflow window kernel (flow.cuh)
  bool flow_active = false;
  int flow_ch = 64;                              // rows per chunk (FMT_FLOW_CH = 32 | 64)
  int flow_max_rows = 256;                       // plans with more token rows run one kernel per op (FMT_FLOW_MAX_ROWS)
  int flow_fixed = 14;                           // FMT_FLOW_FIXED: split-K partial sums are rounded to multiples of 2^-k before they meet in L2 (exact, order-independent sums below 2^(24-k): bitwise reproducible runs); 0 = off
  int flow_poll = 1;                             // FMT_FLOW_POLL: 0 = acquire polls, 1 = relaxed polls + one acquire load
  long long flow_spin_limit = 20000000000ll;     // SM clocks (~10 s) a wait may spin before the kernel traps (FMT_FLOW_SPIN_MS): long enough to survive a time-sliced GPU, short enough to end a protocol bug
  FlowParams flow_params{};
  DevBuf flow_gemms, flow_tmaps, flow_acc, flow_act, flow_flags, flow_trace;
  bool win_grouped = false;                      // the active plan runs fmt_window_kernel<NV, true>
  int win_spg = 1;                               // sequences per group (FMT_WIN_SPG), rows per group = spg * N <= 128
  int win_fuse_gelu = 1;                         // FMT_WIN_FUSE_GELU=0 keeps the separate GELU stage in the grouped schedule
  int win_nf[4] = {0, 0, 0, 0};                  // feature-slice overrides for qkv / proj / fc1 / fc2 (FMT_WIN_NF; 0 = auto)
  int win_pk[4] = {0, 0, 0, 0};                  // K-split overrides for qkv / proj / fc1 / fc2 (FMT_WIN_PK="q,p,1,2"; 0 = auto)
  DevBuf win_params, win_tmaps, win_acc, win_bar, win_trace, win_act;   // win_act: pre-tiled A1 | A2 | Hm operands
  int win_trace_stride = 0;
  int* win_err_host = nullptr;                   // mapped pinned int: which bounded spin tripped (0 = none)
  int* win_err_dev = nullptr;

  cudaGraphExec_t graph_exec = nullptr;
  int graph_nodes = 0;
  long long launches = 0;
  bool capturing = false;
  long long capture_launches = 0;
};

static int dev_alloc(FmtHandle* h, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return 0;
  if (b.p) {
    CUDA_OK(cudaFree(b.p));
    h->ws_bytes -= b.bytes;
    b.p = nullptr;
    b.bytes = 0;
  }
  bytes = (bytes + 255) & ~static_cast<size_t>(255);
  CUDA_OK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  h->ws_bytes += bytes;
  return 0;
}

static inline void count_launch(FmtHandle* h) {
  if (h->capturing) h->capture_launches++;
  else h->launches++;
}
#define LAUNCH_CHECK() CUDA_OK(cudaGetLastError())

// All kernels of the step go through here: optional thread-block cluster + programmatic dependent launch, so that in the
// captured graph the prologue (barrier init, TMEM alloc, weight prefetch) of node i+1 overlaps the tail of node i.
template <typename... KArgs, typename... Args>
static int launch(FmtHandle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (cluster > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = cluster; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (h->use_pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs; cfg.numAttrs = na;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
  count_launch(h);
  return 0;
}

// ------------------------------------------------------------------------------------------------ GEMM launch
static int make_tmap(FmtHandle* h, CUtensorMap* m, const void* ptr, int rows, int cols, int ld_elems, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-3, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%d ptr=%p", (int)r, rows, cols, ld_elems, box_rows, ptr);
  return 0;
}

static int make_tmap_ex(FmtHandle* h, CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int rows, int cols, int ld_elems,
                        int box_cols, int box_rows, CUtensorMapSwizzle sw) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-3, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%dx%d ptr=%p", (int)r, rows, cols, ld_elems, box_cols, box_rows, ptr);
  return 0;
}

template <int BN>
static int launch_tc(FmtHandle* h, const bf16* A, int lda, const bf16* W, int ldw, const EpiParams& ep, int K, cudaStream_t st) {
  using C = TcCfg<BN>;
  static bool attr_set[64] = {};
  if (!attr_set[h->device & 63]) {   // per device: the attribute lives in the device's module image
    CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[h->device & 63] = true;
  }
  CUtensorMap ta, tb;
  FMT_OK(make_tmap(h, &ta, A, ep.M, K, lda, C::BM));
  FMT_OK(make_tmap(h, &tb, W, ep.N, K, ldw, BN));
  const int tiles = ((ep.M + C::BM - 1) / C::BM) * ((ep.N + BN - 1) / BN);
  const int grid = tiles < h->num_sms ? tiles : h->num_sms;
  return launch(h, gemm_tc_kernel<BN, bf16>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, 1, ta, tb, ep, K);
}

template <int BN>
static int launch_tc2(FmtHandle* h, const bf16* A, int lda, const bf16* W, int ldw, const EpiParams& ep, int K, cudaStream_t st, int ksplit = 1) {
  using C = Tc2Cfg<BN>;
  static bool attr_set[64] = {};
  if (!attr_set[h->device & 63]) {
    CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[h->device & 63] = true;
  }
  CUtensorMap ta, tb;
  FMT_OK(make_tmap(h, &ta, A, ep.M, K, lda, C::BMH));
  FMT_OK(make_tmap(h, &tb, W, ep.N, K, ldw, BN / 2));
  const int tiles = ((ep.M + C::BM - 1) / C::BM) * ((ep.N + BN - 1) / BN) * ((ksplit > 1 && ep.kind == EPI_GATE_RES) ? ksplit : 1);
  const int max_pairs = h->num_sms / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  EpiParams ep2 = ep;
  ep2.raster_gm = h->raster_gm;
  ep2.ksplit = (ksplit > 1 && ep.kind == EPI_GATE_RES) ? ksplit : 1;
  // the kernel carries __cluster_dims__(2,1,1); launch() only adds the PDL attribute
  return launch(h, gemm_tc2_kernel<BN, bf16>, dim3(2 * pairs), dim3(C::THREADS), C::SMEM_BYTES, st, 1, ta, tb, ep2, K);
}

static int pick_bn(const FmtHandle* h, int M, int N) {
  const int mt = (M + 127) / 128;
  if (N % 256 == 0 && mt * (N / 256) >= 2 * h->num_sms) return 256;
  if (N % 128 == 0 && mt * (N / 128) >= h->num_sms) return 128;
  return 64;
}

static int gemm_bf16(FmtHandle* h, const bf16* A, int lda, const bf16* W, int ldw, const EpiParams& ep, int K, cudaStream_t st, int force_bn = 0) {
  REQUIRE(ep.N % 32 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_bf16: N %% 32, K/lda/ldw %% 8 required (N=%d K=%d)", ep.N, K);
  // CTA-pair kernel (force_bn 512 / 1024 = pair tiles 256x256 / 256x128): the default once there are enough 256-row tiles
  if (force_bn == 512) return launch_tc2<256>(h, A, lda, W, ldw, ep, K, st);
  if (force_bn == 1024) return launch_tc2<128>(h, A, lda, W, ldw, ep, K, st);
  if (force_bn == 0 && h->use_pair && ep.M >= 512 && ep.N % 256 == 0 && K % 64 == 0) {
    // Wave model calibrated on B200 (tools/gemm_bench.py): time ~ waves x per-wave cost.  A pair tile (256x256 on two SMs)
    // and a single-CTA 128x256 tile cost the same per wave at 32 clips, but at >= 128 clips the pair kernel is 20 % cheaper
    // (half the L2 -> SM operand traffic); a 128x128 tile costs 0.56 of a wave and wins when the wider tiles leave the last
    // wave mostly empty (N = 1024 at 32 clips: 92 pair tiles on 74 pairs).
    const int sms = h->num_sms, pairs = sms / 2, mt128 = (ep.M + 127) / 128, mt256 = (ep.M + 255) / 256;
    const auto waves = [](int tiles, int units) { return static_cast<double>((tiles + units - 1) / units); };
    const double c_pair = waves(mt256 * (ep.N / 256), pairs);
    const double c_256 = waves(mt128 * (ep.N / 256), sms) * (ep.M >= 16384 ? 1.2 : 1.02);
    const double c_128 = waves(mt128 * (ep.N / 128), sms) * 0.56;
    if (ep.kind == EPI_GATE_RES && h->use_splitk && K >= 1024 && c_pair >= 2.0) {   // single-wave problems stay bitwise deterministic
      // x += gate * (.) is additive, so a tile may be cut into 2 K slices whose partial sums meet in x (fp32 RED):
      // N = 1024 at 32 clips = 92 pair tiles on 74 pairs (2 waves) -> 184 half-K units (2.5 half-waves)
      const double c_split = 0.5 * waves(2 * mt256 * (ep.N / 256), pairs) + 0.08;
      if (c_split < c_pair && c_split < c_256 && c_split < c_128) return launch_tc2<256>(h, A, lda, W, ldw, ep, K, st, 2);
    }
    if (c_pair <= c_256 && c_pair <= c_128) return launch_tc2<256>(h, A, lda, W, ldw, ep, K, st);
    return c_128 < c_256 ? launch_tc<128>(h, A, lda, W, ldw, ep, K, st) : launch_tc<256>(h, A, lda, W, ldw, ep, K, st);
  }
  const int bn = force_bn > 0 ? force_bn : pick_bn(h, ep.M, ep.N);
  switch (bn) {
    case 256: return launch_tc<256>(h, A, lda, W, ldw, ep, K, st);
    case 128: return launch_tc<128>(h, A, lda, W, ldw, ep, K, st);
    case 64: return launch_tc<64>(h, A, lda, W, ldw, ep, K, st);
  }
  return set_err(-1, "gemm_bf16: block_n must be 64, 128 or 256 (got %d)", bn);
}

static int gemm_f32(FmtHandle* h, const float* A, int lda, const float* W, int ldw, const EpiParams& ep, int K, cudaStream_t st) {
  REQUIRE(ep.N % 4 == 0 && K % 16 == 0 && lda % 4 == 0 && ldw % 4 == 0, "gemm_f32: N %% 4, K %% 16 required (N=%d K=%d)", ep.N, K);
  dim3 grid((ep.N + 63) / 64, (ep.M + 63) / 64);
  return launch(h, gemm_simt_kernel, grid, dim3(256), 0, st, 1, A, lda, W, ldw, ep, K);
}

template <typename T> struct ModeOps;
template <> struct ModeOps<bf16> {
  static int gemm(FmtHandle* h, const void* A, int lda, const Linear& L, const EpiParams& ep, cudaStream_t st) {
    return gemm_bf16(h, static_cast<const bf16*>(A), lda, L.w16, L.K, ep, L.K, st);
  }
};
template <> struct ModeOps<float> {
  static int gemm(FmtHandle* h, const void* A, int lda, const Linear& L, const EpiParams& ep, cudaStream_t st) {
    return gemm_f32(h, static_cast<const float*>(A), lda, L.w32, L.K, ep, L.K, st);
  }
};

static EpiParams epi(int kind, int M, int N, const float* bias, void* out, int ldo, int out_f32) {
  EpiParams p{};
  p.kind = kind; p.M = M; p.N = N; p.bias = bias; p.out = out; p.ldo = ldo; p.out_f32 = out_f32;
  return p;
}

// ------------------------------------------------------------------------------------------------ the FMT step
template <typename T>
static int enqueue_prepare(FmtHandle* h, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const WindowArgs* wa = static_cast<const WindowArgs*>(h->wargs.p);
  FMT_OK(launch(h, cond_gather_kernel<T>, dim3(h->U), dim3(128), 0, st, 1, wa, s, h->dedup ? static_cast<const int*>(h->uidx_buf.p) : nullptr,
                static_cast<T*>(h->cond.p)));
  EpiParams ep = epi(EPI_STORE, h->U, s.H, h->c_emb.b, h->cemb.p, s.H, 1);
  FMT_OK(ModeOps<T>::gemm(h, h->cond.p, s.Kc, h->c_emb, ep, st));
  return 0;
}

// AdaLN tables for evaluations [e0, e0+n): table[(e-e0)*U + u, :] = Linear_ada(SiLU(c_emb[u] + t_emb[e]))
template <typename T>
static int enqueue_tables(FmtHandle* h, int e0, int n, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const size_t total = static_cast<size_t>(n) * h->U * s.H;
  FMT_OK(launch(h, silu_cond_kernel<T>, dim3(static_cast<unsigned>((total / 4 + 255) / 256)), dim3(256), 0, st, 1,
                static_cast<const float*>(h->cemb.p), static_cast<const float*>(h->temb.p), e0, h->U, s.H, static_cast<T*>(h->silu.p), total));
  EpiParams ep = epi(EPI_STORE, n * h->U, h->NT, h->ada.b, h->table.p, h->NT, 0);
  FMT_OK(ModeOps<T>::gemm(h, h->silu.p, s.H, h->ada, ep, st));
  return 0;
}

template <typename T>
static int launch_lnmod(FmtHandle* h, const T* table_e, long long shift_off, long long scale_off, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const int warps = 4;
  const dim3 grid((h->R + warps - 1) / warps), block(warps * 32);
  const float* X = static_cast<const float*>(h->X.p);
  T* out = static_cast<T*>(h->A1.p);
  const int* urow = h->dedup ? static_cast<const int*>(h->urow_buf.p) : nullptr;
  const long long ldt = h->NT;
  switch (s.H / 128) {
    case 1: return launch(h, lnmod_kernel<T, T, 1>, grid, block, 0, st, 1, X, h->R, s.H, table_e, urow, ldt, shift_off, scale_off, out);
    case 2: return launch(h, lnmod_kernel<T, T, 2>, grid, block, 0, st, 1, X, h->R, s.H, table_e, urow, ldt, shift_off, scale_off, out);
    case 4: return launch(h, lnmod_kernel<T, T, 4>, grid, block, 0, st, 1, X, h->R, s.H, table_e, urow, ldt, shift_off, scale_off, out);
    case 8: return launch(h, lnmod_kernel<T, T, 8>, grid, block, 0, st, 1, X, h->R, s.H, table_e, urow, ldt, shift_off, scale_off, out);
  }
  return set_err(-1, "dim_h %d unsupported by the LayerNorm kernel (128, 256, 512, 1024)", s.H);
}

template <typename T>
static int launch_attn(FmtHandle* h, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const int heads = h->d.num_heads, hd = s.H / heads;
  const int n_seq = s.nb * s.B, total_warps = n_seq * heads * s.N, warps = 4;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  const dim3 grid((total_warps + warps - 1) / warps), block(warps * 32);
  const T* qkv = static_cast<const T*>(h->QKV.p);
  T* out = static_cast<T*>(h->A2.p);
  const int win = h->d.attention_window;
  if constexpr (sizeof(T) == 2) {
    // many sequences: one CTA per (sequence, head) with K / V staged in shared memory (every qkv byte read once)
    const size_t tile_smem = static_cast<size_t>(2) * s.N * hd * sizeof(bf16);   // K, V of one (sequence, head)
    if (n_seq * heads >= 2 * h->num_sms && tile_smem <= 48 * 1024 && (hd == 64 || hd == 128)) {
      const bf16* q16 = reinterpret_cast<const bf16*>(qkv);
      bf16* o16 = reinterpret_cast<bf16*>(out);
      const dim3 tg(n_seq * heads), tb(ATTN_TILE_THREADS);
      if (hd == 64) return launch(h, band_attention_tile_kernel<64>, tg, tb, tile_smem, st, 1, q16, s.N, heads, win, scale, o16);
      return launch(h, band_attention_tile_kernel<128>, tg, tb, tile_smem, st, 1, q16, s.N, heads, win, scale, o16);
    }
  }
  switch (hd) {
    case 32: return launch(h, band_attention_kernel<T, 1>, grid, block, 0, st, 1, qkv, n_seq, s.N, heads, win, scale, out);
    case 64: return launch(h, band_attention_kernel<T, 2>, grid, block, 0, st, 1, qkv, n_seq, s.N, heads, win, scale, out);
    case 128: return launch(h, band_attention_kernel<T, 4>, grid, block, 0, st, 1, qkv, n_seq, s.N, heads, win, scale, out);
  }
  return set_err(-1, "head_dim %d unsupported (32, 64, 128)", hd);
}

// One model evaluation: V[R, W] = FMT.forward over the nb-way batched CFG branches (FMT.py:277-340).  The x-embedder
// operand `ax` has already been written by the producer of the ODE state (init_window / cfg_combine / rk_combine / pack_x).
template <typename T>
static int enqueue_forward(FmtHandle* h, int table_slot, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const FmtDims& d = h->d;
  const int R = h->R, H = s.H;
  const T* table_e = static_cast<const T*>(h->table.p) + static_cast<size_t>(table_slot) * h->U * h->NT;
  {
    EpiParams ep = epi(EPI_POS, R, H, h->x_emb.b, h->X.p, H, 1);
    ep.pos = h->pos; ep.frames = s.N;
    FMT_OK(ModeOps<T>::gemm(h, h->ax.p, s.W, h->x_emb, ep, st));
  }
  for (int i = 0; i < d.depth; ++i) {
    const long long base = static_cast<long long>(i) * 6 * H;
    FMT_OK(launch_lnmod<T>(h, table_e, base + 0, base + H, st));
    FMT_OK(ModeOps<T>::gemm(h, h->A1.p, H, h->qkv[i], epi(EPI_STORE, R, 3 * H, h->qkv[i].b, h->QKV.p, 3 * H, 0), st));
    FMT_OK(launch_attn<T>(h, st));
    {
      EpiParams ep = epi(EPI_GATE_RES, R, H, h->proj[i].b, h->X.p, H, 1);
      ep.gate = table_e; ep.urow = h->dedup ? static_cast<const int*>(h->urow_buf.p) : nullptr; ep.ldg = h->NT; ep.gate_off = base + 2 * H;
      FMT_OK(ModeOps<T>::gemm(h, h->A2.p, H, h->proj[i], ep, st));
    }
    FMT_OK(launch_lnmod<T>(h, table_e, base + 3 * H, base + 4 * H, st));
    FMT_OK(ModeOps<T>::gemm(h, h->A1.p, H, h->fc1[i], epi(EPI_GELU, R, d.mlp_hidden, h->fc1[i].b, h->Hm.p, d.mlp_hidden, 0), st));
    {
      EpiParams ep = epi(EPI_GATE_RES, R, H, h->fc2[i].b, h->X.p, H, 1);
      ep.gate = table_e; ep.urow = h->dedup ? static_cast<const int*>(h->urow_buf.p) : nullptr; ep.ldg = h->NT; ep.gate_off = base + 5 * H;
      FMT_OK(ModeOps<T>::gemm(h, h->Hm.p, d.mlp_hidden, h->fc2[i], ep, st));
    }
  }
  const long long dbase = static_cast<long long>(d.depth) * 6 * H;
  FMT_OK(launch_lnmod<T>(h, table_e, dbase, dbase + H, st));
  FMT_OK(ModeOps<T>::gemm(h, h->A1.p, H, h->dec, epi(EPI_STORE, R, s.W, h->dec.b, h->V.p, s.W, 1), st));
  return 0;
}

template <typename T>
static int launch_combine(FmtHandle* h, int mode, float* dst, const float* dt_ptr, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const size_t n = static_cast<size_t>(s.B) * s.N * s.W;
  return launch(h, cfg_combine_kernel<T>, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1,
                static_cast<const WindowArgs*>(h->wargs.p), s, static_cast<const float*>(h->V.p), mode, dst, dt_ptr, static_cast<T*>(h->ax.p));
}

// ------------------------------------------------------------------------------------------------ persistent window kernel
static bool window_eligible(const FmtHandle* h, const FmtPlan* p) {
  const FmtDims& d = h->d;
  const int R = p->n_branches * p->batch * (d.num_prev_frames + d.frames_per_clip);
  const int nv = d.dim_h / 128;
  return h->use_window != 0 && p->mode == FMT_MODE_BF16 && R <= 256 && p->n_steps * p->n_stages >= 1 && d.depth <= WIN_MAX_DEPTH &&
         (nv == 1 || nv == 2 || nv == 4 || nv == 8) && d.mlp_hidden % 64 == 0 && d.dim_w % 64 == 0;
}

// Builds the device-side description of the window kernel: tensor maps, GEMM item plan, pointers.  Called by fmt_configure
// after the workspace exists.
static int setup_window(FmtHandle* h, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const FmtDims& d = h->d;
  const int R = h->R, H = s.H, M4 = d.mlp_hidden, W = s.W, D = d.depth;
  int Rp = R <= 64 ? 64 : R <= 128 ? 128 : R <= 192 ? 192 : 256;
  const int n_gemms = 2 + 4 * D;
  const int grid = h->num_sms;
  WinParams wp{};
  // grouped schedule: groups of whole sequences (<= 128 rows each), one slice of the SMs per group
  const int n_seq = s.nb * s.B;
  int spg = h->win_spg >= 1 ? h->win_spg : 1;
  if (n_seq % spg != 0 || spg * s.N > 128) spg = 1;
  const bool grouped = h->use_window == 2 && s.N * spg <= 128 && H % 16 == 0 && M4 % 16 == 0 && W % 16 == 0 && grid / (n_seq / spg) >= 8;
  h->win_grouped = grouped;
  if (grouped) {
    wp.G = n_seq / spg; wp.Cg = grid / wp.G; wp.Rg = spg * s.N; wp.RgP = wp.Rg <= 64 ? 64 : 128;
    Rp = wp.G * wp.RgP;                                   // G pre-tiled buffers of RgP rows each
  }
  wp.s = s; wp.R = R; wp.Rp = Rp; wp.depth = D; wp.heads = d.num_heads; wp.window = d.attention_window; wp.mlp_hidden = M4; wp.NT = h->NT;
  wp.n_steps = h->plan.n_steps; wp.n_stages = h->plan.n_stages; wp.n_gemms = n_gemms;

  // accumulator arena: [Pacc (R,H) | QKVacc (R,3H) | Hacc (R,M4) | Vacc (R,W)] fp32, zeroed by one memset per window
  const size_t n_p = static_cast<size_t>(R) * H, n_q = static_cast<size_t>(R) * 3 * H, n_h = static_cast<size_t>(R) * M4, n_v = static_cast<size_t>(R) * W;
  FMT_OK(dev_alloc(h, h->win_acc, (n_p + n_q + n_h + n_v) * 4));
  FMT_OK(dev_alloc(h, h->win_bar, 2048));                 // [0]: grid barrier, [32 * (1 + g)]: barrier of group g (one 128-byte line each)
  FMT_OK(dev_alloc(h, h->win_params, sizeof(WinParams)));
  const int n_maps = n_gemms + 8;
  FMT_OK(dev_alloc(h, h->win_tmaps, static_cast<size_t>(n_maps) * sizeof(CUtensorMap)));
  if (!h->win_err_host) {
    CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&h->win_err_host), sizeof(int), cudaHostAllocMapped));
    *h->win_err_host = 0;
    CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->win_err_dev), h->win_err_host, 0));
  }
  float* arena = static_cast<float*>(h->win_acc.p);
  wp.Pacc = arena; wp.QKVacc = arena + n_p; wp.Hacc = wp.QKVacc + n_q; wp.Vacc = wp.Hacc + n_h;
  wp.X = static_cast<float*>(h->X.p);
  // pre-tiled operands: [K block][Rp][64] bf16 each, zero-filled once (rows >= R are never written)
  const size_t t_a1 = static_cast<size_t>((H + 63) / 64) * Rp * 64, t_hm = static_cast<size_t>((M4 + 63) / 64) * Rp * 64;
  FMT_OK(dev_alloc(h, h->win_act, (2 * t_a1 + t_hm) * sizeof(bf16)));
  CUDA_OK(cudaMemsetAsync(h->win_act.p, 0, h->win_act.bytes, st));
  wp.A1 = static_cast<bf16*>(h->win_act.p); wp.A2 = wp.A1 + t_a1; wp.Hm = wp.A2 + t_a1; wp.ax = static_cast<bf16*>(h->ax.p);
  wp.table = static_cast<const bf16*>(h->table.p);
  wp.b_x = h->x_emb.b; wp.pos = h->pos; wp.b_dec = h->dec.b;
  for (int i = 0; i < D; ++i) { wp.b_qkv[i] = h->qkv[i].b; wp.b_proj[i] = h->proj[i].b; wp.b_fc1[i] = h->fc1[i].b; wp.b_fc2[i] = h->fc2[i].b; }
  wp.x_state = static_cast<float*>(h->xstate.p); wp.kbuf = static_cast<float*>(h->kbuf.p); wp.ddt = static_cast<const float*>(h->ddt.p);
  for (int i = 0; i < h->plan.n_stages * h->plan.n_stages; ++i) wp.rk_a[i] = h->rk_a[i];
  for (int i = 0; i < h->plan.n_stages; ++i) wp.rk_b[i] = h->rk_b[i];
  wp.wargs = static_cast<const WindowArgs*>(h->wargs.p);
  wp.bar_counter = static_cast<unsigned*>(h->win_bar.p);
  wp.err_flag = h->win_err_dev;
  wp.tmaps = static_cast<const CUtensorMap*>(h->win_tmaps.p);
  wp.w_lookahead = 1000;      // ungated: measured 348 us per ODE step against 378 with a one-stage gate (the gate paid off while spilled state made the SIMT stages' loads slow)
  if (const char* e = getenv("FMT_WIN_LA")) wp.w_lookahead = atoi(e);
  if (getenv("FMT_WIN_TRACE") && atoi(getenv("FMT_WIN_TRACE")) != 0) {
    h->win_trace_stride = h->n_eval * (4 + 8 * D);       // upper bound (the grouped schedule with the fused GELU has 4 + 7 * D stages)
    FMT_OK(dev_alloc(h, h->win_trace, static_cast<size_t>(grid) * h->win_trace_stride * 6 * sizeof(long long)));
    CUDA_OK(cudaMemsetAsync(h->win_trace.p, 0, h->win_trace.bytes, st));
    wp.trace = static_cast<long long*>(h->win_trace.p); wp.trace_stride = h->win_trace_stride;
  }

  std::vector<CUtensorMap> maps(n_maps);
  const int A_AX = n_gemms, A_A1 = n_gemms + 1, A_A2 = n_gemms + 2, A_HM = n_gemms + 3;
  const int C_P = n_gemms + 4, C_Q = n_gemms + 5, C_H = n_gemms + 6, C_V = n_gemms + 7;
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  FMT_OK(make_tmap_ex(h, &maps[A_AX], wp.ax, BF, 2, R, W, W, 64, grouped ? wp.RgP : Rp, CU_TENSOR_MAP_SWIZZLE_128B));
  maps[A_A1] = maps[A_AX]; maps[A_A2] = maps[A_AX]; maps[A_HM] = maps[A_AX];   // A1 / A2 / Hm are pre-tiled: fetched with bulk copies, no tensor map
  FMT_OK(make_tmap_ex(h, &maps[C_P], wp.Pacc, F32, 4, R, H, H, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_Q], wp.QKVacc, F32, 4, R, 3 * H, 3 * H, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_H], wp.Hacc, F32, 4, R, M4, M4, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_V], wp.Vacc, F32, 4, R, W, W, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));

  int next_off = 0;
  auto plan_gemm = [&](int g, const Linear& L, int tm_a, int tm_acc, int pk_override) -> int {
    WinGemm& G = wp.gemms[g];
    G.tm_w = g; G.tm_a = tm_a; G.tm_acc = tm_acc; G.N = L.N;
    G.a_tiled = 0;
    if (tm_a == A_A1 || tm_a == A_A2 || tm_a == A_HM) { G.a_tiled = 1; G.tm_a = tm_a == A_A1 ? 0 : tm_a == A_A2 ? 1 : 2; }
    G.n_ft = (L.N + 127) / 128;
    G.nkb = (L.K + 63) / 64;
    REQUIRE(G.n_ft <= grid, "window kernel: %d feature tiles > %d SMs", G.n_ft, grid);
    int pk = grid / G.n_ft;
    if (pk > G.nkb) pk = G.nkb;
    if (pk_override > 0 && pk_override <= pk) pk = pk_override;
    G.pk = pk;
    G.cta_off = next_off;
    next_off = (next_off + G.n_ft * pk) % grid;
    return make_tmap_ex(h, &maps[g], L.w16, BF, 2, L.N, L.K, L.K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  // grouped schedule: slice the output features over the CTAs of a group; K is split only where one CTA would otherwise
  // have to pull the whole K = mlp_hidden operand (fc2)
  auto plan_gemm2 = [&](int g, const Linear& L, int tm_a, int out, int pk_override, int nf_override, int epi) -> int {
    WinGemm& G = wp.gemms[g];
    G.tm_w = g; G.tm_a = tm_a; G.tm_acc = 0; G.out = out; G.N = L.N;
    G.a_tiled = 0;
    if (tm_a == A_A1 || tm_a == A_A2 || tm_a == A_HM) { G.a_tiled = 1; G.tm_a = tm_a == A_A1 ? 0 : tm_a == A_A2 ? 1 : 2; }
    G.nkb = (L.K + 63) / 64;
    REQUIRE(L.N % 16 == 0, "window kernel: N = %d is not a multiple of 16", L.N);
    int pk = pk_override > 0 ? pk_override : (L.K >= 3072 ? 3 : 1);
    if (pk > G.nkb) pk = G.nkb;
    if (pk > 8) pk = 8;
    int nf = 0;
    for (;;) {                                            // smallest slice that covers N with Cg / pk CTAs, at most 128 wide
      const int slots = wp.Cg / pk;
      nf = slots > 0 ? ((L.N + slots - 1) / slots + 15) / 16 * 16 : 1 << 30;
      for (int c : {16, 32, 48, 64, 96, 128})                // slice widths that pack a 24 KB ring slot (192 / nf K blocks) without waste
        if (nf <= c) { nf = c; break; }
      if (nf <= 128) break;
      REQUIRE(pk > 1, "window kernel: N = %d does not fit %d CTAs per group", L.N, wp.Cg);
      --pk;
    }
    if (nf_override >= 16 && nf_override <= 128 && nf_override % 16 == 0 && ((L.N + nf_override - 1) / nf_override) * pk <= wp.Cg) nf = nf_override;
    G.nf = nf; G.n_nt = (L.N + nf - 1) / nf; G.pk = pk; G.n_ft = G.n_nt;
    G.epi = (epi == 1 && pk == 1) ? 1 : 0;
    G.cta_off = next_off;
    next_off = (next_off + G.n_nt * pk) % wp.Cg;
    return make_tmap_ex(h, &maps[g], L.w16, BF, 2, L.N, L.K, L.K, 64, nf, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  if (grouped) {
    FMT_OK(plan_gemm2(0, h->x_emb, A_AX, 0, 0, 0, 0));
    for (int i = 0; i < D; ++i) {
      FMT_OK(plan_gemm2(1 + 4 * i, h->qkv[i], A_A1, 1, h->win_pk[0], h->win_nf[0], 0));
      FMT_OK(plan_gemm2(2 + 4 * i, h->proj[i], A_A2, 0, h->win_pk[1], h->win_nf[1], 0));
      FMT_OK(plan_gemm2(3 + 4 * i, h->fc1[i], A_A1, 2, h->win_pk[2], h->win_nf[2], h->win_fuse_gelu));
      FMT_OK(plan_gemm2(4 + 4 * i, h->fc2[i], A_HM, 0, h->win_pk[3], h->win_nf[3], 0));
    }
    FMT_OK(plan_gemm2(1 + 4 * D, h->dec, A_A1, 3, 0, 0, 0));
    wp.fuse_gelu = wp.gemms[3].epi;
  } else {
    FMT_OK(plan_gemm(0, h->x_emb, A_AX, C_P, 0));
    for (int i = 0; i < D; ++i) {
      FMT_OK(plan_gemm(1 + 4 * i, h->qkv[i], A_A1, C_Q, h->win_pk[0]));
      FMT_OK(plan_gemm(2 + 4 * i, h->proj[i], A_A2, C_P, h->win_pk[1]));
      FMT_OK(plan_gemm(3 + 4 * i, h->fc1[i], A_A1, C_H, h->win_pk[2]));
      FMT_OK(plan_gemm(4 + 4 * i, h->fc2[i], A_HM, C_P, h->win_pk[3]));
    }
    FMT_OK(plan_gemm(1 + 4 * D, h->dec, A_A1, C_V, 0));
  }
  CUDA_OK(cudaMemcpyAsync(h->win_tmaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(h->win_params.p, &wp, sizeof(WinParams), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));   // `maps` / `wp` are stack objects
  return 0;
}

template <int NV, bool GROUPED>
static int launch_window_nv(FmtHandle* h, cudaStream_t st) {
  constexpr int SMEM = GROUPED ? WIN2_SMEM_BYTES : WIN_SMEM_BYTES;
  static bool attr_set[64] = {};
  if (!attr_set[h->device & 63]) {
    CUDA_OK(cudaFuncSetAttribute(fmt_window_kernel<NV, GROUPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set[h->device & 63] = true;
  }
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fmt_window_kernel<NV, GROUPED>, WIN_THREADS, SMEM));
  REQUIRE(occ >= 1, "window kernel does not fit on an SM (%d B smem)", SMEM);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(h->num_sms); cfg.blockDim = dim3(WIN_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: the kernel synchronises across the grid
  attrs[0].val.cooperative = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, fmt_window_kernel<NV, GROUPED>, static_cast<const WinParams*>(h->win_params.p)));
  count_launch(h);
  return 0;
}

static int launch_window(FmtHandle* h, cudaStream_t st) {
  CUDA_OK(cudaMemsetAsync(h->win_acc.p, 0, h->win_acc.bytes, st));
  CUDA_OK(cudaMemsetAsync(h->win_bar.p, 0, h->win_bar.bytes, st));
  switch (h->shape.H / 128) {
    case 1: return h->win_grouped ? launch_window_nv<1, true>(h, st) : launch_window_nv<1, false>(h, st);
    case 2: return h->win_grouped ? launch_window_nv<2, true>(h, st) : launch_window_nv<2, false>(h, st);
    case 4: return h->win_grouped ? launch_window_nv<4, true>(h, st) : launch_window_nv<4, false>(h, st);
    case 8: return h->win_grouped ? launch_window_nv<8, true>(h, st) : launch_window_nv<8, false>(h, st);
  }
  return set_err(-1, "window kernel: dim_h %d unsupported", h->shape.H);
}

// ------------------------------------------------------------------------------------------------ dataflow window kernel (flow.cuh)
static bool flow_eligible(const FmtHandle* h, const FmtPlan* p) {
  const FmtDims& d = h->d;
  const int N = d.num_prev_frames + d.frames_per_clip;
  const int R = p->n_branches * p->batch * N;
  const int nv = d.dim_h / 128;
  const int CH = h->flow_ch;
  const int nsub = (N + CH - 1) / CH;
  const int hd = d.dim_h / (d.num_heads > 0 ? d.num_heads : 1);
  return h->use_window == 3 && p->mode == FMT_MODE_BF16 && R <= h->flow_max_rows && p->n_steps * p->n_stages >= 1 && d.depth <= WIN_MAX_DEPTH &&
         d.dim_h % 128 == 0 && (nv == 1 || nv == 2 || nv == 4 || nv == 8) && d.mlp_hidden % 64 == 0 && d.dim_w % 64 == 0 &&
         (hd == 32 || hd == 64 || hd == 128) && p->n_branches * p->batch * nsub <= FLOW_MAX_CHUNKS && d.attention_window <= CH &&
         p->n_steps * p->n_stages < 256 && 2 + 4 * d.depth < 256;
}

static int setup_flow(FmtHandle* h, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const FmtDims& d = h->d;
  const int R = h->R, H = s.H, M4 = d.mlp_hidden, W = s.W, D = d.depth;
  const int n_gemms = 2 + 4 * D;
  const int grid = h->num_sms;
  const int CH = h->flow_ch;
  FlowParams& fp = h->flow_params;
  fp = FlowParams{};
  fp.s = s; fp.R = R; fp.depth = D; fp.heads = d.num_heads; fp.window = d.attention_window; fp.mlp_hidden = M4; fp.NT = h->NT;
  fp.n_steps = h->plan.n_steps; fp.n_stages = h->plan.n_stages; fp.n_gemms = n_gemms;
  fp.CH = CH; fp.nsub = (s.N + CH - 1) / CH; fp.NPs = fp.nsub * CH; fp.n_chunks = s.nb * s.B * fp.nsub; fp.RP = fp.n_chunks * CH;
  fp.n_tslots = 512 / CH;
  fp.poll_mode = h->flow_poll;
  fp.fx_c = h->flow_fixed > 0 ? 12582912.f / static_cast<float>(1u << h->flow_fixed) : 0.f;   // 1.5 * 2^23 * 2^-k
  fp.spin_limit = h->flow_spin_limit;
  const size_t RP = fp.RP;

  // fp32 arena, zeroed before every launch: X | Pacc | QKVacc x 2 | Hacc | Vacc, padded rows
  const size_t n_x = RP * H, n_q = RP * 3 * H, n_h = RP * M4, n_v = RP * W;
  FMT_OK(dev_alloc(h, h->flow_acc, (2 * n_x + 2 * n_q + n_h + n_v) * 4));
  float* arena = static_cast<float*>(h->flow_acc.p);
  fp.X = arena; fp.Pacc = fp.X + n_x; fp.QKVacc = fp.Pacc + n_x; fp.Hacc = fp.QKVacc + 2 * n_q; fp.Vacc = fp.Hacc + n_h;
  // pre-tiled operands [chunk][K block][CH][64] bf16, zero-filled once (rows past a chunk's valid rows are never written)
  const size_t t_a1 = static_cast<size_t>(fp.n_chunks) * (H / 64) * CH * 64, t_hm = static_cast<size_t>(fp.n_chunks) * (M4 / 64) * CH * 64;
  FMT_OK(dev_alloc(h, h->flow_act, (2 * t_a1 + t_hm) * sizeof(bf16)));
  CUDA_OK(cudaMemsetAsync(h->flow_act.p, 0, h->flow_act.bytes, st));
  fp.A1 = static_cast<bf16*>(h->flow_act.p); fp.A2 = fp.A1 + t_a1; fp.Hm = fp.A2 + t_a1; fp.ax = static_cast<bf16*>(h->ax.p);
  fp.table = static_cast<const bf16*>(h->table.p);
  fp.U = h->U; fp.urow = h->dedup ? static_cast<const int*>(h->urow_buf.p) : nullptr;
  fp.b_x = h->x_emb.b; fp.pos = h->pos; fp.b_dec = h->dec.b;
  for (int i = 0; i < D; ++i) { fp.b_qkv[i] = h->qkv[i].b; fp.b_proj[i] = h->proj[i].b; fp.b_fc1[i] = h->fc1[i].b; fp.b_fc2[i] = h->fc2[i].b; }
  fp.x_state = static_cast<float*>(h->xstate.p); fp.kbuf = static_cast<float*>(h->kbuf.p); fp.ddt = static_cast<const float*>(h->ddt.p);
  for (int i = 0; i < h->plan.n_stages * h->plan.n_stages; ++i) fp.rk_a[i] = h->rk_a[i];
  for (int i = 0; i < h->plan.n_stages; ++i) fp.rk_b[i] = h->rk_b[i];
  fp.wargs = static_cast<const WindowArgs*>(h->wargs.p);
  // release counters: [n_eval][n_gemms][n_chunks] for the GEMM engine and the same for the SIMT engine
  const size_t n_flags = static_cast<size_t>(h->n_eval) * n_gemms * fp.n_chunks * FLOW_FLAG_STRIDE;
  FMT_OK(dev_alloc(h, h->flow_flags, (2 * n_flags + 2 * FLOW_FLAG_STRIDE) * sizeof(unsigned)));   // + the two rendezvous counters of the trace mode
  fp.g_done = static_cast<unsigned*>(h->flow_flags.p); fp.s_done = fp.g_done + n_flags;
  if (!h->win_err_host) {
    CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&h->win_err_host), sizeof(int), cudaHostAllocMapped));
    *h->win_err_host = 0;
    CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->win_err_dev), h->win_err_host, 0));
  }
  fp.err_flag = h->win_err_dev;
  fp.trace = nullptr; fp.trace_eval = -1;
  if (getenv("FMT_WIN_TRACE") && atoi(getenv("FMT_WIN_TRACE")) != 0) {
    FMT_OK(dev_alloc(h, h->flow_trace, (static_cast<size_t>(grid) * 2 * n_gemms * fp.n_chunks * 8 + static_cast<size_t>(grid) * 4) * sizeof(long long)));
    CUDA_OK(cudaMemsetAsync(h->flow_trace.p, 0, h->flow_trace.bytes, st));
    fp.trace = static_cast<long long*>(h->flow_trace.p);
    fp.trace_eval = h->n_eval / 2;
    if (const char* e = getenv("FMT_WIN_TRACE_EVAL")) fp.trace_eval = atoi(e);
  }

  // tensor maps: one per GEMM (weights) + ax + accumulators (Pacc, QKVacc parity 0 / 1, Hacc, Vacc)
  const int TM_AX = n_gemms, C_P = n_gemms + 1, C_Q = n_gemms + 2, C_H = n_gemms + 4, C_V = n_gemms + 5, n_maps = n_gemms + 6;
  std::vector<CUtensorMap> maps(n_maps);
  std::vector<FlowGemm> gemms(n_gemms);
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  FMT_OK(make_tmap_ex(h, &maps[TM_AX], fp.ax, BF, 2, R, W, W, 64, CH, CU_TENSOR_MAP_SWIZZLE_128B));
  FMT_OK(make_tmap_ex(h, &maps[C_P], fp.Pacc, F32, 4, fp.RP, H, H, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_Q], fp.QKVacc, F32, 4, fp.RP, 3 * H, 3 * H, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_Q + 1], fp.QKVacc + n_q, F32, 4, fp.RP, 3 * H, 3 * H, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_H], fp.Hacc, F32, 4, fp.RP, M4, M4, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  FMT_OK(make_tmap_ex(h, &maps[C_V], fp.Vacc, F32, 4, fp.RP, W, W, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  fp.tm_ax = TM_AX;

  int next_off = 0, max_nk = 1;
  auto plan_gemm = [&](int g, const Linear& L, int a_src, int tm_acc, int pk_override, int par_blk) -> int {
    FlowGemm& G = gemms[g];
    G.tm_w = g; G.tm_acc = tm_acc; G.a_src = a_src; G.par_blk = par_blk;
    G.n_ft = (L.N + 127) / 128;
    G.nkb = (L.K + 63) / 64;
    REQUIRE(G.n_ft <= grid, "flow kernel: %d feature tiles > %d SMs", G.n_ft, grid);
    int pk = grid / G.n_ft;
    if (pk > G.nkb) pk = G.nkb;
    if (pk_override > 0 && pk_override <= pk) pk = pk_override;
    while ((G.nkb + pk - 1) / pk > FLOW_MAX_NW - 1 && pk < G.nkb) ++pk;      // an item's weight tiles must fit the ring
    REQUIRE((G.nkb + pk - 1) / pk <= FLOW_MAX_NW - 1 && G.n_ft * pk <= grid, "flow kernel: K = %d does not fit the weight ring", L.K);
    G.pk = pk;
    G.n_items = G.n_ft * pk;
    G.cta_off = next_off;
    next_off = (next_off + G.n_items) % grid;
    const int nk = (G.nkb + pk - 1) / pk;
    if (nk > max_nk) max_nk = nk;
    return make_tmap_ex(h, &maps[g], L.w16, BF, 2, L.N, L.K, L.K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  FMT_OK(plan_gemm(0, h->x_emb, 0, C_P, 0, -1));
  for (int i = 0; i < D; ++i) {
    FMT_OK(plan_gemm(1 + 4 * i, h->qkv[i], 1, C_Q, h->win_pk[0], i));
    FMT_OK(plan_gemm(2 + 4 * i, h->proj[i], 2, C_P, h->win_pk[1], -1));
    FMT_OK(plan_gemm(3 + 4 * i, h->fc1[i], 1, C_H, h->win_pk[2], -1));
    FMT_OK(plan_gemm(4 + 4 * i, h->fc2[i], 3, C_P, h->win_pk[3], -1));
  }
  FMT_OK(plan_gemm(1 + 4 * D, h->dec, 1, C_V, 0, -1));
  // shared memory: 2 staging tiles + activation ring (slots of the largest item's K range) + weight ring with what is left
  fp.a_slot_bytes = max_nk * CH * 128;
  fp.na = CH <= 32 ? 3 : 2;
  if (const char* e = getenv("FMT_FLOW_NA")) { const int v = atoi(e); if (v >= 2 && v <= FLOW_MAX_NA) fp.na = v; }
  const int left = FLOW_SMEM_MAX - 1024 - FLOW_BAR_BYTES - 2 * FLOW_STG_BYTES - fp.na * fp.a_slot_bytes;
  fp.nw = left / FLOW_W_BYTES;
  if (fp.nw > FLOW_MAX_NW) fp.nw = FLOW_MAX_NW;
  if (const char* e = getenv("FMT_FLOW_NW")) { const int v = atoi(e); if (v >= 2 && v <= fp.nw) fp.nw = v; }
  REQUIRE(fp.nw >= max_nk + 1, "flow kernel: weight ring of %d slots cannot hold an item of %d K blocks", fp.nw, max_nk);

  FMT_OK(dev_alloc(h, h->flow_gemms, gemms.size() * sizeof(FlowGemm)));
  FMT_OK(dev_alloc(h, h->flow_tmaps, maps.size() * sizeof(CUtensorMap)));
  CUDA_OK(cudaMemcpyAsync(h->flow_gemms.p, gemms.data(), gemms.size() * sizeof(FlowGemm), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(h->flow_tmaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));   // `maps` / `gemms` are stack objects
  fp.gemms = static_cast<const FlowGemm*>(h->flow_gemms.p);
  fp.tmaps = static_cast<const CUtensorMap*>(h->flow_tmaps.p);
  return 0;
}

static int flow_smem_bytes(const FlowParams& fp) {
  return 1024 + fp.nw * FLOW_W_BYTES + fp.na * fp.a_slot_bytes + 2 * FLOW_STG_BYTES + FLOW_BAR_BYTES;
}

template <int NV>
static int launch_flow_nv(FmtHandle* h, cudaStream_t st, bool probe_only) {
  const int smem = flow_smem_bytes(h->flow_params);
  CUDA_OK(cudaFuncSetAttribute(fmt_flow_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fmt_flow_kernel<NV>, FLOW_THREADS, smem));
  REQUIRE(occ >= 1, "flow kernel does not fit on an SM (%d B smem)", smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(h->num_sms); cfg.blockDim = dim3(FLOW_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: they wait for one another's counters
  attrs[0].val.cooperative = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (probe_only) {
    // one trial launch with zero evaluations (every role loop is empty): under MPS with an SM limit, a green context or any other
    // reason the runtime cannot make num_sms CTAs co-resident, THIS is where it says so - and the plan runs one kernel per op instead
    FlowParams trial = h->flow_params;
    trial.n_steps = 0; trial.trace = nullptr;
    CUDA_OK(cudaLaunchKernelEx(&cfg, fmt_flow_kernel<NV>, trial));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
  }
  CUDA_OK(cudaLaunchKernelEx(&cfg, fmt_flow_kernel<NV>, h->flow_params));
  count_launch(h);
  return 0;
}

static int launch_flow(FmtHandle* h, cudaStream_t st, bool probe_only = false) {
  if (!probe_only) {
    CUDA_OK(cudaMemsetAsync(h->flow_acc.p, 0, h->flow_acc.bytes, st));
    CUDA_OK(cudaMemsetAsync(h->flow_flags.p, 0, h->flow_flags.bytes, st));
  }
  switch (h->shape.H / 128) {
    case 1: return launch_flow_nv<1>(h, st, probe_only);
    case 2: return launch_flow_nv<2>(h, st, probe_only);
    case 4: return launch_flow_nv<4>(h, st, probe_only);
    case 8: return launch_flow_nv<8>(h, st, probe_only);
  }
  return set_err(-1, "flow kernel: dim_h %d unsupported", h->shape.H);
}

// The whole window: prepare -> tables -> S steps x stages -> finalize.  This is what gets captured in the graph.
template <typename T>
static int enqueue_window(FmtHandle* h, cudaStream_t st) {
  const ModelShape& s = h->shape;
  const WindowArgs* wa = static_cast<const WindowArgs*>(h->wargs.p);
  const size_t nx = static_cast<size_t>(s.B) * s.L * s.W, nfull = static_cast<size_t>(s.B) * s.N * s.W;
  float* x_state = static_cast<float*>(h->xstate.p);
  float* y_stage = static_cast<float*>(h->ystage.p);
  float* kbuf = static_cast<float*>(h->kbuf.p);
  T* ax = static_cast<T*>(h->ax.p);
  const float* ddt = static_cast<const float*>(h->ddt.p);
  const int S = h->plan.n_steps, G = h->plan.n_stages;
  const dim3 gx(static_cast<unsigned>((nx + 255) / 256)), blk(256);

  FMT_OK(launch(h, init_window_kernel<T>, dim3(static_cast<unsigned>((nfull + 255) / 256)), blk, 0, st, 1, wa, s, x_state,
                static_cast<float*>(h->prevx.p), ax));
  if (S > 0) FMT_OK(enqueue_prepare<T>(h, st));
  if (h->window_active && S > 0) {
    // small-R plan: every evaluation of the window runs inside ONE persistent kernel (flow.cuh; window.cuh with FMT_WINDOW=1/2)
    FMT_OK(enqueue_tables<T>(h, 0, h->n_eval, st));
    if (h->flow_active) FMT_OK(launch_flow(h, st));
    else FMT_OK(launch_window(h, st));
    FMT_OK(launch(h, finalize_window_kernel, gx, blk, 0, st, 1, wa, s, static_cast<const float*>(x_state), static_cast<float*>(h->prevx.p)));
    return 0;
  }
  for (int step = 0; step < S; ++step) {
    for (int g = 0; g < G; ++g) {
      const int e = step * G + g;
      if (e % h->table_chunk == 0) {
        const int n = (h->n_eval - e) < h->table_chunk ? (h->n_eval - e) : h->table_chunk;
        FMT_OK(enqueue_tables<T>(h, e, n, st));
      }
      if (g > 0) {   // y_g = y0 + dt * sum_j a[g][j] k_j  (also refreshes the x-embedder operand)
        const float* a = &h->rk_a[g * G];
        FMT_OK(launch(h, rk_combine_kernel<T>, gx, blk, 0, st, 1, s, static_cast<const float*>(x_state), y_stage, static_cast<const float*>(kbuf), nx, nx, g,
                      a[0], G > 1 ? a[1] : 0.f, G > 2 ? a[2] : 0.f, G > 3 ? a[3] : 0.f, ddt + step, ax));
      }
      FMT_OK(enqueue_forward<T>(h, e % h->table_chunk, st));
      if (G == 1) FMT_OK(launch_combine<T>(h, 2, x_state, ddt + step, st));           // fused CFG + Euler
      else FMT_OK(launch_combine<T>(h, 1, kbuf + static_cast<size_t>(g) * nx, nullptr, st));
    }
    if (G > 1) {
      const float* b = h->rk_b.data();
      FMT_OK(launch(h, rk_combine_kernel<T>, gx, blk, 0, st, 1, s, static_cast<const float*>(x_state), x_state, static_cast<const float*>(kbuf), nx, nx, G,
                    b[0], G > 1 ? b[1] : 0.f, G > 2 ? b[2] : 0.f, G > 3 ? b[3] : 0.f, ddt + step, ax));
    }
  }
  FMT_OK(launch(h, finalize_window_kernel, gx, blk, 0, st, 1, wa, s, static_cast<const float*>(x_state), static_cast<float*>(h->prevx.p)));
  return 0;
}

static int enqueue_window_mode(FmtHandle* h, cudaStream_t st) {
  return h->plan.mode == FMT_MODE_BF16 ? enqueue_window<bf16>(h, st) : enqueue_window<float>(h, st);
}

// ------------------------------------------------------------------------------------------------ create / destroy
static int upload(FmtHandle* h, const void* src, size_t n_floats, int location, float** out) {
  float* p = nullptr;
  CUDA_OK(cudaMalloc(&p, n_floats * sizeof(float)));
  h->owned.push_back(p);
  CUDA_OK(cudaMemcpy(p, src, n_floats * sizeof(float), location == FMT_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
  *out = p;
  return 0;
}

static int make_linear(FmtHandle* h, Linear& L, const void* w, const void* b, int N, int K, int Kpad, int location) {
  float *w_raw = nullptr;
  FMT_OK(upload(h, w, static_cast<size_t>(N) * K, location, &w_raw));
  FMT_OK(upload(h, b, N, location, &L.b));
  L.N = N; L.K = Kpad;
  const size_t n = static_cast<size_t>(N) * Kpad;
  CUDA_OK(cudaMalloc(&L.w16, n * sizeof(bf16)));
  h->owned.push_back(L.w16);
  pack_weight_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(w_raw, N, K, L.w16, Kpad);
  LAUNCH_CHECK();
  if (Kpad == K) {
    L.w32 = w_raw;
  } else {
    CUDA_OK(cudaMalloc(&L.w32, n * sizeof(float)));
    h->owned.push_back(L.w32);
    pad_weight_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(w_raw, N, K, L.w32, Kpad);
    LAUNCH_CHECK();
  }
  return 0;
}

static int debug_handle(FmtHandle& h) {
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return set_err(-4, "device is sm_%d%d; sm_100a required", prop.major, prop.minor);
  h.device = dev; h.num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_OK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres));
  REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  h.encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return 0;
}

extern "C" {

int32_t fmt_abi_version(void) { return FMT_ABI_VERSION; }
const char* fmt_last_error(void) { return g_err; }

int32_t fmt_create(const FmtDims* dims, const void* const* wp, int32_t n_ptrs, int32_t location, int32_t device, FmtHandle** out) {
  REQUIRE(dims && wp && out, "fmt_create: null argument");
  const FmtDims& d = *dims;
  REQUIRE(n_ptrs == FMT_W_NUM_GLOBAL + d.depth * FMT_W_PER_BLOCK, "fmt_create: expected %d weight pointers, got %d",
          FMT_W_NUM_GLOBAL + d.depth * FMT_W_PER_BLOCK, n_ptrs);
  for (int i = 0; i < n_ptrs; ++i) REQUIRE(wp[i] != nullptr, "fmt_create: weight pointer %d is null", i);
  REQUIRE(d.dim_h % 128 == 0 && d.dim_w % 64 == 0 && d.dim_a % 4 == 0 && d.mlp_hidden % 64 == 0,
          "fmt_create: unsupported dims (dim_h %% 128, dim_w %% 64, mlp_hidden %% 64 required)");
  REQUIRE(d.dim_h % d.num_heads == 0, "fmt_create: dim_h must be divisible by num_heads");
  const int hd = d.dim_h / d.num_heads;
  REQUIRE(hd == 32 || hd == 64 || hd == 128, "fmt_create: head_dim %d unsupported (32, 64, 128)", hd);
  REQUIRE(d.num_prev_frames <= d.frames_per_clip && d.num_prev_frames >= 0 && d.frames_per_clip > 0, "fmt_create: need 0 <= P <= L");
  REQUIRE(d.attention_window >= 0, "fmt_create: attention_window must be >= 0");

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return set_err(-4, "fmt_create: no CUDA device - this library has no CPU fallback");
  REQUIRE(device >= 0 && device < n_dev, "fmt_create: device %d out of range (%d devices)", device, n_dev);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return set_err(-4, "fmt_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
  CUDA_OK(cudaSetDevice(device));

  FmtHandle* h = new FmtHandle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->d = d;
  h->N = d.num_prev_frames + d.frames_per_clip;
  h->Kc = ((d.dim_w + d.dim_a + d.dim_e + 63) / 64) * 64;
  h->NT = d.depth * 6 * d.dim_h + 2 * d.dim_h;
  if (const char* e = getenv("FMT_PDL")) h->use_pdl = atoi(e) != 0;
  if (const char* e = getenv("FMT_WINDOW")) h->use_window = atoi(e);
  if (const char* e = getenv("FMT_FLOW_CH")) { const int v = atoi(e); if (v == 32 || v == 64) h->flow_ch = v; }
  if (const char* e = getenv("FMT_FLOW_MAX_ROWS")) h->flow_max_rows = atoi(e);
  if (const char* e = getenv("FMT_FLOW_POLL")) h->flow_poll = atoi(e);
  if (const char* e = getenv("FMT_DEDUP")) h->use_dedup = atoi(e) != 0;
  if (const char* e = getenv("FMT_FLOW_FIXED")) { const int v = atoi(e); if (v >= 0 && v <= 22) h->flow_fixed = v; }
  if (const char* e = getenv("FMT_FLOW_SPIN_MS")) h->flow_spin_limit = static_cast<long long>(atof(e) * 1.9e6);
  if (const char* e = getenv("FMT_WIN_SPG")) h->win_spg = atoi(e);
  if (const char* e = getenv("FMT_WIN_FUSE_GELU")) h->win_fuse_gelu = atoi(e) != 0;
  if (const char* e = getenv("FMT_WIN_NF")) sscanf(e, "%d,%d,%d,%d", &h->win_nf[0], &h->win_nf[1], &h->win_nf[2], &h->win_nf[3]);
  if (const char* e = getenv("FMT_PAIR")) h->use_pair = atoi(e) != 0;
  if (const char* e = getenv("FMT_SPLITK")) h->use_splitk = atoi(e) != 0;
  if (const char* e = getenv("FMT_RASTER_GM")) { int v = atoi(e); if (v >= 1) h->raster_gm = v; }
  if (const char* e = getenv("FMT_WIN_PK")) sscanf(e, "%d,%d,%d,%d", &h->win_pk[0], &h->win_pk[1], &h->win_pk[2], &h->win_pk[3]);
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      delete h;
      return set_err(-3, "fmt_create: cuTensorMapEncodeTiled not available from the driver");
    }
    h->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  const int H = d.dim_h;
  int rc = 0;
#define TRY(expr) do { rc = (expr); if (rc != 0) { fmt_destroy(h); return rc; } } while (0)
  TRY(make_linear(h, h->x_emb, wp[FMT_W_X_W], wp[FMT_W_X_B], H, d.dim_w, d.dim_w, location));
  TRY(make_linear(h, h->c_emb, wp[FMT_W_C_W], wp[FMT_W_C_B], H, d.dim_w + d.dim_a + d.dim_e, h->Kc, location));
  TRY(make_linear(h, h->dec, wp[FMT_W_DEC_W], wp[FMT_W_DEC_B], d.dim_w, H, H, location));
  TRY(upload(h, wp[FMT_W_T0_W], static_cast<size_t>(H) * 256, location, &h->t0_w));
  TRY(upload(h, wp[FMT_W_T0_B], H, location, &h->t0_b));
  TRY(upload(h, wp[FMT_W_T2_W], static_cast<size_t>(H) * H, location, &h->t2_w));
  TRY(upload(h, wp[FMT_W_T2_B], H, location, &h->t2_b));
  TRY(upload(h, wp[FMT_W_POS], static_cast<size_t>(h->N) * H, location, &h->pos));
  h->qkv.resize(d.depth); h->proj.resize(d.depth); h->fc1.resize(d.depth); h->fc2.resize(d.depth);
  // concatenated AdaLN projection: rows [block0 (6H) | ... | block{depth-1} (6H) | decoder (2H)]
  {
    Linear& A = h->ada;
    A.N = h->NT; A.K = H;
    const size_t n = static_cast<size_t>(h->NT) * H;
    cudaError_t e1 = cudaMalloc(&A.w32, n * sizeof(float));
    cudaError_t e2 = cudaMalloc(&A.w16, n * sizeof(bf16));
    cudaError_t e3 = cudaMalloc(&A.b, static_cast<size_t>(h->NT) * sizeof(float));
    if (A.w32) h->owned.push_back(A.w32);
    if (A.w16) h->owned.push_back(A.w16);
    if (A.b) h->owned.push_back(A.b);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { fmt_destroy(h); return set_err(-2, "fmt_create: cudaMalloc failed for the AdaLN table weights"); }
  }
  const cudaMemcpyKind kind = location == FMT_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  for (int i = 0; i < d.depth; ++i) {
    const void* const* bp = wp + FMT_W_NUM_GLOBAL + i * FMT_W_PER_BLOCK;
    TRY(make_linear(h, h->qkv[i], bp[FMT_WB_QKV_W], bp[FMT_WB_QKV_B], 3 * H, H, H, location));
    TRY(make_linear(h, h->proj[i], bp[FMT_WB_PROJ_W], bp[FMT_WB_PROJ_B], H, H, H, location));
    TRY(make_linear(h, h->fc1[i], bp[FMT_WB_FC1_W], bp[FMT_WB_FC1_B], d.mlp_hidden, H, H, location));
    TRY(make_linear(h, h->fc2[i], bp[FMT_WB_FC2_W], bp[FMT_WB_FC2_B], H, d.mlp_hidden, d.mlp_hidden, location));
    if (cudaMemcpy(h->ada.w32 + static_cast<size_t>(i) * 6 * H * H, bp[FMT_WB_ADA_W], static_cast<size_t>(6) * H * H * sizeof(float), kind) != cudaSuccess ||
        cudaMemcpy(h->ada.b + static_cast<size_t>(i) * 6 * H, bp[FMT_WB_ADA_B], static_cast<size_t>(6) * H * sizeof(float), kind) != cudaSuccess) {
      fmt_destroy(h);
      return set_err(-2, "fmt_create: copying adaLN weights of block %d failed", i);
    }
  }
  if (cudaMemcpy(h->ada.w32 + static_cast<size_t>(d.depth) * 6 * H * H, wp[FMT_W_DEC_ADA_W], static_cast<size_t>(2) * H * H * sizeof(float), kind) != cudaSuccess ||
      cudaMemcpy(h->ada.b + static_cast<size_t>(d.depth) * 6 * H, wp[FMT_W_DEC_ADA_B], static_cast<size_t>(2) * H * sizeof(float), kind) != cudaSuccess) {
    fmt_destroy(h);
    return set_err(-2, "fmt_create: copying decoder adaLN weights failed");
  }
  {
    const size_t n = static_cast<size_t>(h->NT) * H;
    pack_weight_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(h->ada.w32, h->NT, H, h->ada.w16, H);
  }
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    fmt_destroy(h);
    return set_err(-2, "fmt_create: weight packing failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
#undef TRY
  *out = h;
  return 0;
}

int32_t fmt_destroy(FmtHandle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  for (void* p : h->owned) cudaFree(p);
  DevBuf* bufs[] = {&h->cond, &h->cemb, &h->temb, &h->tfreq, &h->th, &h->silu, &h->table, &h->xstate, &h->ystage, &h->kbuf, &h->prevx, &h->ax,
                    &h->X, &h->A1, &h->QKV, &h->A2, &h->Hm, &h->V, &h->ddt, &h->dteval, &h->wargs, &h->st_rs, &h->st_wa, &h->st_we, &h->st_noise, &h->st_rd,
                    &h->win_params, &h->win_tmaps, &h->win_acc, &h->win_bar, &h->win_trace, &h->win_act,
                    &h->flow_gemms, &h->flow_tmaps, &h->flow_acc, &h->flow_act, &h->flow_flags, &h->flow_trace, &h->urow_buf, &h->uidx_buf};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (h->win_err_host) cudaFreeHost(h->win_err_host);
  delete h;
  return 0;
}

int64_t fmt_workspace_bytes(const FmtHandle* h) { return h ? static_cast<int64_t>(h->ws_bytes) : 0; }
int64_t fmt_launch_count(const FmtHandle* hc, int32_t reset) {
  FmtHandle* h = const_cast<FmtHandle*>(hc);
  if (!h) return 0;
  const long long v = h->launches;
  if (reset) h->launches = 0;
  return v;
}
int32_t fmt_graph_kernel_nodes(const FmtHandle* h) { return h ? h->graph_nodes : 0; }
int64_t fmt_debug_window_trace(const FmtHandle* h, int64_t* out, int64_t max_elems) {
  if (h && h->flow_active && h->flow_trace.p) {   // dataflow kernel: [cta][engine][stage][chunk][4] SM-clock stamps
    const int64_t nf = static_cast<int64_t>(h->flow_trace.bytes / sizeof(long long));
    if (out != nullptr && max_elems >= nf) {
      if (cudaMemcpy(out, h->flow_trace.p, nf * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    }
    return nf;
  }
  if (!h || !h->win_trace.p) return 0;
  const int64_t n = static_cast<int64_t>(h->num_sms) * h->win_trace_stride * 6;
  if (out != nullptr && max_elems >= n) {
    if (cudaMemcpy(out, h->win_trace.p, n * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  }
  return n;
}
int32_t fmt_window_kernel_status(const FmtHandle* h) {
  if (!h || !h->window_active) return -1;
  return h->win_err_host ? *h->win_err_host : 0;
}

// Which token rows (branch, clip, frame) see the same condition row [wr | wa | we]?  urow[row] = index of the row's distinct
// condition row, uidx[u] = the first token row that sees distinct row u.  The key is exactly what cond_gather_kernel reads for a
// row: the identity latent unless nulled (per clip), the audio latent (context frames read prev_wa, never nulled, FMT.py:366,388;
// current frames read wa unless nulled), the emotion (nulled; static = per clip, covering the context frames too, FMT.py:325-326;
// dynamic = per frame).
static void distinct_condition_rows(const ModelShape& s, std::vector<int>& urow, std::vector<int>& uidx) {
  const int R = s.nb * s.B * s.N;
  urow.assign(R, 0);
  uidx.clear();
  std::map<std::tuple<int, int, int, int, int, int, int>, int> seen;
  for (int row = 0; row < R; ++row) {
    const int br = row / (s.B * s.N), b = (row / s.N) % s.B, f = row % s.N;
    const bool za = (s.null_a >> br) & 1, zr = (s.null_r >> br) & 1, ze = (s.null_e >> br) & 1, ctx = f < s.P;
    const int kr = zr ? -1 : b;
    const int ka = ctx ? 0 : (za ? 1 : 2), kab = (ctx || !za) ? b : -1, kaf = (ctx || !za) ? f : -1;
    const int ke = ze ? 0 : (s.we_dynamic ? 2 : 1), keb = ze ? -1 : b, kef = (!ze && s.we_dynamic) ? f : -1;
    const auto key = std::make_tuple(kr, ka, kab, kaf, ke, keb, kef);
    auto it = seen.find(key);
    if (it == seen.end()) { it = seen.emplace(key, static_cast<int>(uidx.size())).first; uidx.push_back(row); }
    urow[row] = it->second;
  }
}

static void branch_nulls(ModelShape& s) {
  if (s.nb == 3) { s.null_a = 0b001; s.null_r = 0; s.null_e = 0b101; }            // [uncond | all | audio-only]   FMT.py:360-362
  else if (s.nb == 4) { s.null_a = 0b0011; s.null_r = 0b0001; s.null_e = 0b1011; } // [truly-uncond | uncond | all | audio-only] :382-384
  else { s.null_a = s.null_r = s.null_e = 0; }
}

// Host-only view of the deduplication for tests (no device work): fills urow[n_branches * batch * n_frames], returns U.
int32_t fmt_debug_condition_rows(int32_t n_branches, int32_t batch, int32_t n_frames, int32_t n_prev, int32_t we_dynamic, int32_t* urow_out) {
  if (n_branches < 1 || n_branches > 4 || batch < 1 || n_frames < 1 || n_prev < 0 || n_prev > n_frames) return -1;
  ModelShape s{};
  s.nb = n_branches; s.B = batch; s.N = n_frames; s.P = n_prev; s.we_dynamic = we_dynamic;
  branch_nulls(s);
  std::vector<int> urow, uidx;
  distinct_condition_rows(s, urow, uidx);
  if (urow_out != nullptr) std::copy(urow.begin(), urow.end(), urow_out);
  return static_cast<int32_t>(uidx.size());
}

static bool same_plan(const FmtHandle* h, const FmtPlan* p) {
  if (!h->configured) return false;
  const FmtPlan& q = h->plan;
  if (q.batch != p->batch || q.n_branches != p->n_branches || q.we_dynamic != p->we_dynamic || q.mode != p->mode ||
      q.n_steps != p->n_steps || q.n_stages != p->n_stages)
    return false;
  const int ne = p->n_steps * p->n_stages;
  if (ne > 0 && memcmp(h->t_eval.data(), p->t_eval, ne * sizeof(float)) != 0) return false;
  if (p->n_steps > 0 && memcmp(h->dt.data(), p->dt, p->n_steps * sizeof(float)) != 0) return false;
  if (memcmp(h->rk_a.data(), p->rk_a, p->n_stages * p->n_stages * sizeof(float)) != 0) return false;
  if (memcmp(h->rk_b.data(), p->rk_b, p->n_stages * sizeof(float)) != 0) return false;
  return true;
}

int32_t fmt_configure(FmtHandle* h, const FmtPlan* p, void* stream) {
  REQUIRE(h && p, "fmt_configure: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaSetDevice(h->device));
  REQUIRE(p->batch >= 1, "fmt_configure: batch must be >= 1");
  REQUIRE(p->n_branches == 1 || p->n_branches == 3 || p->n_branches == 4, "fmt_configure: n_branches must be 1, 3 or 4");
  REQUIRE(p->mode == FMT_MODE_BF16 || p->mode == FMT_MODE_FP32_VALIDATE, "fmt_configure: unknown mode %d", p->mode);
  REQUIRE(p->n_steps >= 0 && p->n_stages >= 1 && p->n_stages <= FMT_MAX_STAGES, "fmt_configure: bad n_steps/n_stages");
  REQUIRE(p->rk_a && p->rk_b && (p->n_steps == 0 || (p->t_eval && p->dt)), "fmt_configure: null schedule arrays");
  if (same_plan(h, p)) return 0;

  const FmtDims& d = h->d;
  const int ne = p->n_steps * p->n_stages;
  h->configured = false;
  if (h->graph_exec) { CUDA_OK(cudaStreamSynchronize(st)); cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; h->graph_nodes = 0; }
  h->plan = *p;
  h->t_eval.assign(p->t_eval, p->t_eval + ne);
  h->dt.assign(p->dt, p->dt + p->n_steps);
  h->rk_a.assign(p->rk_a, p->rk_a + p->n_stages * p->n_stages);
  h->rk_b.assign(p->rk_b, p->rk_b + p->n_stages);
  h->plan.t_eval = h->plan.dt = h->plan.rk_a = h->plan.rk_b = nullptr;
  h->n_eval = ne;

  ModelShape& s = h->shape;
  s.B = p->batch; s.nb = p->n_branches; s.N = h->N; s.P = d.num_prev_frames; s.L = d.frames_per_clip;
  s.W = d.dim_w; s.A = d.dim_a; s.E = d.dim_e; s.H = d.dim_h; s.Kc = h->Kc; s.we_dynamic = p->we_dynamic;
  branch_nulls(s);
  h->R = s.nb * s.B * s.N;
  h->U = h->R;
  // Distinct condition rows.  The CFG branches null whole inputs (FMT.py:360-392): the unconditional branch's current frames all see
  // [wr | 0 | 0] - ONE condition row per clip instead of L - and its context frames see [wr | prev_wa | 0], the same rows as the
  // audio-only branch's context frames (prev_wa is never nulled, prev_we is nulled like we).  c_embedder, SiLU and the AdaLN table
  // GEMM only depend on the condition row and the evaluation, so they run on the U distinct rows (121 of 180 per clip with 3-way
  // CFG) and every consumer of a table row goes through urow[].  Round 1's window kernels index the table by token row: plans that
  // could fall back to them keep U = R.
  h->dedup = false;
  {
    const bool old_window_plan = h->use_window != 0 && window_eligible(h, p) && !(h->use_window == 3 && flow_eligible(h, p));
    if (h->use_dedup && s.nb > 1 && !old_window_plan) {
      std::vector<int> urow, uidx;
      distinct_condition_rows(s, urow, uidx);
      if (static_cast<int>(uidx.size()) < h->R) {
        h->U = static_cast<int>(uidx.size());
        h->dedup = true;
        FMT_OK(dev_alloc(h, h->urow_buf, urow.size() * sizeof(int)));
        FMT_OK(dev_alloc(h, h->uidx_buf, uidx.size() * sizeof(int)));
        CUDA_OK(cudaMemcpyAsync(h->urow_buf.p, urow.data(), urow.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(h->uidx_buf.p, uidx.data(), uidx.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));      // `urow` / `uidx` are stack objects
      }
    }
  }
  const size_t ts = p->mode == FMT_MODE_BF16 ? 2 : 4;
  h->tsize = ts;
  const size_t R = h->R, U = h->U, H = s.H;
  // table chunking: the AdaLN tables of as many evaluations as fit in half of the memory that is free right now (what this handle
  // already holds for its tables counts as free: dev_alloc reuses it), at most 48 GB.  ComfyUI keeps the decoder and wav2vec models
  // on the same GPU, so a fixed cap could fail where the reference, with O(1 step) memory, succeeds; FMT_TABLE_GB overrides.
  const size_t per_eval = U * static_cast<size_t>(h->NT) * ts;
  size_t cap = static_cast<size_t>(48) << 30;
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const size_t avail = (free_b + h->table.bytes + h->silu.bytes) / 2;
      if (avail < cap) cap = avail;
    } else {
      (void)cudaGetLastError();
    }
    if (const char* e = getenv("FMT_TABLE_GB")) { const double gb = atof(e); if (gb > 0) cap = static_cast<size_t>(gb * (1u << 30)); }
  }
  size_t chunk = per_eval ? cap / per_eval : 1;
  if (chunk < 1) chunk = 1;
  if (chunk > static_cast<size_t>(ne > 0 ? ne : 1)) chunk = ne > 0 ? ne : 1;
  h->table_chunk = static_cast<int>(chunk);

  const size_t nx = static_cast<size_t>(s.B) * s.L * s.W;
  FMT_OK(dev_alloc(h, h->cond, U * s.Kc * ts));
  FMT_OK(dev_alloc(h, h->cemb, U * H * 4));
  FMT_OK(dev_alloc(h, h->temb, static_cast<size_t>(ne > 0 ? ne : 1) * H * 4));
  FMT_OK(dev_alloc(h, h->tfreq, static_cast<size_t>(ne > 0 ? ne : 1) * 256 * 4));
  FMT_OK(dev_alloc(h, h->th, static_cast<size_t>(ne > 0 ? ne : 1) * H * 4));
  FMT_OK(dev_alloc(h, h->silu, chunk * U * H * ts));
  FMT_OK(dev_alloc(h, h->table, chunk * per_eval));
  FMT_OK(dev_alloc(h, h->xstate, nx * 4));
  FMT_OK(dev_alloc(h, h->ystage, nx * 4));
  FMT_OK(dev_alloc(h, h->kbuf, nx * 4 * p->n_stages));
  FMT_OK(dev_alloc(h, h->prevx, static_cast<size_t>(s.B) * (s.P > 0 ? s.P : 1) * s.W * 4));
  FMT_OK(dev_alloc(h, h->ax, R * s.W * ts));
  FMT_OK(dev_alloc(h, h->X, R * H * 4));
  FMT_OK(dev_alloc(h, h->A1, R * H * ts));
  FMT_OK(dev_alloc(h, h->QKV, R * 3 * H * ts));
  FMT_OK(dev_alloc(h, h->A2, R * H * ts));
  FMT_OK(dev_alloc(h, h->Hm, R * static_cast<size_t>(d.mlp_hidden) * ts));
  FMT_OK(dev_alloc(h, h->V, R * s.W * 4));
  FMT_OK(dev_alloc(h, h->ddt, static_cast<size_t>(p->n_steps > 0 ? p->n_steps : 1) * 4));
  FMT_OK(dev_alloc(h, h->dteval, static_cast<size_t>(ne > 0 ? ne : 1) * 4));
  FMT_OK(dev_alloc(h, h->wargs, sizeof(WindowArgs)));
  CUDA_OK(cudaMemsetAsync(h->prevx.p, 0, h->prevx.bytes, st));
  if (ne > 0) {
    // timestep embeddings of every evaluation (FMT.py:294), computed once per plan
    CUDA_OK(cudaMemcpyAsync(h->dteval.p, h->t_eval.data(), ne * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h->ddt.p, h->dt.data(), p->n_steps * sizeof(float), cudaMemcpyHostToDevice, st));
    timestep_freq_kernel<<<ne, 128, 0, st>>>(static_cast<const float*>(h->dteval.p), ne, static_cast<float*>(h->tfreq.p));
    LAUNCH_CHECK(); count_launch(h);
    const int w1 = ne * s.H;
    small_linear_kernel<<<(w1 * 32 + 255) / 256, 256, 0, st>>>(static_cast<const float*>(h->tfreq.p), h->t0_w, h->t0_b, static_cast<float*>(h->th.p), ne, s.H, 256, 1);
    LAUNCH_CHECK(); count_launch(h);
    small_linear_kernel<<<(w1 * 32 + 255) / 256, 256, 0, st>>>(static_cast<const float*>(h->th.p), h->t2_w, h->t2_b, static_cast<float*>(h->temb.p), ne, s.H, s.H, 0);
    LAUNCH_CHECK(); count_launch(h);
    CUDA_OK(cudaStreamSynchronize(st));   // t_eval/dt host vectors may be reassigned by the next configure
  }

  // dataflow kernel first (FMT_WINDOW=3, the default); its setup fails soft: the plan then runs one kernel per op
  h->flow_active = false;
  if (flow_eligible(h, p) && h->table_chunk == ne) {
    if (setup_flow(h, st) == 0 && launch_flow(h, st, /*probe_only=*/true) == 0) h->flow_active = true;
    else (void)cudaGetLastError();
  }
  h->window_active = h->flow_active || (window_eligible(h, p) && h->table_chunk == ne && !h->dedup);   // round 1's kernels index the table by token row
  if (h->window_active && !h->flow_active) {
    // the persistent kernel synchronises across the grid: it needs one resident CTA on every SM, else the plan runs one kernel per op
    int occ = 0;
    cudaError_t oe = cudaErrorUnknown;
    switch (d.dim_h / 128) {
      case 1: oe = cudaFuncSetAttribute(fmt_window_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM_BYTES);
              if (oe == cudaSuccess) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fmt_window_kernel<1, false>, WIN_THREADS, WIN_SMEM_BYTES); break;
      case 2: oe = cudaFuncSetAttribute(fmt_window_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM_BYTES);
              if (oe == cudaSuccess) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fmt_window_kernel<2, false>, WIN_THREADS, WIN_SMEM_BYTES); break;
      case 4: oe = cudaFuncSetAttribute(fmt_window_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM_BYTES);
              if (oe == cudaSuccess) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fmt_window_kernel<4, false>, WIN_THREADS, WIN_SMEM_BYTES); break;
      case 8: oe = cudaFuncSetAttribute(fmt_window_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM_BYTES);
              if (oe == cudaSuccess) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fmt_window_kernel<8, false>, WIN_THREADS, WIN_SMEM_BYTES); break;
    }
    if (oe != cudaSuccess || occ < 1) { (void)cudaGetLastError(); h->window_active = false; }
  }
  if (h->window_active && !h->flow_active) FMT_OK(setup_window(h, st));

  // capture one window as a CUDA graph
  {
    cudaGraph_t graph = nullptr;
    cudaStream_t cs;
    CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    // make sure the lazily-set function attributes exist before capture (attribute calls are not capturable work but are legal)
    CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    h->capturing = true; h->capture_launches = 0;
    int rc = enqueue_window_mode(h, cs);
    h->capturing = false;
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); cudaStreamDestroy(cs); return rc; }
    if (ce != cudaSuccess) { cudaStreamDestroy(cs); return set_err(-2, "graph capture failed: %s", cudaGetErrorString(ce)); }
    size_t n_nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &n_nodes);
    h->graph_nodes = static_cast<int>(n_nodes);
    ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    cudaStreamDestroy(cs);
    if (ce != cudaSuccess) return set_err(-2, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
  }
  h->configured = true;
  return 0;
}

static int set_wargs(FmtHandle* h, const WindowArgs& a, cudaStream_t st) {
  set_window_args_kernel<<<1, 1, 0, st>>>(static_cast<WindowArgs*>(h->wargs.p), a);
  LAUNCH_CHECK(); count_launch(h);
  return 0;
}

int32_t fmt_sample_clip(FmtHandle* h, const FmtClip* c, void* stream) {
  REQUIRE(h && c, "fmt_sample_clip: null argument");
  REQUIRE(h->configured, "fmt_sample_clip: call fmt_configure first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaSetDevice(h->device));
  const ModelShape& s = h->shape;
  REQUIRE(c->r_s && c->wa && c->we && c->r_d, "fmt_sample_clip: null tensor pointer");
  REQUIRE(c->audio_num_frames >= 1 && c->T_wa >= 1 && c->T_we >= 1, "fmt_sample_clip: empty clip");
  REQUIRE(s.we_dynamic ? c->T_we > 1 : c->T_we == 1, "fmt_sample_clip: T_we=%d does not match plan.we_dynamic=%d", c->T_we, s.we_dynamic);
  const int T = c->audio_num_frames;
  const int n_win = (T + s.L - 1) / s.L;
  REQUIRE(c->noise != nullptr, "fmt_sample_clip: noise (x0 of every window) is required");
  // every window must see at least one real wa (and dynamic we) frame; the reference's F.pad(replicate) fails otherwise
  REQUIRE((n_win - 1) * s.L < c->T_wa, "fmt_sample_clip: wa has %d frames but window %d starts at frame %d", c->T_wa, n_win - 1, (n_win - 1) * s.L);
  REQUIRE(!s.we_dynamic || (n_win - 1) * s.L < c->T_we, "fmt_sample_clip: dynamic we has %d frames but the last window starts at %d", c->T_we, (n_win - 1) * s.L);

  const float *r_s = c->r_s, *wa = c->wa, *we = c->we, *noise = c->noise;
  float* r_d = c->r_d;
  const size_t n_rs = static_cast<size_t>(s.B) * s.W, n_wa = static_cast<size_t>(s.B) * c->T_wa * s.A, n_we = static_cast<size_t>(s.B) * c->T_we * s.E;
  const size_t n_noise = static_cast<size_t>(n_win) * s.B * s.L * s.W, n_rd = static_cast<size_t>(s.B) * T * s.W;
  if (c->location == FMT_LOC_HOST) {
    FMT_OK(dev_alloc(h, h->st_rs, n_rs * 4)); FMT_OK(dev_alloc(h, h->st_wa, n_wa * 4)); FMT_OK(dev_alloc(h, h->st_we, n_we * 4));
    FMT_OK(dev_alloc(h, h->st_noise, n_noise * 4)); FMT_OK(dev_alloc(h, h->st_rd, n_rd * 4));
    CUDA_OK(cudaMemcpyAsync(h->st_rs.p, r_s, n_rs * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h->st_wa.p, wa, n_wa * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h->st_we.p, we, n_we * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h->st_noise.p, noise, n_noise * 4, cudaMemcpyHostToDevice, st));
    r_s = static_cast<float*>(h->st_rs.p); wa = static_cast<float*>(h->st_wa.p); we = static_cast<float*>(h->st_we.p);
    noise = static_cast<float*>(h->st_noise.p); r_d = static_cast<float*>(h->st_rd.p);
  }
  for (int w = 0; w < n_win; ++w) {
    WindowArgs a{};
    a.r_s = r_s; a.wa = wa; a.we = we; a.r_d = r_d;
    a.x0 = noise + static_cast<size_t>(w) * s.B * s.L * s.W;
    a.T_wa = c->T_wa; a.T_we = c->T_we; a.T_out = T;
    a.win_start = w * s.L; a.first_window = (w == 0); a.use_ext = 0;
    a.a_scale = c->a_cfg_scale; a.r_scale = c->r_cfg_scale; a.e_scale = c->e_cfg_scale;
    FMT_OK(set_wargs(h, a, st));
    CUDA_OK(cudaGraphLaunch(h->graph_exec, st));
    h->launches += h->capture_launches;
    if (c->progress) c->progress(w, n_win, c->progress_user);
  }
  if (c->location == FMT_LOC_HOST) {
    CUDA_OK(cudaMemcpyAsync(c->r_d, r_d, n_rd * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    // the persistent kernels report a tripped bounded spin through a mapped status word (the launch fails as well)
    if (h->window_active && h->win_err_host && *h->win_err_host != 0) return set_err(-6, "window kernel reported status 0x%08x", *h->win_err_host);
  }
  return 0;
}

int32_t fmt_velocity(FmtHandle* h, const FmtEval* ev, void* stream) {
  REQUIRE(h && ev, "fmt_velocity: null argument");
  REQUIRE(h->configured, "fmt_velocity: call fmt_configure first");
  REQUIRE(ev->eval_index >= 0 && ev->eval_index < h->n_eval, "fmt_velocity: eval_index %d out of range [0,%d)", ev->eval_index, h->n_eval);
  REQUIRE(ev->x && ev->prev_x && ev->wa && ev->prev_wa && ev->we && ev->r_s && ev->v_out, "fmt_velocity: null tensor pointer");
  REQUIRE(!h->shape.we_dynamic || ev->prev_we, "fmt_velocity: dynamic we requires prev_we (FMT.py:306-307)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaSetDevice(h->device));
  const ModelShape& s = h->shape;
  WindowArgs a{};
  a.r_s = ev->r_s; a.wa = ev->wa; a.we = ev->we; a.x0 = ev->x; a.r_d = nullptr;
  a.prev_x_ext = ev->prev_x; a.prev_wa_ext = ev->prev_wa; a.prev_we_ext = ev->prev_we;
  a.T_wa = s.L; a.T_we = s.we_dynamic ? s.L : 1; a.T_out = s.L; a.win_start = 0; a.first_window = 0; a.use_ext = 1;
  a.a_scale = ev->a_cfg_scale; a.r_scale = ev->r_cfg_scale; a.e_scale = ev->e_cfg_scale;
  FMT_OK(set_wargs(h, a, st));
  if (h->plan.mode == FMT_MODE_BF16) {
    FMT_OK(enqueue_prepare<bf16>(h, st));
    FMT_OK(enqueue_tables<bf16>(h, ev->eval_index, 1, st));
    FMT_OK(launch(h, pack_x_kernel<bf16>, dim3(s.B * s.N), dim3(128), 0, st, 1, static_cast<const WindowArgs*>(h->wargs.p), s, ev->x,
                  static_cast<const float*>(h->prevx.p), static_cast<bf16*>(h->ax.p)));
    FMT_OK(enqueue_forward<bf16>(h, 0, st));
    FMT_OK(launch_combine<bf16>(h, 0, ev->v_out, nullptr, st));
  } else {
    FMT_OK(enqueue_prepare<float>(h, st));
    FMT_OK(enqueue_tables<float>(h, ev->eval_index, 1, st));
    FMT_OK(launch(h, pack_x_kernel<float>, dim3(s.B * s.N), dim3(128), 0, st, 1, static_cast<const WindowArgs*>(h->wargs.p), s, ev->x,
                  static_cast<const float*>(h->prevx.p), static_cast<float*>(h->ax.p)));
    FMT_OK(enqueue_forward<float>(h, 0, st));
    FMT_OK(launch_combine<float>(h, 0, ev->v_out, nullptr, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ diagnostics

// ------------------------------------------------------------------------------------------------ audio projection
// wa = SiLU(LayerNorm(Linear(wav2vec features)))  -  FLOAT.py:338-342, applied by FloatApplyAudioProjection (nodes_vadv.py:147-198)
struct FmtProj {
  FmtHandle h;            // bare handle: device, SM count, tensor-map encoder, launch counters
  Linear lin;
  float *ln_w = nullptr, *ln_b = nullptr;
  float eps = 1e-5f;
  DevBuf x32, x16, y, o;
};

int32_t fmt_proj_create(int32_t in_dim, int32_t out_dim, const float* w, const float* b, const float* ln_w, const float* ln_b, float ln_eps,
                        int32_t location, int32_t device, FmtProj** out) {
  REQUIRE(w && b && ln_w && ln_b && out, "fmt_proj_create: null argument");
  REQUIRE(in_dim > 0 && in_dim % 16 == 0 && out_dim > 0 && out_dim % 32 == 0, "fmt_proj_create: in_dim %% 16 and out_dim %% 32 required (got %d -> %d)",
          in_dim, out_dim);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return set_err(-4, "fmt_proj_create: no CUDA device - this library has no CPU fallback");
  REQUIRE(device >= 0 && device < n_dev, "fmt_proj_create: device %d out of range (%d devices)", device, n_dev);
  CUDA_OK(cudaSetDevice(device));
  FmtProj* p = new FmtProj();
  int rc = debug_handle(p->h);
  if (rc == 0) rc = make_linear(&p->h, p->lin, w, b, out_dim, in_dim, in_dim, location);
  if (rc == 0) rc = upload(&p->h, ln_w, out_dim, location, &p->ln_w);
  if (rc == 0) rc = upload(&p->h, ln_b, out_dim, location, &p->ln_b);
  if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) rc = set_err(-2, "fmt_proj_create: weight packing failed");
  if (rc != 0) { for (void* q : p->h.owned) cudaFree(q); delete p; return rc; }
  p->eps = ln_eps;
  p->h.use_pdl = false;
  *out = p;
  return 0;
}

int32_t fmt_proj_destroy(FmtProj* p) {
  if (!p) return 0;
  cudaSetDevice(p->h.device);
  cudaDeviceSynchronize();
  for (void* q : p->h.owned) cudaFree(q);
  for (DevBuf* bf : {&p->x32, &p->x16, &p->y, &p->o})
    if (bf->p) cudaFree(bf->p);
  delete p;
  return 0;
}

// x: (rows, in_dim) fp32, out: (rows, out_dim) fp32, both at `location`.  FMT_MODE_BF16: bf16 tcgen05 GEMM with fp32 accumulation,
// LayerNorm / SiLU in fp32; FMT_MODE_FP32_VALIDATE: fp32 SIMT GEMM.
int32_t fmt_proj_apply(FmtProj* p, const float* x, int64_t rows, float* out, int32_t mode, int32_t location, void* stream) {
  REQUIRE(p && x && out, "fmt_proj_apply: null argument");
  REQUIRE(rows > 0 && rows < (1ll << 31), "fmt_proj_apply: rows = %lld out of range", static_cast<long long>(rows));
  REQUIRE(mode == FMT_MODE_BF16 || mode == FMT_MODE_FP32_VALIDATE, "fmt_proj_apply: unknown mode %d", mode);
  FmtHandle* h = &p->h;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int K = p->lin.K, N = p->lin.N, M = static_cast<int>(rows);
  const size_t nx = static_cast<size_t>(rows) * K, ny = static_cast<size_t>(rows) * N;
  const float* xd = x;
  if (location == FMT_LOC_HOST) {
    FMT_OK(dev_alloc(h, p->x32, nx * sizeof(float)));
    CUDA_OK(cudaMemcpyAsync(p->x32.p, x, nx * sizeof(float), cudaMemcpyHostToDevice, st));
    xd = static_cast<const float*>(p->x32.p);
  }
  FMT_OK(dev_alloc(h, p->y, ny * sizeof(float)));
  EpiParams ep = epi(EPI_STORE, M, N, p->lin.b, p->y.p, N, 1);
  if (mode == FMT_MODE_BF16) {
    FMT_OK(dev_alloc(h, p->x16, nx * sizeof(bf16)));
    const size_t n8 = nx / 8;                                   // K % 16 == 0
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((n8 + 255) / 256, static_cast<size_t>(h->num_sms) * 16));
    f32_to_bf16_kernel<<<blocks, 256, 0, st>>>(xd, static_cast<bf16*>(p->x16.p), n8);
    LAUNCH_CHECK(); count_launch(h);
    FMT_OK(gemm_bf16(h, static_cast<const bf16*>(p->x16.p), K, p->lin.w16, K, ep, K, st));
  } else {
    FMT_OK(gemm_f32(h, xd, K, p->lin.w32, K, ep, K, st));
  }
  float* od = out;
  if (location == FMT_LOC_HOST) {
    FMT_OK(dev_alloc(h, p->o, ny * sizeof(float)));
    od = static_cast<float*>(p->o.p);
  }
  const int warps = 8;
  ln_silu_kernel<<<static_cast<unsigned>((rows + warps - 1) / warps), warps * 32, 0, st>>>(static_cast<const float*>(p->y.p), rows, N, p->ln_w, p->ln_b, p->eps, od);
  LAUNCH_CHECK(); count_launch(h);
  if (location == FMT_LOC_HOST) {
    CUDA_OK(cudaMemcpyAsync(out, od, ny * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  return 0;
}

int64_t fmt_proj_launch_count(const FmtProj* p, int32_t reset) {
  if (!p) return 0;
  FmtProj* q = const_cast<FmtProj*>(p);
  const long long n = q->h.launches;
  if (reset) q->h.launches = 0;
  return n;
}

int32_t fmt_debug_gemm_bf16(const void* A, const void* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K, int32_t block_n, void* stream) {
  FmtHandle h;
  FMT_OK(debug_handle(h));
  h.use_pdl = false;
  if (const char* e = getenv("FMT_RASTER_GM")) { int v = atoi(e); if (v >= 1) h.raster_gm = v; }
  EpiParams ep = epi(EPI_STORE, M, N, bias, out, N, 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return gemm_bf16(&h, static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), K, ep, K, st, block_n);
}

// `iters` back-to-back launches of one bf16 GEMM kernel variant through a single handle (kernel timing in isolation)
int32_t fmt_debug_gemm_bench(const void* A, const void* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K, int32_t block_n,
                             int32_t iters, void* stream) {
  static FmtHandle* hp = nullptr;       // device queries are slow: keep them out of the timed launches
  if (!hp) { hp = new FmtHandle(); FMT_OK(debug_handle(*hp)); }
  FmtHandle& h = *hp;
  h.use_pdl = false;
  if (const char* e = getenv("FMT_RASTER_GM")) { int v = atoi(e); if (v >= 1) h.raster_gm = v; }
  EpiParams ep = epi(EPI_STORE, M, N, bias, out, N, 0);     // bf16 output, like the qkv / fc1 GEMMs of the step
  for (int i = 0; i < iters; ++i)
    FMT_OK(gemm_bf16(&h, static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), K, ep, K, static_cast<cudaStream_t>(stream), block_n));
  return 0;
}

int32_t fmt_debug_gemm_fp32(const float* A, const float* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K, void* stream) {
  FmtHandle h;
  EpiParams ep = epi(EPI_STORE, M, N, bias, out, N, 1);
  return gemm_f32(&h, A, K, W, K, ep, K, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
