/*
 * fmt_b200.h - C ABI of the B200-native FMT motion-latent sampler (libfmt_b200.so).
 *
 * This is the drop-in boundary for ONE path of set-soft/ComfyUI-FLOAT_Optimized: the Flow-Matching-
 * Transformer ODE sampling loop.  Reference interfaces replaced (paths relative to the reference repo):
 *
 *   src/nodes/nodes_adv.py:545-694   _perform_ode_sampling_loop  -> fmt_sample_clip
 *   src/nodes/models/float/FLOAT.py:172-253  FLOAT.sample        -> fmt_sample_clip (static emotion, 3 branches)
 *   src/nodes/models/float/FMT.py:342-401   forward_with_cfv     -> fmt_velocity
 *   src/nodes/models/float/FMT.py:277-340   forward              -> (inside fmt_velocity / fmt_sample_clip)
 *   torchdiffeq.odeint fixed-grid solvers (nodes_adv.py:658)     -> FmtPlan (Butcher tableau) + fmt_sample_clip
 *
 * Conventions
 *   - plain C types only; every tensor is a dense row-major fp32 array.
 *   - pointers are DEVICE pointers unless the struct says `location = FMT_LOC_HOST`.
 *   - every call returns 0 on success, <0 on error (fmt_last_error() gives the message); no C++ exception
 *     crosses this boundary.  A handle is bound to one CUDA device and is not thread-safe.
 *   - all work is enqueued on the given cudaStream_t (passed as void*); calls do not synchronise unless
 *     documented (host-located outputs synchronise the stream before returning).
 *   - there is NO CPU fallback: on a machine without an sm_100 device fmt_create fails.
 */
#ifndef FMT_B200_H
#define FMT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMT_API __attribute__((visibility("default")))
#define FMT_ABI_VERSION 1

typedef struct FmtHandle FmtHandle;

/* Mirror of the BaseOptions fields the FMT reads (src/nodes/options/base_options.py:10-60). */
typedef struct FmtDims {
  int32_t dim_w;            /* motion latent width (512)            */
  int32_t dim_a;            /* audio latent width (512)             */
  int32_t dim_e;            /* emotion classes (7)                  */
  int32_t dim_h;            /* hidden size (1024)                   */
  int32_t depth;            /* FMT blocks (8)                       */
  int32_t num_heads;        /* attention heads (8)                  */
  int32_t mlp_hidden;       /* int(dim_h * mlp_ratio) (4096)        */
  int32_t num_prev_frames;  /* context frames P (10)                */
  int32_t frames_per_clip;  /* window length L = int(wav2vec_sec*fps) (50) */
  int32_t attention_window; /* band half-width (2), FMT.py:15-19    */
} FmtDims;

/* Order of the fp32 weight pointers handed to fmt_create (nn.Linear layout: (out,in) row-major).
 * Global tensors first, then FMT_W_PER_BLOCK pointers per block. */
enum {
  FMT_W_X_W = 0,      /* x_embedder.proj.weight (H, dim_w)            */
  FMT_W_X_B,          /* x_embedder.proj.bias   (H)                   */
  FMT_W_T0_W,         /* t_embedder.mlp.0.weight (H, 256)             */
  FMT_W_T0_B,
  FMT_W_T2_W,         /* t_embedder.mlp.2.weight (H, H)               */
  FMT_W_T2_B,
  FMT_W_C_W,          /* c_embedder.weight (H, dim_w+dim_a+dim_e), columns [wr | wa | we] */
  FMT_W_C_B,
  FMT_W_POS,          /* pos_embed (P+L, H)                           */
  FMT_W_DEC_ADA_W,    /* decoder.adaLN_modulation.1.weight (2H, H)    */
  FMT_W_DEC_ADA_B,
  FMT_W_DEC_W,        /* decoder.linear.weight (dim_w, H)             */
  FMT_W_DEC_B,
  FMT_W_NUM_GLOBAL
};
enum {
  FMT_WB_QKV_W = 0,   /* blocks.i.attn.qkv.weight (3H, H), rows [q | k | v], head-major */
  FMT_WB_QKV_B,
  FMT_WB_PROJ_W,      /* blocks.i.attn.proj.weight (H, H)             */
  FMT_WB_PROJ_B,
  FMT_WB_FC1_W,       /* blocks.i.mlp.fc1.weight (mlp_hidden, H)      */
  FMT_WB_FC1_B,
  FMT_WB_FC2_W,       /* blocks.i.mlp.fc2.weight (H, mlp_hidden)      */
  FMT_WB_FC2_B,
  FMT_WB_ADA_W,       /* blocks.i.adaLN_modulation.1.weight (6H, H): shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp */
  FMT_WB_ADA_B,
  FMT_W_PER_BLOCK
};

enum { FMT_LOC_DEVICE = 0, FMT_LOC_HOST = 1 };
enum { FMT_MODE_BF16 = 0,       /* bf16 tcgen05/TMA GEMMs, fp32 accumulate / residual / LN / softmax / ODE state */
       FMT_MODE_FP32_VALIDATE = 1 /* same kernels' epilogues + fp32 SIMT GEMMs: the 1e-4 parity mode          */ };

#define FMT_MAX_STAGES 4

/* What forward_with_cfv batches and how the ODE grid is walked.  One plan = one captured CUDA graph. */
typedef struct FmtPlan {
  int32_t batch;          /* clips B                                                                        */
  int32_t n_branches;     /* 1: all scales == 1 (FMT.py:400-401); 3: default CFG; 4: include_r_cfg (:380-399) */
  int32_t we_dynamic;     /* 0: we is (B,1,E) broadcast over all P+L frames; 1: we is (B,T,E)                */
  int32_t mode;           /* FMT_MODE_*                                                                      */
  int32_t n_steps;        /* ODE steps S = nfe-1 (torchdiffeq fixed grid: nfe points)                        */
  int32_t n_stages;       /* explicit RK stages per step (euler 1, midpoint 2, heun2 2, heun3 3, rk4 4)      */
  const float* t_eval;    /* HOST (n_steps*n_stages): time of every function evaluation, fp32               */
  const float* dt;        /* HOST (n_steps): t[i+1]-t[i] in fp32                                            */
  const float* rk_a;      /* HOST (n_stages*n_stages) strictly lower-triangular: y_i = y0 + dt*sum_j a[i][j]*k_j */
  const float* rk_b;      /* HOST (n_stages): y1 = y0 + dt*sum_j b[j]*k_j                                    */
} FmtPlan;

typedef void (*fmt_progress_fn)(int32_t window_index, int32_t n_windows, void* user);

/* One sampler call = _perform_ode_sampling_loop (nodes_adv.py:545-694). */
typedef struct FmtClip {
  int32_t location;       /* FMT_LOC_DEVICE or FMT_LOC_HOST for ALL pointers below                           */
  const float* r_s;       /* (B, dim_w)  reference-identity latent wr                                        */
  const float* wa;        /* (B, T_wa, dim_a) audio latents                                                  */
  const float* we;        /* (B, T_we, dim_e) with T_we == 1 (static) or the dynamic length                  */
  const float* noise;     /* (n_windows, B, L, dim_w): x0 of every window, drawn by the caller in window order
                             (torch.randn(B,L,dim_w, generator) per window, nodes_adv.py:606)                */
  float* r_d;             /* out (B, audio_num_frames, dim_w)                                                */
  int32_t T_wa;
  int32_t T_we;
  int32_t audio_num_frames;
  float a_cfg_scale, r_cfg_scale, e_cfg_scale;
  fmt_progress_fn progress; /* optional: called on the host after each window is enqueued (ProgressBar tick, nodes_adv.py:688) */
  void* progress_user;
} FmtClip;

/* One evaluation of forward_with_cfv (FMT.py:342-401) on explicit window inputs - tests and ncu. */
typedef struct FmtEval {
  const float* x;         /* (B, L, dim_w)                     */
  const float* prev_x;    /* (B, P, dim_w)                     */
  const float* wa;        /* (B, L, dim_a)                     */
  const float* prev_wa;   /* (B, P, dim_a)                     */
  const float* we;        /* (B, 1|L, dim_e) per plan.we_dynamic */
  const float* prev_we;   /* (B, P, dim_e) or NULL when static  */
  const float* r_s;       /* (B, dim_w)                        */
  float* v_out;           /* out (B, P+L, dim_w): the CFG-combined model output, context rows included */
  int32_t eval_index;     /* which t_eval[] entry of the plan  */
  float a_cfg_scale, r_cfg_scale, e_cfg_scale;
} FmtEval;

FMT_API int32_t fmt_abi_version(void);
FMT_API const char* fmt_last_error(void);

/* Packs the reference's fp32 weights into resident bf16 (and keeps fp32 for the validation mode).
 * weight_ptrs: FMT_W_NUM_GLOBAL + depth*FMT_W_PER_BLOCK pointers, all at `location`. */
FMT_API int32_t fmt_create(const FmtDims* dims, const void* const* weight_ptrs, int32_t n_ptrs, int32_t location,
                           int32_t device, FmtHandle** out);
FMT_API int32_t fmt_destroy(FmtHandle* h);

/* Sizes the workspace, computes the timestep embeddings of every evaluation and (re)captures the window
 * graph.  Cheap when the plan equals the current one. */
FMT_API int32_t fmt_configure(FmtHandle* h, const FmtPlan* plan, void* stream);
FMT_API int64_t fmt_workspace_bytes(const FmtHandle* h);

/* Whole clip: all windows, chained on the device through prev_x / prev_wa / prev_we. */
FMT_API int32_t fmt_sample_clip(FmtHandle* h, const FmtClip* clip, void* stream);

/* Single CFG-combined model evaluation (device pointers). */
FMT_API int32_t fmt_velocity(FmtHandle* h, const FmtEval* ev, void* stream);

/* Counters for bench.py: kernels this library launched (graph nodes included) since the last reset. */
FMT_API int64_t fmt_launch_count(const FmtHandle* h, int32_t reset);
/* Number of kernel nodes in the captured window graph (0 before fmt_configure). */
FMT_API int32_t fmt_graph_kernel_nodes(const FmtHandle* h);

/* Persistent window kernel (plans with <= 256 token rows, bf16): -1 = not in use for the current plan, 0 = in use and
 * healthy, > 0 = id of the bounded spin that tripped (the kernel traps instead of hanging the GPU). */
FMT_API int32_t fmt_window_kernel_status(const FmtHandle* h);

/* With FMT_WIN_TRACE=1 in the environment at fmt_configure: copies the per-CTA barrier stamps of the last window,
 * [cta][barrier][arrive, pass] in SM clocks, into `out` (HOST); returns the element count (call with NULL to size). */
FMT_API int64_t fmt_debug_window_trace(const FmtHandle* h, int64_t* out, int64_t max_elems);
/* Host-only (no device work): which token rows (branch, clip, frame) of forward_with_cfv's batched forward (FMT.py:360-392) see the
 * same condition row [wr | wa | we] - the AdaLN tables hold one row per DISTINCT condition row.  Fills urow[n_branches * batch *
 * n_frames] (may be NULL) with the distinct-row index of every token row and returns the number of distinct rows, < 0 on bad
 * arguments. */
FMT_API int32_t fmt_debug_condition_rows(int32_t n_branches, int32_t batch, int32_t n_frames, int32_t n_prev, int32_t we_dynamic,
                                         int32_t* urow_out);

/* ---- audio projection in front of the sampler ----
 * Replaces the nn.Sequential(Linear(in_dim, dim_w), LayerNorm(dim_w), SiLU) that FloatApplyAudioProjection runs
 * (src/nodes/nodes_vadv.py:147-198; module built in src/nodes/models/float/FLOAT.py:338-342 and
 * src/nodes/nodes_vadv_loader.py:228-257): wa = SiLU(LayerNorm(x W^T + b)).  in_dim = 9216 (12 stacked wav2vec layers) or 768. */
typedef struct FmtProj FmtProj;
FMT_API int32_t fmt_proj_create(int32_t in_dim, int32_t out_dim, const float* linear_w /* (out_dim, in_dim) */, const float* linear_b,
                                const float* ln_w, const float* ln_b, float ln_eps, int32_t location, int32_t device, FmtProj** out);
FMT_API int32_t fmt_proj_destroy(FmtProj* p);
/* x (rows, in_dim) fp32 -> out (rows, out_dim) fp32, both at `location` (FMT_LOC_HOST synchronises the stream).  mode as in FmtPlan. */
FMT_API int32_t fmt_proj_apply(FmtProj* p, const float* x, int64_t rows, float* out, int32_t mode, int32_t location, void* stream);
FMT_API int64_t fmt_proj_launch_count(const FmtProj* p, int32_t reset);

/* ---- diagnostic entry points (unit tests of the kernels; not used by the node) ---- */
/* out[M,N] (fp32) = A[M,K] (bf16 bits) @ W[N,K]^T (bf16 bits) + bias[N], through the tcgen05/TMA GEMM. */
FMT_API int32_t fmt_debug_gemm_bf16(const void* A, const void* W, const float* bias, float* out, int32_t M, int32_t N,
                                    int32_t K, int32_t block_n, void* stream);
/* `iters` back-to-back launches of one GEMM variant (bf16 output); block_n as above, 512 / 1024 = CTA-pair kernel. */
FMT_API int32_t fmt_debug_gemm_bench(const void* A, const void* W, const float* bias, float* out, int32_t M, int32_t N,
                                     int32_t K, int32_t block_n, int32_t iters, void* stream);
/* Same through the fp32 SIMT GEMM used by FMT_MODE_FP32_VALIDATE (A, W fp32). */
FMT_API int32_t fmt_debug_gemm_fp32(const float* A, const float* W, const float* bias, float* out, int32_t M, int32_t N,
                                    int32_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FMT_B200_H */
