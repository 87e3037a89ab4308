// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] * W[N,K]^T), bf16 in, fp32 accumulate.
#pragma once
#include "common.cuh"

namespace rtdf {

struct TcEpilogue {
  const float* bias = nullptr;     // [N]
  int act = ACT_NONE;              // applied to acc + bias
  float scale = 1.0f;              // multiplies the activated value
  const float* resid = nullptr;    // fp32 [rows, ldr], added last (may alias out_f32)
  long long ldr = 0;
  float* out_f32 = nullptr;        // optional fp32 output
  long long ld_f32 = 0;
  bf16* out_bf16 = nullptr;        // optional bf16 output
  long long ld_bf16 = 0;
  // Deterministic split-K for skinny GEMMs (64-wide variant, no activation / residual): when set, split s of
  // tc_plan_splits(rows, N, K) writes its partial sum (bias on split 0) to partials[s][rows][N] (fp32) instead of any
  // other output; the consumer adds them up in a fixed order (layernorm_accum_rows).
  float* partials = nullptr;
  // row-LayerNorm fused epilogue (only for the N == 512 full-row variant): y = act(LN(acc + bias))
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float ln_eps = 1e-5f;
  // Walk the output tiles in reverse order (last row block first).  Consecutive kernels of the transformer stack alternate
  // direction (model.cu): a kernel then starts on the rows its predecessor wrote last, which are the ones still in L2.
  bool reverse_tiles = false;
  int a_cache_hint = 0;            // L2 policy of the A-operand loads: 0 default, 1 evict_first, 2 evict_last
  bool profile_as_wide = false;    // per-launch timing: book this launch with the 256-wide tiles (the 128-wide tail of a split GEMM)
  // ---- LayerNorm folded into the GEMMs on either side of it (bf16 transformer layers, model.cu) -------------------
  // y = LN(x) W^T + b  ==  rstd_i * (bf16(x) W'^T - mean_i * c) + d   with  W' = bf16(W diag(gamma)),  c_j = sum_k W'_jk,
  // d_j = b_j + sum_k beta_k W_jk.  The residual GEMM that produces x also emits what the next GEMM needs:
  // Producer (in-place residual update, out_f32 == resid, leading dimension N, N % 64 == 0, one batch):
  //   xb_out    bf16 copy of the updated rows [rows][N]
  //   stats_out [rows][8] float2 partial (sum, sum of squares) of the updated row; slot = n_tile * 2 + column half
  //             (256-wide tiles, N = 1024: all 8 slots are written; the consumer adds the 8 slots in order)
  bf16* xb_out = nullptr;
  float2* stats_out = nullptr;
  // Consumer: v = rstd_i * (acc - mean_i * fold_c[j]) + bias[j] before activation / scale; mean / rstd of row i from the
  // 8 partials of fold_stats (row length fold_len).
  const float2* fold_stats = nullptr;
  const float* fold_c = nullptr;
  int fold_len = 1024;
  float fold_eps = 1e-5f;
  // Row LayerNorm of the UPDATED fp32 output, fused behind the GEMM (256-wide variants, out_f32 written through TMA
  // store / reduce-add, N <= 1024, N % 128 == 0, one batch): every CTA counts its finished tiles per 128-row block in
  // rowln_counters (zero on entry, left zero on exit); the CTA that adds the last N-tile of a block normalises those rows:
  //   rowln_out = LN(out_f32[row, :]) * rowln_gamma + rowln_beta          (bf16 and / or fp32, leading dimension N)
  // Replaces the separate LayerNorm launch between out_proj / fc2 and the next projection.
  const float* rowln_gamma = nullptr;
  const float* rowln_beta = nullptr;
  float rowln_eps = 1e-5f;
  bf16* rowln_out_bf16 = nullptr;
  float* rowln_out_f32 = nullptr;
  int* rowln_counters = nullptr;
};

// A operand as a (k, row, batch) strided view of bf16 memory.  Plain GEMM: batches = 1.
// Implicit-GEMM 1-D conv on channels-last activations: row t of batch b is the contiguous slab
// x[b, stride*t : stride*t + k, 0:C]  =>  k_extent = k*C, row_stride = stride*C, batch_stride = L_in*C.
struct TcOperandA {
  const bf16* ptr = nullptr;
  long long k_extent = 0;
  long long rows_per_batch = 0;
  long long batches = 1;
  long long row_stride = 0;    // elements
  long long batch_stride = 0;  // elements
};

enum TcMode {
  TC_PLAIN = 0,    // A tile (m0, kb*BK)
  TC_POSCONV = 1,  // grouped k=128 conv as shifted-row GEMM: A tile (m0 + kb - 64, g*64), W tile (g*64, kb*64)
};

// variant: 64, 128, 256 -> plain tiles 128 x variant;  512 -> full-row tile with fused LayerNorm (BK 64);
//          513 -> same with BK 32 / 64-byte swizzle (deeper pipeline)
int tc_gemm(cudaStream_t stream, const TcOperandA& A, const bf16* W, int N, int Kw, int mode, int variant,
            const TcEpilogue& epi);

// Number of K splits tc_gemm uses for a (rows, N, K) problem in partials mode: enough to put the idle SMs on the
// weight stream, at most 8, at least two 64-wide k-blocks per split; 1 = do not split.
int tc_plan_splits(long long rows, int N, int Kw);

// A/B switch for timing runs: evaluate ACT_GELU epilogues with ACT_GELU_TANH / ACT_GELU_AS (0 = default)
void tc_set_gelu_variant(int act);
int tc_get_gelu_variant();

// per-launch CUDA-event timing of tc_gemm launches (variant: 64|128|256|512|513, or -1 = all)
void tc_profile_begin();
int tc_profile_end(int variant, double* ms_total, double* flops_total, int* launches);

// Slab-resident grouped positional conv (posconv_tc.cu): x_f32 (B*T,1024) += gelu(conv_k128_g16(x_bf16) + bias);
// w packed [1024][128*64] bf16 (k index = tap*64 + ci).  Any T >= 1 (256-frame tiles).
int posconv_tc(cudaStream_t s, float* x_f32, const bf16* x_bf16, int B, int T, const bf16* w_packed, const float* bias);

}  // namespace rtdf
