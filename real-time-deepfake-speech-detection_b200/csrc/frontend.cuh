// Bandwidth-bound front-end kernels + weight packing helpers.
#pragma once
#include "common.cuh"

namespace rtdf {

// y[b,t] = x[b,t] - coef * x[b,t-1], x[b,-1] := x[b,1]   (reference data/preprocess.py:22-27)
int preemph(cudaStream_t s, const float* x, float* y, int B, int N, float coef);

// Optional per-utterance waveform layer norm (north_star scope; NOT on the reference path, default off)
int wave_layernorm(cudaStream_t s, const float* x, float* y, int B, int N, float eps);

// conv-0 of the XLS-R feature encoder fused with bias + LayerNorm(512) + GELU:
//   wav (B,N) fp32 -> out (B, L1, 512) channels-last, L1 = (N-10)/5+1.   w_t: [10][512] (tap-major)
int conv0_ln_gelu(cudaStream_t s, const float* wav, int B, int N, const float* w_t, const float* bias,
                  const float* gamma, const float* beta, float eps, float* out_f32, bf16* out_bf16);

// conv-0 as a tcgen05 implicit GEMM (bf16 output, large batches): w_packed [512][32] bf16 from conv0_tc_pack_weight
// (w: state-dict layout [512][1][10] fp32); a_scratch: B * L1 * 32 bf16 (the hi/lo-split im2col matrix, 64 B per frame).
int conv0_tc_pack_weight(cudaStream_t s, const float* w, bf16* out);
int conv0_tc_ln_gelu(cudaStream_t s, const float* wav, int B, int N, const bf16* w_packed, const float* bias,
                     const float* gamma, const float* beta, float eps, bf16* a_scratch, bf16* out_bf16);

// conv-0 in extractor_mode="default" (group norm): conv (bias may be NULL) -> GroupNorm(512, 512) over time -> GELU.
// ws: conv0_gn_workspace_floats(B, N) floats of scratch (per-chunk partial statistics + per-utterance scale / shift).
size_t conv0_gn_workspace_floats(int B, int N);
int conv0_gn_gelu(cudaStream_t s, const float* wav, int B, int N, const float* w_t, const float* bias, const float* gamma,
                  const float* beta, float eps, float* ws, float* out_f32, bf16* out_bf16);

// Row LayerNorm: in (rows, C) fp32 or bf16 -> optional fp32 and/or bf16 outputs, optional activation.
// reverse (LayerNorm(1024) fast path only): CTAs walk the rows from the last to the first (see TcEpilogue::reverse_tiles)
int layernorm_rows_f32(cudaStream_t s, const float* in, long long rows, int C, const float* gamma, const float* beta,
                       float eps, int act, float* out_f32, bf16* out_bf16, bool reverse = false);
int layernorm_rows_bf16(cudaStream_t s, const bf16* in, long long rows, int C, const float* gamma, const float* beta,
                        float eps, int act, float* out_f32, bf16* out_bf16);

// LayerNorm(1024) behind a split-K GEMM: x (rows,1024) fp32 += sum_s partials[s] (s = 0..n_splits-1, each (rows,1024),
// added in that order), x is written back, then out = LN(x) as bf16 or fp32.
int layernorm_accum_rows(cudaStream_t s, float* x, const float* partials, int n_splits, long long rows, const float* gamma,
                         const float* beta, float eps, float* out_f32, bf16* out_bf16);

// Folded LayerNorm (gemm_tc.cuh, TcEpilogue::fold_* / xb_out / stats_out): row (1024 wide) -> bf16 copy + per-row
// (sum, sum of squares) in slot 0 of stats[row][8] (slots 1..7 zeroed); n_splits > 0 first folds the K-split partial sums
// into x like layernorm_accum_rows.
int cast_stats_rows(cudaStream_t s, float* x, const float* partials, int n_splits, long long rows, bf16* xb, float2* stats);
// W' = bf16(W diag(gamma)) [n][k], c[j] = sum_k W'[j][k], d[j] = bias[j] + sum_k beta[k] W[j][k]
int fold_ln_weight(cudaStream_t s, const float* w, const float* gamma, const float* beta, const float* bias, int n, int k,
                   bf16* wb, float* c, float* d);

// fp32 verification path of the grouped positional conv: x (B,T,1024) fp32, w packed [1024][128*64]
// (k index = tap*64 + ci), out x += gelu(conv + bias)
int posconv_f32(cudaStream_t s, float* x, const float* xin, int B, int T, const float* w_packed, const float* bias);

// ---- packing -----------------------------------------------------------------------------------
int cast_f32_to_bf16(cudaStream_t s, const float* in, bf16* out, long long n);
int scale_rows_f32(cudaStream_t s, float* w, long long rows, long long cols, float scale);  // in-place
// conv weight [co][ci][k] -> [co][k][ci]   (implicit-GEMM K order = tap-major, channel-minor)
int permute_conv_weight(cudaStream_t s, const float* in, float* out, int co, int ci, int k);
// weight-norm fold for the pos-conv: w[co][ci][k] = g[k] * v[co][ci][k] / ||v[:,:,k]||, written as [co][k][ci]
int posconv_fold_weight(cudaStream_t s, const float* v, const float* g, float* out, int co, int ci, int k);
// transpose a small matrix [r][c] -> [c][r]
int transpose_f32(cudaStream_t s, const float* in, float* out, int r, int c);
// eval-BatchNorm fold: scale = w / sqrt(var + eps), shift = b - mean * scale
int bn_fold(cudaStream_t s, const float* w, const float* b, const float* mean, const float* var, float eps,
            float* scale, float* shift, int n);

}  // namespace rtdf
