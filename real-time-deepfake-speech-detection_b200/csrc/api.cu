// C-ABI wrappers around single kernels (used by the per-kernel parity tests and by bench.py's
// roofline measurements).  Lifecycle / forward entry points live in model.cu.
#include "../../include/rtdf.h"
#include "aasist.cuh"
#include "attention.cuh"
#include "conformer.cuh"
#include "conv_tc.cuh"
#include "frontend.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

using namespace rtdf;

extern "C" {

int rtdf_preemph(const float* x, float* y, int batch, int n, float coef, void* stream) {
  return preemph(static_cast<cudaStream_t>(stream), x, y, batch, n, coef);
}

int rtdf_wave_layernorm(const float* x, float* y, int batch, int n, float eps, void* stream) {
  return wave_layernorm(static_cast<cudaStream_t>(stream), x, y, batch, n, eps);
}

int rtdf_conv0_ln_gelu(const float* wav, int batch, int n, const float* w, const float* bias, const float* gamma,
                       const float* beta, float eps, float* out_f32, void* out_bf16, void* stream) {
  return conv0_ln_gelu(static_cast<cudaStream_t>(stream), wav, batch, n, w, bias, gamma, beta, eps, out_f32,
                       static_cast<bf16*>(out_bf16));
}

long long rtdf_conv0_tc_scratch_bytes(int batch, int n) {
  const long long l1 = n >= 10 ? (n - 10) / 5 + 1 : 0;
  return ((long long)batch * l1 * 32 + 512 * 32) * 2 + 256;
}

int rtdf_conv0_tc_ln_gelu(const float* wav, int batch, int n, const float* w, const float* bias, const float* gamma,
                          const float* beta, float eps, void* scratch, void* out_bf16, void* stream) {
  RTDF_REQUIRE(w && scratch && (reinterpret_cast<uintptr_t>(scratch) & 127) == 0, "rtdf_conv0_tc_ln_gelu: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* wp = static_cast<bf16*>(scratch);                 // [512][32] packed weight, then the im2col matrix
  bf16* a = wp + 512 * 32;
  RTDF_TRY(conv0_tc_pack_weight(s, w, wp));
  return conv0_tc_ln_gelu(s, wav, batch, n, wp, bias, gamma, beta, eps, a, static_cast<bf16*>(out_bf16));
}

long long rtdf_conv0_gn_workspace_floats(int batch, int n) { return (long long)conv0_gn_workspace_floats(batch, n); }

int rtdf_conv0_gn_gelu(const float* wav, int batch, int n, const float* w, const float* bias, const float* gamma,
                       const float* beta, float eps, float* workspace, float* out_f32, void* out_bf16, void* stream) {
  return conv0_gn_gelu(static_cast<cudaStream_t>(stream), wav, batch, n, w, bias, gamma, beta, eps, workspace, out_f32,
                       static_cast<bf16*>(out_bf16));
}

int rtdf_layernorm_rows(const void* in, int in_is_bf16, long long rows, int cols, const float* gamma,
                        const float* beta, float eps, int act, float* out_f32, void* out_bf16, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_is_bf16)
    return layernorm_rows_bf16(s, static_cast<const bf16*>(in), rows, cols, gamma, beta, eps, act, out_f32,
                               static_cast<bf16*>(out_bf16));
  return layernorm_rows_f32(s, static_cast<const float*>(in), rows, cols, gamma, beta, eps, act, out_f32,
                            static_cast<bf16*>(out_bf16));
}

static TcEpilogue make_epi(const float* bias, int act, float scale, const float* resid, float* out_f32, void* out_bf16,
                           int N) {
  TcEpilogue e;
  e.bias = bias;
  e.act = act;
  e.scale = scale;
  e.resid = resid;
  e.ldr = N;
  e.out_f32 = out_f32;
  e.ld_f32 = N;
  e.out_bf16 = static_cast<bf16*>(out_bf16);
  e.ld_bf16 = N;
  return e;
}

int rtdf_gemm_bf16(const void* A, const void* W, int M, int N, int K, const float* bias, int act, float scale,
                   const float* resid, float* out_f32, void* out_bf16, int variant, void* stream) {
  RTDF_REQUIRE(out_f32 || out_bf16, "rtdf_gemm_bf16: no output");
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(A);
  a.k_extent = K;
  a.rows_per_batch = M;
  a.batches = 1;
  a.row_stride = K;
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(W), N, K, TC_PLAIN, variant,
                 make_epi(bias, act, scale, resid, out_f32, out_bf16, N));
}

int rtdf_gemm_bf16_xres(const void* A, const void* W, int M, int N, int K, const float* bias, float* x_inout, void* xb_out,
                        float* stats_out, int variant, void* stream) {
  RTDF_REQUIRE(A && W && x_inout && xb_out && stats_out, "rtdf_gemm_bf16_xres: null argument");
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(A);
  a.k_extent = K;
  a.rows_per_batch = M;
  a.batches = 1;
  a.row_stride = K;
  TcEpilogue e = make_epi(bias, ACT_NONE, 1.0f, x_inout, x_inout, nullptr, N);
  e.xb_out = static_cast<bf16*>(xb_out);
  e.stats_out = reinterpret_cast<float2*>(stats_out);
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(W), N, K, TC_PLAIN, variant, e);
}

int rtdf_gemm_bf16_lnfold(const void* xb, const void* W_folded, int M, int N, int K, const float* c, const float* d,
                          const float* stats, float eps, int act, float* out_f32, void* out_bf16, int variant, void* stream) {
  RTDF_REQUIRE(xb && W_folded && c && d && stats && (out_f32 || out_bf16), "rtdf_gemm_bf16_lnfold: null argument");
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(xb);
  a.k_extent = K;
  a.rows_per_batch = M;
  a.batches = 1;
  a.row_stride = K;
  TcEpilogue e = make_epi(d, act, 1.0f, nullptr, out_f32, out_bf16, N);
  e.fold_c = c;
  e.fold_stats = reinterpret_cast<const float2*>(stats);
  e.fold_len = K;
  e.fold_eps = eps;
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(W_folded), N, K, TC_PLAIN, variant, e);
}

int rtdf_cast_stats_rows(float* x, const float* partials, int n_splits, long long rows, void* xb, float* stats, void* stream) {
  return cast_stats_rows(static_cast<cudaStream_t>(stream), x, partials, n_splits, rows, static_cast<bf16*>(xb),
                         reinterpret_cast<float2*>(stats));
}

int rtdf_fold_ln_weight(const float* w, const float* gamma, const float* beta, const float* bias, int n, int k, void* w_folded,
                        float* c, float* d, void* stream) {
  return fold_ln_weight(static_cast<cudaStream_t>(stream), w, gamma, beta, bias, n, k, static_cast<bf16*>(w_folded), c, d);
}

int rtdf_gemm_plan_splits(int M, int N, int K) { return tc_plan_splits(M, N, K); }

int rtdf_gemm_bf16_splitk(const void* A, const void* W, int M, int N, int K, const float* bias, float* partials,
                          void* stream) {
  RTDF_REQUIRE(A && W && partials, "rtdf_gemm_bf16_splitk: null argument");
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(A);
  a.k_extent = K;
  a.rows_per_batch = M;
  a.batches = 1;
  a.row_stride = K;
  TcEpilogue e;
  e.bias = bias;
  e.partials = partials;
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(W), N, K, TC_PLAIN, 64, e);
}

int rtdf_layernorm_accum_rows(float* x, const float* partials, int n_splits, long long rows, const float* gamma,
                              const float* beta, float eps, float* out_f32, void* out_bf16, void* stream) {
  return layernorm_accum_rows(static_cast<cudaStream_t>(stream), x, partials, n_splits, rows, gamma, beta, eps, out_f32,
                              static_cast<bf16*>(out_bf16));
}

int rtdf_gemm_bf16_rowln(const void* A, const void* W, int M, int N, int K, const float* bias, float* x_inout,
                         const float* gamma, const float* beta, float eps, void* ln_out_bf16, float* ln_out_f32,
                         int32_t* counters, int variant, void* stream) {
  RTDF_REQUIRE(x_inout && counters && gamma && beta, "rtdf_gemm_bf16_rowln: null argument");
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(A);
  a.k_extent = K;
  a.rows_per_batch = M;
  a.batches = 1;
  a.row_stride = K;
  TcEpilogue e = make_epi(bias, ACT_NONE, 1.0f, x_inout, x_inout, nullptr, N);
  e.rowln_gamma = gamma;
  e.rowln_beta = beta;
  e.rowln_eps = eps;
  e.rowln_out_bf16 = static_cast<bf16*>(ln_out_bf16);
  e.rowln_out_f32 = ln_out_f32;
  e.rowln_counters = counters;
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(W), N, K, TC_PLAIN, variant, e);
}

int rtdf_gemm_f32(const float* A, const float* W, int M, int N, int K, const float* bias, int act, float scale,
                  const float* resid, float* out_f32, void* stream) {
  RTDF_REQUIRE(out_f32, "rtdf_gemm_f32: no output");
  SimtOperandA a;
  a.ptr = A;
  a.k_extent = K;
  a.rows_per_batch = M;
  a.batches = 1;
  a.row_stride = K;
  return simt_gemm_f32(static_cast<cudaStream_t>(stream), a, W, N, K, make_epi(bias, act, scale, resid, out_f32, nullptr, N));
}

int rtdf_conv1d_ln_gelu_bf16(const void* x, int batch, int l_in, int k, int stride, const void* w_packed,
                             const float* bias, const float* gamma, const float* beta, float eps, void* y,
                             int variant, void* stream) {
  RTDF_REQUIRE(x && w_packed && y && l_in >= k && k >= 1 && stride >= 1, "rtdf_conv1d_ln_gelu_bf16: bad arguments");
  const int l_out = (l_in - k) / stride + 1;
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(x);
  a.k_extent = (long long)k * 512;
  a.rows_per_batch = l_out;
  a.batches = batch;
  a.row_stride = (long long)stride * 512;
  a.batch_stride = (long long)l_in * 512;
  TcEpilogue e;
  e.bias = bias;
  e.act = ACT_GELU;
  e.ln_gamma = gamma;
  e.ln_beta = beta;
  e.ln_eps = eps;
  e.out_bf16 = static_cast<bf16*>(y);
  e.ld_bf16 = 512;
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(w_packed), 512, k * 512, TC_PLAIN,
                 variant, e);
}

int rtdf_posconv_bf16(float* x_f32, const void* x_bf16, int batch, int n_frames, const void* w_packed,
                      const float* bias, int impl, void* stream) {
  RTDF_REQUIRE(x_f32 && x_bf16 && w_packed && bias, "rtdf_posconv_bf16: bad arguments");
  if (impl == 0)
    return posconv_tc(static_cast<cudaStream_t>(stream), x_f32, static_cast<const bf16*>(x_bf16), batch, n_frames,
                      static_cast<const bf16*>(w_packed), bias);
  TcOperandA a;
  a.ptr = static_cast<const bf16*>(x_bf16);
  a.k_extent = 1024;
  a.rows_per_batch = n_frames;
  a.batches = batch;
  a.row_stride = 1024;
  a.batch_stride = (long long)n_frames * 1024;
  TcEpilogue e;
  e.bias = bias;
  e.act = ACT_GELU;
  e.resid = x_f32;
  e.ldr = 1024;
  e.out_f32 = x_f32;
  e.ld_f32 = 1024;
  return tc_gemm(static_cast<cudaStream_t>(stream), a, static_cast<const bf16*>(w_packed), 1024, 8192, TC_POSCONV, 64, e);
}

int rtdf_posconv_f32(float* x, const float* x_in, int batch, int n_frames, const float* w_packed, const float* bias,
                     void* stream) {
  return posconv_f32(static_cast<cudaStream_t>(stream), x, x_in, batch, n_frames, w_packed, bias);
}

int rtdf_attention(const void* qkv, void* ctx_out, int batch, int n_frames, int heads, int is_bf16, int impl,
                   void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_bf16) {
    if (impl == 0) return attention_ws(s, static_cast<const bf16*>(qkv), static_cast<bf16*>(ctx_out), batch, n_frames, heads);
    if (impl == 2) return attention_tc(s, static_cast<const bf16*>(qkv), static_cast<bf16*>(ctx_out), batch, n_frames, heads);
    return attention_simt_bf16(s, static_cast<const bf16*>(qkv), static_cast<bf16*>(ctx_out), batch, n_frames, heads);
  }
  RTDF_REQUIRE(impl == 1, "rtdf_attention: the tcgen05 kernel takes bf16 inputs");
  return attention_simt_f32(s, static_cast<const float*>(qkv), static_cast<float*>(ctx_out), batch, n_frames, heads);
}

int rtdf_conv_planes_tc(const void* in_hi, const void* in_lo, int ci, long long rows, int hp, int wp,
                        const void* w_hi, const void* w_lo, int co, int n_chunks, const int* shift, const int* sub,
                        int hp_lo, int hp_hi, const float* bias, const float* s1, const float* t1, int act1,
                        const float* resid, const float* s2, const float* t2, int act2, float* out_f32,
                        void* out_hi, void* out_lo, int nsplit, void* stream) {
  RTDF_REQUIRE(shift && n_chunks >= 1 && n_chunks <= kConvTcMaxChunks, "rtdf_conv_planes_tc: bad chunk table");
  ConvTcArgs a;
  a.in_hi = static_cast<const bf16*>(in_hi);
  a.in_lo = static_cast<const bf16*>(in_lo);
  a.ci = ci;
  a.rows = rows;
  a.Hp = hp;
  a.Wp = wp;
  a.w_hi = static_cast<const bf16*>(w_hi);
  a.w_lo = static_cast<const bf16*>(w_lo);
  a.co = co;
  a.n_chunks = n_chunks;
  for (int c = 0; c < n_chunks; ++c) {
    a.shift[c] = shift[c];
    a.sub[c] = sub ? sub[c] : 0;
  }
  a.hp_lo = hp_lo;
  a.hp_hi = hp_hi;
  a.bias = bias; a.s1 = s1; a.t1 = t1; a.act1 = act1; a.resid = resid; a.s2 = s2; a.t2 = t2; a.act2 = act2;
  a.out_f32 = out_f32;
  a.out_hi = static_cast<bf16*>(out_hi);
  a.out_lo = static_cast<bf16*>(out_lo);
  return conv_tc(static_cast<cudaStream_t>(stream), a, nsplit);
}

long long rtdf_launch_count(void) { return launch_count(); }
int rtdf_debug_gelu_variant(int act) {
  tc_set_gelu_variant(act);
  return RTDF_OK;
}
int rtdf_profile_begin(void) {
  tc_profile_begin();
  return RTDF_OK;
}
int rtdf_profile_end(int variant, double* ms_total, double* flops_total, int* launches) {
  return tc_profile_end(variant, ms_total, flops_total, launches);
}

static GatRowWeights gat_w(const rtdf_gat_weights* w) {
  GatRowWeights g;
  g.att_w = w->att_w; g.att_b = w->att_b; g.a11 = w->a11; g.a22 = w->a22; g.a12 = w->a12;
  g.with_t = w->with_t; g.with_b = w->with_b; g.without_t = w->without_t; g.without_b = w->without_b;
  g.bn_s = w->bn_s; g.bn_t = w->bn_t; g.inv_temp = w->inv_temp;
  return g;
}

int rtdf_gat_rows(int d, int dout, const float* x, int batch, int n, int n1, const rtdf_gat_weights* w, float* out,
                  const float* master_in, const rtdf_gat_weights* wm, float* master_out, int impl, void* stream) {
  RTDF_REQUIRE(x && w && out, "rtdf_gat_rows: null argument");
  GraphView v;
  v.ptr = x;
  v.n = n;
  v.batch_stride = (long long)n * d;
  const GatRowWeights g = gat_w(w);
  GatRowWeights gm;
  if (wm) gm = gat_w(wm);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 0)
    return aasist_gat_rows_mma(s, d, dout, v, batch, n1, g, out, (long long)n * dout, master_in, d, wm ? &gm : nullptr, master_out);
  return aasist_gat_rows(s, d, dout, v, batch, n1, g, out, (long long)n * dout, master_in, d, wm ? &gm : nullptr, master_out);
}

int rtdf_graph_pool(const float* h, int batch, int n, int d, const float* w, const float* b, int k, float* out,
                    int32_t* idx, void* stream) {
  GraphView g;
  g.ptr = h;
  g.n = n;
  g.batch_stride = (long long)n * d;
  return aasist_graph_pool(static_cast<cudaStream_t>(stream), d, g, batch, w, b, k, out, idx);
}

// ---- Conformer block kernels (lucidrains ConformerBlock instantiated at reference models/conformer_baseline.py:16-18) ----
int rtdf_conformer_attention(const void* qkv, const float* rel_pos, void* out, int batch, int n, int heads, int dh,
                             int is_bf16, int impl, void* stream) {
  RTDF_REQUIRE(qkv && rel_pos && out, "rtdf_conformer_attention: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!is_bf16)
    return conformer_attention_f32(s, static_cast<const float*>(qkv), rel_pos, static_cast<float*>(out), batch, n, heads, dh);
  if (impl == 0)   // tensor-core kernel only: outside its envelope this reports RTDF_STATUS_UNSUPPORTED instead of falling back
    return conformer_attention_mma(s, static_cast<const bf16*>(qkv), rel_pos, static_cast<bf16*>(out), batch, n, heads, dh);
  return conformer_attention_bf16(s, static_cast<const bf16*>(qkv), rel_pos, static_cast<bf16*>(out), batch, n, heads, dh);
}

int rtdf_conformer_glu_dwconv(const void* in, void* out, int batch, int n, int inner, int k, const float* w,
                              const float* bias, const float* bn_s, const float* bn_t, int is_bf16, void* stream) {
  RTDF_REQUIRE(in && out && w && bias && bn_s && bn_t, "rtdf_conformer_glu_dwconv: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_bf16)
    return conformer_glu_dwconv_bf16(s, static_cast<const bf16*>(in), static_cast<bf16*>(out), batch, n, inner, k, w, bias,
                                     bn_s, bn_t);
  return conformer_glu_dwconv_f32(s, static_cast<const float*>(in), static_cast<float*>(out), batch, n, inner, k, w, bias,
                                  bn_s, bn_t);
}

}  // extern "C"
