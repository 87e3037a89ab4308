/*
 * rtdf.h -- C-ABI of librtdf.so: B200-native (sm_100a) batched evaluation-scoring forward for
 * hungdinhxuan/real-time-deepfake-speech-detection (waveform -> XLS-R -> AASIST | Conformer -> logits).
 *
 * The reference has no FFI/plugin layer: its boundary is the nn.Module contract of the model classes
 * (reference models/xlsr_aasist.py:5-177, models/conformer_baseline.py:31-99, models/fe.py:8-99) as
 * consumed by main.py:199-221 (produce_evaluation_file) and trainer.py:85-132 (Trainer._test).  The
 * entry points below are what a binding for that path calls; the Python mirror of the model classes in
 * real-time-deepfake-speech-detection_b200/models/ binds them with ctypes.
 *
 * Conventions: every function returns 0 on success or a negative rtdf_status; the message is available
 * from rtdf_last_error() (thread-local).  The CALLER owns all input/output/workspace buffers (device
 * memory); the context owns only packed weights.  All work is enqueued on the caller's stream; there are
 * no internal synchronisations and no allocations in the forward calls (CUDA-graph capturable).
 * One context per (device, model instance); a context is not thread-safe.
 */
#ifndef RTDF_H_
#define RTDF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtdf_ctx rtdf_ctx;

enum rtdf_status {
  RTDF_STATUS_OK = 0,
  RTDF_STATUS_INVALID = -1,      /* bad argument / shape                         */
  RTDF_STATUS_CUDA = -2,         /* CUDA runtime / driver error                  */
  RTDF_STATUS_STATE = -3,        /* call order violated (e.g. not finalized)     */
  RTDF_STATUS_UNSUPPORTED = -4   /* configuration outside the implemented path   */
};

enum rtdf_backend { RTDF_BACKEND_AASIST = 0, RTDF_BACKEND_CONFORMER = 1, RTDF_BACKEND_NONE = 2 };
enum rtdf_precision {
  RTDF_PREC_BF16 = 0,            /* tcgen05 bf16 x bf16 -> fp32 GEMMs, fp32 residual stream / statistics */
  RTDF_PREC_FP32 = 1             /* FFMA-only verification mode (max |logit diff| <= 1e-4)               */
};

/* Architecture of one model instance.  Mirrors the reference constructors' kwargs:
 * num_layers (models/fe.py:57), emb_size / heads / kernel_size / n_encoders
 * (models/conformer_baseline.py:38-41). */
typedef struct rtdf_model_desc {
  int backend;        /* rtdf_backend                                               */
  int n_layers;       /* transformer layers kept in ssl_model.model.encoder.layers  */
  int precision;      /* rtdf_precision                                             */
  int conf_emb;       /* Conformer: emb_size (144)                                  */
  int conf_heads;     /* Conformer: heads (4)                                       */
  int conf_kernel;    /* Conformer: depth-wise conv kernel size (31)                */
  int conf_blocks;    /* Conformer: n_encoders (4)                                  */
  int attention_impl; /* 0 = tcgen05 warp-specialised kernel, 1 = SIMT (debug), 2 = tcgen05 tile kernel (A/B) */
  int aasist_conv_impl; /* bf16 mode, residual-block + attention-map convs: 0 = tcgen05 with (hi,lo) bf16 operand
                         * pairs, 3 MMAs per product (~fp32 accuracy); 1 = tcgen05 plain bf16; 2 = fp32 SIMT */
  int gat_impl;       /* bf16 mode, graph-attention rows: 0 = tensor cores (mma.sync, (hi,lo) bf16 pairs), 1 = fp32 SIMT */
} rtdf_model_desc;

/* Kernel-selection regime of the forward calls.  AUTO: batches of at most 512 frames in flight (streaming chunks,
 * batch 1-8 x 1 s) use the weight-streaming tiles (64-wide, deterministic split-K) -- lowest latency, but in bf16 an
 * utterance's logits then depend at rounding level on how many utterances share its batch.  THROUGHPUT: always the
 * large-batch tiles; per-utterance results are independent of batch composition (bit-identical), which is what the
 * sharded scoring sweep (ragged last batch, 1/2/4/8 GPUs) relies on.  The reference has no counterpart (cuBLAS picks
 * kernels by shape too); fp32 mode is batch-invariant in both regimes. */
enum rtdf_regime { RTDF_REGIME_AUTO = 0, RTDF_REGIME_THROUGHPUT = 1 };

/* Optional intermediate outputs of a forward call (device pointers, may be NULL). */
typedef struct rtdf_taps {
  float* feats;    /* (B, T, 1024) XLS-R output features (fe.py:17-21 'x')          */
  float* hidden;   /* (B, 160) AASIST read-out vector (xlsr_aasist.py:171-172)      */
  int32_t* idx_S;  /* (B, 21)  nodes kept by pool_S, descending score               */
  int32_t* idx_T;  /* (B, T'/2) nodes kept by pool_T                                */
  float* layers;   /* (n_layers+1, B*T, 1024) residual stream entering layer 0 and leaving every transformer layer:
                    * the I/O of `ssl_model.model.encoder.layers.N` that the KD forward hooks read
                    * (trainer.py:176-195, main_kd.py:88-141)                      */
} rtdf_taps;

/* ---- lifecycle ------------------------------------------------------------------------------- */
int rtdf_create(rtdf_ctx** out, int device, const rtdf_model_desc* desc);
/* Copies one state-dict entry (fp32, device OR host pointer, contiguous) into the context.  `key` is the
 * reference state-dict key without any "module." prefix (utils.py:13-43), e.g.
 * "ssl_model.model.encoder.layers.3.fc1.weight".  Unknown keys are stored and ignored. */
int rtdf_load_weight(rtdf_ctx* ctx, const char* key, const void* data, const int64_t* shape, int ndim);
/* Packs weights: bf16 casts, fused QKV (1/8 folded into W_q, b_q), weight-norm fold of pos_conv,
 * eval-BatchNorm folds, implicit-GEMM conv weight layout.  Fails listing the first missing key. */
int rtdf_finalize(rtdf_ctx* ctx);
/* Selects the rtdf_regime used by subsequent forward calls on this context (default RTDF_REGIME_AUTO). */
int rtdf_set_regime(rtdf_ctx* ctx, int regime);
void rtdf_destroy(rtdf_ctx* ctx);
const char* rtdf_last_error(void);

/* ---- scoring forward ------------------------------------------------------------------------- */
/* Number of frames T for an N-sample utterance (7 strided convs). */
int rtdf_num_frames(int n_samples);
int rtdf_workspace_bytes(const rtdf_ctx* ctx, int batch, int n_samples, size_t* out);
/* wav: (B,N) fp32 device.  preemph != 0 applies y[t] = x[t] - coef*x[t-1] first (data/preprocess.py:22-27;
 * used by trainer.py:104, not by main.py:208-214).  logits: (B,2) fp32 device.  stream: cudaStream_t. */
int rtdf_forward(rtdf_ctx* ctx, const float* wav, int batch, int n_samples, int preemph, float preemph_coef,
                 float* logits, void* workspace, size_t workspace_bytes, const rtdf_taps* taps, void* stream);
/* Front-end only: wav -> feats (B,T,1024) fp32   (XLSR_FE.extract_feat, models/fe.py:17-21). */
int rtdf_frontend(rtdf_ctx* ctx, const float* wav, int batch, int n_samples, int preemph, float preemph_coef,
                  float* feats, void* workspace, size_t workspace_bytes, void* stream);
/* Back-end only: feats (B,T,1024) fp32 -> logits (B,2). */
int rtdf_backend(rtdf_ctx* ctx, const float* feats, int batch, int n_frames, float* logits, void* workspace,
                 size_t workspace_bytes, const rtdf_taps* taps, void* stream);

/* ---- single kernels (unit parity tests; all pointers are device pointers) --------------------- */
int rtdf_preemph(const float* x, float* y, int batch, int n, float coef, void* stream);
int rtdf_wave_layernorm(const float* x, float* y, int batch, int n, float eps, void* stream);
int rtdf_conv0_ln_gelu(const float* wav, int batch, int n, const float* w_tapmajor /*[10][512]*/, const float* bias,
                       const float* gamma, const float* beta, float eps, float* out_f32 /*or NULL*/,
                       void* out_bf16 /*or NULL*/, void* stream);
/* The same layer on the tensor cores (what the forward uses for large batches in bf16 mode): conv-0 lowered to an
 * implicit GEMM with K = 10 taps padded to 32, waveform and weights split into bf16 (hi, lo) pairs laid side by side along
 * K (three of the four partial products, ~2^-16 relative), bias + LayerNorm(512) + GELU in the tcgen05 tile's epilogue.
 * w: state-dict layout [512][1][10] fp32.  scratch: rtdf_conv0_tc_scratch_bytes(batch, n) bytes, 128-byte aligned. */
long long rtdf_conv0_tc_scratch_bytes(int batch, int n);
int rtdf_conv0_tc_ln_gelu(const float* wav, int batch, int n, const float* w, const float* bias, const float* gamma,
                          const float* beta, float eps, void* scratch, void* out_bf16, void* stream);
/* conv-0 of the feature encoder in fairseq's extractor_mode="default" (wav2vec2-base style; the alternative SURVEY.md
 * App. A.2 step 1 describes): Conv1d(1, 512, 10, stride 5, bias optional -- NULL when conv_bias=False) ->
 * GroupNorm(512 groups, 512 channels), i.e. per-(utterance, channel) statistics over time -> GELU; channels-last output
 * (B, L1, 512).  workspace: rtdf_conv0_gn_workspace_floats(batch, n) floats of device scratch. */
long long rtdf_conv0_gn_workspace_floats(int batch, int n);
int rtdf_conv0_gn_gelu(const float* wav, int batch, int n, const float* w_tapmajor /*[10][512]*/, const float* bias,
                       const float* gamma, const float* beta, float eps, float* workspace, float* out_f32 /*or NULL*/,
                       void* out_bf16 /*or NULL*/, void* stream);
int rtdf_layernorm_rows(const void* in, int in_is_bf16, long long rows, int cols, const float* gamma,
                        const float* beta, float eps, int act, float* out_f32, void* out_bf16, void* stream);
/* D = act(A W^T + bias) * scale + resid.  A: (M,K) bf16, W: (N,K) bf16.  variant: tile width 64|128|256 (one CTA
 * per 128 x variant tile) or 2256 (CTA pair, cta_group::2, per 256 x 256 tile). */
int rtdf_gemm_bf16(const void* A, const void* W, int M, int N, int K, const float* bias, int act, float scale,
                   const float* resid, float* out_f32, void* out_bf16, int variant, void* stream);
/* Residual GEMM with the following LayerNorm fused behind it (the per-layer pattern out_proj -> LN, fc2 -> LN of
 * fairseq's TransformerSentenceEncoderLayer):  x (M,N) fp32 += A W^T + bias, then, per 128-row block as soon as its
 * last N-tile has been added,  ln_out = LN(x row) * gamma + beta  (bf16 and/or fp32).  counters: ceil(M/128) int32,
 * zero on entry (left zero on exit).  variant 256 | 2256 | 64; N % 128 == 0, N <= 1024. */
int rtdf_gemm_bf16_rowln(const void* A, const void* W, int M, int N, int K, const float* bias, float* x_inout,
                         const float* gamma, const float* beta, float eps, void* ln_out_bf16, float* ln_out_f32,
                         int32_t* counters, int variant, void* stream);
/* LayerNorm folded into the GEMMs on either side of it -- how the bf16 transformer layers run (the two LayerNorms of
 * fairseq's TransformerSentenceEncoderLayer never execute as kernels):
 *   LN(x) W^T + b  ==  rstd_i * (bf16(x) W'^T - mean_i * c) + d,   W' = bf16(W diag(gamma)), c_j = sum_k W'_jk, d = b + W beta.
 * rtdf_fold_ln_weight   packs W (n,k fp32), gamma, beta, bias -> W' (bf16), c, d.
 * rtdf_gemm_bf16_xres   residual GEMM x (M,N fp32) += A W^T + bias that also emits bf16(x) and, per row, 8 partial
 *                       (sum, sum of squares) pairs -- stats (M, 8, 2) fp32, slot = 256-column tile * 2 + half; variant 256 |
 *                       2256, N = 1024 fills all 8 slots (smaller N: zero the buffer first).
 * rtdf_cast_stats_rows  the same two outputs from a resident row (N = 1024): xb = bf16(x), totals in slot 0, slots 1..7
 *                       zero; n_splits > 0 first adds the K-split partial sums like rtdf_layernorm_accum_rows.
 * rtdf_gemm_bf16_lnfold out = act(rstd_i * (xb W'^T - mean_i c) + d), row statistics over K columns from the 8 partials. */
int rtdf_fold_ln_weight(const float* w, const float* gamma, const float* beta, const float* bias, int n, int k, void* w_folded,
                        float* c, float* d, void* stream);
int rtdf_gemm_bf16_xres(const void* A, const void* W, int M, int N, int K, const float* bias, float* x_inout, void* xb_out,
                        float* stats_out, int variant, void* stream);
int rtdf_cast_stats_rows(float* x, const float* partials, int n_splits, long long rows, void* xb, float* stats, void* stream);
int rtdf_gemm_bf16_lnfold(const void* xb, const void* W_folded, int M, int N, int K, const float* c, const float* d,
                          const float* stats, float eps, int act, float* out_f32, void* out_bf16, int variant, void* stream);
/* Skinny GEMMs of streaming chunks (batch 1-8 x 1 s: M = 49..392 rows): the op is a weight stream, so K is split over
 * the otherwise idle SMs.  rtdf_gemm_plan_splits gives the number of splits S for a shape (1 = not split);
 * rtdf_gemm_bf16_splitk (S > 1 only) writes split s's partial sum of A W^T (+ bias on split 0) to
 * partials[s][M][N] fp32 (64-wide tiles, plain TMA stores);  rtdf_layernorm_accum_rows (N = 1024) folds them into the
 * residual stream in split order -- x += sum_s partials[s], written back -- and emits LN(x) as bf16 or fp32.  Together
 * they replace  x += out_proj(attn) ; LN  and  x += fc2(h) ; LN  of a transformer layer, bit-reproducibly. */
int rtdf_gemm_plan_splits(int M, int N, int K);
int rtdf_gemm_bf16_splitk(const void* A, const void* W, int M, int N, int K, const float* bias, float* partials,
                          void* stream);
int rtdf_layernorm_accum_rows(float* x, const float* partials, int n_splits, long long rows, const float* gamma,
                              const float* beta, float eps, float* out_f32, void* out_bf16, void* stream);
int rtdf_gemm_f32(const float* A, const float* W, int M, int N, int K, const float* bias, int act, float scale,
                  const float* resid, float* out_f32, void* stream);
/* Strided 1-D conv as implicit GEMM on channels-last bf16 activations, fused bias + LayerNorm(512) + GELU:
 * x (B, L_in, 512) -> y (B, L_out, 512), w packed [512][k][512] bf16.  variant 512 (BK 64) | 513 (BK 32): single
 * 512-column accumulator; 514 (BK 32) | 515 (BK 64): pipelined two-pass tile with 16 epilogue warps; 516: the same
 * as a CTA pair (cta_group::2, 256 rows per cluster: half the weight stream per output row). */
int rtdf_conv1d_ln_gelu_bf16(const void* x, int batch, int l_in, int k, int stride, const void* w_packed,
                             const float* bias, const float* gamma, const float* beta, float eps, void* y,
                             int variant, void* stream);
/* Grouped positional conv (k=128, groups=16, pad 64, last frame dropped) + GELU + residual:
 * x_f32 (B,T,1024) += gelu(conv(x_bf16) + bias).  w packed [1024][128*64] bf16 (k index = tap*64 + ci). */
/* impl 0 = slab-resident kernel (receptive field of a 256-frame tile loaded once, taps = descriptor row offsets),
 * 1 = tap-shifted GEMM (A tile re-fetched per tap; kept for A/B timing). */
int rtdf_posconv_bf16(float* x_f32, const void* x_bf16, int batch, int n_frames, const void* w_packed,
                      const float* bias, int impl, void* stream);
int rtdf_posconv_f32(float* x, const float* x_in, int batch, int n_frames, const float* w_packed, const float* bias,
                     void* stream);
/* qkv (B*T, 3*H*64) [q|k|v] with q pre-scaled -> ctx (B*T, H*64).  impl 0 = tcgen05 warp-specialised (P in
 * tensor memory; T <= 512 frames: two ping-pong TMEM buffers up to 256 frames, one 512-column buffer above),
 * 1 = SIMT (any T that fits shared memory), 2 = tcgen05 one-tile-per-CTA kernel (T <= 256). */
int rtdf_attention(const void* qkv, void* ctx_out, int batch, int n_frames, int heads, int is_bf16, int impl,
                   void* stream);
/* Weights of one graph-attention row pass (device pointers).  *_t matrices are transposed to [D][DO]; bn_s / bn_t are the
 * eval BatchNorm1d folded to scale / shift (NULL for the master row); a22 / a12 NULL for a homogeneous GAT. */
typedef struct rtdf_gat_weights {
  const float *att_w, *att_b, *a11, *a22, *a12, *with_t, *with_b, *without_t, *without_b, *bn_s, *bn_t;
  float inv_temp;
} rtdf_gat_weights;
/* GraphAttentionLayer / HtrgGraphAttentionLayer rows (aasist_modules.py:17-294) on x (B,n,D) -> out (B,n,DO):
 * e_ij = a . tanh(W (x_i*x_j) + b) / temp, softmax_j, aggregation, proj_with_att + proj_without_att, BN, SELU.
 * n1: number of type-1 nodes (quadrant vectors a11 / a22 / a12); n1 = n for a homogeneous GAT.  master_in (B,D) with
 * wm != NULL additionally updates the master node -> master_out (B,DO).  impl 0 = tensor cores, 1 = fp32 SIMT. */
int rtdf_gat_rows(int d, int dout, const float* x, int batch, int n, int n1, const rtdf_gat_weights* w, float* out,
                  const float* master_in, const rtdf_gat_weights* wm, float* master_out, int impl, void* stream);
/* GraphPool (aasist_modules.py:306-338) on h (B,n,D): out (B,k,D), idx (B,k) descending score. */
int rtdf_graph_pool(const float* h, int batch, int n, int d, const float* w, const float* b, int k, float* out,
                    int32_t* idx, void* stream);

/* Conformer block kernels.  The reference instantiates lucidrains' ConformerBlock(dim=144, dim_head=36, heads=4,
 * ff_mult=4, conv_expansion_factor=2, conv_kernel_size=31) at models/conformer_baseline.py:16-18; these two entry
 * points are its two non-GEMM stages.
 * Shaw relative-position multi-head self-attention: qkv (B*n, 3*heads*dh) = [q | k | v] (no bias), rel_pos [1025][dh]
 * fp32 (Embedding, max_pos_emb 512): out (B*n, heads*dh) = softmax_j(dh^-1/2 (q_i.k_j + q_i.R[clamp(i-j,-512,512)+512])) v.
 * is_bf16: 0 = fp32 SIMT kernel; 1 with impl 0 = tensor-core kernel (mma.sync; n <= 208, dh <= 40 even, else
 * RTDF_STATUS_UNSUPPORTED), impl 1 = what the forward uses (tensor cores inside the envelope, bf16 SIMT outside). */
int rtdf_conformer_attention(const void* qkv, const float* rel_pos, void* out, int batch, int n, int heads, int dh,
                             int is_bf16, int impl, void* stream);
/* Convolution module after the first pointwise conv: in (B*n, 2*inner) -> GLU over channels (first half * sigmoid(second
 * half)) -> depth-wise Conv1d(inner, k, groups=inner, "same" padding (k/2, k/2 - (k+1)%2)) + bias -> eval BatchNorm1d
 * folded to bn_s / bn_t -> Swish -> out (B*n, inner).  w: [inner][k] fp32. */
int rtdf_conformer_glu_dwconv(const void* in, void* out, int batch, int n, int inner, int k, const float* w,
                              const float* bias, const float* bn_s, const float* bn_t, int is_bf16, void* stream);

/* Shifted-row tcgen05 convolution on zero-padded channels-last planes (bf16-mode AASIST residual encoder and
 * attention map; replaces the cuDNN convs behind models/aasist_modules.py:340-397, models/xlsr_aasist.py:103).
 * Planes: row m = (b*hp + y)*wp + x holds the channels of one pixel; x = 0 / wp-1 and unused y rows are zero.
 * in_hi/in_lo: (rows, ci) bf16 operand pair (lo = v - hi; in_lo/w_lo may be NULL when nsplit == 1);
 * w_hi/w_lo: [n_chunks][co][min(ci,64)]; chunk c multiplies the plane shifted by shift[c] rows, channel block
 * sub[c].  Epilogue: v = acc + bias; v = v*s1 + t1; act1; v += resid; v = v*s2 + t2; act2 (act: 0 none, 3 SELU);
 * rows with y outside [hp_lo, hp_hi] or x in {0, wp-1} are written as zero.  nsplit: 3 = hi*hi + lo*hi + hi*lo
 * (~fp32 accuracy), 1 = plain bf16. */
int rtdf_conv_planes_tc(const void* in_hi, const void* in_lo, int ci, long long rows, int hp, int wp,
                        const void* w_hi, const void* w_lo, int co, int n_chunks, const int* shift, const int* sub,
                        int hp_lo, int hp_hi, const float* bias, const float* s1, const float* t1, int act1,
                        const float* resid, const float* s2, const float* t2, int act2, float* out_f32,
                        void* out_hi, void* out_lo, int nsplit, void* stream);

/* ---- the two ends of the path ------------------------------ */
/* Fixed-length batch from ragged utterances (reference data/test_set.py:201-248 adjustDuration and
 * adjustDuration_random_start), optionally fused with PreEmphasis (data/preprocess.py:22-27):
 *   out[b][t] = fitted_b[t] - (preemph ? coef * fitted_b[t-1] : 0),  fitted_b[t] = src_b[(starts[b] + t) mod len_b],
 *   fitted_b[-1] := fitted_b[1].  packed: concatenated fp32 samples (device), offsets: int64[batch+1] (device),
 *   starts: int32[batch] or NULL (= 0; the reference draws it with random.randint on the host).  Utterances shorter
 *   than `duration` are tile-repeated (they cross PCIe once), longer ones are cropped at starts[b]. */
int rtdf_fit_duration(const float* packed, const long long* offsets, const int* starts, int batch, int duration,
                      int preemph, float coef, float* out, void* stream);
/* Score sink for one batch: scores[i] = logits[i][1] (main.py:212) and, when labels != NULL, the evaluation
 * accumulators of Trainer._test (trainer.py:104-113) kept on the device: acc[0] += batch * CrossEntropyLoss(
 * weight=class_weight, reduction='mean'), acc[1] += #(argmax == label), acc[2] += batch.  labels: int64[batch],
 * class_weight: fp32[2] or NULL, scores: fp32[batch] or NULL, acc: double[3]. */
int rtdf_score_sink(const float* logits, int batch, const long long* labels, const float* class_weight, float* scores,
                    double* acc, void* stream);
/* Integer ROC for the EER (trainer.py:134-139 calculate_EER): tp[i] / fp[i] = number of positives / negatives whose
 * score is >= scores[i].  NaN scores (gather padding) are skipped and get tp = fp = -1.  labels: int64[n], 1 = bona fide. */
int rtdf_roc_counts(const float* scores, const long long* labels, int n, int32_t* tp, int32_t* fp, void* stream);
/* The two ROC points that bracket the EER: keys[0] = ((tp+fp) << 32 | tp) of the last point with
 * 1 - fp/n_neg - tp/n_pos >= 0 (0 if only the origin qualifies), keys[1] = same packing for the first point past it. */
int rtdf_roc_crossing(const int32_t* tp, const int32_t* fp, int n, long long n_pos, long long n_neg,
                      unsigned long long* keys, void* stream);

/* ---- instrumentation --------------------------------------------------------------------------- */
/* Total number of kernels this library has launched in the calling process. */
long long rtdf_launch_count(void);
/* Timing A/B switch: evaluate the GELU of the tensor-core epilogues with 4 = hardware-tanh form, 5 = A&S 7.1.26
 * erf; 0 restores the default (sigmoid-of-quintic fit of the erf GELU, |err| <= 2.6e-5). */
int rtdf_debug_gelu_variant(int act);
/* Bracket a region: every tcgen05 GEMM launch in it is timed with CUDA events on its own stream.
 * rtdf_profile_end sums duration (ms) and algorithmic FLOPs (2*M*N*K) of the launches of one tile
 * variant (64|128|256|512|513, or -1 for all). */
int rtdf_profile_begin(void);
int rtdf_profile_end(int variant, double* ms_total, double* flops_total, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* RTDF_H_ */
