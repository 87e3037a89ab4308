// Conformer back-end kernels (lucidrains ConformerBlock as used by reference
// models/conformer_baseline.py:8-64): class-token stem, Shaw relative-position MHSA,
// GLU + depth-wise conv + BatchNorm + Swish, classification head.  GEMMs go through tc_gemm / simt_gemm.
#pragma once
#include "common.cuh"

namespace rtdf {

// x[b,0,:] = class_token ; x[b,1+t,:] = SELU(z[b,t,:]*s + t0)        conformer_baseline.py:23-24,59-62
int conformer_stem(cudaStream_t s, const float* z, const float* class_token, int B, int T, int E, float bn_s,
                   float bn_t, float* x);

// qkv: (B*n, 3E) = [q | k | v] (no bias), heads x dh; rel_pos: [1025][dh] (Embedding, max_pos_emb 512).
// out (B*n, E) = softmax(scale * (q k^T + q R_{i-j}^T)) v
int conformer_attention_f32(cudaStream_t s, const float* qkv, const float* rel_pos, float* out, int B, int n, int heads, int dh);
int conformer_attention_bf16(cudaStream_t s, const bf16* qkv, const float* rel_pos, bf16* out, int B, int n, int heads, int dh);

// Same contract on the tensor cores (conformer_attn_mma.cu; bf16, n <= 208, dh <= 40 even); RTDF_ERR_UNSUPPORTED outside
// that envelope.  conformer_attention_bf16 uses it when it applies (RTDF_CONF_ATTN_IMPL=1 forces the SIMT kernel).
int conformer_attention_mma(cudaStream_t s, const bf16* qkv, const float* rel_pos, bf16* out, int B, int n, int heads, int dh);

// in: (B*n, 2*inner) pointwise-conv output; out (B*n, inner) = Swish(BN(depthwise_k(GLU(in)) + bias))
int conformer_glu_dwconv_f32(cudaStream_t s, const float* in, float* out, int B, int n, int inner, int k,
                             const float* w, const float* bias, const float* bn_s, const float* bn_t);
int conformer_glu_dwconv_bf16(cudaStream_t s, const bf16* in, bf16* out, int B, int n, int inner, int k,
                              const float* w, const float* bias, const float* bn_s, const float* bn_t);

// logits[b,:] = W x[b,0,:] + bias          conformer_baseline.py:27-28
int conformer_head(cudaStream_t s, const float* x, int B, int n, int E, const float* w, const float* bias, float* logits);

struct ConformerWs {
  void* featsb = nullptr;  // bf16 copy of feats
  float* z = nullptr;      // (B*T, E)
  float* x = nullptr;      // (B*n, E) fp32 residual stream
  void* yb = nullptr;      // (B*n, E) normalised
  void* h = nullptr;       // (B*n, 4E)
  void* ab = nullptr;      // (B*n, E)
  void* dw = nullptr;      // (B*n, 2E)
};

}  // namespace rtdf
