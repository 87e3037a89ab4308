// AASIST back-end kernels (fp32): stem, residual-block convs, attention pooling, fused graph
// attention (GAT / HS-GAL rows), top-k graph pooling, read-out.  Reference:
// models/xlsr_aasist.py:89-177 and models/aasist_modules.py.
#pragma once
#include "common.cuh"

namespace rtdf {

// z: (B,T,128) = LL(feats).  out (B,1,42,Tp): SELU(BN(max_pool2d(z^T, 3x3)))   xlsr_aasist.py:92-96
int aasist_stem(cudaStream_t s, const float* z, int B, int T, float bn_scale, float bn_shift, float* out);

struct Conv2dArgs {
  const float* in = nullptr;   // (B,Ci,H,W)
  int Ci = 0, H = 0, W = 0;
  const float* w = nullptr;    // packed [Ci][KH][3][Co]
  const float* bias = nullptr; // [Co]
  int Co = 0, KH = 1, pad_h = 0;  // kernel (KH,3), padding (pad_h,1), stride 1
  const float* s1 = nullptr;   // optional per-channel affine after bias (folded eval BatchNorm)
  const float* t1 = nullptr;
  int act1 = ACT_NONE;
  const float* resid = nullptr;  // optional (B,Co,Hout,W) added after act1
  const float* s2 = nullptr;   // optional per-channel affine after the residual add
  const float* t2 = nullptr;
  int act2 = ACT_NONE;
  float* out = nullptr;        // (B,Co,Hout,W), Hout = H + 2*pad_h - KH + 1
};
int aasist_conv2d(cudaStream_t s, const Conv2dArgs& a, int B);

// attention(x) = conv1x1_{128->64}(BN(SELU(conv1x1_{64->128}(x))))           xlsr_aasist.py:103
// w1t: [64][128] (transposed), w2t: [128][64] (transposed), bn folded to s/t [128]
int aasist_attn_map(cudaStream_t s, const float* x, int B, int H, int W, const float* w1t, const float* b1,
                    const float* bn_s, const float* bn_t, const float* w2t, const float* b2, float* wmap);
// e_S[b,h,:] = sum_w x*softmax_w(wmap) + pos_S[h,:];  e_T[b,w,:] = sum_h x*softmax_h(wmap)   :106-118
int aasist_attn_pool(cudaStream_t s, const float* x, const float* wmap, int B, int H, int W, const float* pos_S,
                     float* e_S, float* e_T);

struct GraphView {           // (B, n, D) with arbitrary batch stride (rows contiguous: row stride = D)
  const float* ptr = nullptr;
  int n = 0;
  long long batch_stride = 0;
};

struct GatRowWeights {
  const float* att_w = nullptr;    // [DO][D]  att_proj.weight (or att_projM)
  const float* att_b = nullptr;    // [DO]
  const float* a11 = nullptr;      // [DO] attention vectors; GAT / master rows use a11 only
  const float* a22 = nullptr;
  const float* a12 = nullptr;
  const float* with_t = nullptr;   // [D][DO] proj_with_att.weight^T
  const float* with_b = nullptr;
  const float* without_t = nullptr;  // [D][DO] proj_without_att.weight^T
  const float* without_b = nullptr;
  const float* bn_s = nullptr;     // folded BatchNorm1d (null for the master row)
  const float* bn_t = nullptr;
  float inv_temp = 1.f;
};

// One CTA per (node i, utterance): scores e_ij, softmax_j, aggregation, projections, BN, SELU.
// n1 = number of type-1 nodes (HS-GAL quadrant selection); pass n1 = x.n for a homogeneous GAT.
// If master_in != nullptr an extra row (blockIdx.x == n) updates the master node with wM.
int aasist_gat_rows(cudaStream_t s, int D, int DO, const GraphView& x, int B, int n1, const GatRowWeights& w,
                    float* out, long long out_batch_stride, const float* master_in, long long master_stride,
                    const GatRowWeights* wM, float* master_out);

// Same contract on the tensor cores (gat_mma.cu): one warp per node, pair products built in registers, (hi, lo) bf16
// operand pairs with three mma.sync per product (~fp32 accuracy).  Used by the bf16-mode back-end.
int aasist_gat_rows_mma(cudaStream_t s, int D, int DO, const GraphView& x, int B, int n1, const GatRowWeights& w,
                        float* out, long long out_batch_stride, const float* master_in, long long master_stride,
                        const GatRowWeights* wM, float* master_out);

// HS-GAL type projections: out[:, :n1] = W1 x1 + b1, out[:, n1:] = W2 x2 + b2     aasist_modules.py:159-164
int aasist_type_proj(cudaStream_t s, int D, const GraphView& x1, const GraphView& x2, int B, const float* w1t,
                     const float* b1, const float* w2t, const float* b2, float* out);

// GraphPool: s = sigmoid(w.h + b); keep k = max(floor(n/2),1) best (descending score; ties -> lower
// index first); out[r] = h[idx_r] * s[idx_r].  idx_out (optional): (B,k) int32.   aasist_modules.py:306-338
int aasist_graph_pool(cudaStream_t s, int D, const GraphView& h, int B, const float* w, const float* b, int k,
                      float* out, int* idx_out);

struct ReadoutArgs {
  // branch 1 / 2: pooled T and S graphs, second-layer augmentations, first- and second-layer masters
  GraphView T1, Ta1, S1, T2, Ta2, S2, Sa2;
  const float* m1a = nullptr; const float* m1b = nullptr;   // (B,32)
  const float* m2a = nullptr; const float* m2b = nullptr;
  const float* w = nullptr;   // out_layer.weight [2][160]
  const float* b = nullptr;   // [2]
  float* logits = nullptr;    // (B,2)
  float* hidden = nullptr;    // optional (B,160)
};
int aasist_readout(cudaStream_t s, const ReadoutArgs& a, int B);

}  // namespace rtdf
