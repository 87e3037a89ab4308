// tcgen05 "shifted-row" convolution for the AASIST residual encoder and attention map (bf16 mode).
//
// Activations live in zero-padded channels-last planes: row m = (b*Hp + hp)*Wp + wp holds the C channels of
// one pixel, with wp = 0 / Wp-1 and the unused hp rows kept at zero.  A (KH,3) convolution is then the sum
// over taps of a GEMM whose A operand is the same plane shifted by a constant number of rows, so one smem
// "slab" of 128 + span rows (one TMA load) feeds every tap: the UMMA descriptor of tap c simply starts
// (shift_c - shift_min) rows into the slab.  Weights stay resident in smem for the whole (persistent) CTA.
//
// Precision: operands are bf16 pairs (hi, lo = x - hi); nsplit = 3 issues hi*hi + lo*hi + hi*lo into the same
// fp32 TMEM accumulator (relative error ~2^-16, which keeps the graph top-k decisions of the back-end equal to
// the fp32 reference's), nsplit = 1 is plain bf16.  Replaces the cuDNN convs behind
// models/aasist_modules.py:340-397 (Residual_block) and models/xlsr_aasist.py:103 (attention 1x1 convs).
#pragma once
#include "common.cuh"

namespace rtdf {

constexpr int kConvTcMaxChunks = 8;

struct ConvTcArgs {
  // A operand: planes [rows][ci] (hi / lo bf16); ci in {32, 64, 128}
  const bf16* in_hi = nullptr;
  const bf16* in_lo = nullptr;      // may be null when nsplit == 1
  int ci = 0;
  long long rows = 0;               // B * Hp * Wp
  int Hp = 0, Wp = 0;
  // weights packed [n_chunks][co][kw] (kw = min(ci, 64) channels of one tap / channel block per chunk)
  const bf16* w_hi = nullptr;
  const bf16* w_lo = nullptr;
  int co = 0;                       // 32 | 64 | 128
  int n_chunks = 0;
  int shift[kConvTcMaxChunks] = {0};  // row shift of the A operand for chunk c
  int sub[kConvTcMaxChunks] = {0};    // 64-channel block of the A plane read by chunk c
  int hp_lo = 0, hp_hi = 0;         // output rows with hp outside [hp_lo, hp_hi] or wp in {0, Wp-1} are written as 0
  // epilogue: v = acc + bias; v = v*s1 + t1; act1; v += resid; v = v*s2 + t2; act2
  const float* bias = nullptr;
  const float* s1 = nullptr;
  const float* t1 = nullptr;
  int act1 = ACT_NONE;
  const float* resid = nullptr;     // fp32 plane [rows][co]
  const float* s2 = nullptr;
  const float* t2 = nullptr;
  int act2 = ACT_NONE;
  float* out_f32 = nullptr;         // optional fp32 plane [rows][co]
  bf16* out_hi = nullptr;           // optional bf16 planes
  bf16* out_lo = nullptr;
};

int conv_tc(cudaStream_t s, const ConvTcArgs& a, int nsplit);

// dst_hi/lo[(c*co + o)*kw + k] = split(src[o*stride_o + k*stride_k + chunk_off[c]])
int conv_tc_pack_weight(cudaStream_t s, const float* src, int co, int kw, int n_chunks, long long stride_o,
                        long long stride_k, const long long* chunk_off, bf16* dst_hi, bf16* dst_lo);

// Block-0 convs on the single-channel stem output z (B,42,W) fp32: conv1 (1->32, (2,3), pad (1,1)) + BN2 + SELU
// -> 43-row planes (hi, lo);  conv_downsample (1->32, (1,3), pad (0,1)) -> 42-row fp32 plane.
int conv_tc_block0(cudaStream_t s, const float* z, int B, int W, int Hp, int Wp, const float* w1 /*[32][6]*/,
                   const float* b1, const float* bn_s, const float* bn_t, const float* wd /*[32][3]*/,
                   const float* bd, bf16* y_hi, bf16* y_lo, float* idt);

// e_S / e_T attention pooling on the padded channels-last planes (xlsr_aasist.py:106-118)
int attn_pool_planes(cudaStream_t s, const float* x, const float* wmap, int B, int H, int W, int Hp, int Wp,
                     const float* pos_S, float* e_S, float* e_T);

}  // namespace rtdf
