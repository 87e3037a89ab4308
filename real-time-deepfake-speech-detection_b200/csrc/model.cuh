// Model context: weight store, packed weights, workspace plan, forward orchestration.
#pragma once
#include "layer_stack.cuh"
#include <map>
#include <string>
#include <vector>

#include "../../include/rtdf.h"
#include "aasist.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rtdf {

struct Raw {
  float* p = nullptr;
  std::vector<int64_t> shape;
  long long numel = 0;
};

struct Lin {          // y = W x + b, W: [n][k]
  const float* w = nullptr;
  const bf16* wb = nullptr;
  const float* b = nullptr;
  int n = 0, k = 0;
};

struct Norm {         // LayerNorm affine or folded BatchNorm scale/shift
  const float* g = nullptr;
  const float* b = nullptr;
};

struct FeConv {
  Lin lin;            // packed [512][k*512] (tap-major) ; conv-0: lin.w = [10][512]
  Norm ln;
  int k = 0, stride = 0;
};

struct XlsrLayer {
  Lin qkv, out, fc1, fc2;
  Norm ln1, ln2;
  // RTDF_LN_FOLD=1 (bf16): qkv.wb / fc1.wb hold W' = bf16(W diag(gamma)) of the LayerNorm in front of them; c = column sums of W',
  // d = bias + W beta (folded LayerNorm, gemm_tc.cuh TcEpilogue::fold_*)
  const float* qkv_c = nullptr; const float* qkv_d = nullptr;
  const float* fc1_c = nullptr; const float* fc1_d = nullptr;
};

struct TcConvW {      // weights of one shifted-row tcgen05 conv: [n_chunks][co][min(ci,64)] bf16 (hi, lo)
  const bf16* hi = nullptr;
  const bf16* lo = nullptr;
  int ci = 0, co = 0, n_chunks = 0;
};

struct ResBlockW {
  const float* conv1_w = nullptr; const float* conv1_b = nullptr;
  const float* conv1_raw = nullptr; const float* ds_raw = nullptr;   // state-dict layout [co][ci][kh][3] (block 0)
  TcConvW tc1, tc2, tcd;
  Norm bn2;
  const float* conv2_w = nullptr; const float* conv2_b = nullptr;
  const float* ds_w = nullptr; const float* ds_b = nullptr;
  int ci = 0, co = 0;
};

struct HsGalW {
  const float* t1_wt = nullptr; const float* t1_b = nullptr;
  const float* t2_wt = nullptr; const float* t2_b = nullptr;
  GatRowWeights rows, master;
  int d = 0, dout = 0;
};

struct PoolW { const float* w = nullptr; const float* b = nullptr; };

struct AasistW {
  Lin LL;
  float first_bn_s = 1.f, first_bn_t = 0.f;
  ResBlockW blocks[6];
  Norm first_bn1;
  const float* att_w1t = nullptr; const float* att_b1 = nullptr; Norm att_bn;
  const float* att_w2t = nullptr; const float* att_b2 = nullptr;
  TcConvW att1, att2;                      // attention 1x1 convs; eval BatchNorm folded into att2 / att_b2_folded
  const float* att_b2_folded = nullptr;
  const float* pos_S = nullptr; const float* master1 = nullptr; const float* master2 = nullptr;
  GatRowWeights gat_S, gat_T;
  HsGalW st11, st12, st21, st22;
  PoolW pool_S, pool_T, pool_hS1, pool_hT1, pool_hS2, pool_hT2;
  const float* out_w = nullptr; const float* out_b = nullptr;
};

struct ConformerBlockW {
  Norm ff1_ln, ff2_ln, attn_ln, conv_ln, post_ln;
  Lin ff1_a, ff1_b, ff2_a, ff2_b;     // 144->576, 576->144
  Lin qkv;                            // fused [to_q ; to_kv] (432 x 144), no bias
  Lin attn_out;
  const float* rel_pos = nullptr;     // [1025][dim_head]
  Lin pw1;                            // pointwise 144 -> 576 (GLU halves)
  const float* dw_w = nullptr;        // [288][k]
  const float* dw_b = nullptr;
  Norm dw_bn;                         // folded BatchNorm1d(288)
  Lin pw2;                            // 288 -> 144
};

struct ConformerW {
  Lin LL;
  float first_bn_s = 1.f, first_bn_t = 0.f;
  const float* class_token = nullptr;
  std::vector<ConformerBlockW> blocks;
  Lin fc5;
};

}  // namespace rtdf

struct rtdf_ctx {
  int device = 0;
  rtdf_model_desc d{};
  bool finalized = false;
  bool ln_fold = false;            // RTDF_LN_FOLD=1 (bf16): LayerNorms folded into the neighbouring GEMMs (opt-in experiment)
  int regime = RTDF_REGIME_AUTO;   // rtdf_set_regime
  cudaStream_t side_stream = nullptr;   // AASIST back-end: second stream for the independent graph branches
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::map<std::string, rtdf::Raw> raw;
  std::vector<void*> owned;
  std::vector<void*> scratch;      // fp32 sources of packed bf16 weights, released at the end of rtdf_finalize
  // XLS-R
  rtdf::FeConv fe[7];
  const rtdf::bf16* conv0_tc_w = nullptr;   // conv-0 weight as the hi/lo-split [512][32] bf16 operand of the tcgen05 path
  bool fe_group_norm = false;      // fairseq extractor_mode="default" (GroupNorm after conv-0 only), from the state-dict keys
  rtdf::Norm fp_ln;
  rtdf::Lin proj;
  rtdf::Lin pos;      // packed [1024][8192]
  std::vector<rtdf::XlsrLayer> layers;
  rtdf::Norm enc_ln;
  // streaming chunks of <= 64 frames (bf16): the transformer layers as one persistent kernel (layer_stack.cu)
  bool layer_stack = false;                  // RTDF_LAYER_STACK (read at rtdf_create)
  rtdf::StackLayer* stack_layers = nullptr;  // device array of per-layer pointers (built by rtdf_finalize)
  CUtensorMap* stack_wmaps = nullptr;        // device array [layers][4] of weight tensor maps
  int stack_impl = 0;                        // RTDF_STACK_IMPL (read at rtdf_create): 0 = tcgen05, 1 = mma.sync variant
  int stack_big_boxes = 0;                   // ... of the 3-D kind (one TMA box per operand slice)
  unsigned* stack_sync = nullptr;            // its grid-barrier words
  int* stack_fault = nullptr;                // mapped host word the kernel sets when a barrier times out
  rtdf::AasistW aasist;
  rtdf::ConformerW conf;
};
