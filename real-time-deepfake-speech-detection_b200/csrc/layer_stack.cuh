// The XLS-R transformer layers of a streaming chunk (<= 64 frames in flight: batch 1, 1 s of audio) as ONE persistent kernel.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace rtdf {

// One pre-LN TransformerSentenceEncoderLayer (reference models/fe.py:17-21 -> fairseq wav2vec2 encoder layer):
// bf16 weights [n][k] row-major, q rows pre-scaled by 1/8; fp32 biases and LayerNorm affines.
struct StackLayer {
  const bf16 *wqkv, *wo, *w1, *w2;          // [3072][1024], [1024][1024], [4096][1024], [1024][4096]
  const float *bqkv, *bo, *b1, *b2;
  const float *g1, *be1, *g2, *be2;         // self_attn_layer_norm, final_layer_norm
};

struct StackParams {
  const StackLayer* layers;   // device array [n_layers]
  int n_layers;
  int R, B, T;                // R = B * T <= kStackMaxRows
  float* x;                   // [R][1024] fp32 residual stream (in: after the positional conv; clobbered)
  bf16* xn;                   // [R][1024] normalised rows
  bf16* qkv;                  // [R][3072]
  bf16* att;                  // [R][1024]
  bf16* h;                    // [R][4096]
  float* part;                // [4][R][1024] K-split partial sums
  float* feats;               // [R][1024] encoder.layer_norm(x) (out)
  const float *gF, *bF;       // encoder.layer_norm
  unsigned* sync;             // [2] device words, zero between launches (the kernel resets them on its way out)
  int* fault;                 // mapped host word: set if a grid barrier timed out
  const CUtensorMap* wmaps = nullptr;    // device array [n_layers][4]: TMA maps of wqkv, wo, w1, w2 (layer_stack_build_wmaps)
  int impl = 0;                          // 0 = tcgen05 GEMM phases (default), 1 = mma.sync variant (RTDF_STACK_IMPL=mma, A/B timing)
  int big_boxes = 0;                     // wmaps are the 3-D kind (as reported by layer_stack_build_wmaps)
  unsigned long long* trace = nullptr;   // debug: phase time stamps (RTDF_STACK_TRACE=1)
};

constexpr int kStackMaxRows = 64;

// x -> feats through n_layers layers.  128 co-resident CTAs, grid-wide barriers between the phases (7 per layer), weights
// streamed from HBM exactly once (TMA boxes issued one barrier ahead -> tcgen05 MMAs; RTDF_STACK_IMPL=mma: straight into
// mma.sync B fragments).  Not a PDL launch; safe inside stream capture.
int layer_stack_bf16(cudaStream_t s, const StackParams& p);

// host_out[4 * l + {0,1,2,3}] = TMA maps of layer l's wqkv, wo, w1, w2 (a CTA's slice = 32 / 16 / 32 / 32 rows x 1024 columns: one
// 3-D box, or sixteen 2-D boxes if the 3-D encode is refused -> *big_boxes); the caller copies the array to device memory
// (64-byte aligned) and passes it as StackParams::wmaps with StackParams::big_boxes.
int layer_stack_build_wmaps(const StackLayer* host_layers, int n_layers, CUtensorMap* host_out, int* big_boxes);

}  // namespace rtdf
