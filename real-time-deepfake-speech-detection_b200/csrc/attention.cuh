// Multi-head self-attention cores for the XLS-R transformer (16 heads x 64; tcgen05 kernel: T <= 512 frames).
#pragma once
#include "common.cuh"

namespace rtdf {

// qkv: (B*T, 3*H*64) rows = [q | k | v], q already scaled by 1/8 (folded into W_q at pack time).
// ctx: (B*T, H*64).  Softmax statistics in fp32.  Replaces fairseq MultiheadAttention's core
// (bmm -> softmax -> bmm), reference models/fe.py:19.

// tcgen05 path (default): persistent warp-specialised kernel -- TMA producer warp, MMA issuer warp, two softmax
// warpgroups ping-ponging over two TMEM buffers; S = Q K^T (SS), P written back to TMEM, O = P V (TS).  T <= 256 frames
// use two 256-column buffers; 256 < T <= 512 one 512-column buffer (single softmax warpgroup, K / V in two TMA boxes).
int attention_ws(cudaStream_t s, const bf16* qkv, bf16* ctx, int B, int T, int H, bool reverse = false);

// earlier tile kernel, kept for A/B runs: one CTA per (query tile of 128, head, utterance): S = Q K^T in TMEM, fp32 softmax in
// registers, P (bf16) staged in swizzled smem, O = P V in TMEM.
int attention_tc(cudaStream_t s, const bf16* qkv, bf16* ctx, int B, int T, int H);

// SIMT fp32-accumulate path (verification mode; also usable with bf16 I/O).
int attention_simt_f32(cudaStream_t s, const float* qkv, float* ctx, int B, int T, int H);
int attention_simt_bf16(cudaStream_t s, const bf16* qkv, bf16* ctx, int B, int T, int H);

}  // namespace rtdf
