// Bandwidth-bound front-end kernels: coalesced, vectorised, warp-shuffle reductions.
#include "frontend.cuh"
#include "gemm_tc.cuh"

#include <stdlib.h>

namespace rtdf {

// ------------------------------------------------------------------------------------------------
// pre-emphasis (reference data/preprocess.py:22-27)
// ------------------------------------------------------------------------------------------------
__global__ void preemph_kernel(const float* __restrict__ x, float* __restrict__ y, int N, float coef) {
  const int b = blockIdx.y;
  const float* xb = x + (long long)b * N;
  float* yb = y + (long long)b * N;
  const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t0 >= N) return;
  if ((N & 3) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(xb + t0);
    const float prev = t0 == 0 ? v.y : xb[t0 - 1];  // reflect: x[-1] := x[1]
    float4 o;
    o.x = v.x - coef * prev;
    o.y = v.y - coef * v.x;
    o.z = v.z - coef * v.y;
    o.w = v.w - coef * v.z;
    *reinterpret_cast<float4*>(yb + t0) = o;
  } else {
    for (int t = t0; t < min(t0 + 4, N); ++t) {
      const float prev = t == 0 ? xb[N > 1 ? 1 : 0] : xb[t - 1];
      yb[t] = xb[t] - coef * prev;
    }
  }
}

int preemph(cudaStream_t s, const float* x, float* y, int B, int N, float coef) {
  RTDF_REQUIRE(x && y && B > 0 && N > 1, "preemph: bad arguments");
  RTDF_REQUIRE(x != y, "preemph: in-place operation is not supported");
  dim3 grid(ceil_div(ceil_div(N, 4), 256), B);
  preemph_kernel<<<grid, 256, 0, s>>>(x, y, N, coef);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// per-utterance waveform layer norm (zero mean / unit variance over time); block per utterance
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) wave_ln_kernel(const float* __restrict__ x, float* __restrict__ y, int N, float eps) {
  __shared__ float red[32];
  __shared__ float stat[2];
  const float* xb = x + (long long)blockIdx.x * N;
  float* yb = y + (long long)blockIdx.x * N;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += xb[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) stat[0] = t / N;
  }
  __syncthreads();
  const float mean = stat[0];
  float q = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = xb[i] - mean;
    q += d * d;
  }
  q = warp_sum(q);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) stat[1] = rsqrtf(t / N + eps);
  }
  __syncthreads();
  const float rstd = stat[1];
  for (int i = threadIdx.x; i < N; i += blockDim.x) yb[i] = (xb[i] - mean) * rstd;
}

int wave_layernorm(cudaStream_t s, const float* x, float* y, int B, int N, float eps) {
  RTDF_REQUIRE(x && y && B > 0 && N > 0, "wave_layernorm: bad arguments");
  wave_ln_kernel<<<B, 1024, 0, s>>>(x, y, N, eps);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// conv-0 (C_in = 1, k = 10, stride 5) + bias + LayerNorm(512) + GELU, channels-last output.
// One warp produces 4 consecutive frames; lane owns channels {j*128 + lane*4 + e}.
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__device__ __forceinline__ void store4(TOut* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

constexpr int kConv0FramesPerCta = 256;   // 8 warps x 8 groups of 4 frames: the 26 KB parameter block is loaded once

template <typename TOut, bool kTanhGelu = false>
__global__ void __launch_bounds__(256, 2)
conv0_kernel(const float* __restrict__ wav, int N, int L1, const float* __restrict__ w_t,
             const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
             float eps, TOut* __restrict__ out, int frames_per_cta) {
  __shared__ __align__(16) float sw[13][512];  // 10 taps | bias | gamma | beta
  for (int i = threadIdx.x; i < 10 * 512; i += 256) sw[0][i] = w_t[i];
  for (int i = threadIdx.x; i < 512; i += 256) {
    sw[10][i] = bias ? bias[i] : 0.f;
    sw[11][i] = gamma[i];
    sw[12][i] = beta[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const float* xb = wav + (long long)b * N;
  // lane owns channels {j*128 + lane*4 + e}: 16 channels as 8 packed pairs
  for (int g = warp; g < frames_per_cta / 4; g += 8) {
    const int t0 = blockIdx.x * frames_per_cta + g * 4;
    if (t0 >= L1) break;
    // 4 frames need samples [5*t0, 5*t0 + 25)
    float xv = 0.f;
    {
      const int i = 5 * t0 + lane;
      if (lane < 25 && i < N) xv = xb[i];
    }
    float2 acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bv = *reinterpret_cast<const float4*>(&sw[10][j * 128 + lane * 4]);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        acc[f][2 * j] = make_float2(bv.x, bv.y);
        acc[f][2 * j + 1] = make_float2(bv.z, bv.w);
      }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      float4 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(&sw[k][j * 128 + lane * 4]);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float x = __shfl_sync(0xffffffffu, xv, 5 * f + k);
        const float2 x2 = make_float2(x, x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[f][2 * j] = fma2(x2, make_float2(w[j].x, w[j].y), acc[f][2 * j]);
          acc[f][2 * j + 1] = fma2(x2, make_float2(w[j].z, w[j].w), acc[f][2 * j + 1]);
        }
      }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      float2 s2 = acc[f][0];
#pragma unroll
      for (int c = 1; c < 8; ++c) s2 = add2(s2, acc[f][c]);
      const float mean = warp_sum(s2.x + s2.y) * (1.0f / 512.0f);
      const float2 nm2 = make_float2(-mean, -mean);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        acc[f][c] = add2(acc[f][c], nm2);
        q2 = fma2(acc[f][c], acc[f][c], q2);
      }
      const float rstd = rsqrtf(warp_sum(q2.x + q2.y) * (1.0f / 512.0f) + eps);
      const float2 rstd2 = make_float2(rstd, rstd);
      const int t = t0 + f;
      if (t < L1) {
        TOut* o = out + ((long long)b * L1 + t) * 512;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 gv = *reinterpret_cast<const float4*>(&sw[11][j * 128 + lane * 4]);
          const float4 bv = *reinterpret_cast<const float4*>(&sw[12][j * 128 + lane * 4]);
          float2 y0 = fma2(acc[f][2 * j], mul2(make_float2(gv.x, gv.y), rstd2), make_float2(bv.x, bv.y));
          float2 y1 = fma2(acc[f][2 * j + 1], mul2(make_float2(gv.z, gv.w), rstd2), make_float2(bv.z, bv.w));
          if (sizeof(TOut) == 2) {   // bf16 output: the fitted GELU (|err| <= 2.6e-5) is far below the rounding step
            y0 = kTanhGelu ? gelu2_tanh(y0) : gelu2(y0);
            y1 = kTanhGelu ? gelu2_tanh(y1) : gelu2(y1);
          } else {
            y0 = make_float2(gelu_erf(y0.x), gelu_erf(y0.y));
            y1 = make_float2(gelu_erf(y1.x), gelu_erf(y1.y));
          }
          store4<TOut>(o + j * 128 + lane * 4, y0.x, y0.y, y1.x, y1.y);
        }
      }
    }
  }
}

int conv0_ln_gelu(cudaStream_t s, const float* wav, int B, int N, const float* w_t, const float* bias,
                  const float* gamma, const float* beta, float eps, float* out_f32, bf16* out_bf16) {
  RTDF_REQUIRE(wav && w_t && gamma && beta && N >= 10 && B > 0 && B <= 65535, "conv0: bad arguments");
  RTDF_REQUIRE((out_f32 != nullptr) != (out_bf16 != nullptr), "conv0: exactly one output must be given");
  const int L1 = (N - 10) / 5 + 1;
  // streaming chunks: 32 frames per CTA put the few thousand frames on ~100 SMs instead of ~13
  const int fpc = (long long)B * L1 >= 2LL * kNumSMs * kConv0FramesPerCta ? kConv0FramesPerCta : 32;
  dim3 grid(ceil_div(L1, fpc), B);
  if (out_f32)
    conv0_kernel<float><<<grid, 256, 0, s>>>(wav, N, L1, w_t, bias, gamma, beta, eps, out_f32, fpc);
  else if (tc_get_gelu_variant() == ACT_GELU_TANH)
    conv0_kernel<bf16, true><<<grid, 256, 0, s>>>(wav, N, L1, w_t, bias, gamma, beta, eps, out_bf16, fpc);
  else
    conv0_kernel<bf16><<<grid, 256, 0, s>>>(wav, N, L1, w_t, bias, gamma, beta, eps, out_bf16, fpc);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// conv-0 on the tensor cores (large batches): the layer as the implicit GEMM  out[m][n] = sum_k A[m][k] W[n][k],
// m = (utterance, frame), k = tap, n = channel, with K = 10 padded to 32 and both operands split into bf16 (hi, lo)
// pairs laid side by side along K so that ONE bf16 GEMM carries three of the four partial products (~2^-16 relative,
// the raw waveform is never rounded to 8 mantissa bits):
//     A row = [ hi(x_0..x_9) | lo(x_0..x_9) | hi(x_0..x_9) | 0 0 ]      W row = [ hi(w) | hi(w) | lo(w) | 0 0 ]
// The strided slabs x[5t .. 5t+10) are 10 bytes apart in bf16 -- below TMA's 16-byte stride granularity -- so the
// A matrix (64 B per frame) is materialised by conv0_im2col_kernel; bias + LayerNorm(512) + GELU run in the epilogue of
// the full-row tcgen05 tile (tc_gemm variant 514: K-block 32, 64-byte swizzle).
// ------------------------------------------------------------------------------------------------
__global__ void conv0_pack_w_kernel(const float* __restrict__ w /*[512][10]*/, bf16* __restrict__ out /*[512][32]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 512 * 32) return;
  const int n = i >> 5, e = i & 31;
  float v = 0.f;
  if (e < 30) {
    const float x = w[n * 10 + e % 10];
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    v = e < 20 ? hi : x - hi;
  }
  out[i] = __float2bfloat16_rn(v);
}

int conv0_tc_pack_weight(cudaStream_t s, const float* w, bf16* out) {
  RTDF_REQUIRE(w && out, "conv0_tc_pack_weight: null argument");
  conv0_pack_w_kernel<<<64, 256, 0, s>>>(w, out);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

__global__ void __launch_bounds__(256)
conv0_im2col_kernel(const float* __restrict__ wav, int N, int L1, long long rows, bf16* __restrict__ A) {
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= rows) return;
  const long long b = m / L1;
  const int t = (int)(m - b * L1);
  const float* x = wav + b * N + 5 * t;
  uint32_t hi[5], lo[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float x0 = x[2 * j], x1 = x[2 * j + 1];
    const float h0 = __bfloat162float(__float2bfloat16_rn(x0)), h1 = __bfloat162float(__float2bfloat16_rn(x1));
    hi[j] = pack_bf16x2(h0, h1);
    lo[j] = pack_bf16x2(x0 - h0, x1 - h1);
  }
  uint4* dst = reinterpret_cast<uint4*>(A + m * 32);
  dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  dst[1] = make_uint4(hi[4], lo[0], lo[1], lo[2]);
  dst[2] = make_uint4(lo[3], lo[4], hi[0], hi[1]);
  dst[3] = make_uint4(hi[2], hi[3], hi[4], 0u);
}

int conv0_tc_ln_gelu(cudaStream_t s, const float* wav, int B, int N, const bf16* w_packed, const float* bias,
                     const float* gamma, const float* beta, float eps, bf16* a_scratch, bf16* out_bf16) {
  RTDF_REQUIRE(wav && w_packed && gamma && beta && a_scratch && out_bf16 && N >= 10 && B > 0, "conv0_tc: bad arguments");
  const int L1 = (N - 10) / 5 + 1;
  const long long rows = (long long)B * L1;
  RTDF_REQUIRE(rows < (1LL << 31), "conv0_tc: too many frames");
  conv0_im2col_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(wav, N, L1, rows, a_scratch);
  RTDF_LAUNCH_CHECK();
  TcOperandA a;
  a.ptr = a_scratch;
  a.k_extent = 32;
  a.rows_per_batch = rows;
  a.batches = 1;
  a.row_stride = 32;
  TcEpilogue e;
  e.bias = bias;
  e.act = ACT_GELU;
  e.ln_gamma = gamma;
  e.ln_beta = beta;
  e.ln_eps = eps;
  e.out_bf16 = out_bf16;
  e.ld_bf16 = 512;
  return tc_gemm(s, a, w_packed, 512, 32, TC_PLAIN, 514, e);
}

// ------------------------------------------------------------------------------------------------
// conv-0 in fairseq's extractor_mode="default" (wav2vec2-base style; SURVEY.md App. A.2 step 1, alternative):
//   conv (bias optional) -> GroupNorm(512 groups, 512 channels) = per-(utterance, channel) normalisation over TIME
//   -> GELU.  The statistics span the whole utterance, so the layer is two sweeps over the (cheap, K = 10) conv:
//     1. conv0_gn_stats_kernel     per-CTA partial (sum, sum of squares) of every channel over 256 frames
//     2. conv0_gn_finalize_kernel  partials added in chunk order in fp64 -> scale = gamma*rstd, shift = beta - mean*scale
//     3. conv0_gn_apply_kernel     conv recomputed, y = gelu(v*scale + shift), channels-last output
//   The pre-norm activation (B, L1, 512) fp32 = 1.7 GB at batch 64 never exists in memory.  thread = 2 channels.
// ------------------------------------------------------------------------------------------------
constexpr int kGnFrames = 256;                 // frames per CTA
constexpr int kGnSamples = 5 * kGnFrames + 5;  // samples they read

__device__ __forceinline__ void gn_stage(const float* __restrict__ xb, int N, int t0, float* sx) {
  for (int i = threadIdx.x; i < kGnSamples; i += 256) {
    const int j = 5 * t0 + i;
    sx[i] = j < N ? xb[j] : 0.f;
  }
}

__global__ void __launch_bounds__(256)
conv0_gn_stats_kernel(const float* __restrict__ wav, int N, int L1, const float* __restrict__ w_t,
                      const float* __restrict__ bias, float* __restrict__ partial /*[B][chunks][2][512]*/) {
  __shared__ float sx[kGnSamples];
  const int b = blockIdx.y, t0 = blockIdx.x * kGnFrames, c = threadIdx.x * 2;
  gn_stage(wav + (long long)b * N, N, t0, sx);
  float w0[10], w1[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) { w0[k] = w_t[k * 512 + c]; w1[k] = w_t[k * 512 + c + 1]; }
  const float b0 = bias ? bias[c] : 0.f, b1 = bias ? bias[c + 1] : 0.f;
  __syncthreads();
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const int nf = min(kGnFrames, L1 - t0);
  for (int f = 0; f < nf; ++f) {
    float v0 = b0, v1 = b1;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float x = sx[5 * f + k];
      v0 = fmaf(x, w0[k], v0);
      v1 = fmaf(x, w1[k], v1);
    }
    s0 += v0; s1 += v1;
    q0 = fmaf(v0, v0, q0); q1 = fmaf(v1, v1, q1);
  }
  float* p = partial + ((long long)b * gridDim.x + blockIdx.x) * 1024;
  p[c] = s0; p[c + 1] = s1;
  p[512 + c] = q0; p[512 + c + 1] = q1;
}

__global__ void __launch_bounds__(512)
conv0_gn_finalize_kernel(const float* __restrict__ partial, int chunks, int L1, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, float* __restrict__ scale_shift /*[B][2][512]*/) {
  const int b = blockIdx.x, c = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < chunks; ++i) {
    const float* p = partial + ((long long)b * chunks + i) * 1024;
    s += (double)p[c];
    q += (double)p[512 + c];
  }
  const double mean = s / L1;
  const double var = fmax(q / L1 - mean * mean, 0.0);
  const float sc = gamma[c] * (float)(1.0 / sqrt(var + (double)eps));
  scale_shift[(long long)b * 1024 + c] = sc;
  scale_shift[(long long)b * 1024 + 512 + c] = beta[c] - (float)mean * sc;
}

template <typename TOut>
__global__ void __launch_bounds__(256)
conv0_gn_apply_kernel(const float* __restrict__ wav, int N, int L1, const float* __restrict__ w_t,
                      const float* __restrict__ bias, const float* __restrict__ scale_shift, TOut* __restrict__ out) {
  __shared__ float sx[kGnSamples];
  const int b = blockIdx.y, t0 = blockIdx.x * kGnFrames, c = threadIdx.x * 2;
  gn_stage(wav + (long long)b * N, N, t0, sx);
  float w0[10], w1[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) { w0[k] = w_t[k * 512 + c]; w1[k] = w_t[k * 512 + c + 1]; }
  const float b0 = bias ? bias[c] : 0.f, b1 = bias ? bias[c + 1] : 0.f;
  const float* ss = scale_shift + (long long)b * 1024;
  const float sc0 = ss[c], sc1 = ss[c + 1], sh0 = ss[512 + c], sh1 = ss[512 + c + 1];
  __syncthreads();
  const int nf = min(kGnFrames, L1 - t0);
  TOut* o = out + ((long long)b * L1 + t0) * 512 + c;
  for (int f = 0; f < nf; ++f) {
    float v0 = b0, v1 = b1;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float x = sx[5 * f + k];
      v0 = fmaf(x, w0[k], v0);
      v1 = fmaf(x, w1[k], v1);
    }
    v0 = fmaf(v0, sc0, sh0);
    v1 = fmaf(v1, sc1, sh1);
    if (sizeof(TOut) == 2) {
      const float2 g = gelu2(make_float2(v0, v1));     // fitted GELU: |err| <= 2.6e-5, below the bf16 rounding step
      *reinterpret_cast<uint32_t*>(o + (long long)f * 512) = pack_bf16x2(g.x, g.y);
    } else {
      *reinterpret_cast<float2*>(o + (long long)f * 512) = make_float2(gelu_erf(v0), gelu_erf(v1));
    }
  }
}

size_t conv0_gn_workspace_floats(int B, int N) {
  const int L1 = N >= 10 ? (N - 10) / 5 + 1 : 0;
  return (size_t)B * ceil_div(L1, kGnFrames) * 1024 + (size_t)B * 1024;
}

int conv0_gn_gelu(cudaStream_t s, const float* wav, int B, int N, const float* w_t, const float* bias, const float* gamma,
                  const float* beta, float eps, float* ws, float* out_f32, bf16* out_bf16) {
  RTDF_REQUIRE(wav && w_t && gamma && beta && ws && N >= 10 && B > 0 && B <= 65535, "conv0_gn: bad arguments");
  RTDF_REQUIRE((out_f32 != nullptr) != (out_bf16 != nullptr), "conv0_gn: exactly one output must be given");
  const int L1 = (N - 10) / 5 + 1;
  const int chunks = ceil_div(L1, kGnFrames);
  float* partial = ws;
  float* scale_shift = ws + (size_t)B * chunks * 1024;
  dim3 grid(chunks, B);
  conv0_gn_stats_kernel<<<grid, 256, 0, s>>>(wav, N, L1, w_t, bias, partial);
  RTDF_LAUNCH_CHECK();
  conv0_gn_finalize_kernel<<<B, 512, 0, s>>>(partial, chunks, L1, gamma, beta, eps, scale_shift);
  RTDF_LAUNCH_CHECK();
  if (out_f32) conv0_gn_apply_kernel<float><<<grid, 256, 0, s>>>(wav, N, L1, w_t, bias, scale_shift, out_f32);
  else conv0_gn_apply_kernel<bf16><<<grid, 256, 0, s>>>(wav, N, L1, w_t, bias, scale_shift, out_bf16);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// Row LayerNorm.  Fast path: C % 128 == 0 and C <= 1024, one warp per row, row cached in registers
// (single global read), two-pass statistics.  Generic path: any C, three strided passes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

template <typename TIn>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const TIn* __restrict__ in, long long rows, int C, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, int act, float* __restrict__ out_f32,
               bf16* __restrict__ out_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int groups = C >> 7;
  const TIn* p = in + row * C;
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < groups) {
      v[j] = load4(p + j * 128 + lane * 4);
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < groups) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < groups) {
      const int c0 = j * 128 + lane * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + c0);
      const float4 bt = *reinterpret_cast<const float4*>(beta + c0);
      const float a = apply_act((v[j].x - mean) * rstd * g.x + bt.x, act);
      const float b = apply_act((v[j].y - mean) * rstd * g.y + bt.y, act);
      const float c = apply_act((v[j].z - mean) * rstd * g.z + bt.z, act);
      const float d = apply_act((v[j].w - mean) * rstd * g.w + bt.w, act);
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c0) = make_float4(a, b, c, d);
      if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + row * C + c0) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
    }
}

template <typename TIn>
__global__ void __launch_bounds__(256)
ln_rows_generic_kernel(const TIn* __restrict__ in, long long rows, int C, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, int act, float* __restrict__ out_f32,
                       bf16* __restrict__ out_bf16) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TIn* p = in + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f32(p[c]);
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = to_f32(p[c]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  for (int c = lane; c < C; c += 32) {
    const float y = apply_act((to_f32(p[c]) - mean) * rstd * gamma[c] + beta[c], act);
    if (out_f32) out_f32[row * C + c] = y;
    if (out_bf16) out_bf16[row * C + c] = __float2bfloat16_rn(y);
  }
}

// Specialised LayerNorm(1024), fp32 in, no activation (the 49 LayerNorms of the transformer stack): everything is
// compile-time, R rows per warp are in flight together, 4 warps per CTA.
template <int R, bool kBf16Out>
__global__ void __launch_bounds__(128)
ln1024_kernel(const float* __restrict__ in, long long rows, const float* __restrict__ gamma, const float* __restrict__ beta,
              float eps, float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int reverse) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const bool plain_loads = (reverse & 2) != 0;     // bit 1: read x with default caching instead of evict-first (ld.global.cs)
  reverse &= 1;
  // reverse: the first CTAs take the LAST rows -- the ones the residual GEMM in front wrote last and L2 still holds
  const long long blk = reverse ? (long long)gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const long long row0 = (blk * 4 + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  float4 v[R][8];
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[u][j] = row0 + u >= rows ? make_float4(0.f, 0.f, 0.f, 0.f)
                : plain_loads ? *reinterpret_cast<const float4*>(in + (row0 + u) * 1024 + j * 128 + lane * 4)
                              : __ldcs(reinterpret_cast<const float4*>(in + (row0 + u) * 1024 + j * 128 + lane * 4));
  float mean[R], rstd[R];
#pragma unroll
  for (int u = 0; u < R; ++u) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += (v[u][j].x + v[u][j].y) + (v[u][j].z + v[u][j].w);
    mean[u] = warp_sum(s) * (1.0f / 1024.0f);
  }
#pragma unroll
  for (int u = 0; u < R; ++u) {
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = v[u][j].x - mean[u], b = v[u][j].y - mean[u], c = v[u][j].z - mean[u], d = v[u][j].w - mean[u];
      q += a * a + b * b + c * c + d * d;
    }
    rstd[u] = rsqrtf(warp_sum(q) * (1.0f / 1024.0f) + eps);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c0 = j * 128 + lane * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c0);
    const float4 bt = *reinterpret_cast<const float4*>(beta + c0);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (row0 + u >= rows) continue;
      const float a = (v[u][j].x - mean[u]) * rstd[u] * g.x + bt.x;
      const float b = (v[u][j].y - mean[u]) * rstd[u] * g.y + bt.y;
      const float c = (v[u][j].z - mean[u]) * rstd[u] * g.z + bt.z;
      const float d = (v[u][j].w - mean[u]) * rstd[u] * g.w + bt.w;
      if (kBf16Out) *reinterpret_cast<uint2*>(out_bf16 + (row0 + u) * 1024 + c0) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
      else *reinterpret_cast<float4*>(out_f32 + (row0 + u) * 1024 + c0) = make_float4(a, b, c, d);
    }
  }
}

// Same normalisation behind a split-K GEMM: the row of the residual stream first takes the K-split partial sums
// (x += sum_s partials[s], added in split order, so the result does not depend on which CTA finished first), is
// written back, and is normalised from registers.
template <bool kBf16Out>
__global__ void __launch_bounds__(128)
ln1024_accum_kernel(float* __restrict__ x, const float* __restrict__ partials, int n_splits, long long split_stride,
                    long long rows, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                    float* __restrict__ out_f32, bf16* __restrict__ out_bf16) {
  // one row per CTA (rows are few in this regime): thread t owns columns [8t, 8t+8)
  __shared__ float red[2][4];
  pdl_launch_dependents();
  pdl_wait();
  const long long row = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  float* xr = x + row * 1024 + t * 8;
  float4 v0 = *reinterpret_cast<const float4*>(xr), v1 = *reinterpret_cast<const float4*>(xr + 4);
  const float* pr = partials + row * 1024 + t * 8;
#pragma unroll 4
  for (int sp = 0; sp < n_splits; ++sp) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(pr + sp * split_stride));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(pr + sp * split_stride + 4));
    v0.x += a.x; v0.y += a.y; v0.z += a.z; v0.w += a.w;
    v1.x += b.x; v1.y += b.y; v1.z += b.z; v1.w += b.w;
  }
  *reinterpret_cast<float4*>(xr) = v0;
  *reinterpret_cast<float4*>(xr + 4) = v1;
  float s = warp_sum((v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w));
  if (lane == 0) red[0][wid] = s;
  __syncthreads();
  const float mean = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) * (1.0f / 1024.0f);
  const float d0 = v0.x - mean, d1 = v0.y - mean, d2 = v0.z - mean, d3 = v0.w - mean;
  const float d4 = v1.x - mean, d5 = v1.y - mean, d6 = v1.z - mean, d7 = v1.w - mean;
  float q = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 + d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7);
  if (lane == 0) red[1][wid] = q;
  __syncthreads();
  const float rstd = rsqrtf(((red[1][0] + red[1][1]) + (red[1][2] + red[1][3])) * (1.0f / 1024.0f) + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + t * 8), g1 = *reinterpret_cast<const float4*>(gamma + t * 8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + t * 8), b1 = *reinterpret_cast<const float4*>(beta + t * 8 + 4);
  const float y0 = d0 * rstd * g0.x + b0.x, y1 = d1 * rstd * g0.y + b0.y, y2 = d2 * rstd * g0.z + b0.z, y3 = d3 * rstd * g0.w + b0.w;
  const float y4 = d4 * rstd * g1.x + b1.x, y5 = d5 * rstd * g1.y + b1.y, y6 = d6 * rstd * g1.z + b1.z, y7 = d7 * rstd * g1.w + b1.w;
  if (kBf16Out) {
    *reinterpret_cast<uint4*>(out_bf16 + row * 1024 + t * 8) =
        make_uint4(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3), pack_bf16x2(y4, y5), pack_bf16x2(y6, y7));
  } else {
    *reinterpret_cast<float4*>(out_f32 + row * 1024 + t * 8) = make_float4(y0, y1, y2, y3);
    *reinterpret_cast<float4*>(out_f32 + row * 1024 + t * 8 + 4) = make_float4(y4, y5, y6, y7);
  }
}

// Folded-LayerNorm companion of the two kernels above (gemm_tc.cuh, TcEpilogue::fold_*): the GEMM that follows applies
// the normalisation in its epilogue, so this kernel only (optionally) folds the K-split partial sums into the residual
// row, and emits the row as bf16 plus its (sum, sum of squares) in slot 0 of the 8 statistics slots (slots 1..7 zero).
template <bool kAccum>
__global__ void __launch_bounds__(128)
cast_stats1024_kernel(float* __restrict__ x, const float* __restrict__ partials, int n_splits, long long split_stride,
                      bf16* __restrict__ xb, float2* __restrict__ stats) {
  __shared__ float red[2][4];
  pdl_launch_dependents();
  pdl_wait();
  const long long row = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  float* xr = x + row * 1024 + t * 8;
  float4 v0 = *reinterpret_cast<const float4*>(xr), v1 = *reinterpret_cast<const float4*>(xr + 4);
  if (kAccum) {
    const float* pr = partials + row * 1024 + t * 8;
#pragma unroll 4
    for (int sp = 0; sp < n_splits; ++sp) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(pr + sp * split_stride));
      const float4 b = __ldcs(reinterpret_cast<const float4*>(pr + sp * split_stride + 4));
      v0.x += a.x; v0.y += a.y; v0.z += a.z; v0.w += a.w;
      v1.x += b.x; v1.y += b.y; v1.z += b.z; v1.w += b.w;
    }
    *reinterpret_cast<float4*>(xr) = v0;
    *reinterpret_cast<float4*>(xr + 4) = v1;
  }
  const float s = warp_sum((v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w));
  const float q = warp_sum(v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y + v1.z * v1.z +
                           v1.w * v1.w);
  if (lane == 0) { red[0][wid] = s; red[1][wid] = q; }
  *reinterpret_cast<uint4*>(xb + row * 1024 + t * 8) =
      make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
  __syncthreads();
  if (t < 8)
    stats[row * 8 + t] = t == 0 ? make_float2((red[0][0] + red[0][1]) + (red[0][2] + red[0][3]),
                                              (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]))
                                : make_float2(0.f, 0.f);
}

int cast_stats_rows(cudaStream_t s, float* x, const float* partials, int n_splits, long long rows, bf16* xb, float2* stats) {
  RTDF_REQUIRE(x && xb && stats && rows > 0 && (n_splits == 0 || partials), "cast_stats_rows: bad arguments");
  const unsigned g = (unsigned)rows;
  if (n_splits > 0)
    RTDF_CHECK_CUDA(launch_pdl(cast_stats1024_kernel<true>, dim3(g), dim3(128), 0, s, x, partials, n_splits, rows * 1024, xb, stats));
  else
    RTDF_CHECK_CUDA(launch_pdl(cast_stats1024_kernel<false>, dim3(g), dim3(128), 0, s, x, partials, n_splits, rows * 1024, xb, stats));
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// W' = bf16(W diag(gamma)), c_j = sum_k W'_jk (of the ROUNDED values: exactly what the tensor core multiplies with),
// d_j = bias_j + sum_k beta_k W_jk.  One warp per output row.
__global__ void __launch_bounds__(256)
fold_ln_weight_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ bias, int n, int k, bf16* __restrict__ wb, float* __restrict__ c,
                      float* __restrict__ d) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  float cs = 0.f, ds = 0.f;
  for (int kk = lane; kk < k; kk += 32) {
    const float wv = w[(long long)j * k + kk];
    const bf16 r = __float2bfloat16_rn(wv * gamma[kk]);
    wb[(long long)j * k + kk] = r;
    cs += __bfloat162float(r);
    ds = fmaf(beta[kk], wv, ds);
  }
  cs = warp_sum(cs);
  ds = warp_sum(ds);
  if (lane == 0) {
    c[j] = cs;
    d[j] = (bias ? bias[j] : 0.f) + ds;
  }
}

int fold_ln_weight(cudaStream_t s, const float* w, const float* gamma, const float* beta, const float* bias, int n, int k,
                   bf16* wb, float* c, float* d) {
  RTDF_REQUIRE(w && gamma && beta && wb && c && d && n > 0 && k > 0, "fold_ln_weight: bad arguments");
  fold_ln_weight_kernel<<<ceil_div(n, 8), 256, 0, s>>>(w, gamma, beta, bias, n, k, wb, c, d);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int layernorm_accum_rows(cudaStream_t s, float* x, const float* partials, int n_splits, long long rows, const float* gamma,
                         const float* beta, float eps, float* out_f32, bf16* out_bf16) {
  RTDF_REQUIRE(x && partials && n_splits >= 1 && rows > 0 && gamma && beta && ((out_f32 != nullptr) != (out_bf16 != nullptr)),
               "layernorm_accum_rows: bad arguments");
  const unsigned g = (unsigned)rows;
  if (out_bf16)
    RTDF_CHECK_CUDA(launch_pdl(ln1024_accum_kernel<true>, dim3(g), dim3(128), 0, s, x, partials, n_splits, rows * 1024, rows,
                               gamma, beta, eps, out_f32, out_bf16));
  else
    RTDF_CHECK_CUDA(launch_pdl(ln1024_accum_kernel<false>, dim3(g), dim3(128), 0, s, x, partials, n_splits, rows * 1024, rows,
                               gamma, beta, eps, out_f32, out_bf16));
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

static int ln_plain_loads() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_LN_LOAD");
    v = (e && e[0] == '1') ? 2 : 0;
  }
  return v;
}

static int ln_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_LN_VARIANT");
    v = e ? atoi(e) : 2;
  }
  return v;
}

template <typename TIn>
static int ln_launch(cudaStream_t s, const TIn* in, long long rows, int C, const float* gamma, const float* beta,
                     float eps, int act, float* out_f32, bf16* out_bf16, bool reverse = false) {
  RTDF_REQUIRE(in && gamma && beta && rows > 0 && C > 0 && (out_f32 || out_bf16), "layernorm_rows: bad arguments");
  if (sizeof(TIn) == 4 && C == 1024 && act == ACT_NONE && (out_f32 != nullptr) != (out_bf16 != nullptr) && ln_variant() > 0) {
    const float* inf = reinterpret_cast<const float*>(in);
    const int R = ln_variant() >= 2 ? 2 : 1;
    const unsigned g = (unsigned)((rows + 4 * R - 1) / (4 * R));
    if (R == 2) {
      if (out_bf16) RTDF_CHECK_CUDA(launch_pdl(ln1024_kernel<2, true>, dim3(g), dim3(128), 0, s, inf, rows, gamma, beta, eps, out_f32, out_bf16, (reverse ? 1 : 0) | ln_plain_loads()));
      else RTDF_CHECK_CUDA(launch_pdl(ln1024_kernel<2, false>, dim3(g), dim3(128), 0, s, inf, rows, gamma, beta, eps, out_f32, out_bf16, (reverse ? 1 : 0) | ln_plain_loads()));
    } else {
      if (out_bf16) RTDF_CHECK_CUDA(launch_pdl(ln1024_kernel<1, true>, dim3(g), dim3(128), 0, s, inf, rows, gamma, beta, eps, out_f32, out_bf16, (reverse ? 1 : 0) | ln_plain_loads()));
      else RTDF_CHECK_CUDA(launch_pdl(ln1024_kernel<1, false>, dim3(g), dim3(128), 0, s, inf, rows, gamma, beta, eps, out_f32, out_bf16, (reverse ? 1 : 0) | ln_plain_loads()));
    }
    RTDF_LAUNCH_CHECK();
    return RTDF_OK;
  }
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if ((C & 127) == 0 && C <= 1024)
    RTDF_CHECK_CUDA(launch_pdl(ln_rows_kernel<TIn>, dim3(grid), dim3(256), 0, s, in, rows, C, gamma, beta, eps, act, out_f32,
                               out_bf16));
  else
    ln_rows_generic_kernel<TIn><<<grid, 256, 0, s>>>(in, rows, C, gamma, beta, eps, act, out_f32, out_bf16);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int layernorm_rows_f32(cudaStream_t s, const float* in, long long rows, int C, const float* gamma, const float* beta,
                       float eps, int act, float* out_f32, bf16* out_bf16, bool reverse) {
  return ln_launch<float>(s, in, rows, C, gamma, beta, eps, act, out_f32, out_bf16, reverse);
}
int layernorm_rows_bf16(cudaStream_t s, const bf16* in, long long rows, int C, const float* gamma, const float* beta,
                        float eps, int act, float* out_f32, bf16* out_bf16) {
  return ln_launch<bf16>(s, in, rows, C, gamma, beta, eps, act, out_f32, out_bf16);
}

// ------------------------------------------------------------------------------------------------
// fp32 grouped positional conv (verification mode): one warp per output element, lanes split K = 8192.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
posconv_f32_kernel(float* __restrict__ x, const float* __restrict__ xin, int B, int T,
                   const float* __restrict__ w, const float* __restrict__ bias) {
  const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long total = (long long)B * T * 1024;
  if (gw >= total) return;
  const int lane = threadIdx.x & 31;
  const int co = (int)(gw % 1024);
  const int t = (int)((gw / 1024) % T);
  const int b = (int)(gw / (1024LL * T));
  const int g = co >> 6;
  const float* wr = w + (long long)co * 8192;
  const float* xb = xin + (long long)b * T * 1024 + g * 64;
  float acc = 0.f;
  for (int k = 0; k < 128; ++k) {
    const int tt = t + k - 64;
    if (tt < 0 || tt >= T) continue;  // warp-uniform
    const float* xr = xb + (long long)tt * 1024;
    acc = fmaf(xr[lane], wr[k * 64 + lane], acc);
    acc = fmaf(xr[lane + 32], wr[k * 64 + lane + 32], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) x[gw] += gelu_erf(acc + bias[co]);
}

int posconv_f32(cudaStream_t s, float* x, const float* xin, int B, int T, const float* w_packed, const float* bias) {
  RTDF_REQUIRE(x && xin && x != xin && w_packed && bias, "posconv_f32: bad arguments");
  const long long total = (long long)B * T * 1024;
  posconv_f32_kernel<<<(unsigned)((total + 7) / 8), 256, 0, s>>>(x, xin, B, T, w_packed, bias);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// packing helpers (run once at finalize)
// ------------------------------------------------------------------------------------------------
__global__ void cast_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
int cast_f32_to_bf16(cudaStream_t s, const float* in, bf16* out, long long n) {
  cast_kernel<<<(unsigned)min((n + 255) / 256, (long long)kNumSMs * 16), 256, 0, s>>>(in, out, n);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

__global__ void scale_kernel(float* w, long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    w[i] *= scale;
}
int scale_rows_f32(cudaStream_t s, float* w, long long rows, long long cols, float scale) {
  const long long n = rows * cols;
  scale_kernel<<<(unsigned)min((n + 255) / 256, (long long)kNumSMs * 16), 256, 0, s>>>(w, n, scale);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

__global__ void permute_conv_kernel(const float* __restrict__ in, float* __restrict__ out, int co, int ci, int k) {
  const long long n = (long long)co * ci * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ci);
    const int kk = (int)((i / ci) % k);
    const int o = (int)(i / ((long long)ci * k));
    out[i] = in[((long long)o * ci + c) * k + kk];
  }
}
int permute_conv_weight(cudaStream_t s, const float* in, float* out, int co, int ci, int k) {
  const long long n = (long long)co * ci * k;
  permute_conv_kernel<<<(unsigned)min((n + 255) / 256, (long long)kNumSMs * 16), 256, 0, s>>>(in, out, co, ci, k);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// one block per tap k: norm over (co, ci) of v[:, :, k], then scaled, permuted write
__global__ void __launch_bounds__(1024)
posconv_fold_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ out, int co, int ci, int k) {
  __shared__ float red[32];
  __shared__ float scale_s;
  const int kk = blockIdx.x;
  const long long n = (long long)co * ci;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i * k + kk];
    s += x * x;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = red[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) scale_s = g[kk] / sqrtf(t);
  }
  __syncthreads();
  const float sc = scale_s;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % ci);
    const int o = (int)(i / ci);
    out[((long long)o * k + kk) * ci + c] = v[i * k + kk] * sc;
  }
}
int posconv_fold_weight(cudaStream_t s, const float* v, const float* g, float* out, int co, int ci, int k) {
  posconv_fold_kernel<<<k, 1024, 0, s>>>(v, g, out, co, ci, k);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int r, int c) {
  const int n = r * c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int rr = i / c, cc = i % c;
    out[cc * r + rr] = in[i];
  }
}
int transpose_f32(cudaStream_t s, const float* in, float* out, int r, int c) {
  transpose_kernel<<<ceil_div(r * c, 256), 256, 0, s>>>(in, out, r, c);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

__global__ void bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sc = w[i] / sqrtf(var[i] + eps);
  scale[i] = sc;
  shift[i] = b[i] - mean[i] * sc;
}
int bn_fold(cudaStream_t s, const float* w, const float* b, const float* mean, const float* var, float eps,
            float* scale, float* shift, int n) {
  bn_fold_kernel<<<ceil_div(n, 128), 128, 0, s>>>(w, b, mean, var, eps, scale, shift, n);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace rtdf
