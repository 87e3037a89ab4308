// The two ends of the scoring path, as HBM-bound kernels:
//   * fit_duration: ragged utterances -> fixed-length (B, duration) batch, tile-repeat / crop
//     (reference data/test_set.py:201-248 adjustDuration[_random_start]) fused with PreEmphasis
//     (data/preprocess.py:22-27), so short utterances cross PCIe once and are tiled in HBM;
//   * score_sink: out[:,1] -> device-resident score vector + weighted-CE / accuracy accumulators
//     (main.py:210-212, trainer.py:104-113) without the reference's per-batch .cpu()/.item() syncs;
//   * roc_counts + roc_crossing: integer ROC (TP/FP at every threshold) and the two ROC points that
//     bracket the EER (trainer.py:134-139 calculate_EER), bit-exact integer work.
#include "../../include/rtdf.h"
#include "common.cuh"

using namespace rtdf;

namespace {

// ------------------------------------------------------------------------------------------------
// fit_duration: out[b][t] = src_b[(start_b + t) mod len_b]; optional pre-emphasis on the fitted signal
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fit_duration_kernel(const float* __restrict__ packed,
                                                           const long long* __restrict__ offsets,
                                                           const int* __restrict__ starts, float* __restrict__ out,
                                                           int duration, int preemph, float coef) {
  const int b = blockIdx.y;
  const long long lo = offsets[b];
  const int len = (int)(offsets[b + 1] - lo);
  const float* src = packed + lo;
  float* dst = out + (long long)b * duration;
  const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t0 >= duration) return;
  if (len <= 0) {  // empty utterance: silence (the reference would divide by zero; callers reject it on the host)
    for (int t = t0; t < min(t0 + 4, duration); ++t) dst[t] = 0.f;
    return;
  }
  const int st = starts ? starts[b] : 0;
  int p = (int)(((long long)st + t0) % len);
  // sample before the 4-vector, in fitted coordinates: fitted[-1] := fitted[1] (reflect padding)
  float prev = 0.f;
  if (preemph) {
    const long long q = t0 == 0 ? (long long)st + (duration > 1 ? 1 : 0) : (long long)st + t0 - 1;
    prev = src[q % len];
  }
  float v[4];
  const int n = min(4, duration - t0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < n) {
      v[i] = src[p];
      p = p + 1 == len ? 0 : p + 1;
    }
  }
  if (preemph) {
    float last = prev;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < n) {
        const float cur = v[i];
        v[i] = cur - coef * last;
        last = cur;
      }
    }
  }
  if (n == 4 && (duration & 3) == 0) {
    *reinterpret_cast<float4*>(dst + t0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int i = 0; i < n; ++i) dst[t0 + i] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// score_sink: one CTA per batch call
// acc[0] += B * (sum_i w[y_i] nll_i / sum_i w[y_i])   (nn.CrossEntropyLoss(weight), reduction 'mean', times batch size:
//                                                      trainer.py:108,112), acc[1] += #correct, acc[2] += B
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_sink_kernel(const float* __restrict__ logits, int B,
                                                         const long long* __restrict__ labels,
                                                         const float* __restrict__ class_w, float* __restrict__ scores,
                                                         double* __restrict__ acc) {
  __shared__ double red[3][8];
  double wl = 0.0, ws = 0.0, nc = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float a = logits[2 * i], c = logits[2 * i + 1];
    if (scores) scores[i] = c;                       // bona-fide score, main.py:212
    if (labels) {
      const int y = labels[i] != 0;
      const float m = fmaxf(a, c);
      const float lse = m + logf(expf(a - m) + expf(c - m));
      const float nll = lse - (y ? c : a);
      const float w = class_w ? class_w[y] : 1.0f;
      wl += (double)(w * nll);
      ws += (double)w;
      const int pred = c > a ? 1 : 0;                // torch.max: first index wins a tie
      nc += pred == y ? 1.0 : 0.0;
    }
  }
  if (!labels || !acc) return;
  for (int o = 16; o > 0; o >>= 1) {
    wl += __shfl_xor_sync(0xffffffffu, wl, o);
    ws += __shfl_xor_sync(0xffffffffu, ws, o);
    nc += __shfl_xor_sync(0xffffffffu, nc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = wl;
    red[1][threadIdx.x >> 5] = ws;
    red[2][threadIdx.x >> 5] = nc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = 0; i < 8; ++i) { a0 += red[0][i]; a1 += red[1][i]; a2 += red[2][i]; }
    acc[0] += a1 > 0 ? (double)B * (a0 / a1) : 0.0;
    acc[1] += a2;
    acc[2] += (double)B;
  }
}

// ------------------------------------------------------------------------------------------------
// roc_counts: tp[i] = #{j : label_j = 1, s_j >= s_i}, fp[i] = #{j : label_j = 0, s_j >= s_i}
// Thread per threshold i, scores streamed through shared memory in tiles; the label rides in the tile as the
// counter increment (1 for a positive, 1<<32 for a negative) so the inner loop is one compare + one predicated add.
// NaN scores (padding of the score gather) are ignored on both sides.
// ------------------------------------------------------------------------------------------------
constexpr int kRocTile = 2048;
__global__ void __launch_bounds__(256) roc_counts_kernel(const float* __restrict__ scores,
                                                         const long long* __restrict__ labels, int n,
                                                         int* __restrict__ tp, int* __restrict__ fp) {
  __shared__ float s_sc[kRocTile];
  __shared__ unsigned long long s_inc[kRocTile];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float mine = i < n ? scores[i] : 0.f;
  unsigned long long cnt = 0;
  for (int base = 0; base < n; base += kRocTile) {
    const int m = min(kRocTile, n - base);
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const float v = scores[base + j];
      s_sc[j] = v;
      s_inc[j] = v != v ? 0ull : (labels[base + j] != 0 ? 1ull : (1ull << 32));
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < m; ++j) cnt += s_sc[j] >= mine ? s_inc[j] : 0ull;
  }
  if (i < n) {
    const bool bad = mine != mine;
    tp[i] = bad ? -1 : (int)(cnt & 0xffffffffull);
    fp[i] = bad ? -1 : (int)(cnt >> 32);
  }
}

// roc_crossing: with P positives and Q negatives, g(point) = 1 - fp/Q - tp/P is >= 0  <=>  fp*P + tp*Q <= P*Q.
// Points are ordered along the curve by (tp + fp).  out[0..1] = (tp,fp) of the LAST point with g >= 0 (the origin
// (0,0) if none), out[2..3] = (tp,fp) of the FIRST point with g < 0.  Keys are packed (tp+fp, tp) so one 64-bit
// atomicMax / atomicMin per thread settles it.
__global__ void __launch_bounds__(256) roc_crossing_kernel(const int* __restrict__ tp, const int* __restrict__ fp, int n,
                                                           long long P, long long Q,
                                                           unsigned long long* __restrict__ keys /*[2]: max, min*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || tp[i] < 0) return;
  const long long t = tp[i], f = fp[i];
  const unsigned long long key = ((unsigned long long)(t + f) << 32) | (unsigned long long)t;
  if (f * P + t * Q <= P * Q) atomicMax(keys, key);
  else atomicMin(keys + 1, key);
}

}  // namespace

extern "C" {

int rtdf_fit_duration(const float* packed, const long long* offsets, const int* starts, int batch, int duration,
                      int preemph, float coef, float* out, void* stream) {
  RTDF_REQUIRE(packed && offsets && out && batch >= 1 && duration >= 1, "rtdf_fit_duration: bad arguments");
  RTDF_REQUIRE(batch <= 65535, "rtdf_fit_duration: batch %d too large", batch);
  dim3 grid(ceil_div(ceil_div(duration, 4), 256), batch);
  fit_duration_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, offsets, starts, out, duration,
                                                                          preemph, coef);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int rtdf_score_sink(const float* logits, int batch, const long long* labels, const float* class_weight, float* scores,
                    double* acc, void* stream) {
  RTDF_REQUIRE(logits && batch >= 1, "rtdf_score_sink: bad arguments");
  RTDF_REQUIRE(scores || (labels && acc), "rtdf_score_sink: nothing to do");
  score_sink_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, batch, labels, class_weight, scores, acc);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int rtdf_roc_counts(const float* scores, const long long* labels, int n, int32_t* tp, int32_t* fp, void* stream) {
  RTDF_REQUIRE(scores && labels && tp && fp && n >= 1, "rtdf_roc_counts: bad arguments");
  roc_counts_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, labels, n, tp, fp);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int rtdf_roc_crossing(const int32_t* tp, const int32_t* fp, int n, long long n_pos, long long n_neg,
                      unsigned long long* keys, void* stream) {
  RTDF_REQUIRE(tp && fp && keys && n >= 1 && n_pos >= 1 && n_neg >= 1, "rtdf_roc_crossing: needs both classes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned long long init[2] = {0ull, ~0ull};
  RTDF_CHECK_CUDA(cudaMemcpyAsync(keys, init, sizeof(init), cudaMemcpyHostToDevice, s));
  roc_crossing_kernel<<<ceil_div(n, 256), 256, 0, s>>>(tp, fp, n, n_pos, n_neg, keys);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // extern "C"
