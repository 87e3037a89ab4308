// Conformer back-end kernels (fp32 arithmetic, fp32 or bf16 I/O).
#include "conformer.cuh"

#include <stdlib.h>

namespace rtdf {

__global__ void conformer_stem_kernel(const float* __restrict__ z, const float* __restrict__ tok, int T, int E,
                                      float sc, float sh, float* __restrict__ x, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = (int)(i % E);
  const int r = (int)((i / E) % (T + 1));
  const long long b = i / ((long long)E * (T + 1));
  x[i] = r == 0 ? tok[e] : selu_f(z[(b * T + r - 1) * E + e] * sc + sh);
}

int conformer_stem(cudaStream_t s, const float* z, const float* class_token, int B, int T, int E, float bn_s,
                   float bn_t, float* x) {
  RTDF_REQUIRE(z && class_token && x, "conformer_stem: bad arguments");
  const long long total = (long long)B * (T + 1) * E;
  conformer_stem_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(z, class_token, T, E, bn_s, bn_t, x, total);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// MHSA with Shaw relative-position bias.  CTA = (head of an utterance, tile of 32 queries); K, V and
// the (2n-1)-row window of the relative-position table live in smem; one warp per query.
// ------------------------------------------------------------------------------------------------
template <typename T_>
__global__ void __launch_bounds__(256)
conformer_attn_kernel(const T_* __restrict__ qkv, const float* __restrict__ rel, T_* __restrict__ out, int n,
                      int heads, int dh, float scale) {
  // K, V and the relative-position rows are kept in the I/O type (bf16 in bf16 mode: exact for K / V, the table is rounded
  // like in the tensor-core kernel), so that utterances of up to 512 frames fit; only the n + 31 table rows this CTA's 32
  // queries can reach are loaded: row (i - j) + n - 1 for i in [i0, i0 + 32), j in [0, n)  ->  [i0, i0 + n + 30].
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int P = dh + (sizeof(T_) == 2 ? 2 : 1);
  const int np = (n + 31) & ~31;
  const int i0 = blockIdx.y * 32;
  const int ne = n + 31;
  T_* sK = reinterpret_cast<T_*>(sm_raw);  // [n][P]
  T_* sV = sK + (size_t)n * P;             // [n][dh]
  T_* sE = sV + (size_t)n * dh;            // [n + 31][P]   (table row i0 + r)
  float* sQ = reinterpret_cast<float*>(sm_raw + ((((size_t)n * P + (size_t)n * dh + (size_t)ne * P) * sizeof(T_) + 15) & ~(size_t)15));  // [8][64]
  float* sP = sQ + 8 * 64;                 // [8][np]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int E = heads * dh, ld = 3 * E;
  const T_* base = qkv + (long long)b * n * ld;
  for (int i = threadIdx.x; i < n * dh; i += 256) {
    const int t = i / dh, d = i % dh;
    sK[t * P + d] = base[(long long)t * ld + E + h * dh + d];
    sV[t * dh + d] = base[(long long)t * ld + 2 * E + h * dh + d];
  }
  for (int i = threadIdx.x; i < ne * dh; i += 256) {
    const int r = i / dh, d = i % dh;
    int relpos = i0 + r - (n - 1);
    relpos = max(-512, min(512, relpos)) + 512;
    sE[r * P + d] = from_f32<T_>(rel[relpos * dh + d]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = sQ + warp * 64;
  float* pr = sP + warp * np;
  for (int qi = warp; qi < 32; qi += 8) {
    const int i = blockIdx.y * 32 + qi;
    if (i >= n) break;
    for (int d = lane; d < dh; d += 32) q[d] = to_f32(base[(long long)i * ld + h * dh + d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < np; j += 32) {
      float sc = -INFINITY;
      if (j < n) {
        const T_* kr = sK + j * P;
        const T_* er = sE + (i - i0 - j + n - 1) * P;
        float dk = 0.f, de = 0.f;
        for (int d = 0; d < dh; ++d) {
          dk = fmaf(q[d], to_f32(kr[d]), dk);
          de = fmaf(q[d], to_f32(er[d]), de);
        }
        sc = dk * scale + de * scale;
      }
      pr[j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < np; j += 32) {
      const float e = j < n ? expf(pr[j] - mx) : 0.f;
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < n; ++j) o = fmaf(pr[j], to_f32(sV[j * dh + d]), o);
      out[((long long)b * n + i) * E + h * dh + d] = from_f32<T_>(o * inv);
    }
    __syncwarp();
  }
}

template <typename T_>
static int attn_launch(cudaStream_t s, const T_* qkv, const float* rel, T_* out, int B, int n, int heads, int dh) {
  RTDF_REQUIRE(qkv && rel && out && B > 0 && n > 0 && dh >= 1 && dh <= 64, "conformer_attention: bad arguments");
  const int np = (n + 31) & ~31;
  const int P = dh + (sizeof(T_) == 2 ? 2 : 1);
  const size_t smem = ((((size_t)n * P + (size_t)n * dh + (size_t)(n + 31) * P) * sizeof(T_) + 15) & ~(size_t)15) +
                      ((size_t)8 * 64 + (size_t)8 * np) * sizeof(float);
  RTDF_REQUIRE(smem <= 220 * 1024, "conformer_attention: sequence of %d tokens too long (bf16 mode: up to ~900, fp32: ~460)", n);
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&conformer_attn_kernel<T_>), (size_t)smem));
  dim3 grid(B * heads, ceil_div(n, 32));
  conformer_attn_kernel<T_><<<grid, 256, smem, s>>>(qkv, rel, out, n, heads, dh, 1.0f / sqrtf((float)dh));
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}
int conformer_attention_f32(cudaStream_t s, const float* qkv, const float* rel_pos, float* out, int B, int n, int heads, int dh) {
  return attn_launch<float>(s, qkv, rel_pos, out, B, n, heads, dh);
}
int conformer_attention_bf16(cudaStream_t s, const bf16* qkv, const float* rel_pos, bf16* out, int B, int n, int heads, int dh) {
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("RTDF_CONF_ATTN_IMPL");
    impl = (e && e[0] == '1') ? 1 : 0;
  }
  if (impl == 0) {
    const int r = conformer_attention_mma(s, qkv, rel_pos, out, B, n, heads, dh);
    if (r != RTDF_ERR_UNSUPPORTED) return r;
  }
  return attn_launch<bf16>(s, qkv, rel_pos, out, B, n, heads, dh);
}

// ------------------------------------------------------------------------------------------------
// GLU -> depth-wise conv (k taps, "same" padding) -> BatchNorm (folded) -> Swish.  Channels-last:
// thread = channel, CTA = 32 time steps of one utterance, GLU'd halo tile staged in smem.
// ------------------------------------------------------------------------------------------------
constexpr int kDwTT = 32;
constexpr int kDwMaxK = 63;

template <typename T_>
__global__ void __launch_bounds__(512)
glu_dwconv_kernel(const T_* __restrict__ in, T_* __restrict__ out, int n, int inner, int k,
                  const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ bn_s,
                  const float* __restrict__ bn_t) {
  extern __shared__ float tile[];  // [kDwTT + k - 1][inner]
  const int b = blockIdx.y, t0 = blockIdx.x * kDwTT;
  const int padl = k / 2;          // (k/2, k/2 - (k+1)%2): odd k -> symmetric
  const int rows = kDwTT + k - 1;
  const T_* inb = in + (long long)b * n * 2 * inner;
  for (int i = threadIdx.x; i < rows * inner; i += blockDim.x) {
    const int r = i / inner, c = i % inner;
    const int t = t0 - padl + r;
    float v = 0.f;
    if (t >= 0 && t < n) {
      const float a = to_f32(inb[(long long)t * 2 * inner + c]);
      const float g = to_f32(inb[(long long)t * 2 * inner + inner + c]);
      v = a * sigmoid_f(g);
    }
    tile[i] = v;
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= inner) return;
  float wk[kDwMaxK];
#pragma unroll
  for (int j = 0; j < kDwMaxK; ++j) wk[j] = j < k ? w[c * k + j] : 0.f;
  const float bi = bias[c], sc = bn_s[c], sh = bn_t[c];
  for (int tt = 0; tt < kDwTT; ++tt) {
    const int t = t0 + tt;
    if (t >= n) break;
    float acc = bi;
#pragma unroll
    for (int j = 0; j < kDwMaxK; ++j)
      if (j < k) acc = fmaf(wk[j], tile[(tt + j) * inner + c], acc);
    out[((long long)b * n + t) * inner + c] = from_f32<T_>(swish_f(acc * sc + sh));
  }
}

// Same contract with the tap count known at compile time: every thread (= channel) produces 8 consecutive time steps
// per pass from K + 7 shared-memory reads (6.5 FMAs per LDS instead of 1), 64 time steps per CTA.
constexpr int kDwTT2 = 64;
template <typename T_, int K>
__global__ void __launch_bounds__(512)
glu_dwconv_k_kernel(const T_* __restrict__ in, T_* __restrict__ out, int n, int inner, const float* __restrict__ w,
                    const float* __restrict__ bias, const float* __restrict__ bn_s, const float* __restrict__ bn_t) {
  extern __shared__ float tile[];  // [kDwTT2 + K - 1][inner]
  const int b = blockIdx.y, t0 = blockIdx.x * kDwTT2;
  constexpr int padl = K / 2;
  constexpr int rows = kDwTT2 + K - 1;
  const T_* inb = in + (long long)b * n * 2 * inner;
  for (int i = threadIdx.x; i < rows * inner; i += blockDim.x) {
    const int r = i / inner, c = i % inner;
    const int t = t0 - padl + r;
    float v = 0.f;
    if (t >= 0 && t < n) {
      const float a = to_f32(inb[(long long)t * 2 * inner + c]);
      const float g = to_f32(inb[(long long)t * 2 * inner + inner + c]);
      v = a * sigmoid_f(g);
    }
    tile[i] = v;
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= inner) return;
  float wk[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wk[j] = w[c * K + j];
  const float bi = bias[c], sc = bn_s[c], sh = bn_t[c];
  for (int tt0 = 0; tt0 < kDwTT2 && t0 + tt0 < n; tt0 += 8) {
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = bi;
#pragma unroll
    for (int j = 0; j < K + 7; ++j) {
      const float v = tile[(tt0 + j) * inner + c];
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (j - o >= 0 && j - o < K) acc[o] = fmaf(wk[j - o], v, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int t = t0 + tt0 + o;
      if (t < n) out[((long long)b * n + t) * inner + c] = from_f32<T_>(swish_f(acc[o] * sc + sh));
    }
  }
}

template <typename T_>
static int dw_launch(cudaStream_t s, const T_* in, T_* out, int B, int n, int inner, int k, const float* w,
                     const float* bias, const float* bn_s, const float* bn_t) {
  RTDF_REQUIRE(in && out && w && bias && bn_s && bn_t, "glu_dwconv: bad arguments");
  RTDF_REQUIRE(k % 2 == 1 && k <= kDwMaxK && inner <= 512, "glu_dwconv: unsupported kernel %d / width %d", k, inner);
  if (k == 31 && (size_t)(kDwTT2 + 30) * inner * sizeof(float) <= 200 * 1024) {      // the reference's kernel_size
    const size_t smem2 = (size_t)(kDwTT2 + 30) * inner * sizeof(float);
    RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&glu_dwconv_k_kernel<T_, 31>), (size_t)smem2));
    glu_dwconv_k_kernel<T_, 31><<<dim3(ceil_div(n, kDwTT2), B), ((inner + 31) / 32) * 32, smem2, s>>>(in, out, n, inner, w, bias,
                                                                                                     bn_s, bn_t);
    RTDF_LAUNCH_CHECK();
    return RTDF_OK;
  }
  const size_t smem = (size_t)(kDwTT + k - 1) * inner * sizeof(float);
  RTDF_REQUIRE(smem <= 200 * 1024, "glu_dwconv: tile too large");
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&glu_dwconv_kernel<T_>), (size_t)smem));
  const int threads = ((inner + 31) / 32) * 32;
  glu_dwconv_kernel<T_><<<dim3(ceil_div(n, kDwTT), B), threads, smem, s>>>(in, out, n, inner, k, w, bias, bn_s, bn_t);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}
int conformer_glu_dwconv_f32(cudaStream_t s, const float* in, float* out, int B, int n, int inner, int k,
                             const float* w, const float* bias, const float* bn_s, const float* bn_t) {
  return dw_launch<float>(s, in, out, B, n, inner, k, w, bias, bn_s, bn_t);
}
int conformer_glu_dwconv_bf16(cudaStream_t s, const bf16* in, bf16* out, int B, int n, int inner, int k,
                              const float* w, const float* bias, const float* bn_s, const float* bn_t) {
  return dw_launch<bf16>(s, in, out, B, n, inner, k, w, bias, bn_s, bn_t);
}

__global__ void conformer_head_kernel(const float* __restrict__ x, int n, int E, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ logits) {
  const int b = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tok = x + (long long)b * n * E;
  float s = 0.f;
  for (int e = lane; e < E; e += 32) s = fmaf(w[k * E + e], tok[e], s);
  s = warp_sum(s);
  if (lane == 0) logits[b * 2 + k] = s + bias[k];
}

int conformer_head(cudaStream_t s, const float* x, int B, int n, int E, const float* w, const float* bias, float* logits) {
  RTDF_REQUIRE(x && w && bias && logits, "conformer_head: bad arguments");
  conformer_head_kernel<<<B, 64, 0, s>>>(x, n, E, w, bias, logits);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace rtdf
