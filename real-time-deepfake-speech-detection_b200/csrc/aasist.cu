// AASIST back-end kernels, fp32 end to end (protects top-k parity; ~1 % of the path's FLOPs).
#include "aasist.cuh"

namespace rtdf {

// ------------------------------------------------------------------------------------------------
// stem: (B,T,128) -> transpose -> max_pool2d(3,3) -> BN2d(1) -> SELU     (xlsr_aasist.py:92-96)
// ------------------------------------------------------------------------------------------------
__global__ void stem_kernel(const float* __restrict__ z, int T, int Tp, float sc, float sh, float* __restrict__ out,
                            long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % 42);
  const int tp = (int)((i / 42) % Tp);
  const int b = (int)(i / (42LL * Tp));
  const float* p = z + ((long long)b * T + 3 * tp) * 128 + 3 * f;
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) m = fmaxf(m, p[j * 128 + k]);
  out[((long long)b * 42 + f) * Tp + tp] = selu_f(m * sc + sh);
}

int aasist_stem(cudaStream_t s, const float* z, int B, int T, float bn_scale, float bn_shift, float* out) {
  const int Tp = T / 3;
  RTDF_REQUIRE(z && out && B > 0 && Tp >= 1, "aasist_stem: bad arguments (T = %d)", T);
  const long long total = (long long)B * 42 * Tp;
  stem_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(z, T, Tp, bn_scale, bn_shift, out, total);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// direct conv2d, kernel (KH,3), pad (pad_h,1), stride 1, NCHW fp32.  CTA = (row tile, utterance);
// thread = 4 output channels x 8 consecutive pixels; input channels streamed through smem in
// chunks of 16 (weights chunk + input patch), float4 weight reads, float4/float2 patch reads.
// ------------------------------------------------------------------------------------------------
constexpr int kConvCC = 16;

template <int KH>
__global__ void __launch_bounds__(320)
conv2d_kernel(const Conv2dArgs a, int Hout, int RT, int WG, int Wp) {
  extern __shared__ __align__(16) float csm[];
  const int CC = min(kConvCC, a.Ci);
  const int PR = RT + KH - 1;                 // patch rows
  float* sw = csm;                            // [CC][KH][3][Co]
  float* sin = csm + kConvCC * KH * 3 * a.Co; // [CC][PR][Wp]
  const int cog = a.Co >> 2;
  const int item = threadIdx.x;
  const bool active = item < cog * RT * WG;
  const int cg = item % cog;
  const int wg = (item / cog) % WG;
  const int r = item / (cog * WG);
  const int h0 = blockIdx.x * RT, b = blockIdx.y;
  const float* inb = a.in + (long long)b * a.Ci * a.H * a.W;
  float acc[8][4];
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;

  for (int c0 = 0; c0 < a.Ci; c0 += CC) {
    const int wcount = CC * KH * 3 * a.Co;
    const float* wsrc = a.w + (long long)c0 * KH * 3 * a.Co;
    for (int i = threadIdx.x; i < wcount; i += blockDim.x) sw[i] = wsrc[i];
    const int pcount = CC * PR * Wp;
    for (int i = threadIdx.x; i < pcount; i += blockDim.x) {
      const int col = i % Wp;
      const int rr = (i / Wp) % PR;
      const int ci = i / (Wp * PR);
      const int hin = h0 - a.pad_h + rr, win = col - 1;
      float v = 0.f;
      if (hin >= 0 && hin < a.H && win >= 0 && win < a.W) v = inb[((long long)(c0 + ci) * a.H + hin) * a.W + win];
      sin[i] = v;
    }
    __syncthreads();
    if (active) {
      for (int ci = 0; ci < CC; ++ci) {
#pragma unroll
        for (int kh = 0; kh < KH; ++kh) {
          const float* prow = sin + (ci * PR + r + kh) * Wp + wg * 8;
          const float4 i0 = *reinterpret_cast<const float4*>(prow);
          const float4 i1 = *reinterpret_cast<const float4*>(prow + 4);
          const float2 i2 = *reinterpret_cast<const float2*>(prow + 8);
          const float in[10] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y};
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 wv = *reinterpret_cast<const float4*>(sw + ((ci * KH + kh) * 3 + kw) * a.Co + cg * 4);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              acc[p][0] = fmaf(in[p + kw], wv.x, acc[p][0]);
              acc[p][1] = fmaf(in[p + kw], wv.y, acc[p][1]);
              acc[p][2] = fmaf(in[p + kw], wv.z, acc[p][2]);
              acc[p][3] = fmaf(in[p + kw], wv.w, acc[p][3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  const int h = h0 + r;
  if (!active || h >= Hout) return;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = cg * 4 + c;
    const float bias = a.bias ? a.bias[co] : 0.f;
    const float s1 = a.s1 ? a.s1[co] : 1.f, t1 = a.t1 ? a.t1[co] : 0.f;
    const float s2 = a.s2 ? a.s2[co] : 1.f, t2 = a.t2 ? a.t2[co] : 0.f;
    const long long o = (((long long)b * a.Co + co) * Hout + h) * a.W + wg * 8;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (wg * 8 + p >= a.W) break;
      float v = acc[p][c] + bias;
      if (a.s1) v = v * s1 + t1;
      v = apply_act(v, a.act1);
      if (a.resid) v += a.resid[o + p];
      if (a.s2) v = v * s2 + t2;
      v = apply_act(v, a.act2);
      a.out[o + p] = v;
    }
  }
}

int aasist_conv2d(cudaStream_t s, const Conv2dArgs& a, int B) {
  RTDF_REQUIRE(a.in && a.w && a.out && B > 0 && B <= 65535, "conv2d: bad arguments");
  RTDF_REQUIRE((a.KH == 1 || a.KH == 2) && a.Co % 4 == 0 && a.Co <= 64, "conv2d: unsupported shape");
  RTDF_REQUIRE(a.Ci == 1 || a.Ci % kConvCC == 0, "conv2d: Ci must be 1 or a multiple of 16");
  const int Hout = a.H + 2 * a.pad_h - a.KH + 1;
  const int WG = ceil_div(a.W, 8);
  const int Wp = WG * 8 + 4;
  const int cog = a.Co / 4;
  RTDF_REQUIRE(cog * WG <= 320, "conv2d: row too wide (W = %d)", a.W);
  int RT = 320 / (cog * WG);
  if (RT > Hout) RT = Hout;
  if (RT > 8) RT = 8;
  const int threads = ((cog * RT * WG + 31) / 32) * 32;
  const size_t smem = ((size_t)kConvCC * a.KH * 3 * a.Co + (size_t)kConvCC * (RT + a.KH - 1) * Wp) * sizeof(float);
  dim3 grid(ceil_div(Hout, RT), B);
  if (a.KH == 1) {
    RTDF_CHECK_CUDA(cudaFuncSetAttribute(conv2d_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    conv2d_kernel<1><<<grid, threads, smem, s>>>(a, Hout, RT, WG, Wp);
  } else {
    RTDF_CHECK_CUDA(cudaFuncSetAttribute(conv2d_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    conv2d_kernel<2><<<grid, threads, smem, s>>>(a, Hout, RT, WG, Wp);
  }
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// attention map: per pixel 64 -> 128 (SELU, BN) -> 64.   CTA = (row h, utterance), 128 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_map_kernel(const float* __restrict__ x, int H, int W, int Wq, const float* __restrict__ w1t,
                const float* __restrict__ b1, const float* __restrict__ bn_s, const float* __restrict__ bn_t,
                const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ wmap) {
  extern __shared__ __align__(16) float asm_[];
  float* xs = asm_;             // [64][Wq]
  float* hid = asm_ + 64 * Wq;  // [128][Wq]
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  for (int i = t; i < 64 * Wq; i += 128) {
    const int w = i % Wq, c = i / Wq;
    xs[i] = w < W ? x[(((long long)b * 64 + c) * H + h) * W + w] : 0.f;
  }
  __syncthreads();
  for (int wb = 0; wb < Wq; wb += 32) {
    float acc[32];
    const float bj = b1[t];
#pragma unroll
    for (int p = 0; p < 32; ++p) acc[p] = bj;
    for (int c = 0; c < 64; ++c) {
      const float wv = w1t[c * 128 + t];
      const float* xr = xs + c * Wq + wb;
#pragma unroll
      for (int p = 0; p < 32; ++p) acc[p] = fmaf(wv, xr[p], acc[p]);
    }
    const float sc = bn_s[t], sh = bn_t[t];
#pragma unroll
    for (int p = 0; p < 32; ++p) hid[t * Wq + wb + p] = selu_f(acc[p]) * sc + sh;
  }
  __syncthreads();
  const int co = t & 63, half = t >> 6;
  for (int wb = 0, ch = 0; wb < Wq; wb += 32, ++ch) {
    if ((ch & 1) != half) continue;
    float acc[32];
    const float bj = b2[co];
#pragma unroll
    for (int p = 0; p < 32; ++p) acc[p] = bj;
    for (int j = 0; j < 128; ++j) {
      const float wv = w2t[j * 64 + co];
      const float* hr = hid + j * Wq + wb;
#pragma unroll
      for (int p = 0; p < 32; ++p) acc[p] = fmaf(wv, hr[p], acc[p]);
    }
    float* o = wmap + (((long long)b * 64 + co) * H + h) * W + wb;
#pragma unroll
    for (int p = 0; p < 32; ++p)
      if (wb + p < W) o[p] = acc[p];
  }
}

int aasist_attn_map(cudaStream_t s, const float* x, int B, int H, int W, const float* w1t, const float* b1,
                    const float* bn_s, const float* bn_t, const float* w2t, const float* b2, float* wmap) {
  RTDF_REQUIRE(x && wmap && W >= 1 && W <= 128, "attn_map: bad arguments (W = %d)", W);
  const int Wq = ((W + 31) / 32) * 32;
  const size_t smem = (size_t)(64 + 128) * Wq * sizeof(float);
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(attn_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  attn_map_kernel<<<dim3(H, B), 128, smem, s>>>(x, H, W, Wq, w1t, b1, bn_s, bn_t, w2t, b2, wmap);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// dual-softmax attention pooling.  CTA = (channel, utterance).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_pool_kernel(const float* __restrict__ x, const float* __restrict__ wmap, int H, int W,
                 const float* __restrict__ pos_S, float* __restrict__ e_S, float* __restrict__ e_T) {
  extern __shared__ float psm[];
  float* xs = psm;
  float* ws = psm + H * W;
  const int c = blockIdx.x, b = blockIdx.y;
  const long long base = ((long long)b * 64 + c) * H * W;
  for (int i = threadIdx.x; i < H * W; i += 128) {
    xs[i] = x[base + i];
    ws[i] = wmap[base + i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < H; h += 4) {  // softmax over w, per row
    float mx = -INFINITY;
    for (int w = lane; w < W; w += 32) mx = fmaxf(mx, ws[h * W + w]);
    mx = warp_max(mx);
    float se = 0.f, sx = 0.f;
    for (int w = lane; w < W; w += 32) {
      const float e = expf(ws[h * W + w] - mx);
      se += e;
      sx = fmaf(xs[h * W + w], e, sx);
    }
    se = warp_sum(se);
    sx = warp_sum(sx);
    if (lane == 0) e_S[((long long)b * H + h) * 64 + c] = sx / se + pos_S[h * 64 + c];
  }
  for (int w = threadIdx.x; w < W; w += 128) {  // softmax over h, per column
    float mx = -INFINITY;
    for (int h = 0; h < H; ++h) mx = fmaxf(mx, ws[h * W + w]);
    float se = 0.f, sx = 0.f;
    for (int h = 0; h < H; ++h) {
      const float e = expf(ws[h * W + w] - mx);
      se += e;
      sx = fmaf(xs[h * W + w], e, sx);
    }
    e_T[((long long)b * W + w) * 64 + c] = sx / se;
  }
}

int aasist_attn_pool(cudaStream_t s, const float* x, const float* wmap, int B, int H, int W, const float* pos_S,
                     float* e_S, float* e_T) {
  RTDF_REQUIRE(x && wmap && e_S && e_T && pos_S, "attn_pool: bad arguments");
  const size_t smem = (size_t)2 * H * W * sizeof(float);
  RTDF_REQUIRE(smem <= 48 * 1024, "attn_pool: map too large");
  attn_pool_kernel<<<dim3(64, B), 128, smem, s>>>(x, wmap, H, W, pos_S, e_S, e_T);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// fused graph-attention row: never materialises the (n,n,D) pairwise tensor.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxNodes = 256;

template <int D, int DO>
__global__ void __launch_bounds__(256)
gat_rows_kernel(const GraphView x, int n1, const GatRowWeights w, float* __restrict__ out, long long out_bs,
                const float* __restrict__ master_in, long long master_stride, const GatRowWeights wM,
                float* __restrict__ master_out) {
  extern __shared__ __align__(16) float gsm[];
  constexpr int WP = D + 4;
  const int n = x.n;
  float* sx = gsm;                 // [n][D]
  float* sW = sx + kMaxNodes * D;  // [DO][WP]
  float* sa = sW + DO * WP;        // [3][DO]
  float* sb = sa + 3 * DO;         // [DO]
  float* se = sb + DO;             // [kMaxNodes]
  float* sv = se + kMaxNodes;      // [D] query vector
  float* sax = sv + D;             // [D] aggregated
  const int i = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const bool is_master = (i == n);
  const GatRowWeights& ww = is_master ? wM : w;
  const float* xb = x.ptr + (long long)b * x.batch_stride;
  for (int k = t; k < n * D; k += 256) sx[k] = xb[k];
  for (int k = t; k < DO * D; k += 256) sW[(k / D) * WP + (k % D)] = ww.att_w[k];
  for (int k = t; k < DO; k += 256) {
    sa[k] = ww.a11[k];
    sa[DO + k] = ww.a22 ? ww.a22[k] : 0.f;
    sa[2 * DO + k] = ww.a12 ? ww.a12[k] : 0.f;
    sb[k] = ww.att_b[k];
  }
  if (t < D) sv[t] = is_master ? master_in[(long long)b * master_stride + t] : xb[(long long)i * D + t];
  __syncthreads();

  const int oq = t & 3;
  for (int j0 = 0; j0 < n; j0 += 64) {
    const int j = j0 + (t >> 2);
    float e = 0.f;
    if (j < n) {
      float prod[D];
#pragma unroll
      for (int d = 0; d < D; ++d) prod[d] = sv[d] * sx[j * D + d];
      const float* av = sa + (is_master ? 0 : ((i < n1) == (j < n1) ? (i < n1 ? 0 : DO) : 2 * DO));
#pragma unroll 2
      for (int ii = 0; ii < DO / 4; ++ii) {
        const int o = oq + 4 * ii;
        const float* wr = sW + o * WP;
        float dot = sb[o];
#pragma unroll
        for (int d = 0; d < D; d += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wr + d);
          dot = fmaf(w4.x, prod[d], dot);
          dot = fmaf(w4.y, prod[d + 1], dot);
          dot = fmaf(w4.z, prod[d + 2], dot);
          dot = fmaf(w4.w, prod[d + 3], dot);
        }
        e = fmaf(av[o], tanhf(dot), e);
      }
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    if (oq == 0 && j < n) se[j] = e * ww.inv_temp;
  }
  __syncthreads();
  if (t < 32) {  // softmax over j
    float mx = -INFINITY;
    for (int j = t; j < n; j += 32) mx = fmaxf(mx, se[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = t; j < n; j += 32) {
      const float ev = expf(se[j] - mx);
      se[j] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = t; j < n; j += 32) se[j] *= inv;
  }
  __syncthreads();
  if (t < D) {
    float a = 0.f;
    for (int j = 0; j < n; ++j) a = fmaf(se[j], sx[j * D + t], a);
    sax[t] = a;
  }
  __syncthreads();
  if (t < DO) {
    float y = ww.with_b[t] + ww.without_b[t];
    float y2 = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      y = fmaf(ww.with_t[d * DO + t], sax[d], y);
      y2 = fmaf(ww.without_t[d * DO + t], sv[d], y2);
    }
    y += y2;
    if (is_master) {
      master_out[(long long)b * DO + t] = y;
    } else {
      out[(long long)b * out_bs + (long long)i * DO + t] = selu_f(y * ww.bn_s[t] + ww.bn_t[t]);
    }
  }
}

template <int D, int DO>
static int gat_launch(cudaStream_t s, const GraphView& x, int B, int n1, const GatRowWeights& w, float* out,
                      long long out_bs, const float* master_in, long long master_stride, const GatRowWeights* wM,
                      float* master_out) {
  const size_t smem = ((size_t)kMaxNodes * D + DO * (D + 4) + 4 * DO + kMaxNodes + 2 * D) * sizeof(float);
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&gat_rows_kernel<D, DO>), (size_t)smem));
  const bool has_master = master_in != nullptr;
  dim3 grid(x.n + (has_master ? 1 : 0), B);
  gat_rows_kernel<D, DO><<<grid, 256, smem, s>>>(x, n1, w, out, out_bs, master_in, master_stride,
                                                  has_master ? *wM : w, master_out);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int aasist_gat_rows(cudaStream_t s, int D, int DO, const GraphView& x, int B, int n1, const GatRowWeights& w,
                    float* out, long long out_batch_stride, const float* master_in, long long master_stride,
                    const GatRowWeights* wM, float* master_out) {
  RTDF_REQUIRE(x.ptr && out && x.n >= 1 && x.n <= kMaxNodes && B > 0 && B <= 65535, "gat_rows: bad arguments (n = %d)", x.n);
  RTDF_REQUIRE(!master_in || (wM && master_out), "gat_rows: master row needs weights and an output");
  if (D == 64 && DO == 64) return gat_launch<64, 64>(s, x, B, n1, w, out, out_batch_stride, master_in, master_stride, wM, master_out);
  if (D == 64 && DO == 32) return gat_launch<64, 32>(s, x, B, n1, w, out, out_batch_stride, master_in, master_stride, wM, master_out);
  if (D == 32 && DO == 32) return gat_launch<32, 32>(s, x, B, n1, w, out, out_batch_stride, master_in, master_stride, wM, master_out);
  set_error("gat_rows: unsupported dims %d -> %d", D, DO);
  return RTDF_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
__global__ void type_proj_kernel(int D, const GraphView x1, const GraphView x2, const float* __restrict__ w1t,
                                 const float* __restrict__ b1, const float* __restrict__ w2t,
                                 const float* __restrict__ b2, float* __restrict__ out) {
  __shared__ float sv[64];
  const int i = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const bool first = i < x1.n;
  const float* src = first ? x1.ptr + (long long)b * x1.batch_stride + (long long)i * D
                           : x2.ptr + (long long)b * x2.batch_stride + (long long)(i - x1.n) * D;
  sv[t] = src[t];
  __syncthreads();
  const float* wt = first ? w1t : w2t;
  float y = first ? b1[t] : b2[t];
  for (int d = 0; d < D; ++d) y = fmaf(wt[d * D + t], sv[d], y);
  out[((long long)b * (x1.n + x2.n) + i) * D + t] = y;
}

int aasist_type_proj(cudaStream_t s, int D, const GraphView& x1, const GraphView& x2, int B, const float* w1t,
                     const float* b1, const float* w2t, const float* b2, float* out) {
  RTDF_REQUIRE(D == 32 || D == 64, "type_proj: unsupported D");
  type_proj_kernel<<<dim3(x1.n + x2.n, B), D, 0, s>>>(D, x1, x2, w1t, b1, w2t, b2, out);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// GraphPool: rank-by-count top-k (descending, ties -> lower index first), scale, gather.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
graph_pool_kernel(int D, const GraphView h, const float* __restrict__ w, const float* __restrict__ bptr, int k,
                  float* __restrict__ out, int* __restrict__ idx_out) {
  __shared__ float ss[kMaxNodes];
  __shared__ int sidx[kMaxNodes];
  const int b = blockIdx.x, t = threadIdx.x, n = h.n;
  const float* hb = h.ptr + (long long)b * h.batch_stride;
  for (int i = t; i < n; i += 128) {
    float z = bptr[0];
    for (int d = 0; d < D; ++d) z = fmaf(w[d], hb[(long long)i * D + d], z);
    ss[i] = 1.0f / (1.0f + expf(-z));
  }
  __syncthreads();
  for (int i = t; i < n; i += 128) {
    const float si = ss[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = ss[j];
      rank += (sj > si) || (sj == si && j < i);
    }
    if (rank < k) sidx[rank] = i;
  }
  __syncthreads();
  for (int e = t; e < k * D; e += 128) {
    const int r = e / D, d = e % D;
    const int src = sidx[r];
    out[((long long)b * k + r) * D + d] = hb[(long long)src * D + d] * ss[src];
  }
  if (idx_out)
    for (int i = t; i < k; i += 128) idx_out[(long long)b * k + i] = sidx[i];
}

int aasist_graph_pool(cudaStream_t s, int D, const GraphView& h, int B, const float* w, const float* b, int k,
                      float* out, int* idx_out) {
  RTDF_REQUIRE(h.ptr && out && h.n >= 1 && h.n <= kMaxNodes && k >= 1 && k <= h.n, "graph_pool: bad arguments");
  graph_pool_kernel<<<B, 128, 0, s>>>(D, h, w, b, k, out, idx_out);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// ------------------------------------------------------------------------------------------------
// read-out (xlsr_aasist.py:134-175), including the reference's `out_S1 + 1` quirk (:138).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gv(const GraphView& g, int b, int node, int c) {
  return g.ptr[(long long)b * g.batch_stride + node * 32 + c];
}

__global__ void __launch_bounds__(160) readout_kernel(const ReadoutArgs a) {
  __shared__ float hid[160];
  const int b = blockIdx.x, t = threadIdx.x, seg = t >> 5, c = t & 31;
  float v;
  if (seg < 2) {
    const int n = a.T1.n;
    float mx = 0.f, sum = 0.f;
    for (int i = 0; i < n; ++i) {
      const float x = fmaxf(gv(a.T1, b, i, c) + gv(a.Ta1, b, i, c), gv(a.T2, b, i, c) + gv(a.Ta2, b, i, c));
      mx = fmaxf(mx, fabsf(x));
      sum += x;
    }
    v = seg == 0 ? mx : sum / n;
  } else if (seg < 4) {
    const int n = a.S1.n;
    float mx = 0.f, sum = 0.f;
    for (int i = 0; i < n; ++i) {
      const float x = fmaxf(gv(a.S1, b, i, c) + 1.0f, gv(a.S2, b, i, c) + gv(a.Sa2, b, i, c));
      mx = fmaxf(mx, fabsf(x));
      sum += x;
    }
    v = seg == 2 ? mx : sum / n;
  } else {
    v = fmaxf(a.m1a[b * 32 + c] + a.m1b[b * 32 + c], a.m2a[b * 32 + c] + a.m2b[b * 32 + c]);
  }
  hid[t] = v;
  if (a.hidden) a.hidden[(long long)b * 160 + t] = v;
  __syncthreads();
  if (t < 64) {
    const int k = t >> 5;
    float s = 0.f;
    for (int i = c; i < 160; i += 32) s = fmaf(a.w[k * 160 + i], hid[i], s);
    s = warp_sum(s);
    if (c == 0) a.logits[b * 2 + k] = s + a.b[k];
  }
}

int aasist_readout(cudaStream_t s, const ReadoutArgs& a, int B) {
  RTDF_REQUIRE(a.logits && a.w && a.b && a.T1.ptr && a.S1.ptr, "readout: bad arguments");
  readout_kernel<<<B, 160, 0, s>>>(a);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace rtdf
