// Graph-attention rows on the tensor cores (bf16 mode): GraphAttentionLayer / HtrgGraphAttentionLayer of
// reference models/aasist_modules.py:17-294.
//
// The pairwise attention map  e_ij = a . tanh(W (x_i (.) x_j) + b) / temp  is a GEMM whose A operand -- the
// element-wise products of node pairs, (n*n, D) per utterance -- never exists in memory: each warp owns one
// query node i, builds the A fragments of 16-pair tiles in registers from x_i and the shared-memory copy of the
// graph, and multiplies them with W (DO x D) held in shared memory in fragment order.  fp32 accuracy is kept by
// splitting both operands into (hi, lo) bf16 pairs and issuing three MMAs per product (hi*hi + lo*hi + hi*lo,
// relative error ~2^-16), the same scheme as the shifted-row convolutions of conv_tc.cu: the back-end feeds a
// top-k node selection, so it must not run at plain bf16 accuracy.  tanh, the dot with the attention vector, the
// softmax over j, the aggregation, both output projections, BatchNorm and SELU stay in the same warp.
//
// Warp-level mma.sync (m16n8k16) rather than tcgen05: a tile here is 16 pairs x 64 outputs x 64 inputs, the
// whole layer is 2.3 GFLOP per 64 utterances, and the operand has to be produced by the threads themselves.
#include "aasist.cuh"

namespace rtdf {

namespace {

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (hi, lo) split of two fp32 values into packed bf16x2 registers
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
  const float r0 = v0 - __bfloat162float(h0), r1 = v1 - __bfloat162float(h1);
  __nv_bfloat162 H, L;
  H.x = h0; H.y = h1;
  L = __floats2bfloat162_rn(r0, r1);
  hi = *reinterpret_cast<uint32_t*>(&H);
  lo = *reinterpret_cast<uint32_t*>(&L);
}

// tanh(x) = 1 - 2 / (exp(2x) + 1): two MUFU ops, absolute error ~1e-7 (the products carry ~1.5e-5 already)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

constexpr int kWarps = 8;
constexpr int kMaxN = 256;      // nodes per graph (per-warp score row in shared memory)

template <int D, int DO>
__global__ void __launch_bounds__(kWarps * 32)
gat_rows_mma_kernel(const GraphView x, int n1, const GatRowWeights w, float* __restrict__ out, long long out_bs,
                    const float* __restrict__ master_in, long long master_stride, const GatRowWeights wM,
                    float* __restrict__ master_out, int n_iblocks) {
  constexpr int KS = D / 16, NT = DO / 8;
  constexpr int XS = D + 8;           // row stride of the graph copy: 8 mod 32 words -> conflict-free 64-bit fragment loads
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = x.n;
  uint2* sBhi = reinterpret_cast<uint2*>(smem_raw);            // [NT*KS][32] fragment-ordered W (hi)
  uint2* sBlo = sBhi + NT * KS * 32;                            // (lo)
  float* sa = reinterpret_cast<float*>(sBlo + NT * KS * 32);    // [3][DO]
  float* sb = sa + 3 * DO;                                      // [DO]
  float* swarp = sb + DO;                                       // per warp: se[kMaxN], sax[D], sv[D]
  float* sx = swarp + kWarps * (kMaxN + 2 * D);                 // [n][XS]
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, b = blockIdx.y;
  const bool is_master = (int)blockIdx.x == n_iblocks;
  const GatRowWeights& ww = is_master ? wM : w;
  const float* xb = x.ptr + (long long)b * x.batch_stride;

  for (int k = t; k < n * (D / 4); k += kWarps * 32) {
    const int r = k / (D / 4), c4 = k % (D / 4);
    *reinterpret_cast<float4*>(sx + r * XS + c4 * 4) = *reinterpret_cast<const float4*>(xb + (long long)r * D + c4 * 4);
  }
  {
    __nv_bfloat16* bh = reinterpret_cast<__nv_bfloat16*>(sBhi);
    __nv_bfloat16* bl = reinterpret_cast<__nv_bfloat16*>(sBlo);
    for (int k = t; k < DO * D; k += kWarps * 32) {
      const int o = k / D, kk = k % D;
      const float v = ww.att_w[k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      // B fragment of m16n8k16 (col-major B = row-major W): lane = (o % 8) * 4 + (k % 8) / 2, regs {k%16 < 8, >= 8}
      const int frag = (o / 8) * KS + kk / 16, kl = kk % 16;
      const int ln = (o % 8) * 4 + (kl % 8) / 2;
      const int e = ((frag * 32 + ln) * 2 + kl / 8) * 2 + (kl & 1);
      bh[e] = h;
      bl[e] = l;
    }
  }
  for (int k = t; k < DO; k += kWarps * 32) {
    sa[k] = ww.a11[k];
    sa[DO + k] = ww.a22 ? ww.a22[k] : 0.f;
    sa[2 * DO + k] = ww.a12 ? ww.a12[k] : 0.f;
    sb[k] = ww.att_b[k];
  }
  __syncthreads();

  const int i = is_master ? n : (int)blockIdx.x * kWarps + wid;
  if (is_master ? wid != 0 : i >= n) return;
  float* se = swarp + wid * (kMaxN + 2 * D);
  float* sax = se + kMaxN;
  float* sv = sax + D;
  for (int d = lane; d < D; d += 32) sv[d] = is_master ? master_in[(long long)b * master_stride + d] : sx[i * XS + d];
  __syncwarp();

  const int g = lane >> 2, c2 = (lane & 3) * 2;
  float xi[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    xi[ks][0] = sv[ks * 16 + c2];
    xi[ks][1] = sv[ks * 16 + c2 + 1];
    xi[ks][2] = sv[ks * 16 + c2 + 8];
    xi[ks][3] = sv[ks * 16 + c2 + 9];
  }
  const bool i_first = i < n1;
  const int ntile = (n + 15) >> 4;
  for (int jt = 0; jt < ntile; jt += 2) {
    const bool two = jt + 1 < ntile;
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[m][nt][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m == 1 && !two) continue;
        const int ja = min((jt + m) * 16 + g, n - 1), jb = min((jt + m) * 16 + g + 8, n - 1);
        const float2 a0 = *reinterpret_cast<const float2*>(sx + ja * XS + ks * 16 + c2);
        const float2 a2 = *reinterpret_cast<const float2*>(sx + ja * XS + ks * 16 + c2 + 8);
        const float2 b0 = *reinterpret_cast<const float2*>(sx + jb * XS + ks * 16 + c2);
        const float2 b2 = *reinterpret_cast<const float2*>(sx + jb * XS + ks * 16 + c2 + 8);
        split2(xi[ks][0] * a0.x, xi[ks][1] * a0.y, ah[m][0], al[m][0]);
        split2(xi[ks][0] * b0.x, xi[ks][1] * b0.y, ah[m][1], al[m][1]);
        split2(xi[ks][2] * a2.x, xi[ks][3] * a2.y, ah[m][2], al[m][2]);
        split2(xi[ks][2] * b2.x, xi[ks][3] * b2.y, ah[m][3], al[m][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint2 bh = sBhi[(nt * KS + ks) * 32 + lane];
        const uint2 bl = sBlo[(nt * KS + ks) * 32 + lane];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          if (m == 1 && !two) continue;
          mma_bf16(acc[m][nt], ah[m], bh.x, bh.y);
          mma_bf16(acc[m][nt], al[m], bh.x, bh.y);
          mma_bf16(acc[m][nt], ah[m], bl.x, bl.y);
        }
      }
    }
    // e_ij = a(i,j) . tanh(acc + b) / temp for the two rows this thread holds in each tile
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (m == 1 && !two) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = (jt + m) * 16 + g + 8 * h;
        const float* av = sa + (is_master ? 0 : (i_first == (j < n1) ? (i_first ? 0 : DO) : 2 * DO));
        float e = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float2 a2 = *reinterpret_cast<const float2*>(av + nt * 8 + c2);
          const float2 b2 = *reinterpret_cast<const float2*>(sb + nt * 8 + c2);
          e = fmaf(a2.x, tanh_fast(acc[m][nt][2 * h] + b2.x), e);
          e = fmaf(a2.y, tanh_fast(acc[m][nt][2 * h + 1] + b2.y), e);
        }
        e += __shfl_xor_sync(0xffffffffu, e, 1);
        e += __shfl_xor_sync(0xffffffffu, e, 2);
        if ((lane & 3) == 0 && j < n) se[j] = e * ww.inv_temp;
      }
    }
  }
  __syncwarp();
  {  // softmax over j
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, se[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float ev = expf(se[j] - mx);
      se[j] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < n; j += 32) se[j] *= inv;
  }
  __syncwarp();
  for (int d = lane; d < D; d += 32) {
    float a = 0.f;
    for (int j = 0; j < n; ++j) a = fmaf(se[j], sx[j * XS + d], a);
    sax[d] = a;
  }
  __syncwarp();
  for (int o = lane; o < DO; o += 32) {
    float y = ww.with_b[o] + ww.without_b[o];
    float y2 = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      y = fmaf(ww.with_t[d * DO + o], sax[d], y);
      y2 = fmaf(ww.without_t[d * DO + o], sv[d], y2);
    }
    y += y2;
    if (is_master) master_out[(long long)b * DO + o] = y;
    else out[(long long)b * out_bs + (long long)i * DO + o] = selu_f(y * ww.bn_s[o] + ww.bn_t[o]);
  }
}

template <int D, int DO>
int launch(cudaStream_t s, const GraphView& x, int B, int n1, const GatRowWeights& w, float* out, long long out_bs,
           const float* master_in, long long master_stride, const GatRowWeights* wM, float* master_out) {
  constexpr int KS = D / 16, NT = DO / 8;
  const size_t smem = (size_t)2 * NT * KS * 32 * sizeof(uint2) + (size_t)4 * DO * sizeof(float) +
                      (size_t)kWarps * (kMaxN + 2 * D) * sizeof(float) + (size_t)x.n * (D + 8) * sizeof(float);
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&gat_rows_mma_kernel<D, DO>), (size_t)smem));
  const bool has_master = master_in != nullptr;
  const int n_iblocks = ceil_div(x.n, kWarps);
  dim3 grid(n_iblocks + (has_master ? 1 : 0), B);
  gat_rows_mma_kernel<D, DO><<<grid, kWarps * 32, smem, s>>>(x, n1, w, out, out_bs, master_in, master_stride,
                                                             has_master ? *wM : w, master_out, n_iblocks);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace

int aasist_gat_rows_mma(cudaStream_t s, int D, int DO, const GraphView& x, int B, int n1, const GatRowWeights& w,
                        float* out, long long out_batch_stride, const float* master_in, long long master_stride,
                        const GatRowWeights* wM, float* master_out) {
  RTDF_REQUIRE(x.ptr && out && x.n >= 1 && x.n <= kMaxN && B > 0 && B <= 65535, "gat_rows_mma: bad arguments (n = %d)", x.n);
  RTDF_REQUIRE(!master_in || (wM && master_out), "gat_rows_mma: master row needs weights and an output");
  RTDF_REQUIRE((reinterpret_cast<uintptr_t>(x.ptr) & 15) == 0 && (x.batch_stride % 4) == 0, "gat_rows_mma: unaligned graph");
  if (D == 64 && DO == 64) return launch<64, 64>(s, x, B, n1, w, out, out_batch_stride, master_in, master_stride, wM, master_out);
  if (D == 64 && DO == 32) return launch<64, 32>(s, x, B, n1, w, out, out_batch_stride, master_in, master_stride, wM, master_out);
  if (D == 32 && DO == 32) return launch<32, 32>(s, x, B, n1, w, out, out_batch_stride, master_in, master_stride, wM, master_out);
  set_error("gat_rows_mma: unsupported dims %d -> %d", D, DO);
  return RTDF_ERR_UNSUPPORTED;
}

}  // namespace rtdf
