// Conformer MHSA with Shaw relative-position bias on the tensor cores (bf16 mode).
// lucidrains conformer.Attention as used by reference models/conformer_baseline.py:16-18:
//   dots = (q k^T) * scale + (q . rel_pos_emb[clamp(i - j, +-512) + 512]) * scale ; softmax_j ; out = attn v
// with heads x dim_head = 4 x 36 and n = T + 1 tokens (class token first).
//
// CTA = (utterance, head, half of the query tiles), 8 warps, one 16-query tile per warp at a time.  K, the window of
// the relative-position table the sequence can reach (2n - 1 rows) and V^T sit in shared memory as bf16 B operands;
// the three contractions are mma.sync m16n8k16 (+ one k8 step: 36 = 16 + 16 + 4, padded to 8):
//   S  = Q K^T                 (16 x n)
//   QE = Q E^T                 (16 x (n + 15)): row i of the tile needs E rows i - j + n - 1, j = 0..n-1; the tile's
//                               rows together span n + 15 consecutive table rows.  QE goes through a per-warp smem
//                               tile and comes back skewed: pos[i][j] = QE[i][(i - i0) + n - 1 - j].
//   O  = softmax(S + pos) V    P stays in registers: the fp32 accumulator fragment of two adjacent 8-column tiles IS
//                               the bf16 A fragment of the next k16 step.
// The SIMT kernel this replaces (one warp per query, fp32 FMAs fed by three LDS per two FMAs) took 624 us per block at
// B = 64 for 0.2 GFLOP.
#include "conformer.cuh"

namespace rtdf {

namespace {

constexpr int kDP = 40;        // head dim padded to 16 + 16 + 8; smem row stride of K / E in bf16 (80 B: conflict-free)
constexpr int kWarps = 8;

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

// one 16 x 8 output tile of A (16 x 40, fragments in registers) times rows [r0, r0 + 8) of a [rows][kDP] bf16 matrix
__device__ __forceinline__ void qk_tile(float (&c)[4], const uint32_t (&qa)[10], const bf16* __restrict__ m, int r0,
                                        int g, int c2) {
  const uint32_t* row = reinterpret_cast<const uint32_t*>(m + (size_t)(r0 + g) * kDP);
  c[0] = c[1] = c[2] = c[3] = 0.f;
  mma16816(c, qa[0], qa[1], qa[2], qa[3], row[c2 >> 1], row[(c2 >> 1) + 4]);
  mma16816(c, qa[4], qa[5], qa[6], qa[7], row[8 + (c2 >> 1)], row[12 + (c2 >> 1)]);
  mma1688(c, qa[8], qa[9], row[16 + (c2 >> 1)]);
}

template <int NT>   // NT = number of 8-key tiles (n <= 8 NT), NT even
__global__ void __launch_bounds__(kWarps * 32)
conformer_attn_mma_kernel(const bf16* __restrict__ qkv, const float* __restrict__ rel, bf16* __restrict__ out, int n,
                          int heads, int dh, float scale, int q_splits) {
  constexpr int NK = NT * 8;              // padded key count
  constexpr int VS = NK + 8;              // V^T row stride (bf16): (NK + 8) / 2 words = 12 mod 32 for NK = 208 -> conflict-free
  constexpr int QES = (NK + 16 + 31) / 32 * 32 + 8;   // QE tile row stride (fp32 words), 8 mod 32: conflict-free float2 stores
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);              // [NK][kDP]
  bf16* sE = sK + NK * kDP;                                  // [2 NK + 16][kDP]   rows r = i - j + n - 1 (+ slack)
  bf16* sVt = sE + (2 * NK + 16) * kDP;                      // [kDP][VS]
  float* sQE = reinterpret_cast<float*>(sVt + kDP * VS);     // [kWarps][16][QES]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int bh = blockIdx.x / q_splits, split = blockIdx.x % q_splits;
  const int b = bh / heads, h = bh % heads;
  const int E = heads * dh, ld = 3 * E;
  const bf16* base = qkv + (long long)b * n * ld;
  const bf16 zero = __float2bfloat16_rn(0.f);
  // ---- stage K, E window, V^T (zero padded) -------------------------------------------------------------------
  for (int i = t; i < NK * kDP; i += kWarps * 32) {
    const int j = i / kDP, d = i % kDP;
    sK[i] = (j < n && d < dh) ? base[(long long)j * ld + E + h * dh + d] : zero;
  }
  for (int i = t; i < (2 * NK + 16) * kDP; i += kWarps * 32) {
    const int r = i / kDP, d = i % kDP;
    int relpos = r - (n - 1);
    relpos = max(-512, min(512, relpos)) + 512;
    sE[i] = (r < 2 * n - 1 && d < dh) ? __float2bfloat16_rn(rel[relpos * dh + d]) : zero;
  }
  for (int i = t; i < kDP * NK; i += kWarps * 32) {
    const int j = i / kDP, d = i % kDP;     // read row-major (coalesced), write transposed
    sVt[d * VS + j] = (j < n && d < dh) ? base[(long long)j * ld + 2 * E + h * dh + d] : zero;
  }
  __syncthreads();

  const int g = lane >> 2, c2 = (lane & 3) * 2;
  float* qe = sQE + warp * 16 * QES;
  const int n_qt = (n + 15) >> 4;
  for (int qt = split * kWarps + warp; qt < n_qt; qt += q_splits * kWarps) {
    const int i0 = qt * 16;
    // ---- Q fragments (rows i0 + g, i0 + g + 8; k = 0..39) from global -------------------------------------------
    uint32_t qa[10];
    {
      const int ra = min(i0 + g, n - 1), rb = min(i0 + g + 8, n - 1);
      const bf16* qra = base + (long long)ra * ld + h * dh;
      const bf16* qrb = base + (long long)rb * ld + h * dh;
      auto ld2 = [&](const bf16* p, int k) -> uint32_t {     // elements k, k+1 (k even); zero beyond dh
        return k < dh ? *reinterpret_cast<const uint32_t*>(p + k) : 0u;
      };
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        qa[4 * ks + 0] = ld2(qra, ks * 16 + c2);
        qa[4 * ks + 1] = ld2(qrb, ks * 16 + c2);
        qa[4 * ks + 2] = ld2(qra, ks * 16 + c2 + 8);
        qa[4 * ks + 3] = ld2(qrb, ks * 16 + c2 + 8);
      }
      qa[8] = ld2(qra, 32 + c2);
      qa[9] = ld2(qrb, 32 + c2);
    }
    // ---- QE tile: table rows r = i0 + c, c = 0 .. NK + 15  ->  smem -------------------------------------------
#pragma unroll 1
    for (int nt = 0; nt < NT + 2; ++nt) {
      float c[4];
      qk_tile(c, qa, sE, i0 + nt * 8, g, c2);
      *reinterpret_cast<float2*>(qe + g * QES + nt * 8 + c2) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(qe + (g + 8) * QES + nt * 8 + c2) = make_float2(c[2], c[3]);
    }
    __syncwarp();
    // ---- S = (Q K^T + skewed QE) * scale, masked ------------------------------------------------------------------
    float s[NT][4];
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      qk_tile(s[nt], qa, sK, nt * 8, g, c2);
      const int j = nt * 8 + c2;
      // pos[i][j] = QE[i][(i - i0) + n - 1 - j]
      const float* pa = qe + g * QES + g + n - 1 - j;
      const float* pb = qe + (g + 8) * QES + g + 8 + n - 1 - j;
      s[nt][0] = j < n ? (s[nt][0] + pa[0]) * scale : -INFINITY;
      s[nt][1] = j + 1 < n ? (s[nt][1] + pa[-1]) * scale : -INFINITY;
      s[nt][2] = j < n ? (s[nt][2] + pb[0]) * scale : -INFINITY;
      s[nt][3] = j + 1 < n ? (s[nt][3] + pb[-1]) * scale : -INFINITY;
      mxa = fmaxf(mxa, fmaxf(s[nt][0], s[nt][1]));
      mxb = fmaxf(mxb, fmaxf(s[nt][2], s[nt][3]));
    }
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1));
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
    // ---- P = exp(S - max) as bf16 A fragments; row sums of the rounded values --------------------------------------
    const float kL2e = 1.4426950408889634f;
    const float ma = mxa * kL2e, mb = mxb * kL2e;
    float suma = 0.f, sumb = 0.f;
    uint32_t pfrag[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const __nv_bfloat162 pa = __floats2bfloat162_rn(ex2_approx(fmaf(s[nt][0], kL2e, -ma)), ex2_approx(fmaf(s[nt][1], kL2e, -ma)));
      const __nv_bfloat162 pb = __floats2bfloat162_rn(ex2_approx(fmaf(s[nt][2], kL2e, -mb)), ex2_approx(fmaf(s[nt][3], kL2e, -mb)));
      suma += __low2float(pa) + __high2float(pa);
      sumb += __low2float(pb) + __high2float(pb);
      pfrag[nt][0] = *reinterpret_cast<const uint32_t*>(&pa);
      pfrag[nt][1] = *reinterpret_cast<const uint32_t*>(&pb);
    }
    suma += __shfl_xor_sync(0xffffffffu, suma, 1);
    suma += __shfl_xor_sync(0xffffffffu, suma, 2);
    sumb += __shfl_xor_sync(0xffffffffu, sumb, 1);
    sumb += __shfl_xor_sync(0xffffffffu, sumb, 2);
    // ---- O = P V ----------------------------------------------------------------------------------------------------
    float o[kDP / 8][4];
#pragma unroll
    for (int dt = 0; dt < kDP / 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks) {
#pragma unroll
      for (int dt = 0; dt < kDP / 8; ++dt) {
        const uint32_t* vrow = reinterpret_cast<const uint32_t*>(sVt + (size_t)(dt * 8 + g) * VS + ks * 16);
        mma16816(o[dt], pfrag[2 * ks][0], pfrag[2 * ks][1], pfrag[2 * ks + 1][0], pfrag[2 * ks + 1][1], vrow[c2 >> 1],
                 vrow[(c2 >> 1) + 4]);
      }
    }
    const float inva = 1.0f / suma, invb = 1.0f / sumb;
    const int ia = i0 + g, ib = i0 + g + 8;
#pragma unroll
    for (int dt = 0; dt < kDP / 8; ++dt) {
      const int d = dt * 8 + c2;
      if (d < dh) {
        if (ia < n) *reinterpret_cast<uint32_t*>(out + ((long long)b * n + ia) * E + h * dh + d) = pack_bf16x2(o[dt][0] * inva, o[dt][1] * inva);
        if (ib < n) *reinterpret_cast<uint32_t*>(out + ((long long)b * n + ib) * E + h * dh + d) = pack_bf16x2(o[dt][2] * invb, o[dt][3] * invb);
      }
    }
    __syncwarp();       // the QE tile is rewritten by the next query tile
  }
}

template <int NT>
int launch(cudaStream_t s, const bf16* qkv, const float* rel, bf16* out, int B, int n, int heads, int dh) {
  constexpr int NK = NT * 8, VS = NK + 8, QES = (NK + 16 + 31) / 32 * 32 + 8;
  const size_t smem = (size_t)(NK * kDP + (2 * NK + 16) * kDP + kDP * VS) * sizeof(bf16) + (size_t)kWarps * 16 * QES * sizeof(float);
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&conformer_attn_mma_kernel<NT>), (size_t)smem));
  const int n_qt = (n + 15) / 16;
  const int q_splits = n_qt > kWarps ? 2 : 1;
  conformer_attn_mma_kernel<NT><<<B * heads * q_splits, kWarps * 32, smem, s>>>(qkv, rel, out, n, heads, dh,
                                                                               1.0f / sqrtf((float)dh), q_splits);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace

// n <= 208 tokens, dh <= 40 (even), (heads * dh) even.  Returns RTDF_ERR_UNSUPPORTED outside that envelope (the caller
// falls back to the SIMT kernel).
int conformer_attention_mma(cudaStream_t s, const bf16* qkv, const float* rel_pos, bf16* out, int B, int n, int heads, int dh) {
  RTDF_REQUIRE(qkv && rel_pos && out && B > 0 && n > 0, "conformer_attention_mma: bad arguments");
  if (dh > kDP || (dh & 1) || n > 208) return RTDF_ERR_UNSUPPORTED;
  if (n <= 64) return launch<8>(s, qkv, rel_pos, out, B, n, heads, dh);
  if (n <= 112) return launch<14>(s, qkv, rel_pos, out, B, n, heads, dh);
  return launch<26>(s, qkv, rel_pos, out, B, n, heads, dh);
}

}  // namespace rtdf
