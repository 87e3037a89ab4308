// The transformer layers of a streaming chunk as one persistent kernel (bf16 mode, <= 64 frames in flight).
//
// Reference path: models/fe.py:17-21 (XLSR_FE.extract_feat -> fairseq Wav2Vec2Model encoder, 24 pre-LN layers + the final
// LayerNorm).  At batch 1 x 1 s of audio the layers are 49 rows against 604 MB of bf16 weights: a weight-streaming problem
// whose cost as separate kernels is the ~170 launch boundaries, not the bytes.  Here 128 co-resident CTAs walk the phases
//
//     LN1 | QKV | attention | out_proj | +x, LN2 | fc1 + GELU | fc2 (K split in 4) | +x, LN1 of the next layer | ...
//
// with a grid-wide barrier (one release-add + acquire-poll on a global word, in an arrive and a wait half) between them.
// Everything is summed in a fixed order (accumulators in index order, fc2's four K chunks through global partials in chunk
// order): replays are bit-identical.  A barrier that never completes (a CTA that is not resident) makes the kernel return and
// set a host-visible flag; the next forward call raises.
//
// Two variants of the GEMM phases (a CTA = an 8 - 32 column slice of the output over K = 1024):
//   * layer_stack_tc_kernel (default, second half of this file): TMA boxes + tcgen05 MMAs, mma.sync attention;
//   * layer_stack_mma_kernel (RTDF_STACK_IMPL=mma, the first version, kept for A/B timing): the activations are copied to
//     shared memory in mma.sync A-fragment order with cp.async, the weights never touch shared memory -- every thread loads
//     its B fragments for the whole phase as 16-byte global loads (8 consecutive k of one weight row; the k permutation inside
//     a 32-wide block is applied to A and B alike, so the products pair up correctly) between the two halves of the barrier in
//     front of the phase.  The 8 warps split K; their partial tiles are summed through shared memory in warp order.
//     Attention is fp32 SIMT: (utterance, head, query split) items, one query row per warp at a time.
// Measurements of every step: profiles/r02_layer_stack.txt.
#include "layer_stack.cuh"

#include "ptx.cuh"
#include "tma_host.h"

namespace rtdf {

namespace {

constexpr int kCtas = 128;
constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kABytes = 32 * 4 * 2 * 512;              // A operand: [k-block 32][m-tile 4][row half 2][lane 32] x 16 B
constexpr int kRedStride = 40;                         // floats per row of a warp's partial tile (8 mod 32: conflict-free float2)
constexpr int kRedBytes = kWarps * 64 * kRedStride * 4;
constexpr int kSmemBytes = kABytes + kRedBytes;        // 212,992
constexpr unsigned kSpinLimit = 1u << 24;              // barrier polls before the kernel gives up (seconds)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {   // weights: read once, keep out of L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier in two halves: every CTA adds 1 to the word (arrive), then polls it until all gridDim.x additions of this
// round are in (wait).  Between the halves the threads issue the weight loads of the next GEMM phase, so the HBM stream runs
// under the barrier.  wait: false = gave up (a CTA never arrived: the grid was not co-resident); the kernel then returns and
// the host finds the flag.
__device__ __forceinline__ void barrier_arrive(const StackParams& p, unsigned& target, unsigned long long* tr = nullptr) {
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (tr) tr[1] = clock64();
    target += gridDim.x;
    red_release_gpu_add(p.sync, 1u);
    if (tr) tr[2] = clock64();
  }
}
__device__ __forceinline__ bool barrier_wait(const StackParams& p, unsigned target, volatile int* s_abort,
                                             unsigned long long* tr = nullptr) {
  if (threadIdx.x == 0) {
    if (tr) tr[3] = clock64();
    unsigned spins = 0;
    while (ld_acquire_gpu(p.sync) < target) {
      if (++spins > kSpinLimit) {
        *s_abort = 1;
        *reinterpret_cast<volatile int*>(p.fault) = 1;
        __threadfence_system();
        break;
      }
    }
    if (tr) {
      tr[4] = clock64();
      tr[6] = spins;
    }
  }
  __syncthreads();
  if (tr && threadIdx.x == 0) tr[5] = clock64();
  return *s_abort == 0;
}

// ---- GEMM phase pieces ------------------------------------------------------------------------------------------------
// B fragments of this thread for the whole phase: warp w owns k-blocks [4w, 4w + 4) of the CTA's 1024-wide K range,
// thread (g, t) of the warp the 8 values k = 32 kb + 8 t .. + 7 of weight row n0 + 8 nt + g.
template <int NT>
__device__ __forceinline__ void load_w(uint4 (&w)[4][NT], const bf16* __restrict__ W, int ldw, int n0, int kbase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const bf16* ptr = W + (size_t)(n0 + g) * ldw + kbase + warp * 128 + t * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) w[i][nt] = ldg_stream_v4(ptr + (size_t)nt * 8 * ldw + i * 32);
}

// rows [0, R) x 1024 columns of A (bf16, row stride lda) -> shared memory in fragment order: the 16-byte piece
// (row, k = 32 kb + 8 t ..) lands where lane 4 (row % 8) + t reads it for m-tile row / 16, half (row / 8) % 2, i.e. a k-block
// is 64 rows x 64 B, rows 64 B apart.  (A TMA box of 32 columns x 64 rows writes exactly this, but 64-byte box rows run at half
// the speed of the cp.async loop: 4,600 - 5,800 against 2,600 - 3,300 cycles per phase, profiles/r02_layer_stack.txt.)
__device__ __forceinline__ void load_a_cp(unsigned char* sA, const bf16* A, int lda, int R) {
  const uint32_t base = ptx::smem_u32(sA);
  for (int q = threadIdx.x; q < R * 128; q += kThreads) {
    const int row = q >> 7, c16 = q & 127;
    const int kb = c16 >> 2, t = c16 & 3;
    cp_async_16(base + (kb * 64 + row) * 64 + t * 16, A + (size_t)row * lda + c16 * 8);
  }
  cp_async_wait_all();
  __syncthreads();
}
template <int NT>
__device__ __forceinline__ void mma_phase(float (&acc)[4][NT][4], const uint4 (&w)[4][NT], const unsigned char* sA, int mtiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kb = warp * 4 + i;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt < mtiles) {
        const uint4 lo = *reinterpret_cast<const uint4*>(sA + ((((kb * 4 + mt) * 2 + 0) * 32) + lane) * 16);   // row g
        const uint4 hi = *reinterpret_cast<const uint4*>(sA + ((((kb * 4 + mt) * 2 + 1) * 32) + lane) * 16);   // row g + 8
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          mma16816(acc[mt][nt], lo.x, hi.x, lo.y, hi.y, w[i][nt].x, w[i][nt].y);
          mma16816(acc[mt][nt], lo.z, hi.z, lo.w, hi.w, w[i][nt].z, w[i][nt].w);
        }
      }
    }
  }
}

// sum the 8 warps' partial tiles (warp order) and hand (row, col, value) to the phase's epilogue
template <int NT, class Epi>
__device__ __forceinline__ void reduce_phase(const float (&acc)[4][NT][4], float* red, int mtiles, int R, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    if (mt < mtiles) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float* q = red + (warp * 64 + mt * 16 + g) * kRedStride + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(q) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<float2*>(q + 8 * kRedStride) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
    }
  }
  __syncthreads();
  constexpr int COLS = NT * 8;
  for (int e = threadIdx.x; e < R * COLS; e += kThreads) {
    const int row = e / COLS, col = e % COLS;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) v += red[(w * 64 + row) * kRedStride + col];
    epi(row, col, v);
  }
}

// ---- LayerNorm phase: one row per CTA ------------------------------------------------------------------------------------
// v = x[row] (+ bias + sum of np partials, in order); x[row] = v (if write_x); LayerNorm(v) -> xn (bf16) or feats (fp32)
struct LnWeights { float4 bias, g, b; };   // this thread's 4 columns of the projection bias and the LayerNorm affine
__device__ __forceinline__ LnWeights ln_weights(const float* __restrict__ bias, const float* __restrict__ gam, const float* __restrict__ bet) {
  const int c = threadIdx.x * 4;
  LnWeights w;
  w.bias = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  w.g = __ldg(reinterpret_cast<const float4*>(gam + c));
  w.b = __ldg(reinterpret_cast<const float4*>(bet + c));
  return w;
}
__device__ __forceinline__ void ln_row(const StackParams& p, int row, const LnWeights& lw, int np, bool write_x, bf16* xn_out,
                                       float* f_out, float* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = tid * 4;
  float4 v = __ldcg(reinterpret_cast<const float4*>(p.x + (size_t)row * 1024 + c));
  v.x += lw.bias.x; v.y += lw.bias.y; v.z += lw.bias.z; v.w += lw.bias.w;
  for (int q = 0; q < np; ++q) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p.part + ((size_t)q * p.R + row) * 1024 + c));
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  float s = warp_sum((v.x + v.y) + (v.z + v.w));
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) tot += s_red[w];
  const float mean = tot * (1.0f / 1024.0f);
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  float sq = warp_sum((dx * dx + dy * dy) + (dz * dz + dw * dw));
  if (lane == 0) s_red[kWarps + warp] = sq;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) var += s_red[kWarps + w];
  const float rstd = rsqrtf(var * (1.0f / 1024.0f) + 1e-5f);
  const float4 gg = lw.g, bb = lw.b;
  const float y0 = dx * rstd * gg.x + bb.x, y1 = dy * rstd * gg.y + bb.y, y2 = dz * rstd * gg.z + bb.z, y3 = dw * rstd * gg.w + bb.w;
  if (write_x) __stcg(reinterpret_cast<float4*>(p.x + (size_t)row * 1024 + c), v);
  if (xn_out) {
    uint2 o;
    o.x = pack_bf16x2(y0, y1);
    o.y = pack_bf16x2(y2, y3);
    __stcg(reinterpret_cast<uint2*>(xn_out + (size_t)row * 1024 + c), o);
  }
  if (f_out) __stcg(reinterpret_cast<float4*>(f_out + (size_t)row * 1024 + c), make_float4(y0, y1, y2, y3));
}

__device__ __forceinline__ void ln_row(const StackParams& p, int row, const float* __restrict__ bias, int np,
                                       const float* __restrict__ gam, const float* __restrict__ bet, bool write_x, bf16* xn_out,
                                       float* f_out, float* s_red) {
  ln_row(p, row, ln_weights(bias, gam, bet), np, write_x, xn_out, f_out, s_red);
}

// ---- attention phase (<= 64 keys per utterance): fp32 SIMT ------------------------------------------------------------------
__device__ __forceinline__ void unpack8(float* dst, const uint4& raw) {
  const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    dst[2 * e] = __uint_as_float(r[e] << 16);
    dst[2 * e + 1] = __uint_as_float(r[e] & 0xffff0000u);
  }
}

__device__ __forceinline__ void attention_phase(const StackParams& p, float* scratch, unsigned long long* tr = nullptr) {
  float* Ks = scratch;              // [64][65]
  float* Vs = Ks + 64 * 65;         // [64][64]
  float* Qs = Vs + 64 * 64;         // [64][64] the item's query rows
  float* Ps = Qs + 64 * 64;         // [8][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T, n_bh = p.B * 16;
  int qsplit = (int)gridDim.x / n_bh;
  qsplit = qsplit < 1 ? 1 : (qsplit > T ? T : qsplit);
  const int rows_per = (T + qsplit - 1) / qsplit;
  const int items = n_bh * qsplit;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int bh = it / qsplit, qs = it % qsplit, b = bh >> 4, h = bh & 15;
    const int q0 = qs * rows_per, q1 = min(T, q0 + rows_per);
    if (q0 >= q1) continue;
    const int nq = q1 - q0;
    __syncthreads();   // the previous item's readers are done with the scratch
    const bf16* base = p.qkv + (size_t)b * T * 3072 + h * 64;
    // K rows, V rows, then the query rows, 8 pieces of 16 B each: at most (2 * 64 + 64) * 8 / 256 = 6 pieces per thread, all
    // loads in flight before the first unpack
    constexpr int kPieces = 6;
    const int n_pieces = (2 * T + nq) * 8;
    if (tr && tid == 0) tr[0] = clock64();
    uint4 raw[kPieces];
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int q = tid + i * kThreads;
      if (q < n_pieces) {
        const int rr = q >> 3, c8 = q & 7;
        const int sel = rr < T ? 1 : (rr < 2 * T ? 2 : 0), j = sel == 1 ? rr : (sel == 2 ? rr - T : q0 + rr - 2 * T);
        raw[i] = __ldcg(reinterpret_cast<const uint4*>(base + (size_t)j * 3072 + sel * 1024 + c8 * 8));
      }
    }
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int q = tid + i * kThreads;
      if (q < n_pieces) {
        const int rr = q >> 3, c8 = q & 7;
        const int sel = rr < T ? 1 : (rr < 2 * T ? 2 : 0), j = sel == 1 ? rr : (sel == 2 ? rr - T : q0 + rr - 2 * T);
        unpack8(sel == 1 ? Ks + j * 65 + c8 * 8 : (sel == 2 ? Vs + j * 64 + c8 * 8 : Qs + (j - q0) * 64 + c8 * 8), raw[i]);
      }
    }
    __syncthreads();
    if (tr && tid == 0) tr[1] = clock64();
    float* prow = Ps + warp * 64;
    for (int r = warp; r < nq; r += kWarps) {
      const float* qrow = Qs + r * 64;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) {
        const float qd = qrow[d];
        s0 = fmaf(qd, Ks[lane * 65 + d], s0);
        s1 = fmaf(qd, Ks[(lane + 32) * 65 + d], s1);
      }
      s0 = lane < T ? s0 : -INFINITY;
      s1 = lane + 32 < T ? s1 : -INFINITY;
      const float mx = warp_max(fmaxf(s0, s1));
      const float e0 = lane < T ? expf(s0 - mx) : 0.f, e1 = lane + 32 < T ? expf(s1 - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      prow[lane] = e0 * inv;
      prow[lane + 32] = e1 * inv;
      __syncwarp();
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      int j = 0;
      for (; j + 1 < T; j += 2) {
        const float pa = prow[j], pb = prow[j + 1];
        o0 = fmaf(pa, Vs[j * 64 + lane], o0);
        o1 = fmaf(pa, Vs[j * 64 + lane + 32], o1);
        o2 = fmaf(pb, Vs[(j + 1) * 64 + lane], o2);
        o3 = fmaf(pb, Vs[(j + 1) * 64 + lane + 32], o3);
      }
      if (j < T) {
        const float pa = prow[j];
        o0 = fmaf(pa, Vs[j * 64 + lane], o0);
        o1 = fmaf(pa, Vs[j * 64 + lane + 32], o1);
      }
      bf16* out = p.att + ((size_t)b * T + q0 + r) * 1024 + h * 64;
      out[lane] = __float2bfloat16_rn(o0 + o2);
      out[lane + 32] = __float2bfloat16_rn(o1 + o3);
      __syncwarp();
    }
    if (tr && tid == 0) tr[2] = clock64();
  }
}

// barrier with the statement(s) `prefetch` (weight loads of the phase behind it) issued between its arrive and wait halves
#define STACK_BARRIER(prefetch)                            \
  do {                                                     \
    barrier_arrive(p, target);                             \
    prefetch;                                              \
    if (!barrier_wait(p, target, &s_abort)) return;        \
  } while (0)
// RTDF_STACK_TRACE=1 (debug): SM clock of CTA 0 / CTA 127 at the phase boundaries of layer 1
#define STAMP(i)                                                                                   \
  do {                                                                                             \
    if (p.trace && l == 1 && tid == 0 && (cta == 0 || cta == 127)) p.trace[(cta ? 32 : 0) + (i)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(kThreads, 1)
layer_stack_mma_kernel(const StackParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;
  float* red = reinterpret_cast<float*>(smem + kABytes);
  __shared__ float s_red[2 * kWarps];
  __shared__ int s_abort;
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int R = p.R, mtiles = (R + 15) >> 4;
  if (tid == 0) s_abort = 0;
  for (int i = tid; i < kABytes / 16; i += kThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);   // rows >= R stay zero
  unsigned target = 0;
  const StackLayer* __restrict__ L = p.layers;

  if (cta < R) ln_row(p, cta, nullptr, 0, L[0].g1, L[0].be1, false, p.xn, nullptr, s_red);

  for (int l = 0; l < p.n_layers; ++l) {
    const StackLayer W = L[l];
    const bool last = l + 1 == p.n_layers;
    uint4 wq[4][3];
    uint4 wo[4][1];
    uint4 w1[4][4];
    uint4 w2[4][4];
    // ---- q, k, v = LN1(x) Wqkv^T + b: 24 columns per CTA
    STAMP(0);
    STACK_BARRIER(load_w<3>(wq, W.wqkv, 1024, cta * 24, 0));
    STAMP(1);
    load_a_cp(sA, p.xn, 1024, R);
    STAMP(2);
    {
      float acc[4][3][4];
      mma_phase<3>(acc, wq, sA, mtiles);
      STAMP(3);
      reduce_phase<3>(acc, red, mtiles, R, [&](int row, int col, float v) {
        const int n = cta * 24 + col;
        p.qkv[(size_t)row * 3072 + n] = __float2bfloat16_rn(v + __ldg(W.bqkv + n));
      });
    }
    STAMP(4);
    STACK_BARRIER((void)0);
    STAMP(5);
    attention_phase(p, red);
    STAMP(6);
    // ---- attention output projection: 8 columns per CTA, partial slot 0 (bias and residual are added by the LayerNorm phase)
    STACK_BARRIER(load_w<1>(wo, W.wo, 1024, cta * 8, 0));
    STAMP(7);
    load_a_cp(sA, p.att, 1024, R);
    STAMP(8);
    {
      float acc[4][1][4];
      mma_phase<1>(acc, wo, sA, mtiles);
      STAMP(9);
      reduce_phase<1>(acc, red, mtiles, R, [&](int row, int col, float v) { p.part[(size_t)row * 1024 + cta * 8 + col] = v; });
    }
    STAMP(10);
    STACK_BARRIER((void)0);
    STAMP(11);
    if (cta < R) ln_row(p, cta, W.bo, 1, W.g2, W.be2, true, p.xn, nullptr, s_red);
    STAMP(12);
    // ---- h = GELU(LN2(x) W1^T + b1): 32 columns per CTA
    STACK_BARRIER(load_w<4>(w1, W.w1, 1024, cta * 32, 0));
    STAMP(13);
    load_a_cp(sA, p.xn, 1024, R);
    STAMP(14);
    {
      float acc[4][4][4];
      mma_phase<4>(acc, w1, sA, mtiles);
      STAMP(15);
      reduce_phase<4>(acc, red, mtiles, R, [&](int row, int col, float v) {
        const int n = cta * 32 + col;
        p.h[(size_t)row * 4096 + n] = __float2bfloat16_rn(gelu_erf(v + __ldg(W.b1 + n)));
      });
    }
    STAMP(16);
    // ---- fc2: 32 columns x one quarter of K = 4096 per CTA, partial slot = K quarter
    STACK_BARRIER(load_w<4>(w2, W.w2, 4096, (cta & 31) * 32, (cta >> 5) * 1024));
    STAMP(17);
    load_a_cp(sA, p.h + (cta >> 5) * 1024, 4096, R);
    STAMP(18);
    {
      float acc[4][4][4];
      mma_phase<4>(acc, w2, sA, mtiles);
      STAMP(19);
      reduce_phase<4>(acc, red, mtiles, R, [&](int row, int col, float v) {
        p.part[((size_t)(cta >> 5) * R + row) * 1024 + (cta & 31) * 32 + col] = v;
      });
    }
    STAMP(20);
    STACK_BARRIER((void)0);
    STAMP(21);
    if (cta < R) {
      if (last) ln_row(p, cta, W.b2, 4, p.gF, p.bF, false, nullptr, p.feats, s_red);
      else ln_row(p, cta, W.b2, 4, L[l + 1].g1, L[l + 1].be1, true, p.xn, nullptr, s_red);
    }
    STAMP(22);
  }
  // leave the barrier words at zero for the next launch: the last CTA out clears them
  __syncthreads();
  if (tid == 0) {
    const unsigned old = atomicAdd(p.sync + 1, 1u);
    if (old == gridDim.x - 1) {
      p.sync[0] = 0;
      p.sync[1] = 0;
      __threadfence();
    }
  }
}

// ---- attention phase on mma.sync (tcgen05 variant) --------------------------------------------------------------------------
// One item = (utterance, head, query split of <= 16 rows).  K, V, Q of the item arrive as bf16 through cp.async; warp w owns keys
// [8 w, 8 w + 8) of S = Q K^T (4 HMMAs), the row maxima / sums cross the warps through shared memory, P (bf16, un-normalised)
// goes back to shared memory, and warp w owns dims [8 w, 8 w + 8) of O = P V (<= 4 HMMAs, V through ldmatrix.trans).
// Every K / V element is read from shared memory once per item instead of once per query row.
constexpr int kAttLd = 72;   // bf16 per shared-memory row (64 + 8: 16-byte rows land in distinct banks for ldmatrix)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

__device__ __forceinline__ void attention_phase_mma(const StackParams& p, unsigned char* scratch, unsigned long long* tr) {
  bf16* Qs = reinterpret_cast<bf16*>(scratch);      // [16][72]
  bf16* Ks = Qs + 16 * kAttLd;                      // [64][72]
  bf16* Vs = Ks + 64 * kAttLd;                      // [64][72]
  bf16* Ps = Vs + 64 * kAttLd;                      // [16][72]
  float* rmax = reinterpret_cast<float*>(Ps + 16 * kAttLd);   // [8 warps][16 rows]
  float* rsum = rmax + kWarps * 16;                            // [8 warps][16 rows]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int T = p.T, n_bh = p.B * 16;
  int qsplit = (int)gridDim.x / n_bh;
  qsplit = qsplit < 1 ? 1 : (qsplit > T ? T : qsplit);
  const int rows_per = (T + qsplit - 1) / qsplit;   // <= 12 for B T <= 64
  const int items = n_bh * qsplit;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int bh = it / qsplit, qs = it % qsplit, b = bh >> 4, h = bh & 15;
    const int q0 = qs * rows_per, q1 = min(T, q0 + rows_per);
    if (q0 >= q1) continue;
    const int nq = q1 - q0;
    __syncthreads();   // the previous item's readers are done with the scratch
    if (tr && tid == 0) tr[0] = clock64();
    const bf16* base = p.qkv + (size_t)b * T * 3072 + h * 64;
    for (int q = tid; q < (2 * T + nq) * 8; q += kThreads) {   // K rows, V rows, query rows: 8 pieces of 16 B each
      const int rr = q >> 3, c8 = q & 7;
      const int sel = rr < T ? 1 : (rr < 2 * T ? 2 : 0), j = sel == 1 ? rr : (sel == 2 ? rr - T : q0 + rr - 2 * T);
      bf16* dst = sel == 1 ? Ks + j * kAttLd : (sel == 2 ? Vs + j * kAttLd : Qs + (j - q0) * kAttLd);
      cp_async_16(ptx::smem_u32(dst + c8 * 8), base + (size_t)j * 3072 + sel * 1024 + c8 * 8);
    }
    for (int q = tid; q < (64 - T) * 8; q += kThreads)   // V rows of the padding keys: P is 0 there, 0 x garbage must stay 0
      *reinterpret_cast<uint4*>(Vs + (T + (q >> 3)) * kAttLd + (q & 7) * 8) = make_uint4(0, 0, 0, 0);
    cp_async_wait_all();
    __syncthreads();
    if (tr && tid == 0) tr[1] = clock64();
    // S = Q K^T for keys [8 warp, 8 warp + 8)
    const int key0 = warp * 8;
    float sc[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)
    if (key0 < T) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4], bb[2];
        ldmatrix_x4(a, ptx::smem_u32(Qs + (lane & 15) * kAttLd + ks * 16 + (lane >> 4) * 8));
        ldmatrix_x2(bb, ptx::smem_u32(Ks + (key0 + (lane & 7)) * kAttLd + ks * 16 + ((lane >> 3) & 1) * 8));
        mma16816(acc, a[0], a[1], a[2], a[3], bb[0], bb[1]);
      }
      const int ka = key0 + 2 * t;
      sc[0] = ka < T ? acc[0] : -INFINITY;
      sc[1] = ka + 1 < T ? acc[1] : -INFINITY;
      sc[2] = ka < T ? acc[2] : -INFINITY;
      sc[3] = ka + 1 < T ? acc[3] : -INFINITY;
    }
    float m0 = fmaxf(sc[0], sc[1]), m1 = fmaxf(sc[2], sc[3]);
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    if (t == 0) {
      rmax[warp * 16 + g] = m0;
      rmax[warp * 16 + g + 8] = m1;
    }
    __syncthreads();
    float M0 = -INFINITY, M1 = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      M0 = fmaxf(M0, rmax[w * 16 + g]);
      M1 = fmaxf(M1, rmax[w * 16 + g + 8]);
    }
    // keys [0, 8) always hold a real key, so M0 / M1 are finite for rows that exist (rows >= nq carry garbage, never stored)
    const float e0 = expf(sc[0] - M0), e1 = expf(sc[1] - M0), e2 = expf(sc[2] - M1), e3 = expf(sc[3] - M1);
    float s0 = e0 + e1, s1 = e2 + e3;
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    if (t == 0) {
      rsum[warp * 16 + g] = s0;
      rsum[warp * 16 + g + 8] = s1;
    }
    *reinterpret_cast<uint32_t*>(Ps + g * kAttLd + key0 + 2 * t) = pack_bf16x2(e0, e1);
    *reinterpret_cast<uint32_t*>(Ps + (g + 8) * kAttLd + key0 + 2 * t) = pack_bf16x2(e2, e3);
    __syncthreads();
    float L0 = 0.f, L1 = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      L0 += rsum[w * 16 + g];
      L1 += rsum[w * 16 + g + 8];
    }
    // O = P V for dims [8 warp, 8 warp + 8)
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    const int ksteps = (T + 15) >> 4;
    for (int kk = 0; kk < ksteps; ++kk) {
      uint32_t a[4], bb[2];
      ldmatrix_x4(a, ptx::smem_u32(Ps + (lane & 15) * kAttLd + kk * 16 + (lane >> 4) * 8));
      ldmatrix_x2_trans(bb, ptx::smem_u32(Vs + (kk * 16 + (lane & 15)) * kAttLd + warp * 8));
      mma16816(o, a[0], a[1], a[2], a[3], bb[0], bb[1]);
    }
    bf16* out = p.att + ((size_t)b * T + q0) * 1024 + h * 64 + warp * 8 + 2 * t;
    if (g < nq) *reinterpret_cast<uint32_t*>(out + (size_t)g * 1024) = pack_bf16x2(o[0] / L0, o[1] / L0);
    if (g + 8 < nq) *reinterpret_cast<uint32_t*>(out + (size_t)(g + 8) * 1024) = pack_bf16x2(o[2] / L1, o[3] / L1);
    if (tr && tid == 0) tr[2] = clock64();
  }
}

// =====================================================================================================================
// tcgen05 variant (default): the GEMM phases on the 5th-generation tensor cores.
//   A (<= 64 rows x 1024, bf16) arrives as eight TMA boxes of 64 rows x 2 k-chunks of 64 columns (128-byte swizzled rows; rows
//   >= R are zero-filled by the TMA unit), the CTA's weight slice (32 or 16 rows x 1024) as ONE box issued between the two
//   halves of the grid barrier in front of the phase or -- where an attention / LayerNorm phase sits in between -- a whole
//   phase earlier, so the HBM stream runs under the barriers.  (3-D tensor maps (column, row, k-chunk); sixteen 2-D boxes per
//   slice if the driver refuses them.)  Eight threads issue 8 MMAs each (M = 64, N = 32 | 16, K = 16) over their own two
//   k-chunks into their own accumulator (32 columns of tensor memory each) -- a single thread paces such small MMAs at ~50
//   cycles apiece, the tensor pipe takes them at ~24 -- and the read-back sums the eight accumulators in order: no weight
//   registers, no shared-memory reduction.
// =====================================================================================================================
constexpr int kTcABytes = 16 * 8192;                   // 16 k-chunks x (64 rows x 128 B)
constexpr int kTcWBytes = 16 * 32 * 128;               // 16 k-chunks x (32 rows x 128 B)
constexpr int kTcIssuers = 8;                          // MMA-issuing threads per GEMM phase (measured: profiles/r02_layer_stack.txt)
constexpr int kTcSmemBytes = kTcABytes + kTcWBytes + 1024;   // + slack for the 1024-byte alignment of the tiles

struct TcState {
  uint32_t sA, sW;             // shared-memory addresses (1024-byte aligned)
  uint32_t abar, wbar, dbar;   // mbarriers: A landed (one per issuer, 8 bytes apart), W landed, MMAs retired (count = issuers)
  uint32_t tmem;               // accumulator base (32 columns per issuer)
  uint32_t pa, pw, pd;         // their phase parities (uniform over the CTA)
  int big;                     // tensor maps are the 3-D (columns, rows, k-chunks) kind: one TMA box per operand slice
};

// warp 0, between the two halves of a grid barrier (or right after the previous phase's MMAs): this CTA's weight rows
// [n0, n0 + N) x [k0, k0 + 1024) -> sW
// (Issuing a slice BEFORE an arrive -- fc2's weights, right behind fc1's MMAs -- makes that arrive's release wait for the box as
// well, from whichever warp it is issued: profiles/r02_layer_stack.txt.  It still wins over issuing them behind the arrive.)
template <int N>
__device__ __forceinline__ void tc_issue_w(const TcState& st, const CUtensorMap* map, int k0, int n0) {
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) ptx::mbar_expect_tx(st.wbar, 16u * N * 128u);
    __syncwarp();
    if (st.big) {   // one box: 64 columns x N rows x 16 k-chunks
      if (threadIdx.x == 0) ptx::tma_load_3d(st.sW, map, st.wbar, 0, n0, k0 >> 6);
    } else if (threadIdx.x < 16) {
      ptx::tma_load_2d(st.sW + threadIdx.x * N * 128, map, st.wbar, k0 + threadIdx.x * 64, n0);
    }
  }
}

// one GEMM phase of a participating CTA: A boxes, MMAs, accumulator read-back.  epi(row, first column, 16 fp32 values + bias);
// bias (nullable) points at the CTA's first column and is fetched before the accumulator is ready.
// after_mma(): runs in warp 0 once this phase's MMAs have retired (the weight buffer is free again), ahead of the read-back.
template <int N, class Epi, class After>
__device__ __forceinline__ void tc_gemm_phase(TcState& st, const CUtensorMap* mapA, int k0, int R, const float* __restrict__ bias,
                                              unsigned long long* tr, Epi epi, After after_mma) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ni = kTcIssuers, per = 16 / ni;   // MMA issuers, k-chunks per issuer
  if (threadIdx.x < 16) {   // one A box (k-chunk) per lane; the chunks of issuer i complete on mbarrier abar + 8 i
    if (tr && threadIdx.x == 0) tr[0] = clock64();
    // the rows were written by other CTAs' generic-proxy stores and are ordered before this point by the grid barrier; the
    // proxy fence orders them before the async-proxy (TMA) reads below
    // (restricted to the global space: the unrestricted form costs 800 cycles more per phase, profiles/r02_layer_stack.txt)
    asm volatile("fence.proxy.async.global;" ::: "memory");
    if (st.big) {   // one box per issuer: 64 columns x 64 rows x `per` k-chunks
      if (threadIdx.x < ni) {
        const uint32_t mybar = st.abar + 8 * threadIdx.x;
        ptx::mbar_expect_tx(mybar, (uint32_t)per * 8192u);
        ptx::tma_load_3d(st.sA + threadIdx.x * per * 8192, mapA, mybar, 0, 0, (k0 >> 6) + threadIdx.x * per);
      }
    } else {
      const uint32_t mybar = st.abar + 8 * (threadIdx.x / per);
      if (threadIdx.x % per == 0) ptx::mbar_expect_tx(mybar, (uint32_t)per * 8192u);
      __syncwarp(0xffffu);
      ptx::tma_load_2d(st.sA + threadIdx.x * 8192, mapA, mybar, k0 + threadIdx.x * 64, 0);
    }
    if (tr && threadIdx.x == 0) tr[1] = clock64();
  }
  // MMA issuers: lane 0 of warps 0 .. ni-1, each over its own k-chunks into its own accumulator (columns 32 i ..): one thread
  // issues an N = 32 MMA every ~50 cycles, the tensor pipe takes them faster.  Descriptors advance by plain additions to the
  // address field (addresses < 256 KB: no carry out of its 14 bits)
  if (lane == 0 && warp < ni) {
    ptx::mbar_wait(st.wbar, st.pw);
    if (tr && warp == 0) tr[2] = clock64();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(64, N);
    const uint64_t adesc = ptx::umma_desc_sw128(st.sA), bdesc = ptx::umma_desc_sw128(st.sW);
    ptx::mbar_wait(st.abar + 8 * warp, st.pa);
    if (tr && warp == ni - 1) tr[3] = clock64();
    ptx::tc_fence_after();
    const uint64_t abase = adesc + (uint64_t)((warp * per * 8192) >> 4), bbase = bdesc + (uint64_t)((warp * per * N * 128) >> 4);
#pragma unroll
    for (int j = 0; j < per; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::mma_bf16_ss(st.tmem + 32 * warp, abase + (uint64_t)((j * 8192 + k * 32) >> 4), bbase + (uint64_t)((j * N * 128 + k * 32) >> 4),
                         idesc, (j | k) != 0);
    ptx::mma_commit(st.dbar);
    if (tr && warp == 0) tr[4] = clock64();
  }
  // read-back.  An M = 64 accumulator keeps row r at TMEM lane 32 (r / 16) + r % 16: warp w reads lane quarter w % 4, i.e. rows
  // 16 (w % 4) + lane for lane < 16, and 16 columns (N = 32: all eight warps; N = 16: warps 0-3).
  // (M = 128 MMAs over the same tiles -- upper 64 rows garbage -- take 40 instead of 24 cycles each: the A fetch from shared
  // memory paces them.)
  const int q = warp & 3, c0 = N == 32 ? (warp >> 2) * 16 : 0, row = lane < 16 ? q * 16 + lane : 64;
  const bool reader = N == 32 ? true : warp < 4;
  if (reader) {
    float4 b4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) b4[i] = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    ptx::mbar_wait(st.dbar, st.pd);
    __syncwarp();
    ptx::tc_fence_after();
    if (warp == 0) after_mma();
    if (tr && threadIdx.x == 0) tr[5] = clock64();
    uint32_t r[16];
    ptx::tmem_ld16(st.tmem + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[4 * i] = __uint_as_float(r[4 * i]) + b4[i].x;
      v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4[i].y;
      v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4[i].z;
      v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4[i].w;
    }
#pragma unroll
    for (int a = 1; a < ni; ++a) {   // the other issuers' accumulators, in order
      ptx::tmem_ld16(st.tmem + (static_cast<uint32_t>(q * 32) << 16) + 32 * a + c0, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(r[i]);
    }
    if (row < R) epi(row, c0, v);
    ptx::tc_fence_before();
    if (tr && threadIdx.x == 0) tr[6] = clock64();
  }
  st.pa ^= 1u;
  st.pw ^= 1u;
  st.pd ^= 1u;
}

__device__ __forceinline__ void store16_bf16(bf16* dst, const float (&v)[16]) {
  uint4 o[2];
  uint32_t* w = reinterpret_cast<uint32_t*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  reinterpret_cast<uint4*>(dst)[0] = o[0];
  reinterpret_cast<uint4*>(dst)[1] = o[1];
}
__device__ __forceinline__ void store16_f32(float* dst, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// debug trace slots of GEMM phase i (CTA 0, layer 1): 7 stamps each behind the 64 phase-boundary slots
#define TR(i) ((p.trace && l == 1 && cta == 0) ? p.trace + 64 + 8 * (i) : nullptr)
#define TC_BARRIER(prefetch)                                    \
  do {                                                          \
    ptx::fence_proxy_async_smem();                              \
    barrier_arrive(p, target, btr);                             \
    prefetch;                                                   \
    if (!barrier_wait(p, target, &s_abort, btr)) goto finish;   \
    btr = nullptr;                                              \
  } while (0)

__global__ void __launch_bounds__(kThreads, 1)
layer_stack_tc_kernel(const StackParams p, const __grid_constant__ CUtensorMap map_xn, const __grid_constant__ CUtensorMap map_att,
                      const __grid_constant__ CUtensorMap map_h) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ float s_red[2 * kWarps];
  __shared__ int s_abort;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) unsigned long long s_bars[10];
  const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int R = p.R;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  TcState st;
  st.sA = ptx::smem_u32(smem);
  st.sW = st.sA + kTcABytes;
  st.abar = ptx::smem_u32(&s_bars[0]);
  st.wbar = ptx::smem_u32(&s_bars[8]);
  st.dbar = ptx::smem_u32(&s_bars[9]);
  st.pa = st.pw = st.pd = 0;
  st.big = p.big_boxes;
  if (tid == 0) {
    s_abort = 0;
    for (int i = 0; i < 8; ++i) ptx::mbar_init(st.abar + 8 * i, 1);
    ptx::mbar_init(st.wbar, 1);
    ptx::mbar_init(st.dbar, kTcIssuers);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&map_xn);
    ptx::prefetch_tmap(&map_att);
    ptx::prefetch_tmap(&map_h);
  }
  if (warp == 3) {
    ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 256);
    ptx::tmem_relinquish();
  }
  for (int i = tid; i < (kTcABytes + kTcWBytes) / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();   // ... before the TMA unit writes there
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  st.tmem = s_tmem;
  unsigned target = 0;
  unsigned long long* btr = nullptr;   // debug trace slots of the next barrier
  const StackLayer* __restrict__ L = p.layers;
  const CUtensorMap* __restrict__ wm = p.wmaps;   // [layer][qkv, out, fc1, fc2]
  const bool qkv_cta = cta < 96, out_cta = cta < 64;   // 3072 / 32 and 1024 / 16 column slices

  // A phase's weight boxes are issued one phase ahead wherever a phase without weights sits in between (attention, LayerNorm):
  // they have left the TMA queue by the time the phase's A boxes enter it
  if (qkv_cta) tc_issue_w<32>(st, wm + 0, 0, cta * 32);
  if (cta < R) ln_row(p, cta, nullptr, 0, L[0].g1, L[0].be1, false, p.xn, nullptr, s_red);

  for (int l = 0; l < p.n_layers; ++l) {
    const StackLayer W = L[l];
    const bool last = l + 1 == p.n_layers;
    // ---- q, k, v = LN1(x) Wqkv^T + b: 32 columns per CTA, 96 CTAs
    STAMP(0);
    TC_BARRIER((void)0);
    STAMP(1);
    if (qkv_cta)
      tc_gemm_phase<32>(st, &map_xn, 0, R, W.bqkv + cta * 32, TR(0), [&](int row, int c0, const float (&v)[16]) {
        store16_bf16(p.qkv + (size_t)row * 3072 + cta * 32 + c0, v);
      }, [] {});
    STAMP(4);
    btr = TR(6);
    TC_BARRIER(if (out_cta) tc_issue_w<16>(st, wm + 4 * l + 1, 0, cta * 16));   // lands during the attention phase
    STAMP(5);
    attention_phase_mma(p, smem, TR(4));
    STAMP(6);
    // ---- attention output projection: 16 columns per CTA, 64 CTAs, partial slot 0 (bias / residual: LayerNorm phase)
    TC_BARRIER((void)0);
    STAMP(7);
    if (out_cta)
      tc_gemm_phase<16>(st, &map_att, 0, R, nullptr, TR(1), [&](int row, int c0, const float (&v)[16]) {
        store16_f32(p.part + (size_t)row * 1024 + cta * 16 + c0, v);
      }, [] {});
    STAMP(10);
    LnWeights lw;   // the LayerNorm phases' weights are fetched (from HBM) under the barrier in front of them
    TC_BARRIER(tc_issue_w<32>(st, wm + 4 * l + 2, 0, cta * 32); if (cta < R) lw = ln_weights(W.bo, W.g2, W.be2));   // W1 lands during the LayerNorm phase
    STAMP(11);
    if (cta < R) ln_row(p, cta, lw, 1, true, p.xn, nullptr, s_red);
    STAMP(12);
    // ---- h = GELU(LN2(x) W1^T + b1): 32 columns per CTA
    TC_BARRIER((void)0);
    STAMP(13);
    tc_gemm_phase<32>(st, &map_xn, 0, R, W.b1 + cta * 32, TR(2), [&](int row, int c0, const float (&v)[16]) {
      float y[16];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 g = gelu2(make_float2(v[i], v[i + 1]));
        y[i] = g.x;
        y[i + 1] = g.y;
      }
      store16_bf16(p.h + (size_t)row * 4096 + cta * 32 + c0, y);
    }, [&] { tc_issue_w<32>(st, wm + 4 * l + 3, (cta >> 5) * 1024, (cta & 31) * 32); });   // fc2's weights: no phase in between to hide them
    STAMP(16);
    // ---- fc2: 32 columns x one quarter of K = 4096 per CTA, partial slot = K quarter
    btr = TR(5);
    TC_BARRIER((void)0);
    STAMP(17);
    tc_gemm_phase<32>(st, &map_h, (cta >> 5) * 1024, R, nullptr, TR(3), [&](int row, int c0, const float (&v)[16]) {
      store16_f32(p.part + ((size_t)(cta >> 5) * R + row) * 1024 + (cta & 31) * 32 + c0, v);
    }, [] {});
    STAMP(20);
    TC_BARRIER(if (!last && qkv_cta) tc_issue_w<32>(st, wm + 4 * (l + 1) + 0, 0, cta * 32);   /* next layer's Wqkv */
               if (cta < R) lw = last ? ln_weights(W.b2, p.gF, p.bF) : ln_weights(W.b2, L[l + 1].g1, L[l + 1].be1));
    STAMP(21);
    if (cta < R) {
      if (last) ln_row(p, cta, lw, 4, false, nullptr, p.feats, s_red);
      else ln_row(p, cta, lw, 4, true, p.xn, nullptr, s_red);
    }
    STAMP(22);
  }
  __syncthreads();
  if (tid == 0) {   // leave the barrier words at zero for the next launch: the last CTA out clears them
    const unsigned old = atomicAdd(p.sync + 1, 1u);
    if (old == gridDim.x - 1) {
      p.sync[0] = 0;
      p.sync[1] = 0;
      __threadfence();
    }
  }
finish:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(st.tmem, 256);
  }
}

}  // namespace

// bf16 matrix [rows][k] as a TMA tensor: 3-D (64 columns, rows, k / 64 chunks) with boxes of (64, box_rows, box_chunks) -- the
// chunk dimension's stride (128 B) is smaller than the row stride, which the encoder accepts on the drivers this was run on --
// or plain 2-D (k, rows) with boxes of (64, box_rows)
static int stack_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t k, uint32_t box_rows, uint32_t box_chunks, bool big) {
  if (big) {
    const uint64_t dims[3] = {64, rows, k / 64};
    const uint64_t strides[2] = {k * 2, 128};
    const uint32_t box[3] = {64, box_rows, box_chunks};
    return make_tmap_bf16(out, base, 3, dims, strides, box, TMAP_SW128);
  }
  const uint64_t dims[2] = {k, rows};
  const uint64_t strides[1] = {k * 2};
  const uint32_t box[2] = {64, box_rows};
  return make_tmap_bf16(out, base, 2, dims, strides, box, TMAP_SW128);
}

int layer_stack_build_wmaps(const StackLayer* host_layers, int n_layers, CUtensorMap* host_out, int* big_boxes) {
  // RTDF_STACK_BOXES=small (read per context): one 2-D box per k-chunk (A/B timing; also the automatic choice if the 3-D encode fails)
  const char* e = getenv("RTDF_STACK_BOXES");
  const int want = (e && e[0] == 's') ? 0 : 1;
  for (int big = want; big >= 0; --big) {
    int rc = RTDF_OK;
    for (int l = 0; l < n_layers && rc == RTDF_OK; ++l) {
      const bf16* w[4] = {host_layers[l].wqkv, host_layers[l].wo, host_layers[l].w1, host_layers[l].w2};
      const uint64_t n[4] = {3072, 1024, 4096, 1024}, k[4] = {1024, 1024, 1024, 4096};
      const uint32_t rows[4] = {32, 16, 32, 32};
      for (int i = 0; i < 4 && rc == RTDF_OK; ++i) rc = stack_tmap(&host_out[4 * l + i], w[i], n[i], k[i], rows[i], 16, big != 0);
    }
    if (rc == RTDF_OK) {
      *big_boxes = big;
      return RTDF_OK;
    }
    if (big == 0) return rc;
  }
  return RTDF_ERR_CUDA;
}

int layer_stack_bf16(cudaStream_t s, const StackParams& p) {
  RTDF_REQUIRE(p.R >= 1 && p.R <= kStackMaxRows && p.R == p.B * p.T, "layer_stack: %d rows (B=%d, T=%d) outside 1..%d", p.R, p.B,
               p.T, kStackMaxRows);
  RTDF_REQUIRE(p.n_layers >= 1 && p.layers && p.wmaps && p.sync && p.fault, "layer_stack: bad arguments");
  static int trace_mode = -1;   // RTDF_STACK_TRACE=1: phase timing of layer 1 on stderr (synchronises: not for captured forwards)
  if (trace_mode < 0) {
    const char* e = getenv("RTDF_STACK_TRACE");
    trace_mode = (e && e[0] == '1') ? 1 : 0;
  }
  StackParams q = p;
  q.trace = nullptr;
  if (trace_mode) {
    void* t = nullptr;
    RTDF_CHECK_CUDA(cudaMalloc(&t, 128 * sizeof(unsigned long long)));
    RTDF_CHECK_CUDA(cudaMemsetAsync(t, 0, 128 * sizeof(unsigned long long), s));
    q.trace = static_cast<unsigned long long*>(t);
  }
  if (p.impl == 1) {
    RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&layer_stack_mma_kernel), (size_t)kSmemBytes));
    layer_stack_mma_kernel<<<kCtas, kThreads, kSmemBytes, s>>>(q);
  } else {
    CUtensorMap maps[3];   // activations [R][k]: boxes of 64 columns x 64 rows (x 2 k-chunks = one MMA issuer's share)
    const void* base[3] = {p.xn, p.att, p.h};
    const int kdim[3] = {1024, 1024, 4096};
    for (int i = 0; i < 3; ++i) RTDF_TRY(stack_tmap(&maps[i], base[i], (uint64_t)p.R, (uint64_t)kdim[i], 64, 16 / kTcIssuers, p.big_boxes != 0));
    RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&layer_stack_tc_kernel), (size_t)kTcSmemBytes));
    layer_stack_tc_kernel<<<kCtas, kThreads, kTcSmemBytes, s>>>(q, maps[0], maps[1], maps[2]);
  }
  RTDF_LAUNCH_CHECK();
  if (trace_mode) {
    unsigned long long h[128];
    RTDF_CHECK_CUDA(cudaStreamSynchronize(s));
    RTDF_CHECK_CUDA(cudaMemcpy(h, q.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(q.trace);
    static const char* names[23] = {"(start)", "barrier", "load A", "mma qkv", "reduce+store", "barrier", "attention", "barrier",
                                    "load A", "mma out", "reduce+store", "barrier", "ln2", "barrier", "load A", "mma fc1",
                                    "reduce+store", "barrier", "load A", "mma fc2", "reduce+store", "barrier", "ln1 next"};
    for (int c = 0; c < 2; ++c) {
      fprintf(stderr, "layer_stack trace (R=%d) CTA %d, SM cycles per phase of layer 1:", p.R, c ? 127 : 0);
      for (int i = 1; i < 23; ++i)
        if (h[c * 32 + i]) {   // (the tcgen05 variant stamps whole GEMM phases: their time shows under "reduce+store")
          int j = i - 1;
          while (j > 0 && !h[c * 32 + j]) --j;
          fprintf(stderr, " %s=%lld", names[i], (long long)(h[c * 32 + i] - h[c * 32 + j]));
        }
      fprintf(stderr, " | layer total %lld\n", (long long)(h[c * 32 + 22] - h[c * 32]));
    }
    for (int b = 0; b < 2; ++b) {
      const unsigned long long* t = h + 104 + 8 * b;
      if (t[0])
        fprintf(stderr, "  grid barrier behind the %s phase, CTA 0, cycles from its start: CTA synchronised %lld | arrival sent %lld | "
                        "polling from %lld | all arrived seen %lld (%llu polls) | CTA released %lld\n", b ? "qkv" : "fc1",
                (long long)(t[1] - t[0]), (long long)(t[2] - t[0]), (long long)(t[3] - t[0]), (long long)(t[4] - t[0]), t[6],
                (long long)(t[5] - t[0]));
    }
    if (h[96]) fprintf(stderr, "  attention phase, CTA 0: K / V / Q in shared memory after %lld cycles, rows done after %lld\n",
                       (long long)(h[97] - h[96]), (long long)(h[98] - h[96]));
    if (h[64]) {
      static const char* gn[4] = {"qkv", "out", "fc1", "fc2"};
      for (int g = 0; g < 4; ++g) {
        const unsigned long long* t = h + 64 + 8 * g;
        fprintf(stderr, "  %s phase, CTA 0, cycles from its start: A boxes issued %lld | W landed %lld | A landed %lld | MMAs issued %lld | "
                        "accumulator ready %lld | read-back + stores done %lld\n", gn[g], (long long)(t[1] - t[0]), (long long)(t[2] - t[0]),
                (long long)(t[3] - t[0]), (long long)(t[4] - t[0]), (long long)(t[5] - t[0]), (long long)(t[6] - t[0]));
      }
    }
  }
  return RTDF_OK;
}

}  // namespace rtdf
