// Shifted-row tcgen05 convolution on zero-padded channels-last planes (see conv_tc.cuh).
//   warp 0      : TMA producer -- weights once per CTA, then one A slab (128 + span rows) per 128-row tile
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer; per tile n_chunks x nsplit x KW/16 MMAs whose
//                 A descriptors start (shift_c - shift_min) rows into the slab; double-buffered accumulators
//   warps 2..9  : epilogue, one output row and half of the columns per thread: bias / folded BN / SELU / residual / second BN+SELU,
//                 zero at padding positions, fp32 plane + bf16 (hi, lo) planes
#include "conv_tc.cuh"

#include <stdlib.h>

#include "ptx.cuh"
#include "tma_host.h"

namespace rtdf {

using namespace ptx;

namespace {

constexpr int kMaxSmem = 232448;
constexpr int kTileM = 128;
constexpr int kMaxStages = 4;
constexpr int kThreads = 320;       // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kEpiThreads = 256;

struct KParams {
  long long rows;
  int Hp, Wp, n_chunks, n_sub, shift_min, slab_rows, n_stages, total_tiles;
  int n_boxes, box_rows;         // a slab of more than 256 rows (planes wider than 124 columns) arrives as two TMA boxes
  int a_off[kConvTcMaxChunks];   // byte offset of chunk c's first A row inside the hi (or lo) region of a stage
  int hp_lo, hp_hi;
  const float *bias, *s1, *t1, *resid, *s2, *t2;
  int act1, act2;
  float* out_f32;
  bf16* out_hi;
  bf16* out_lo;
  int debug;     // RTDF_CONVTC_DEBUG bit mask -- timing experiments only (wrong results): 1 = no residual loads, 2 = no fp32 store,
                 // 4 = no hi / lo stores, 8 = no MMAs, 16 = no slab loads
};

__device__ __forceinline__ float selu_fast(float x) {
  const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
  return x > 0.f ? scale * x : (scale * alpha) * (ex2_approx(x * 1.4426950408889634f) - 1.0f);
}

// The A descriptor of a tap starts an arbitrary number of rows into the slab.  This relies on the swizzle being a
// function of the shared-memory ADDRESS bits (chunk index ^= address bits [7,10)), which is what TMA wrote and what
// the tensor core applies (verified on B200: the base-offset field must stay 0).
template <int KW>
__device__ __forceinline__ uint64_t a_desc(uint32_t addr) {
  return KW == 64 ? umma_desc_sw128(addr) : umma_desc_sw64(addr);
}

template <int KW, int CO, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
               const KParams p) {
  constexpr int kRowBytes = KW * 2;
  constexpr int kWChunkBytes = CO * KW * 2;
  constexpr int kTmemCols = 2 * CO < 32 ? 32 : 2 * CO;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int slab_bytes = p.slab_rows * kRowBytes;                       // one 64-channel block, hi or lo
  const int w_half_bytes = p.n_chunks * kWChunkBytes;
  const int w_bytes = w_half_bytes * (NS == 3 ? 2 : 1);
  const int a_half_bytes = p.n_sub * slab_bytes;
  const int stage_bytes = a_half_bytes * (NS > 1 ? 2 : 1);
  const int off_stages = w_bytes;
  const int off_bars = off_stages + p.n_stages * stage_bytes;
  const int off_params = off_bars + 256;
  const uint32_t bar_base = smem_base + off_bars;
  const uint32_t wfull_bar = bar_base;
  auto full_bar = [&](int s) { return bar_base + 8u * (1 + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (1 + kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (1 + 2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 + 2 * kMaxStages + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (5 + 2 * kMaxStages);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + off_bars + 8 * (5 + 2 * kMaxStages));
  float* s_par = reinterpret_cast<float*>(smem_gen + off_params);      // bias | s1 | t1 | s2 | t2  (CO each)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapAhi);
    prefetch_tmap(&mapWhi);
    if (NS > 1) prefetch_tmap(&mapAlo);
    if (NS == 3) prefetch_tmap(&mapWlo);
    mbar_init(wfull_bar, 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads / 32);   // one elected arrive per epilogue warp (256 per-thread arrives on one
                                                     // shared-memory word serialise: ~1 us per tile)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < CO; i += kThreads) {
    s_par[i] = p.bias ? p.bias[i] : 0.f;
    s_par[CO + i] = p.s1 ? p.s1[i] : 1.f;
    s_par[2 * CO + i] = p.t1 ? p.t1[i] : 0.f;
    s_par[3 * CO + i] = p.s2 ? p.s2[i] : 1.f;
    s_par[4 * CO + i] = p.t2 ? p.t2[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_expect_tx(wfull_bar, w_bytes);
      for (int c = 0; c < p.n_chunks; ++c) {
        tma_load_2d(smem_base + c * kWChunkBytes, &mapWhi, wfull_bar, 0, c * CO);
        if (NS == 3) tma_load_2d(smem_base + w_half_bytes + c * kWChunkBytes, &mapWlo, wfull_bar, 0, c * CO);
      }
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int s = it % p.n_stages;
        const uint32_t ph = (it / p.n_stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (p.debug & 16) {
          mbar_arrive(full_bar(s));
          continue;
        }
        mbar_expect_tx(full_bar(s), stage_bytes);
        const uint32_t dst = smem_base + off_stages + s * stage_bytes;
        const int row0 = t * kTileM + p.shift_min;        // may be negative / run past the end: TMA zero-fills
        for (int b = 0; b < p.n_sub; ++b)
          for (int bx = 0; bx < p.n_boxes; ++bx) {
            const int off = bx * p.box_rows * kRowBytes;
            tma_load_2d(dst + b * slab_bytes + off, &mapAhi, full_bar(s), b * 64, row0 + bx * p.box_rows);
            if (NS > 1) tma_load_2d(dst + a_half_bytes + b * slab_bytes + off, &mapAlo, full_bar(s), b * 64, row0 + bx * p.box_rows);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, CO);
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty_bar(a), aph ^ 1);
        const int s = it % p.n_stages;
        const uint32_t ph = (it / p.n_stages) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * CO;
        const uint32_t a_hi = smem_base + off_stages + s * stage_bytes;
        const uint32_t a_lo = a_hi + a_half_bytes;
        uint32_t first = 1;
        for (int c = 0; c < p.n_chunks && !(p.debug & 8); ++c) {
          const uint32_t ah = a_hi + p.a_off[c], al = a_lo + p.a_off[c];
          const uint32_t wh = smem_base + c * kWChunkBytes, wl = wh + w_half_bytes;
#pragma unroll
          for (int k = 0; k < KW / 16; ++k) {
            const uint64_t bh = KW == 64 ? umma_desc_sw128(wh + k * 32) : umma_desc_sw64(wh + k * 32);
            mma_bf16_ss(d_tmem, a_desc<KW>(ah + k * 32), bh, idesc, first ^ 1u);
            first = 0;
            if (NS > 1) mma_bf16_ss(d_tmem, a_desc<KW>(al + k * 32), bh, idesc, 1u);
            if (NS == 3) {
              const uint64_t bl = KW == 64 ? umma_desc_sw128(wl + k * 32) : umma_desc_sw64(wl + k * 32);
              mma_bf16_ss(d_tmem, a_desc<KW>(ah + k * 32), bl, idesc, 1u);
            }
          }
        }
        mma_commit(empty_bar(s));
        mma_commit(tfull_bar(a));
      }
      pdl_launch_dependents();
    }
  } else {
    // ===== epilogue: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4; one output row per thread =====
    const int q = warp & 3;
    const int c_begin = ((warp - 2) >> 2) * (CO / 2), c_end = c_begin + CO / 2;
    const float* s_bias = s_par;
    const float* s_s1 = s_par + CO;
    const float* s_t1 = s_par + 2 * CO;
    const float* s_s2 = s_par + 3 * CO;
    const float* s_t2 = s_par + 4 * CO;
    const bool has1 = p.s1 != nullptr, has2 = p.s2 != nullptr;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const long long m = (long long)t * kTileM + q * 32 + lane;
      const bool in_range = m < p.rows;
      const int wp = (int)(m % p.Wp);
      const int hp = (int)((m / p.Wp) % p.Hp);
      const bool ok = in_range && wp >= 1 && wp <= p.Wp - 2 && hp >= p.hp_lo && hp <= p.hp_hi;
      // the residual does not depend on this tile's MMAs: its loads are in flight while the accumulator is awaited
      constexpr int kChunks = CO / 32;              // 16-column chunks per thread (this warp's half of the CO columns)
      float4 rs[kChunks][4];
      if (p.resid && ok && !(p.debug & 1)) {
        const float4* rp = reinterpret_cast<const float4*>(p.resid + m * CO + c_begin);
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch)
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[ch][j] = rp[ch * 4 + j];
      } else {
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch)
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[ch][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(tfull_bar(a), aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * CO;
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        const int c = c_begin + ch * 16;
        uint32_t r[16];
        tmem_ld16(t_row + c, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(r[i]) + s_bias[c + i];
          if (has1) x = fmaf(x, s_s1[c + i], s_t1[c + i]);
          if (p.act1 == ACT_SELU) x = selu_fast(x);
          x += reinterpret_cast<const float*>(rs[ch])[i];
          if (has2) x = fmaf(x, s_s2[c + i], s_t2[c + i]);
          if (p.act2 == ACT_SELU) x = selu_fast(x);
          v[i] = ok ? x : 0.f;
        }
        if (in_range) {
          if (p.out_f32 && !(p.debug & 2)) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + m * CO + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (p.out_hi && !(p.debug & 4)) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
              lo[j] = pack_bf16x2(v[2 * j] - __low2float(h2), v[2 * j + 1] - __high2float(h2));
            }
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + m * CO + c);
            oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            oh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            if (p.out_lo) {
              uint4* ol = reinterpret_cast<uint4*>(p.out_lo + m * CO + c);
              ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(a));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int KW, int CO, int NS>
int launch(cudaStream_t stream, const ConvTcArgs& a) {
  KParams p{};
  p.rows = a.rows;
  p.Hp = a.Hp;
  p.Wp = a.Wp;
  p.n_chunks = a.n_chunks;
  p.n_sub = (a.ci + 63) / 64;
  int smin = a.shift[0], smax = a.shift[0];
  for (int c = 1; c < a.n_chunks; ++c) {
    smin = a.shift[c] < smin ? a.shift[c] : smin;
    smax = a.shift[c] > smax ? a.shift[c] : smax;
  }
  p.shift_min = smin;
  p.slab_rows = ((kTileM + (smax - smin) + 15) / 16) * 16;
  p.n_boxes = 1;
  p.box_rows = p.slab_rows;
  if (p.slab_rows > 256) {       // two boxes of half the slab each (a multiple of 8 rows: whole swizzle atoms)
    p.n_boxes = 2;
    p.box_rows = ((p.slab_rows + 1) / 2 + 7) / 8 * 8;
    p.slab_rows = 2 * p.box_rows;
  }
  RTDF_REQUIRE(p.box_rows <= 256, "conv_tc: tap span %d rows too wide for two TMA boxes (plane width %d)", smax - smin, a.Wp);
  const int slab_bytes = p.slab_rows * KW * 2;
  for (int c = 0; c < a.n_chunks; ++c) {
    RTDF_REQUIRE(a.sub[c] >= 0 && a.sub[c] < p.n_sub, "conv_tc: chunk %d reads channel block %d of %d", c, a.sub[c], p.n_sub);
    p.a_off[c] = a.sub[c] * slab_bytes + (a.shift[c] - smin) * KW * 2;
  }
  const int w_bytes = a.n_chunks * CO * KW * 2 * (NS == 3 ? 2 : 1);
  const int stage_bytes = p.n_sub * slab_bytes * (NS > 1 ? 2 : 1);
  const int fixed = w_bytes + 256 + 5 * CO * 4 + 1024;
  int stages = (kMaxSmem - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  RTDF_REQUIRE(stages >= 1, "conv_tc: %d bytes of weights + %d-byte stages do not fit in shared memory", w_bytes, stage_bytes);
  p.n_stages = stages;
  const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
  const long long tiles = (a.rows + kTileM - 1) / kTileM;
  RTDF_REQUIRE(tiles < (1LL << 30), "conv_tc: too many rows");
  p.total_tiles = (int)tiles;
  p.hp_lo = a.hp_lo;
  p.hp_hi = a.hp_hi;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("RTDF_CONVTC_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  p.bias = a.bias; p.s1 = a.s1; p.t1 = a.t1; p.act1 = a.act1; p.resid = a.resid;
  p.s2 = a.s2; p.t2 = a.t2; p.act2 = a.act2;
  p.out_f32 = a.out_f32; p.out_hi = a.out_hi; p.out_lo = a.out_lo;

  const TmapSwizzle sw = KW == 64 ? TMAP_SW128 : TMAP_SW64;
  CUtensorMap mAh, mAl, mWh, mWl;
  {
    uint64_t dims[2] = {(uint64_t)a.ci, (uint64_t)a.rows};
    uint64_t strides[1] = {(uint64_t)a.ci * 2};
    uint32_t box[2] = {(uint32_t)KW, (uint32_t)p.box_rows};
    RTDF_TRY(make_tmap_bf16(&mAh, a.in_hi, 2, dims, strides, box, sw));
    if (NS > 1) RTDF_TRY(make_tmap_bf16(&mAl, a.in_lo, 2, dims, strides, box, sw));
    else mAl = mAh;
  }
  {
    uint64_t dims[2] = {(uint64_t)KW, (uint64_t)a.n_chunks * CO};
    uint64_t strides[1] = {(uint64_t)KW * 2};
    uint32_t box[2] = {(uint32_t)KW, (uint32_t)CO};
    RTDF_TRY(make_tmap_bf16(&mWh, a.w_hi, 2, dims, strides, box, sw));
    if (NS == 3) RTDF_TRY(make_tmap_bf16(&mWl, a.w_lo, 2, dims, strides, box, sw));
    else mWl = mWh;
  }
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&conv_tc_kernel<KW, CO, NS>), (size_t)smem));
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  RTDF_CHECK_CUDA(launch_pdl(conv_tc_kernel<KW, CO, NS>, dim3(grid), dim3(kThreads), smem, stream, mAh, mAl, mWh, mWl, p));
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

template <int KW, int CO>
int launch_ns(cudaStream_t s, const ConvTcArgs& a, int nsplit) {
  return nsplit == 3 ? launch<KW, CO, 3>(s, a) : launch<KW, CO, 1>(s, a);
}

// ---- weight packing: fp32 -> (hi, lo) bf16 chunks ------------------------------------------------
struct PackOffsets { long long off[kConvTcMaxChunks]; };

__global__ void pack_weight_kernel(const float* __restrict__ src, int co, int kw, int n_chunks, long long stride_o,
                                   long long stride_k, PackOffsets po, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_chunks * co * kw;
  if (i >= total) return;
  const int k = (int)(i % kw);
  const int o = (int)((i / kw) % co);
  const int c = (int)(i / ((long long)kw * co));
  const float v = src[o * stride_o + k * stride_k + po.off[c]];
  const bf16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- block 0 (single input channel): one warp per plane row, lane = output channel -----------------
__global__ void __launch_bounds__(256)
block0_kernel(const float* __restrict__ z, int W, int Hp, int Wp, long long rows, const float* __restrict__ w1,
              const float* __restrict__ b1, const float* __restrict__ bn_s, const float* __restrict__ bn_t,
              const float* __restrict__ wd, const float* __restrict__ bd, bf16* __restrict__ y_hi,
              bf16* __restrict__ y_lo, float* __restrict__ idt) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= rows) return;
  const int co = threadIdx.x & 31;
  const int wp = (int)(m % Wp);
  const int hp = (int)((m / Wp) % Hp);
  const long long b = m / ((long long)Wp * Hp);
  const float* zb = z + b * 42 * W;
  const int w = wp - 1;
  auto at = [&](int h, int ww) -> float { return (h >= 0 && h < 42 && ww >= 0 && ww < W) ? zb[h * W + ww] : 0.f; };
  float y = 0.f, d = 0.f;
  if (w >= 0 && w < W) {
    if (hp <= 42) {          // conv1 output row h = hp (43 rows): taps z[h + kh - 1][w + kw - 1]
      float acc = b1[co];
#pragma unroll
      for (int kh = 0; kh < 2; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) acc = fmaf(at(hp + kh - 1, w + kw - 1), w1[co * 6 + kh * 3 + kw], acc);
      y = selu_f(fmaf(acc, bn_s[co], bn_t[co]));
    }
    if (hp >= 1 && hp <= 42) {   // downsample output row h = hp - 1
      float acc = bd[co];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) acc = fmaf(at(hp - 1, w + kw - 1), wd[co * 3 + kw], acc);
      d = acc;
    }
  }
  const bf16 h = __float2bfloat16_rn(y);
  y_hi[m * 32 + co] = h;
  y_lo[m * 32 + co] = __float2bfloat16_rn(y - __bfloat162float(h));
  idt[m * 32 + co] = d;
}

// ---- attention pooling on planes: CTA = (row h | column w, utterance), 4 x 64 threads, online softmax ----------
__global__ void __launch_bounds__(256)
attn_pool_planes_kernel(const float* __restrict__ x, const float* __restrict__ wmap, int H, int W, int Hp, int Wp,
                        const float* __restrict__ pos_S, float* __restrict__ e_S, float* __restrict__ e_T) {
  __shared__ float red[3][4][64];
  const int c = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int b = blockIdx.y;
  const bool over_w = (int)blockIdx.x < H;
  const int fixed = over_w ? blockIdx.x : blockIdx.x - H;
  const int n = over_w ? W : H;
  float mx = -INFINITY, se = 0.f, sx = 0.f;
  for (int i = grp; i < n; i += 4) {
    const int h = over_w ? fixed : i, w = over_w ? i : fixed;
    const long long idx = (((long long)b * Hp + h + 1) * Wp + w + 1) * 64 + c;
    const float v = wmap[idx], xv = x[idx];
    if (v > mx) {
      const float sc = expf(mx - v);
      se *= sc;
      sx *= sc;
      mx = v;
    }
    const float e = expf(v - mx);
    se += e;
    sx = fmaf(xv, e, sx);
  }
  red[0][grp][c] = mx;
  red[1][grp][c] = se;
  red[2][grp][c] = sx;
  __syncthreads();
  if (grp == 0) {
    float M = red[0][0][c];
#pragma unroll
    for (int g = 1; g < 4; ++g) M = fmaxf(M, red[0][g][c]);
    float S = 0.f, X = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float sc = red[0][g][c] == -INFINITY ? 0.f : expf(red[0][g][c] - M);
      S = fmaf(red[1][g][c], sc, S);
      X = fmaf(red[2][g][c], sc, X);
    }
    if (over_w) e_S[((long long)b * H + fixed) * 64 + c] = X / S + pos_S[fixed * 64 + c];
    else e_T[((long long)b * W + fixed) * 64 + c] = X / S;
  }
}

}  // namespace

int conv_tc(cudaStream_t s, const ConvTcArgs& a, int nsplit) {
  RTDF_REQUIRE(a.in_hi && a.w_hi && (a.out_f32 || a.out_hi), "conv_tc: null operand");
  RTDF_REQUIRE(nsplit == 1 || (nsplit == 3 && a.in_lo && a.w_lo), "conv_tc: nsplit must be 1, or 3 with lo operands");
  RTDF_REQUIRE(a.n_chunks >= 1 && a.n_chunks <= kConvTcMaxChunks, "conv_tc: 1..%d chunks", kConvTcMaxChunks);
  RTDF_REQUIRE(a.rows > 0 && a.Hp > 0 && a.Wp > 2, "conv_tc: bad plane geometry");
  RTDF_REQUIRE(!a.out_lo || a.out_hi, "conv_tc: out_lo needs out_hi");
  const int kw = a.ci < 64 ? a.ci : 64;
  if (kw == 32 && a.co == 32) return launch_ns<32, 32>(s, a, nsplit);
  if (kw == 32 && a.co == 64) return launch_ns<32, 64>(s, a, nsplit);
  if (kw == 64 && a.co == 64 && (a.ci == 64 || a.ci == 128)) return launch_ns<64, 64>(s, a, nsplit);
  if (kw == 64 && a.co == 128 && a.ci == 64) return launch_ns<64, 128>(s, a, nsplit);
  set_error("conv_tc: unsupported channel counts ci %d co %d", a.ci, a.co);
  return RTDF_ERR_UNSUPPORTED;
}

int conv_tc_pack_weight(cudaStream_t s, const float* src, int co, int kw, int n_chunks, long long stride_o,
                        long long stride_k, const long long* chunk_off, bf16* dst_hi, bf16* dst_lo) {
  RTDF_REQUIRE(src && dst_hi && n_chunks >= 1 && n_chunks <= kConvTcMaxChunks, "conv_tc_pack_weight: bad arguments");
  PackOffsets po{};
  for (int c = 0; c < n_chunks; ++c) po.off[c] = chunk_off[c];
  const long long total = (long long)n_chunks * co * kw;
  pack_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, co, kw, n_chunks, stride_o, stride_k, po, dst_hi, dst_lo);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int conv_tc_block0(cudaStream_t s, const float* z, int B, int W, int Hp, int Wp, const float* w1, const float* b1,
                   const float* bn_s, const float* bn_t, const float* wd, const float* bd, bf16* y_hi, bf16* y_lo,
                   float* idt) {
  RTDF_REQUIRE(z && w1 && b1 && bn_s && bn_t && wd && bd && y_hi && y_lo && idt, "conv_tc_block0: null argument");
  RTDF_REQUIRE(Hp >= 44 && Wp == W + 2 && B > 0, "conv_tc_block0: bad plane geometry");
  const long long rows = (long long)B * Hp * Wp;
  block0_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(z, W, Hp, Wp, rows, w1, b1, bn_s, bn_t, wd, bd, y_hi, y_lo, idt);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int attn_pool_planes(cudaStream_t s, const float* x, const float* wmap, int B, int H, int W, int Hp, int Wp,
                     const float* pos_S, float* e_S, float* e_T) {
  RTDF_REQUIRE(x && wmap && pos_S && e_S && e_T && B > 0 && B <= 65535, "attn_pool_planes: bad arguments");
  attn_pool_planes_kernel<<<dim3(H + W, B), 256, 0, s>>>(x, wmap, H, W, Hp, Wp, pos_S, e_S, e_T);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace rtdf
