// Fused attention tiles for the XLS-R transformer layers.
#include "attention.cuh"

#include <stdlib.h>
#include "ptx.cuh"
#include "tma_host.h"

namespace rtdf {

using namespace ptx;

// =================================================================================================
// tcgen05 kernel.  smem map (1024-aligned):
//   [0, 32K)    V  (keys x 64 d, 128-byte rows, TMA SWIZZLE_128B)  -> MN-major B operand of O = P V
//   [32K, 48K)  Q  (128 queries x 64 d)                            -> A operand of S = Q K^T
//   [48K, 80K)  K  (keys x 64 d)                                   -> K-major B operand of S
//   [32K, 96K)  P  (128 x 256 bf16 as 4 K-blocks of 128 x 64), written after S completed, so it may
//               alias Q and K.
// TMEM: 256 columns; S occupies [0, Tk_pad), O re-uses [0, 64) once every S row has been read.
// =================================================================================================
constexpr int kAttThreads = 128;
constexpr int kAttSmem = 96 * 1024 + 1024 + 64;

__global__ void __launch_bounds__(kAttThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, bf16* __restrict__ ctx, int T, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sV = base, sQ = base + 32768, sK = base + 49152, sP = base + 32768;
  uint8_t* sP_gen = gen + 32768;
  const uint32_t bars = base + 98304;
  const uint32_t bar_load = bars, bar_s = bars + 8, bar_o = bars + 16, tmem_slot = bars + 24;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 98304 + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int Tk = (T + 15) & ~15;           // keys padded to the MMA N / K granularity
  const int kboxes = (Tk + 127) / 128;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapQKV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_load, 16384u * (1 + 2 * kboxes));
    const int HD = H * 64;
    tma_load_3d(sQ, &mapQKV, bar_load, h * 64, q0, b);
    for (int i = 0; i < kboxes; ++i) {
      tma_load_3d(sK + i * 16384, &mapQKV, bar_load, HD + h * 64, i * 128, b);
      tma_load_3d(sV + i * 16384, &mapQKV, bar_load, 2 * HD + h * 64, i * 128, b);
    }
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, Tk);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      mma_bf16_ss(tmem, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + k * 32), idesc, k != 0);
    mma_commit(bar_s);
  }
  __syncwarp();
  mbar_wait(bar_s, 0);
  tc_fence_after();

  // ---- softmax: one S row per thread ----------------------------------------------------------
  const int r = warp * 32 + lane;                       // row inside the query tile == TMEM lane
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int nchunks = Tk / 16;
  float mx = -INFINITY;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[16];
    tmem_ld16(t_row + c * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c * 16 + i < T) mx = fmaxf(mx, __uint_as_float(v[i]));
  }
  const float kLog2e = 1.4426950408889634f;
  const float mxs = mx * kLog2e;
  float sum = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[16];
    tmem_ld16(t_row + c * 16, v);
    tmem_ld_wait();
    float p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      p[i] = (c * 16 + i < T) ? exp2f(fmaf(__uint_as_float(v[i]), kLog2e, -mxs)) : 0.f;
      // accumulate the bf16-rounded probability so that numerator and denominator agree
      p[i] = __bfloat162float(__float2bfloat16_rn(p[i]));
      sum += p[i];
    }
    // swizzled K-major store: K-block (c / 4), 16-byte chunks (c % 4) * 2 and +1 of row r
    const int kb = c >> 2;
    const int ch = (c & 3) * 2;
    uint8_t* rowp = sP_gen + kb * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
    *reinterpret_cast<uint4*>(rowp + (((ch) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
    *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(p[8], p[9]), pack_bf16x2(p[10], p[11]), pack_bf16x2(p[12], p[13]), pack_bf16x2(p[14], p[15]));
  }
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, /*a_mn=*/0, /*b_mn=*/1);
    for (int ks = 0; ks < nchunks; ++ks)
      mma_bf16_ss(tmem, umma_desc_sw128(sP + (ks >> 2) * 16384 + (ks & 3) * 32), umma_desc_sw128(sV + ks * 2048),
                  idesc, ks != 0);
    mma_commit(bar_o);
  }
  __syncwarp();
  mbar_wait(bar_o, 0);
  tc_fence_after();

  // ---- epilogue: O / rowsum -> bf16 ctx ----------------------------------------------------------
  const float inv = 1.0f / sum;
  const int q = q0 + r;
#pragma unroll
  for (int c = 0; c < 64; c += 32) {
    uint32_t v[32];
    tmem_ld32(t_row + c, v);
    tmem_ld_wait();
    if (q < T) {
      bf16* dst = ctx + ((long long)b * T + q) * (H * 64) + h * 64 + c;
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(
            pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv),
            pack_bf16x2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv),
            pack_bf16x2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv),
            pack_bf16x2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int attention_tc(cudaStream_t s, const bf16* qkv, bf16* ctx, int B, int T, int H) {
  RTDF_REQUIRE(qkv && ctx && B > 0 && H > 0, "attention_tc: bad arguments");
  RTDF_REQUIRE(T >= 1 && T <= 256, "attention_tc: T = %d frames unsupported (1..256; <= 5.1 s of audio)", T);
  RTDF_REQUIRE(B <= 65535, "attention_tc: batch too large");
  CUtensorMap map;
  uint64_t dims[3] = {(uint64_t)3 * H * 64, (uint64_t)T, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)3 * H * 64 * 2, (uint64_t)T * 3 * H * 64 * 2};
  uint32_t box[3] = {64, 128, 1};
  RTDF_TRY(make_tmap_bf16(&map, qkv, 3, dims, strides, box, TMAP_SW128));
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
  dim3 grid(ceil_div(T, 128), H, B);
  attention_tc_kernel<<<grid, kAttThreads, kAttSmem, s>>>(map, ctx, T, H);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// =================================================================================================
// Warp-specialised persistent kernel (default).  One CTA per SM loops over (utterance, head, 128-query tile):
//   warp 0      : TMA producer, ring of stages {Q 128x64 | K Tk x 64 | V Tk x 64} (SWIZZLE_128B)
//   warp 1      : single-thread tcgen05.mma issuer:  S = Q K^T (SS form) into TMEM buffer g = item & 1, and
//                 O = P V with P read from TENSOR MEMORY (TS form) and V as MN-major B operand
//   warps 2..5 / 6..9 : two softmax warpgroups, one per TMEM buffer (ping-pong): row max, exp2, bf16 probabilities
//                 written back over the S columns with tcgen05.st (P never touches shared memory), O / rowsum -> ctx
// TMEM buffer g (256 columns): S fp32 [0, Tk) ; P bf16x2 [0, Tk/2) (in place) ; O fp32 [128, 192).
// S(i+1) is issued before PV(i), so both warpgroups' softmax passes and the tensor pipe overlap.
// =================================================================================================
constexpr int kWsThreads = 320;
constexpr int kWsMaxStages = 6;

struct AttWsParams {
  int T, H, Tk, n_qt, total_items, n_stages, stage_bytes, kv_bytes;
  // T <= 256: two TMEM buffers of 256 columns (ping-pong), K / V in one TMA box each, O at column 128.
  // 256 < T <= 512 (5.1 .. 10.2 s of audio): S of one query tile fills all 512 columns, so there is ONE buffer, one
  // softmax warpgroup works, K / V arrive as two 256-row boxes (rows past T zero-filled by TMA), S = two N <= 256 MMAs
  // per k-step, O sits at column 448 (S columns consumed before P V is issued; P covers [0, Tk/2 <= 256)).
  int n_buf, buf_cols, o_col, kv_boxes;
  int eager_issue; // MMA thread issues whichever of (P V of the older tile, S of the next tile) is ready first (RTDF_ATTN_EAGER)
  int reverse;     // items from the last utterance to the first (the rows the QKV projection wrote last are still in L2)
  // RTDF_ATTN_DEBUG bit mask -- timing experiments only, the output is wrong when any bit is set (tools/attention_experiment.py):
  // 1 = no row-max pass, 2 = exp pass on the first chunk only, 4 = no TMA loads, 8 = no S MMAs, 16 = no PV MMAs, 32 = no O store
  int debug;
};

template <bool kLong>      // kLong: 256 < T <= 512 (single 512-column TMEM buffer); compile-time so the usual path keeps constants
__global__ void __launch_bounds__(kWsThreads, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV,
                    bf16* __restrict__ ctx, const AttWsParams p) {
  constexpr int kNb = kLong ? 1 : 2, kBufCols = kLong ? 512 : 256, kOCol = kLong ? 448 : 128, kKvBoxes = kLong ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + p.n_stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kWsMaxStages + s); };
  auto sfull_bar = [&](int g) { return bars + 8u * (2 * kWsMaxStages + g); };
  auto pfull_bar = [&](int g) { return bars + 8u * (2 * kWsMaxStages + 2 + g); };
  auto ofull_bar = [&](int g) { return bars + 8u * (2 * kWsMaxStages + 4 + g); };
  auto tempty_bar = [&](int g) { return bars + 8u * (2 * kWsMaxStages + 6 + g); };
  const uint32_t tmem_slot = bars + 8u * (2 * kWsMaxStages + 8);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + p.n_stages * p.stage_bytes + 8 * (2 * kWsMaxStages + 8));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&mapQ);
    prefetch_tmap(&mapKV);
    for (int s = 0; s < kWsMaxStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(sfull_bar(g), 1);
      mbar_init(pfull_bar(g), 4);       // one elected arrive per softmax warp (128 per-thread arrives on one
      mbar_init(ofull_bar(g), 1);       // shared-memory word serialise: ~1 us of the 16 us the empty pipeline costs)
      mbar_init(tempty_bar(g), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot_gen;
  const int HD = p.H * 64;
  const int n_local = p.total_items > (int)blockIdx.x ? (p.total_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int i = 0; i < n_local; ++i) {
        const int item = p.reverse ? p.total_items - 1 - ((int)blockIdx.x + i * (int)gridDim.x) : (int)blockIdx.x + i * (int)gridDim.x;
        const int qt = item % p.n_qt, h = (item / p.n_qt) % p.H, b = item / (p.n_qt * p.H);
        const int s = i % p.n_stages;
        const uint32_t ph = (i / p.n_stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (p.debug & 4) {
          mbar_arrive(full_bar(s));
          continue;
        }
        mbar_expect_tx(full_bar(s), 16384u + 2u * p.kv_bytes);
        const uint32_t sQ = base + s * p.stage_bytes, sK = sQ + 16384, sV = sK + p.kv_bytes;
        tma_load_3d(sQ, &mapQ, full_bar(s), h * 64, qt * 128, b);
#pragma unroll
        for (int bx = 0; bx < kKvBoxes; ++bx) {
          tma_load_3d(sK + bx * 32768, &mapKV, full_bar(s), HD + h * 64, bx * 256, b);
          tma_load_3d(sV + bx * 32768, &mapKV, full_bar(s), 2 * HD + h * 64, bx * 256, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const int n1 = p.Tk < 256 ? p.Tk : 256, n2 = p.Tk - n1;      // S columns [0, n1) and [256, 256 + n2)
      const uint32_t idesc_s = umma_idesc_bf16(128, n1);
      const uint32_t idesc_s2 = n2 > 0 ? umma_idesc_bf16(128, n2) : 0u;
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, /*a_mn=*/0, /*b_mn=*/1);
      const int ksteps = p.Tk / 16;
      constexpr int nb = kNb;
      auto issue_pv = [&](int j) {
        const int g = j % nb, s = j % p.n_stages;
        mbar_wait(pfull_bar(g), (j / nb) & 1);
        tc_fence_after();
        const uint32_t sV = base + s * p.stage_bytes + 16384 + p.kv_bytes;
        const uint32_t tP = tmem + g * kBufCols, tO = tP + kOCol;
        for (int ks = 0; ks < ksteps && !(p.debug & 16); ++ks)
          mma_bf16_ts(tO, tP + ks * 8, umma_desc_sw128(sV + ks * 2048), idesc_o, ks != 0);
        mma_commit(empty_bar(s));     // Q, K, V of this stage are no longer read
        mma_commit(ofull_bar(g));
      };
      auto issue_s = [&](int i) {
        const int g = i % nb, s = i % p.n_stages;
        tc_fence_after();
        const uint32_t sQ = base + s * p.stage_bytes, sK = sQ + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(p.debug & 8)) {
            mma_bf16_ss(tmem + g * kBufCols, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + k * 32), idesc_s, k != 0);
            if (kLong && n2 > 0)
              mma_bf16_ss(tmem + g * kBufCols + 256, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + 32768 + k * 32),
                          idesc_s2, k != 0);
          }
        mma_commit(sfull_bar(g));
      };
      if (nb == 2 && p.eager_issue) {
        // Event-driven issue order: the thread polls what each buffer is waiting for and issues whichever is ready -- the
        // P V of the older tile as soon as its probabilities are in tensor memory, the S of the next tile as soon as its
        // buffer's O has been read and its stage has landed -- instead of blocking on one barrier while the other
        // buffer's work is ready (head-of-line blocking of the fixed S(i), P V(i-1) order).
        int next_s = 0, next_pv = 0;
        while (next_pv < n_local) {
          bool progressed = false;
          if (next_pv < next_s && mbar_test_wait(pfull_bar(next_pv & 1), (next_pv >> 1) & 1)) {
            const int j = next_pv, g = j & 1, s = j % p.n_stages;
            tc_fence_after();
            const uint32_t sV = base + s * p.stage_bytes + 16384 + p.kv_bytes;
            const uint32_t tP = tmem + g * kBufCols, tO = tP + kOCol;
            for (int ks = 0; ks < ksteps && !(p.debug & 16); ++ks)
              mma_bf16_ts(tO, tP + ks * 8, umma_desc_sw128(sV + ks * 2048), idesc_o, ks != 0);
            mma_commit(empty_bar(s));
            mma_commit(ofull_bar(g));
            ++next_pv;
            progressed = true;
          }
          if (next_s < n_local && next_s - next_pv < 2 &&
              mbar_test_wait(tempty_bar(next_s & 1), ((next_s >> 1) & 1) ^ 1) &&
              mbar_test_wait(full_bar(next_s % p.n_stages), (next_s / p.n_stages) & 1)) {
            issue_s(next_s);
            ++next_s;
            progressed = true;
          }
          (void)progressed;
        }
      } else {
        for (int i = 0; i < n_local; ++i) {
          const int g = i % nb, s = i % p.n_stages;
          mbar_wait(tempty_bar(g), ((i / nb) & 1) ^ 1);   // warpgroup g has read O of item i - n_buf
          mbar_wait(full_bar(s), (i / p.n_stages) & 1);
          issue_s(i);
          if (nb == 1) issue_pv(i);          // single buffer: S(i+1) has to wait for O(i) anyway
          else if (i > 0) issue_pv(i - 1);   // ping-pong: S(i) is issued ahead of P V(i-1), so that both warpgroups' softmax
                                             // passes run side by side (issuing P V(i-1) first serialises them: 66 vs 50 us)
        }
        if (nb == 2 && n_local > 0) issue_pv(n_local - 1);
      }
      pdl_launch_dependents();
    }
  } else {
    // ===== softmax warpgroups =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                 // row inside the query tile
    const uint32_t t_row = tmem + (static_cast<uint32_t>(q * 32) << 16) + g * kBufCols;
    const int T = p.T;
    constexpr int nb = kNb;
    const float kLog2e = 1.4426950408889634f;
    for (int i = g; i < n_local && g < nb; i += nb) {
      const int item = p.reverse ? p.total_items - 1 - ((int)blockIdx.x + i * (int)gridDim.x) : (int)blockIdx.x + i * (int)gridDim.x;
      const int qt = item % p.n_qt, h = (item / p.n_qt) % p.H, b = item / (p.n_qt * p.H);
      const uint32_t ph = (i / nb) & 1;
      const bool warp_active = qt * 128 + q * 32 < T;        // warp-uniform: any valid query row in this warp
      mbar_wait(sfull_bar(g), ph);
      __syncwarp();
      tc_fence_after();
      float sum = 1.f;
      if (warp_active) {
        // Two passes over the S row in 32-column chunks, software-pipelined: the tcgen05.ld of chunk c+1 is in
        // flight while chunk c is reduced / exponentiated (the loads' ~latency was the top stall of the
        // one-load-one-wait version: long_scoreboard 33 % of samples, MUFU pipe 29 % busy).
        const int nch = (p.Tk + 31) >> 5;
        uint32_t va[32], vb[32];
        float mx = -INFINITY, mx1 = -INFINITY;
        auto chunk_max = [&](const uint32_t (&v)[32], int c) {
          if (c * 32 + 32 <= T) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              mx = max3(mx, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              mx1 = max3(mx1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
        };
        if (!(p.debug & 1)) tmem_ld32(t_row, va);
        for (int c = 0; c < nch && !(p.debug & 1); c += 2) {
          tmem_ld_wait_regs(va);
          if (c + 1 < nch) tmem_ld32(t_row + (c + 1) * 32, vb);
          chunk_max(va, c);
          if (c + 1 < nch) {
            tmem_ld_wait_regs(vb);
            if (c + 2 < nch) tmem_ld32(t_row + (c + 2) * 32, va);
            chunk_max(vb, c + 1);
          }
        }
        tmem_ld32(t_row, va);             // first chunk of the second pass
        const float mxs = fmaxf(mx, mx1) * kLog2e;
        sum = 0.f;
        auto chunk_exp = [&](const uint32_t (&v)[32], int c) {
          uint32_t pk[16];
          const bool fullc = c * 32 + 32 <= T;      // warp-uniform
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), kLog2e, -mxs));
            float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), kLog2e, -mxs));
            if (!fullc) {
              if (c * 32 + 2 * j >= T) e0 = 0.f;
              if (c * 32 + 2 * j + 1 >= T) e1 = 0.f;
            }
            // accumulate the bf16-rounded probabilities so that numerator and denominator agree
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
            sum += __low2float(h2) + __high2float(h2);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          // P chunk c overwrites S columns [16c, 16c+16): already consumed, and behind the chunk in flight
          tmem_st8(t_row + c * 16, pk);
          tmem_st8(t_row + c * 16 + 8, pk + 8);
        };
        for (int c = 0; c < ((p.debug & 2) ? 1 : nch); c += 2) {
          tmem_ld_wait_regs(va);
          if (c + 1 < nch) tmem_ld32(t_row + (c + 1) * 32, vb);
          chunk_exp(va, c);
          if (c + 1 < nch) {
            tmem_ld_wait_regs(vb);
            if (c + 2 < nch) tmem_ld32(t_row + (c + 2) * 32, va);
            chunk_exp(vb, c + 1);
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull_bar(g));
      mbar_wait(ofull_bar(g), ph);
      __syncwarp();
      tc_fence_after();
      if (warp_active && !(p.debug & 32)) {
        const float inv = 1.0f / sum;
        const int tq = qt * 128 + r;
        uint32_t v0[32], v1[32];
        tmem_ld32(t_row + kOCol, v0);
        tmem_ld32(t_row + kOCol + 32, v1);
        tmem_ld_wait();
        if (tq < T) {
          bf16* dst = ctx + ((long long)b * T + tq) * HD + h * 64;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(
                pack_bf16x2(__uint_as_float(v0[j]) * inv, __uint_as_float(v0[j + 1]) * inv),
                pack_bf16x2(__uint_as_float(v0[j + 2]) * inv, __uint_as_float(v0[j + 3]) * inv),
                pack_bf16x2(__uint_as_float(v0[j + 4]) * inv, __uint_as_float(v0[j + 5]) * inv),
                pack_bf16x2(__uint_as_float(v0[j + 6]) * inv, __uint_as_float(v0[j + 7]) * inv));
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(dst + 32 + j) = make_uint4(
                pack_bf16x2(__uint_as_float(v1[j]) * inv, __uint_as_float(v1[j + 1]) * inv),
                pack_bf16x2(__uint_as_float(v1[j + 2]) * inv, __uint_as_float(v1[j + 3]) * inv),
                pack_bf16x2(__uint_as_float(v1[j + 4]) * inv, __uint_as_float(v1[j + 5]) * inv),
                pack_bf16x2(__uint_as_float(v1[j + 6]) * inv, __uint_as_float(v1[j + 7]) * inv));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(g));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int attention_ws(cudaStream_t s, const bf16* qkv, bf16* ctx, int B, int T, int H, bool reverse) {
  RTDF_REQUIRE(qkv && ctx && B > 0 && H > 0, "attention_ws: bad arguments");
  RTDF_REQUIRE(T >= 1 && T <= 512, "attention_ws: T = %d frames unsupported (1..512; <= 10.2 s of audio)", T);
  AttWsParams p;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("RTDF_ATTN_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  p.T = T;
  p.H = H;
  p.reverse = reverse ? 1 : 0;
  {
    static int eager = -1;
    if (eager < 0) {
      const char* e = getenv("RTDF_ATTN_EAGER");
      eager = (e && e[0] == '1') ? 1 : 0;
    }
    p.eager_issue = eager;
  }
  p.Tk = (T + 15) & ~15;
  p.n_qt = ceil_div(T, 128);
  const long long items = (long long)B * H * p.n_qt;
  RTDF_REQUIRE(items < (1LL << 30), "attention_ws: batch too large");
  p.total_items = (int)items;
  const bool long_mode = T > 256;
  p.n_buf = long_mode ? 1 : 2;
  p.buf_cols = long_mode ? 512 : 256;
  p.o_col = long_mode ? 448 : 128;
  p.kv_boxes = long_mode ? 2 : 1;
  p.kv_bytes = long_mode ? 65536 : p.Tk * 128;
  p.stage_bytes = 16384 + 2 * p.kv_bytes;
  const int budget = 232448 - 1024 - 256;
  p.n_stages = budget / p.stage_bytes;
  if (p.n_stages > kWsMaxStages) p.n_stages = kWsMaxStages;
  RTDF_REQUIRE(p.n_stages >= (long_mode ? 1 : 2), "attention_ws: stage of %d bytes does not fit", p.stage_bytes);
  const size_t smem = (size_t)p.n_stages * p.stage_bytes + 256 + 1024;
  CUtensorMap mapQ, mapKV;
  uint64_t dims[3] = {(uint64_t)3 * H * 64, (uint64_t)T, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)3 * H * 64 * 2, (uint64_t)T * 3 * H * 64 * 2};
  uint32_t boxq[3] = {64, 128, 1};
  uint32_t boxkv[3] = {64, (uint32_t)(long_mode ? 256 : p.Tk), 1};
  RTDF_TRY(make_tmap_bf16(&mapQ, qkv, 3, dims, strides, boxq, TMAP_SW128));
  RTDF_TRY(make_tmap_bf16(&mapKV, qkv, 3, dims, strides, boxkv, TMAP_SW128));
  const int grid = p.total_items < kNumSMs ? p.total_items : kNumSMs;
  if (long_mode) {
    RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&attention_ws_kernel<true>), (size_t)smem));
    RTDF_CHECK_CUDA(launch_pdl(attention_ws_kernel<true>, dim3(grid), dim3(kWsThreads), smem, s, mapQ, mapKV, ctx, p));
  } else {
    RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&attention_ws_kernel<false>), (size_t)smem));
    RTDF_CHECK_CUDA(launch_pdl(attention_ws_kernel<false>, dim3(grid), dim3(kWsThreads), smem, s, mapQ, mapKV, ctx, p));
  }
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

// =================================================================================================
// SIMT kernel: K/V of one head staged in smem as fp32; one warp per query at a time.
// =================================================================================================
constexpr int kSimtQPerCta = 32;

template <typename T_>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T_* __restrict__ qkv, T_* __restrict__ ctx, int T, int H) {
  extern __shared__ float sm[];
  float* sK = sm;                         // [T][65]
  float* sV = sK + (size_t)T * 65;        // [T][64]
  float* sQ = sV + (size_t)T * 64;        // [8][64]
  float* sPr = sQ + 8 * 64;               // [8][Tp]
  const int Tp = (T + 31) & ~31;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int ld = 3 * H * 64;
  const T_* base = qkv + (long long)b * T * ld;
  for (int i = threadIdx.x; i < T * 64; i += 256) {
    const int t = i >> 6, d = i & 63;
    sK[t * 65 + d] = to_f32(base[(long long)t * ld + H * 64 + h * 64 + d]);
    sV[t * 64 + d] = to_f32(base[(long long)t * ld + 2 * H * 64 + h * 64 + d]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = sQ + warp * 64;
  float* pr = sPr + warp * Tp;
  for (int qi = warp; qi < kSimtQPerCta; qi += 8) {
    const int t = blockIdx.y * kSimtQPerCta + qi;
    if (t >= T) break;
    q[lane] = to_f32(base[(long long)t * ld + h * 64 + lane]);
    q[lane + 32] = to_f32(base[(long long)t * ld + h * 64 + lane + 32]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Tp; j += 32) {
      float sc = -INFINITY;
      if (j < T) {
        sc = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) sc = fmaf(q[d], sK[j * 65 + d], sc);
      }
      pr[j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tp; j += 32) {
      const float e = j < T ? expf(pr[j] - mx) : 0.f;
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float p = pr[j];
      o0 = fmaf(p, sV[j * 64 + lane], o0);
      o1 = fmaf(p, sV[j * 64 + lane + 32], o1);
    }
    const float inv = 1.0f / sum;
    T_* dst = ctx + ((long long)b * T + t) * (H * 64) + h * 64;
    dst[lane] = from_f32<T_>(o0 * inv);
    dst[lane + 32] = from_f32<T_>(o1 * inv);
    __syncwarp();
  }
}

template <typename T_>
static int simt_launch(cudaStream_t s, const T_* qkv, T_* ctx, int B, int T, int H) {
  RTDF_REQUIRE(qkv && ctx && B > 0 && T > 0 && H > 0, "attention_simt: bad arguments");
  const int Tp = (T + 31) & ~31;
  const size_t smem = ((size_t)T * 65 + (size_t)T * 64 + 8 * 64 + 8 * Tp) * sizeof(float);
  RTDF_REQUIRE(smem <= 220 * 1024, "attention_simt: T = %d too long for the smem-resident K/V kernel", T);
  RTDF_CHECK_CUDA(raise_max_dyn_smem(reinterpret_cast<const void*>(&attention_simt_kernel<T_>), (size_t)smem));
  dim3 grid(B * H, ceil_div(T, kSimtQPerCta));
  attention_simt_kernel<T_><<<grid, 256, smem, s>>>(qkv, ctx, T, H);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int attention_simt_f32(cudaStream_t s, const float* qkv, float* ctx, int B, int T, int H) {
  return simt_launch<float>(s, qkv, ctx, B, T, H);
}
int attention_simt_bf16(cudaStream_t s, const bf16* qkv, bf16* ctx, int B, int T, int H) {
  return simt_launch<bf16>(s, qkv, ctx, B, T, H);
}

}  // namespace rtdf
