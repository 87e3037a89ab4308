// Grouped positional convolution of the XLS-R encoder (fairseq encoder.pos_conv: Conv1d(1024, 1024, k=128, pad=64,
// groups=16, weight-normed) + drop of the last frame + GELU, added to the residual stream) as a slab-resident tcgen05
// kernel.
//
// One work item = (utterance, 256-frame tile, group).  The 64 input channels of a group are one 128-byte swizzle row,
// so the item's whole receptive field -- frames [t0-64, t0+256+63], zero-filled by TMA outside the utterance -- is one
// 384-row slab (48 KB) in shared memory, loaded ONCE.  Tap k of the convolution is then just the same slab read 128
// bytes further down: the A descriptor of tap k starts k rows into the slab (the 128B swizzle is a function of the
// shared-memory address bits, so a row-granular start is legal; conv_tc.cu relies on the same property).  Only the
// weights stream: 8 KB per tap, 4 taps per pipeline stage, shared through L2 by all CTAs working on the same group.
// Against the tap-shifted GEMM it replaces (A tile re-fetched from L2 for every tap: 24 KB per 128x64x64 block,
// L2-feed bound at 550 us for B=64) the L2 traffic drops 6x and the kernel becomes tensor-pipe bound.
//
//   warp 0      : TMA producer (slab double-buffered across items, 4-stage weight ring)
//   warp 1      : TMEM allocator + single-thread MMA issuer: per tap 2 row-halves x 4 k-steps of M128 N64 K16
//   warps 2..9  : epilogue: TMEM -> +bias -> GELU -> += into the fp32 residual stream (rows t < T only)
#include "gemm_tc.cuh"

#include "ptx.cuh"
#include "tma_host.h"

namespace rtdf {

using namespace ptx;

namespace {

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kTileFrames = 256;                 // output frames per item (two M=128 halves)
constexpr int kTaps = 128, kPad = 64, kGroupCh = 64, kGroups = 16;
constexpr int kSlabBox = 192;                    // rows per TMA box (<= 256); two boxes per slab
constexpr int kSlabRows = 2 * kSlabBox;          // 384 >= 256 + 127
constexpr int kSlabBytes = kSlabRows * 128;      // 49152
constexpr int kTapBytes = kGroupCh * 128;        // 64 output channels x 64 input channels bf16 = 8192
constexpr int kTapsPerStage = 4;
constexpr int kStageBytes = kTapsPerStage * kTapBytes;   // 32768
constexpr int kStages = 4;
constexpr int kOffStages = 2 * kSlabBytes;
constexpr int kOffBars = kOffStages + kStages * kStageBytes;
constexpr int kOffBias = kOffBars + 256;
constexpr int kSmemBytes = kOffBias + kGroupCh * 4 + 1024;   // + alignment slack
static_assert(kSmemBytes <= 232448, "pos-conv shared memory plan exceeds 227 KB");

struct Params {
  int B, T, n_ft, items;
  const float* bias;
  float* x;            // (B*T, 1024) fp32 residual stream, updated in place
};

__global__ void __launch_bounds__(kThreads, 1)
posconv_slab_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kOffBars;
  auto slab_full = [&](int i) { return bar_base + 8u * i; };
  auto slab_empty = [&](int i) { return bar_base + 8u * (2 + i); };
  auto w_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (4 + kStages + s); };
  auto tfull = [&](int a) { return bar_base + 8u * (4 + 2 * kStages + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (6 + 2 * kStages + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (8 + 2 * kStages);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBars + 8 * (8 + 2 * kStages));
  float* s_bias = reinterpret_cast<float*>(smem_gen + kOffBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(slab_full(i), 1);
      mbar_init(slab_empty(i), 1);
      mbar_init(tfull(i), 1);
      mbar_init(tempty(i), kEpiThreads / 32);   // one elected arrive per epilogue warp
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 256);      // 2 accumulator buffers x 2 row-halves x 64 columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // item -> (group, frame tile, utterance); group fastest so that co-resident CTAs stream the same few weight tiles
  auto decode = [&](int item, int& g, int& t0, int& b) {
    g = item % kGroups;
    const int r = item / kGroups;
    t0 = (r % p.n_ft) * kTileFrames;
    b = r / p.n_ft;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, wc = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        int g, t0, b;
        decode(item, g, t0, b);
        const int sb = it & 1;
        mbar_wait(slab_empty(sb), ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(slab_full(sb), kSlabBytes);
        const uint32_t slab = smem_base + sb * kSlabBytes;
        tma_load_3d(slab, &mapA, slab_full(sb), g * kGroupCh, t0 - kPad, b);
        tma_load_3d(slab + kSlabBox * 128, &mapA, slab_full(sb), g * kGroupCh, t0 - kPad + kSlabBox, b);
        for (int st = 0; st < kTaps / kTapsPerStage; ++st, ++wc) {
          const int s = wc % kStages;
          mbar_wait(w_empty(s), ((wc / kStages) & 1) ^ 1);
          mbar_expect_tx(w_full(s), kStageBytes);
          const uint32_t dst = smem_base + kOffStages + s * kStageBytes;
#pragma unroll
          for (int j = 0; j < kTapsPerStage; ++j)
            tma_load_2d(dst + j * kTapBytes, &mapW, w_full(s), (st * kTapsPerStage + j) * kGroupCh, g * kGroupCh);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kGroupCh);
      uint32_t it = 0, wc = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        int g, t0, b;
        decode(item, g, t0, b);
        const int halves = p.T - t0 > 128 ? 2 : 1;
        const int a = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(tempty(a), ph ^ 1);
        mbar_wait(slab_full(a), ph);
        tc_fence_after();
        const uint32_t slab = smem_base + a * kSlabBytes;
        const uint32_t d_tmem = tmem_base + a * 128;
        for (int st = 0; st < kTaps / kTapsPerStage; ++st, ++wc) {
          const int s = wc % kStages;
          mbar_wait(w_full(s), (wc / kStages) & 1);
          tc_fence_after();
          const uint32_t wst = smem_base + kOffStages + s * kStageBytes;
#pragma unroll
          for (int j = 0; j < kTapsPerStage; ++j) {
            const int tap = st * kTapsPerStage + j;
            for (int h = 0; h < halves; ++h) {
              const uint32_t arow = slab + (h * 128 + tap) * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_bf16_ss(d_tmem + h * kGroupCh, umma_desc_sw128(arow + k * 32),
                            umma_desc_sw128(wst + j * kTapBytes + k * 32), idesc, (tap | k) != 0 ? 1u : 0u);
            }
          }
          mma_commit(w_empty(s));
        }
        mma_commit(slab_empty(a));
        mma_commit(tfull(a));
      }
      pdl_launch_dependents();
    }
  } else {
    // epilogue: TMEM lane quarter = warp % 4 (rows), column half = (warp - 2) / 4
    const int q = warp & 3;
    const int c_begin = ((warp - 2) >> 2) * 32;
    uint32_t it = 0;
    int g_prev = -1;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      int g, t0, b;
      decode(item, g, t0, b);
      const int halves = p.T - t0 > 128 ? 2 : 1;
      const int a = it & 1;
      if (g != g_prev) {               // (re)load this group's bias; only the epilogue warps touch s_bias
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
        if (threadIdx.x - 64 < kGroupCh) s_bias[threadIdx.x - 64] = p.bias[g * kGroupCh + threadIdx.x - 64];
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
        g_prev = g;
      }
      mbar_wait(tfull(a), (it >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      for (int h = 0; h < halves; ++h) {
        const int t = t0 + h * 128 + q * 32 + lane;
        const bool ok = t < p.T;
        float* xrow = p.x + ((long long)b * p.T + (ok ? t : 0)) * 1024 + g * kGroupCh;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 128 + h * kGroupCh;
#pragma unroll
        for (int c = c_begin; c < c_begin + 32; c += 16) {
          uint32_t r[16];
          tmem_ld16(t_row + c, r);
          float4 rs[4];
          if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = reinterpret_cast<const float4*>(xrow + c)[j];
          }
          tmem_ld_wait();
          float2 o2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o2[i] = gelu2(make_float2(__uint_as_float(r[2 * i]) + s_bias[c + 2 * i],
                                      __uint_as_float(r[2 * i + 1]) + s_bias[c + 2 * i + 1]));
          if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 v = rs[j];
              v.x += o2[2 * j].x; v.y += o2[2 * j].y; v.z += o2[2 * j + 1].x; v.w += o2[2 * j + 1].y;
              reinterpret_cast<float4*>(xrow + c)[j] = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(a));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int posconv_tc(cudaStream_t s, float* x_f32, const bf16* x_bf16, int B, int T, const bf16* w_packed, const float* bias) {
  RTDF_REQUIRE(x_f32 && x_bf16 && w_packed && bias && B >= 1 && T >= 1, "posconv_tc: bad arguments");
  Params p{};
  p.B = B;
  p.T = T;
  p.n_ft = ceil_div(T, kTileFrames);
  const long long items = (long long)B * p.n_ft * kGroups;
  RTDF_REQUIRE(items < (1LL << 30), "posconv_tc: too many work items");
  p.items = (int)items;
  p.bias = bias;
  p.x = x_f32;
  CUtensorMap mA, mW;
  {
    uint64_t dims[3] = {1024, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {1024 * 2, (uint64_t)T * 1024 * 2};
    uint32_t box[3] = {(uint32_t)kGroupCh, (uint32_t)kSlabBox, 1};
    RTDF_TRY(make_tmap_bf16(&mA, x_bf16, 3, dims, strides, box, TMAP_SW128));
  }
  {
    uint64_t dims[2] = {(uint64_t)kTaps * kGroupCh, 1024};
    uint64_t strides[1] = {(uint64_t)kTaps * kGroupCh * 2};
    uint32_t box[2] = {(uint32_t)kGroupCh, (uint32_t)kGroupCh};
    RTDF_TRY(make_tmap_bf16(&mW, w_packed, 2, dims, strides, box, TMAP_SW128));
  }
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(posconv_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  RTDF_CHECK_CUDA(launch_pdl(posconv_slab_kernel, dim3(grid), dim3(kThreads), (size_t)kSmemBytes, s, mA, mW, p));
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

}  // namespace rtdf
