// Fused attention tiles for the XLS-R transformer layers.
#include "attention.cuh"
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
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(attention_simt_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
