// tcgen05 GEMM kernel (sm_100a), persistent: one CTA per SM loops over 128 x BN output tiles.
//   warp 0      : TMA producer  (A tile 128 x BK, W tile BN x BK per stage, 128B/64B swizzle)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulators live in TMEM,
//                 double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..   : epilogue -- tcgen05.ld the accumulator (one row per thread), fused bias /
//                 activation / residual / LayerNorm, vectorised global stores.
// smem stages are recycled through full/empty mbarriers; MMA completion is signalled with
// tcgen05.commit.  Replaces the cuBLAS/cuDNN calls behind fairseq's Linear / Conv1d layers
// (reference models/fe.py:19).
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tma_host.h"

#include <vector>

namespace rtdf {

using namespace ptx;

struct TcKernelParams {
  int rows_per_batch;
  int N;
  int num_kb;
  int tiles_n, tiles_m, total_tiles;
  int a_kb_col_step, a_kb_row_step, a_row_off, a_col_per_ntile;
  TcEpilogue epi;
};

// ---- optional per-launch timing (CUDA events on the launching stream), used by bench.py's roofline leg ----
struct ProfRec { cudaEvent_t a, b; double flops; int variant; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

void tc_profile_begin() {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = true;
}
int tc_profile_end(int variant, double* ms_total, double* flops_total, int* launches) {
  g_prof_on = false;
  double ms = 0, fl = 0;
  int n = 0;
  for (auto& r : g_prof) {
    RTDF_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    RTDF_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    if (variant < 0 || r.variant == variant) { ms += t; fl += r.flops; ++n; }
  }
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  if (ms_total) *ms_total = ms;
  if (flops_total) *flops_total = fl;
  if (launches) *launches = n;
  return RTDF_OK;
}

constexpr int BM = 128;

template <int BN, int BK>
struct TcCfg {
  static constexpr bool kLN = (BN == 512);
  static constexpr int kAcc = kLN ? 1 : 2;             // TMEM accumulator stages
  static constexpr int kEpiWarps = kLN ? 4 : 8;        // epilogue warps (2 per TMEM lane quarter when 8)
  static constexpr int kThreads = 64 + 32 * kEpiWarps; // warp 0: TMA, warp 1: MMA, rest: epilogue
  static constexpr int kTmemCols = kLN ? 512 : (2 * BN < 32 ? 32 : 2 * BN);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kParamBytes = kLN ? 3 * 512 * 4 : 2 * BN * 4;  // LN: bias|gamma|beta ; plain: bias[2][BN]
  static constexpr int kBudget = 224 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - kParamBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kParamBytes;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int bk) {
  if (bk == 64) return umma_desc_sw128(addr);
  // SWIZZLE_64B: 8-row x 64-byte atoms, 512 B apart
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// ---- plain epilogue: one accumulator row per thread, 32 columns per step; bias comes from smem ----
__device__ __forceinline__ void epilogue_store32(const TcEpilogue& e, const uint32_t* acc, const float* s_bias,
                                                 long long row, int col, int N, bool row_ok) {
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(acc[i]) + s_bias[i];
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = gelu_fast(o[i]);
  } else if (e.act == ACT_SWISH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = swish_f(o[i]);
  } else if (e.act == ACT_SELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = selu_f(o[i]);
  }
  if (e.scale != 1.0f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] *= e.scale;
  }
  if (!row_ok) return;
  const bool full = col + 32 <= N;
  if (e.resid) {
    const float* r = e.resid + row * e.ldr + col;
    if (full) {
      float4 t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(r + 4 * i);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[4 * i] += t[i].x; o[4 * i + 1] += t[i].y; o[4 * i + 2] += t[i].z; o[4 * i + 3] += t[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (col + i + 4 <= N) {
          const float4 t = *reinterpret_cast<const float4*>(r + i);
          o[i] += t.x; o[i + 1] += t.y; o[i + 2] += t.z; o[i + 3] += t.w;
        }
    }
  }
  if (e.out_f32) {
    float* p = e.out_f32 + row * e.ld_f32 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (full || col + i + 4 <= N) *reinterpret_cast<float4*>(p + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
  }
  if (e.out_bf16) {
    bf16* p = e.out_bf16 + row * e.ld_bf16 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 8)
      if (full || col + i + 8 <= N)
        *reinterpret_cast<uint4*>(p + i) = make_uint4(pack_bf16x2(o[i], o[i + 1]), pack_bf16x2(o[i + 2], o[i + 3]),
                                                      pack_bf16x2(o[i + 4], o[i + 5]), pack_bf16x2(o[i + 6], o[i + 7]));
  }
}

// Persistent kernel: CTA b processes tiles b, b + gridDim.x, ...  (n-tile fastest so that concurrently
// running CTAs share A rows in L2).  The smem ring runs across tile boundaries; the accumulator is
// double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
template <int BN, int BK>
__global__ void __launch_bounds__(TcCfg<BN, BK>::kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const TcKernelParams p) {
  using Cfg = TcCfg<BN, BK>;
  constexpr int kStages = Cfg::kStages;
  constexpr bool kLN = Cfg::kLN;
  constexpr int kAcc = Cfg::kAcc;
  constexpr int kEpiThreads = Cfg::kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout: full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * Cfg::kStageBytes + 8 * (2 * kStages + 4));
  float* s_params = reinterpret_cast<float*>(smem_gen + kStages * Cfg::kStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (kLN) {
    for (int i = threadIdx.x; i < 512; i += Cfg::kThreads) {
      s_params[i] = p.epi.bias ? p.epi.bias[i] : 0.f;
      s_params[512 + i] = p.epi.ln_gamma[i];
      s_params[1024 + i] = p.epi.ln_beta[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int tiles_n = p.tiles_n, tiles_m = p.tiles_m;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n_tile = t % tiles_n, m_tile = (t / tiles_n) % tiles_m, batch = t / (tiles_n * tiles_m);
        const int m0 = m_tile * BM, n0 = n_tile * BN;
        const int a_col0 = n_tile * p.a_col_per_ntile;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
          const uint32_t a_dst = smem_base + s * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          tma_load_3d(a_dst, &mapA, full_bar(s), a_col0 + kb * p.a_kb_col_step,
                      m0 + p.a_row_off + kb * p.a_kb_row_step, batch);
#pragma unroll
          for (int h = 0; h < (BN + 255) / 256; ++h)
            tma_load_2d(b_dst + h * 256 * BK * 2, &mapB, full_bar(s), kb * BK, n0 + h * 256);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr int kInstrN = BN > 256 ? 256 : BN;
      constexpr uint32_t idesc = umma_idesc_bf16(BM, kInstrN);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
        const int a = lt % kAcc;
        const uint32_t aph = (lt / kAcc) & 1;
        mbar_wait(tempty_bar(a), aph ^ 1);     // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (kLN ? 0 : a * BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_desc(a_addr + k * 32, BK);
#pragma unroll
            for (int h = 0; h < (BN + 255) / 256; ++h) {
              const uint64_t bdesc = make_desc(b_addr + h * 256 * BK * 2 + k * 32, BK);
              mma_bf16_ss(d_tmem + h * 256, adesc, bdesc, idesc, (kb | k) != 0);
            }
          }
          mma_commit(empty_bar(s));  // smem slot reusable once these MMAs retire
        }
        mma_commit(tfull_bar(a));    // accumulator stage complete
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4; with 8 warps the column range is split in two =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;  // 0 .. kEpiThreads-1
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const int n_tile = t % tiles_n, m_tile = (t / tiles_n) % tiles_m, batch = t / (tiles_n * tiles_m);
      const int m0 = m_tile * BM, n0 = n_tile * BN;
      const int a = lt % kAcc;
      const uint32_t aph = (lt / kAcc) & 1;
      const int row_local = m0 + q * 32 + lane;
      const bool row_ok = row_local < p.rows_per_batch;
      const long long row = static_cast<long long>(batch) * p.rows_per_batch + row_local;
      if (!kLN) {
        // stage this tile's bias slice (double-buffered by accumulator stage)
        float* sb = s_params + a * BN;
        for (int i = et; i < BN; i += kEpiThreads) sb[i] = (p.epi.bias && n0 + i < p.N) ? p.epi.bias[n0 + i] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      mbar_wait(tfull_bar(a), aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (kLN ? 0 : a * BN);
      if (!kLN) {
        constexpr int kColsPerWarp = BN / (Cfg::kEpiWarps / 4);
        const float* sb = s_params + a * BN;
#pragma unroll 1
        for (int c = half * kColsPerWarp; c < (half + 1) * kColsPerWarp; c += 32) {
          if (n0 + c >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          epilogue_store32(p.epi, r, sb + c, row, n0 + c, p.N, row_ok);
        }
      } else {
        // y = act(LayerNorm_512(acc + bias)); two-pass statistics straight out of TMEM
        const float* s_bias = s_params;
        const float* s_gamma = s_params + 512;
        const float* s_beta = s_params + 1024;
        float sum = 0.f;
        for (int c = 0; c < 512; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += __uint_as_float(r[i]) + s_bias[c + i];
        }
        const float mean = sum * (1.0f / 512.0f);
        float ssq = 0.f;
        for (int c = 0; c < 512; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = __uint_as_float(r[i]) + s_bias[c + i] - mean;
            ssq += d * d;
          }
        }
        const float rstd = rsqrtf(ssq * (1.0f / 512.0f) + p.epi.ln_eps);
        for (int c = 0; c < 512; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            o[i] = (__uint_as_float(r[i]) + s_bias[c + i] - mean) * rstd * s_gamma[c + i] + s_beta[c + i];
          if (p.epi.act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = gelu_fast(o[i]);
          }
          if (row_ok) {
            if (p.epi.out_bf16) {
              bf16* dst = p.epi.out_bf16 + row * p.epi.ld_bf16 + c;
#pragma unroll
              for (int i = 0; i < 32; i += 8)
                *reinterpret_cast<uint4*>(dst + i) =
                    make_uint4(pack_bf16x2(o[i], o[i + 1]), pack_bf16x2(o[i + 2], o[i + 3]),
                               pack_bf16x2(o[i + 4], o[i + 5]), pack_bf16x2(o[i + 6], o[i + 7]));
            }
            if (p.epi.out_f32) {
              float* dst = p.epi.out_f32 + row * p.epi.ld_f32 + c;
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
            }
          }
        }
      }
      // all TMEM reads of this stage are complete (tmem_ld_wait above): hand the stage back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(a));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int BK>
static int launch_variant(cudaStream_t stream, const TcOperandA& A, const bf16* W, int N, int Kw, int mode,
                          const TcEpilogue& epi) {
  using Cfg = TcCfg<BN, BK>;
  static_assert(Cfg::kStages >= 2, "need at least two stages");
  CUtensorMap mapA, mapB;
  {
    uint64_t dims[3] = {(uint64_t)A.k_extent, (uint64_t)A.rows_per_batch, (uint64_t)A.batches};
    uint64_t strides[2] = {(uint64_t)A.row_stride * 2, (uint64_t)(A.batches > 1 ? A.batch_stride : A.row_stride * A.rows_per_batch) * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    RTDF_TRY(make_tmap_bf16(&mapA, A.ptr, 3, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  {
    uint64_t dims[2] = {(uint64_t)Kw, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)Kw * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN > 256 ? 256 : BN)};
    RTDF_TRY(make_tmap_bf16(&mapB, W, 2, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  TcKernelParams p;
  p.rows_per_batch = (int)A.rows_per_batch;
  p.N = N;
  p.num_kb = ceil_div(Kw, BK);
  p.tiles_n = ceil_div(N, BN);
  p.tiles_m = ceil_div((int)A.rows_per_batch, BM);
  const long long total = (long long)p.tiles_n * p.tiles_m * A.batches;
  RTDF_REQUIRE(total < (1LL << 31), "tc_gemm: too many tiles");
  p.total_tiles = (int)total;
  if (mode == TC_POSCONV) {
    p.a_kb_col_step = 0; p.a_kb_row_step = 1; p.a_row_off = -64; p.a_col_per_ntile = 64;
  } else {
    p.a_kb_col_step = BK; p.a_kb_row_step = 0; p.a_row_off = 0; p.a_col_per_ntile = 0;
  }
  p.epi = epi;
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  ProfRec rec{};
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.a));
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.b));
    rec.flops = 2.0 * (double)A.rows_per_batch * (double)A.batches * (double)N * (double)Kw;
    rec.variant = BN + (BK == 32 ? 1 : 0);
    RTDF_CHECK_CUDA(cudaEventRecord(rec.a, stream));
  }
  tc_gemm_kernel<BN, BK><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(mapA, mapB, p);
  RTDF_LAUNCH_CHECK();
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventRecord(rec.b, stream));
    g_prof.push_back(rec);
  }
  return RTDF_OK;
}

int tc_gemm(cudaStream_t stream, const TcOperandA& A, const bf16* W, int N, int Kw, int mode, int variant,
            const TcEpilogue& epi) {
  RTDF_REQUIRE(A.ptr && W, "tc_gemm: null operand");
  RTDF_REQUIRE((reinterpret_cast<uintptr_t>(A.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "tc_gemm: operands must be 16-byte aligned");
  RTDF_REQUIRE(Kw % 8 == 0 && A.row_stride % 8 == 0 && (A.batches == 1 || A.batch_stride % 8 == 0),
               "tc_gemm: strides must be multiples of 8 elements (TMA 16-byte rule)");
  RTDF_REQUIRE(N % 8 == 0, "tc_gemm: N must be a multiple of 8");
  RTDF_REQUIRE(A.rows_per_batch > 0 && A.batches > 0 && A.batches <= 65535, "tc_gemm: bad row/batch counts");
  if (mode == TC_POSCONV) RTDF_REQUIRE(variant == 64 && Kw == 128 * 64, "tc_gemm: posconv needs variant 64, K 8192");
  switch (variant) {
    case 64: return launch_variant<64, 64>(stream, A, W, N, Kw, mode, epi);
    case 128: return launch_variant<128, 64>(stream, A, W, N, Kw, mode, epi);
    case 256: return launch_variant<256, 64>(stream, A, W, N, Kw, mode, epi);
    case 512:
    case 513:
      RTDF_REQUIRE(N == 512 && epi.ln_gamma && epi.ln_beta, "tc_gemm: LN variant needs N == 512 and LN parameters");
      return variant == 512 ? launch_variant<512, 64>(stream, A, W, N, Kw, mode, epi)
                            : launch_variant<512, 32>(stream, A, W, N, Kw, mode, epi);
    default:
      set_error("tc_gemm: unknown variant %d", variant);
      return RTDF_ERR_INVALID;
  }
}

}  // namespace rtdf
