// tcgen05 GEMM kernel (sm_100a), persistent: one CTA per SM loops over 128 x BN output tiles.
//   warp 0      : TMA producer  (A tile 128 x BK, W tile BN x BK per stage, 128B/64B swizzle)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulators live in TMEM,
//                 double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..9  : epilogue -- tcgen05.ld the accumulator (one row per thread), fused bias / activation /
//                 scale / LayerNorm, results staged in 128B-swizzled smem and written with TMA stores
//                 (fp32 residual updates use the TMA reduce-add, so the SM never reads the residual).
// smem stages are recycled through full/empty mbarriers; MMA completion is signalled with
// tcgen05.commit.  Replaces the cuBLAS/cuDNN calls behind fairseq's Linear / Conv1d layers
// (reference models/fe.py:19).
#include "gemm_tc.cuh"

#include <atomic>
#include <mutex>
#include "ptx.cuh"
#include "tma_host.h"

#include <stdlib.h>
#include <vector>

namespace rtdf {

using namespace ptx;

enum EpiMode {
  EPI_DIRECT = 0,       // per-thread global loads/stores (any combination of outputs / residual)
  EPI_TMA_F32 = 1,      // out_f32 = v                (TMA store)
  EPI_TMA_F32_ADD = 2,  // out_f32 += v               (TMA reduce-add; in-place residual)
  EPI_TMA_BF16 = 3,     // out_bf16 = v               (TMA store)
  EPI_XRES = 4,         // out_f32 = x_old + v (x_old read by the SM, TMA store) + bf16 copy + per-row partial statistics
};

struct TcKernelParams {
  int rows_per_batch;
  int N;
  int num_kb;
  int tiles_n, tiles_m, total_tiles;
  int a_kb_col_step, a_kb_row_step, a_row_off, a_col_per_ntile;
  int epi_mode;
  int debug = 0;           // RTDF_GEMM_DEBUG bit mask (timing experiments only): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue
  // CTA-pair kernel, wave quantisation: work units [0, tail_start) are whole 256-wide tiles; each tile from tail_start on --
  // the ones that would form a partly empty last round -- is cut into tail_sub column slices (units of 256 / tail_sub
  // columns, own accumulator, own epilogue), so that the last round costs a fraction of a tile time.  tail_sub = 1: off.
  int tail_start = 0, tail_sub = 1, total_units = 0;
  int a_hint = 0;          // L2 policy of the A-operand loads: 0 default, 1 evict_first (the activation is dead after this GEMM:
                           // do not let it push the residual stream / the outputs out of L2), 2 evict_last
  int reverse = 0;         // walk the tiles from the last to the first (TcEpilogue::reverse_tiles): the rows the previous kernel
                           // wrote last are then read first, while they are still in L2
  int stream_mode = 0;     // streaming-chunk launches under programmatic dependent launch: dependents are released at the
                           // start, and the weight (W) tiles of the first ring round are requested BEFORE waiting for the
                           // previous kernel (they do not depend on it), so the weight stream of kernel k+1 overlaps kernel k
  int k_splits = 1;        // split-K: tile t covers k-blocks [split * kb_per_split, ...) and reduce-adds its partial
  int kb_per_split = 0;
  TcEpilogue epi;
};

// ---- optional per-launch timing (CUDA events on the launching stream), used by bench.py's roofline leg ----
struct ProfRec { cudaEvent_t a, b; double flops; int variant; };
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;     // launches may come from several host threads (one context each)

void tc_profile_begin() {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = true;
}
int tc_profile_end(int variant, double* ms_total, double* flops_total, int* launches) {
  g_prof_on = false;
  std::lock_guard<std::mutex> lock(g_prof_mu);
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

static bool tail_slices_enabled() {
  static int v = -1;
  if (v < 0) {
    // opt-in (r02 experiment, measured SLOWER: out_proj 38 -> 46 us, fc2 90 -> 122 us, profiles/r02_tail_split_ab.txt):
    // a 64-column slice of a pair tile costs 0.7 of the whole tile -- M = 256 MMAs are bound by the A-operand feed, not N
    const char* e = getenv("RTDF_TAIL_SLICES");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static bool stream_prefetch_enabled() {
  static int v = -1;
  if (v < 0) {
    // opt-in experiment (r02): release the dependents at kernel start and request the first ring round of weight tiles
    // before waiting for the previous kernel.  Measured SLOWER for streaming chunks (batch 1, 1 s: p50 1.44 vs 1.37 ms,
    // profiles/r02_stream_prefetch_ab.txt): the early dependents only find room on the SMs the running GEMM left idle.
    const char* e = getenv("RTDF_STREAM_PREFETCH");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

constexpr int BM = 128;
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA

template <int BN, int BK>
struct TcCfg {
  static constexpr bool kLN = (BN == 512);
  static constexpr int kAcc = kLN ? 1 : 2;             // TMEM accumulator stages
  static constexpr int kEpiWarps = 8;                  // 2 warps per TMEM lane quarter (column halves)
  static constexpr int kThreads = 64 + 32 * kEpiWarps; // warp 0: TMA, warp 1: MMA, rest: epilogue
  static constexpr int kTmemCols = kLN ? 512 : (2 * BN < 32 ? 32 : 2 * BN);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 4096;                 // one 32-row x 128-byte box per warp
  // LN: bias|gamma|beta [512] + partial stats [2 bufs][2 halves][128 rows] float2 ; plain: bias[2][BN]
  static constexpr int kParamBytes = kLN ? (3 * 512 * 4 + 2 * 2 * 128 * 8) : 2 * BN * 4;
  static constexpr int kFixed = kStagingBytes + 256 /*barriers*/ + kParamBytes;
  static constexpr int kStagesRaw = (kMaxSmem - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kNeeded = kStages * kStageBytes + kFixed;         // measured from a 1024-aligned base
  static constexpr int kSmemBytes = kNeeded + 1024 > kMaxSmem ? kMaxSmem : kNeeded + 1024;
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

// Per-thread (= per output row) state of a folded LayerNorm: v = acc * rstd + (nmr * c_j + bias_j), nmr = -mean * rstd
struct LnFold {
  const float* c = nullptr;   // global, column 0 of the output (nullptr: no fold)
  float rstd = 1.f, nmr = 0.f;
};

__device__ __forceinline__ LnFold ln_fold_row(const TcEpilogue& e, long long row, bool row_ok) {
  LnFold f;
  if (!e.fold_stats) return f;
  f.c = e.fold_c;
  if (row_ok) {
    const float4* st = reinterpret_cast<const float4*>(e.fold_stats + row * 8);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {      // fixed order: the statistics do not depend on who wrote which slot first
      const float4 v = st[i];
      s += v.x; q += v.y;
      s += v.z; q += v.w;
    }
    const float inv_n = 1.0f / (float)e.fold_len;
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    f.rstd = rsqrtf(var + e.fold_eps);
    f.nmr = -mean * f.rstd;
  }
  return f;
}

// v = act(acc + bias) * scale for 32 consecutive columns of this thread's row (col = first of them in the output)
__device__ __forceinline__ void epilogue_math32(const TcEpilogue& e, const uint32_t* acc, const float* s_bias, float* o,
                                                const LnFold& fold = LnFold(), int col = 0) {
  float2* o2 = reinterpret_cast<float2*>(o);
  const float2* b2 = reinterpret_cast<const float2*>(s_bias);
  if (fold.c) {
    const float4* c4 = reinterpret_cast<const float4*>(fold.c + col);     // same address in every lane: broadcast loads
    const float2 r2 = make_float2(fold.rstd, fold.rstd), m2 = make_float2(fold.nmr, fold.nmr);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 cv = __ldg(c4 + i);
      o2[2 * i] = fma2(make_float2(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1])), r2,
                       fma2(m2, make_float2(cv.x, cv.y), b2[2 * i]));
      o2[2 * i + 1] = fma2(make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])), r2,
                           fma2(m2, make_float2(cv.z, cv.w), b2[2 * i + 1]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      o2[i] = add2(make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1])), b2[i]);
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o2[i] = gelu2(o2[i]);
  } else if (e.act == ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o2[i] = gelu2_tanh(o2[i]);
  } else if (e.act == ACT_GELU_AS) {
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
}

// direct (non-TMA) output path: per-thread row-strided global accesses
__device__ __forceinline__ void epilogue_direct32(const TcEpilogue& e, float* o, long long row, int col, int N, bool row_ok) {
  if (!row_ok) return;
  const bool full = col + 32 <= N;
  if (e.resid) {
    const float* r = e.resid + row * e.ldr + col;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      if (full || col + i + 4 <= N) {
        const float4 t = *reinterpret_cast<const float4*>(r + i);
        o[i] += t.x; o[i + 1] += t.y; o[i + 2] += t.z; o[i + 3] += t.w;
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

// Stage one 128-byte row per lane into this warp's 32 x 128 B box (SWIZZLE_128B pattern expected by the
// TMA store: 16-byte chunk j of row r lives at chunk j ^ (r & 7)).
__device__ __forceinline__ void stage_row_f32(uint8_t* box, int lane, const float* o) {
  uint8_t* rowp = box + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
}
__device__ __forceinline__ void stage_row_bf16(uint8_t* box, int lane, const float* o /*64 values*/) {
  uint8_t* rowp = box + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) =
        make_uint4(pack_bf16x2(o[8 * j], o[8 * j + 1]), pack_bf16x2(o[8 * j + 2], o[8 * j + 3]),
                   pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), pack_bf16x2(o[8 * j + 6], o[8 * j + 7]));
}

// Epilogue of one 128 x BN accumulator tile for one warp (32 rows, half of the columns): bias / activation / scale,
// then TMA store (bf16 | fp32), TMA reduce-add (in-place fp32 residual) or direct global stores.
template <int BN>
__device__ __forceinline__ void epilogue_plain_tile(const TcKernelParams& p, const CUtensorMap* mapC, int mode, uint32_t t_row,
                                                    const float* sb, int half, int lane, uint8_t* box_gen, uint32_t box,
                                                    int n0, int row0, int batch, long long row, bool row_ok,
                                                    const LnFold& fold, int n_tile, int n_limit = 0x7fffffff) {
  constexpr int kColsPerWarp = BN / 2;
  const int c_begin = half * kColsPerWarp, c_end = c_begin + kColsPerWarp;
  if (n_limit > p.N) n_limit = p.N;       // columns from n_limit on are not part of this work unit
  if (mode == EPI_XRES) {
    // x_new = x_old + v: x_old is read by this thread (its own row), x_new goes back through TMA stores, its bf16 copy and
    // this thread's partial (sum, sum of squares) over its columns go out with plain stores.
    if constexpr (kColsPerWarp >= 32) {
      const float* xr = p.epi.resid + row * p.epi.ldr + n0;
      bf16* xb = p.epi.xb_out + row * (long long)p.N + n0;
      float s_sum = 0.f, s_sq = 0.f;
      float4 xo[8];
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) xo[i] = *reinterpret_cast<const float4*>(xr + c_begin + 4 * i);
      }
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c, r);
        tmem_ld_wait();
        __align__(16) float o[32];
        epilogue_math32(p.epi, r, sb + c, o, fold, n0 + c);
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[4 * i] += xo[i].x; o[4 * i + 1] += xo[i].y; o[4 * i + 2] += xo[i].z; o[4 * i + 3] += xo[i].w;
          }
          if (c + 32 < c_end) {       // next chunk's x_old: in flight while this one is staged and stored
#pragma unroll
            for (int i = 0; i < 8; ++i) xo[i] = *reinterpret_cast<const float4*>(xr + c + 32 + 4 * i);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s_sum += o[i];
            s_sq = fmaf(o[i], o[i], s_sq);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            *reinterpret_cast<uint4*>(xb + c + i) = make_uint4(pack_bf16x2(o[i], o[i + 1]), pack_bf16x2(o[i + 2], o[i + 3]),
                                                               pack_bf16x2(o[i + 4], o[i + 5]), pack_bf16x2(o[i + 6], o[i + 7]));
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        stage_row_f32(box_gen, lane, o);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < p.rows_per_batch) {
          tma_store_3d(mapC, box, n0 + c, row0, batch);
          tma_store_commit();
        }
      }
      if (row_ok) p.epi.stats_out[row * 8 + n_tile * 2 + half] = make_float2(s_sum, s_sq);
    }
  } else if (mode == EPI_TMA_BF16 && kColsPerWarp >= 64) {
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 64) {
      if (n0 + c >= n_limit) break;  // warp-uniform
      uint32_t r0[32], r1[32];
      tmem_ld32(t_row + c, r0);
      tmem_ld32(t_row + c + 32, r1);
      tmem_ld_wait();
      __align__(16) float o[64];
      epilogue_math32(p.epi, r0, sb + c, o, fold, n0 + c);
      epilogue_math32(p.epi, r1, sb + c + 32, o + 32, fold, n0 + c + 32);
      if (lane == 0) tma_store_wait_read<0>();   // previous store out of this box has drained
      __syncwarp();
      stage_row_bf16(box_gen, lane, o);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && row0 < p.rows_per_batch) {
        tma_store_3d(mapC, box, n0 + c, row0, batch);
        tma_store_commit();
      }
    }
  } else {
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      if (n0 + c >= n_limit) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld32(t_row + c, r);
      tmem_ld_wait();
      __align__(16) float o[32];
      epilogue_math32(p.epi, r, sb + c, o, fold, n0 + c);
      if (mode == EPI_TMA_F32 || mode == EPI_TMA_F32_ADD) {
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        stage_row_f32(box_gen, lane, o);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < p.rows_per_batch) {
          if (mode == EPI_TMA_F32_ADD) tma_reduce_add_3d(mapC, box, n0 + c, row0, batch);
          else tma_store_3d(mapC, box, n0 + c, row0, batch);
          tma_store_commit();
        }
      } else {
        epilogue_direct32(p.epi, o, row, n0 + c, p.N, row_ok);
      }
    }
  }
}

// ---- fused row LayerNorm behind a residual GEMM (TcEpilogue::rowln_*) --------------------------------------------
// Called by all epilogue warps of a CTA after the epilogue of one tile.  rb = 128-row block of that tile.
template <int kEpiThreads>
__device__ __forceinline__ void rowln_after_tile(const TcKernelParams& p, int rb, int row_begin, int et, int lane,
                                                 volatile int* s_flag) {
  const TcEpilogue& e = p.epi;
  if (lane == 0) tma_store_wait_all<0>();      // this warp's TMA stores / reduce-adds of the tile have been performed
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  if (et == 0) {
    const int old = atomicAdd(e.rowln_counters + rb, 1);
    const int last = old == p.tiles_n * p.k_splits - 1;      // every N-tile (and K split) of the block has been added
    if (last) e.rowln_counters[rb] = 0;       // ready for the next launch
    __threadfence();
    *s_flag = last;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  if (!*s_flag) return;
  // this CTA added the last N-tile of rows [row_begin, row_begin + 128): normalise them (warp per row, 2 rows in flight)
  const int N = p.N, groups = N >> 7;
  const float inv_n = 1.0f / (float)N;
  const int w = et >> 5;
  constexpr int kWarps = kEpiThreads / 32;
  for (int r = w; r < 128; r += 2 * kWarps) {
    float4 v[2][8];
    long long rows[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      rows[u] = (long long)row_begin + r + u * kWarps;
      const bool ok = r + u * kWarps < 128 && rows[u] < p.rows_per_batch;
      if (!ok) rows[u] = -1;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < groups && ok) v[u][j] = __ldcg(reinterpret_cast<const float4*>(e.out_f32 + rows[u] * N + j * 128 + lane * 4));
        else v[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (rows[u] < 0) continue;       // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[u][j].x + v[u][j].y + v[u][j].z + v[u][j].w;
      const float mean = warp_sum(s) * inv_n;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < groups) {
          const float a = v[u][j].x - mean, b = v[u][j].y - mean, c = v[u][j].z - mean, d = v[u][j].w - mean;
          q += a * a + b * b + c * c + d * d;
        }
      const float rstd = rsqrtf(warp_sum(q) * inv_n + e.rowln_eps);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < groups) {
          const int c0 = j * 128 + lane * 4;
          const float4 g = *reinterpret_cast<const float4*>(e.rowln_gamma + c0);
          const float4 bt = *reinterpret_cast<const float4*>(e.rowln_beta + c0);
          const float a = (v[u][j].x - mean) * rstd * g.x + bt.x;
          const float b = (v[u][j].y - mean) * rstd * g.y + bt.y;
          const float c = (v[u][j].z - mean) * rstd * g.z + bt.z;
          const float d = (v[u][j].w - mean) * rstd * g.w + bt.w;
          if (e.rowln_out_f32) *reinterpret_cast<float4*>(e.rowln_out_f32 + rows[u] * N + c0) = make_float4(a, b, c, d);
          if (e.rowln_out_bf16)
            *reinterpret_cast<uint2*>(e.rowln_out_bf16 + rows[u] * N + c0) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
        }
    }
  }
}

// Persistent kernel: CTA b processes tiles b, b + gridDim.x, ...  (n-tile fastest so that concurrently
// running CTAs share A rows in L2).  The smem ring runs across tile boundaries.
template <int BN, int BK>
__global__ void __launch_bounds__(TcCfg<BN, BK>::kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, const TcKernelParams p) {
  using Cfg = TcCfg<BN, BK>;
  constexpr int kStages = Cfg::kStages;
  constexpr bool kLN = Cfg::kLN;
  constexpr int kAcc = Cfg::kAcc;
  constexpr int kEpiThreads = Cfg::kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (threadIdx.x == 0 && (smem_base - smem_u32(smem_raw)) + Cfg::kNeeded > Cfg::kSmemBytes) {
    printf("rtdf: tc_gemm smem layout does not fit (base misaligned by %u)\n", smem_base - smem_u32(smem_raw));
    __trap();
  }
  // layout: stages | epilogue staging boxes (1024-aligned) | barriers (256 B) | params
  constexpr int kOffStaging = kStages * Cfg::kStageBytes;
  constexpr int kOffBars = kOffStaging + Cfg::kStagingBytes;
  constexpr int kOffParams = kOffBars + 256;
  const uint32_t bar_base = smem_base + kOffBars;
  // barriers: full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBars + 8 * (2 * kStages + 4));
  float* s_params = reinterpret_cast<float*>(smem_gen + kOffParams);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    if (p.epi_mode != EPI_DIRECT) prefetch_tmap(&mapC);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), Cfg::kEpiWarps);   // one elected arrive per epilogue warp
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
  const bool stream_mode = p.stream_mode != 0;
  if (!stream_mode) pdl_wait();   // the previous kernel's output (our A operand / residual) is complete and visible
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int tiles_n = p.tiles_n, tiles_m = p.tiles_m;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t pre = 0;
      if (stream_mode) {
        // Weight tiles of the first ring round: requested before the wait on the previous kernel.  The stage barrier
        // expects the bytes of both operands; the A tile follows after the wait.
        for (int t = blockIdx.x; t < p.total_tiles && pre < (uint32_t)kStages; t += gridDim.x) {
          const int tile = p.reverse ? p.total_tiles - 1 - t : t;
        const int split = tile % p.k_splits, tt = tile / p.k_splits;
          const int n0 = (tt % tiles_n) * BN;
          const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
          for (int kb = kb0; kb < kb1 && pre < (uint32_t)kStages; ++kb, ++pre) {
            mbar_expect_tx(full_bar(pre), Cfg::kStageBytes);
            const uint32_t b_dst = smem_base + pre * Cfg::kStageBytes + Cfg::kABytes;
#pragma unroll
            for (int h = 0; h < (BN + 255) / 256; ++h)
              tma_load_2d(b_dst + h * 256 * BK * 2, &mapB, full_bar(pre), kb * BK, n0 + h * 256);
          }
        }
        pdl_wait();
      }
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tile = p.reverse ? p.total_tiles - 1 - t : t;
        const int split = tile % p.k_splits, tt = tile / p.k_splits;
        const int n_tile = tt % tiles_n, m_tile = (tt / tiles_n) % tiles_m, batch = tt / (tiles_n * tiles_m);
        const int m0 = m_tile * BM, n0 = n_tile * BN;
        const int a_col0 = n_tile * p.a_col_per_ntile;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          const uint32_t a_dst = smem_base + s * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          if (it >= pre) {
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
          }
          if (p.a_hint)
            tma_load_3d_hint(a_dst, &mapA, full_bar(s), a_col0 + kb * p.a_kb_col_step,
                             m0 + p.a_row_off + kb * p.a_kb_row_step, batch,
                             p.a_hint == 2 ? l2_policy_evict_last() : l2_policy_evict_first());
          else
            tma_load_3d(a_dst, &mapA, full_bar(s), a_col0 + kb * p.a_kb_col_step,
                        m0 + p.a_row_off + kb * p.a_kb_row_step, batch);
          if (it >= pre) {
#pragma unroll
            for (int h = 0; h < (BN + 255) / 256; ++h)
              tma_load_2d(b_dst + h * 256 * BK * 2, &mapB, full_bar(s), kb * BK, n0 + h * 256);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && stream_mode) pdl_launch_dependents();   // dependents wait for our completion at their pdl_wait()
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
        const int kb0 = ((p.reverse ? p.total_tiles - 1 - t : t) % p.k_splits) * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
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
              mma_bf16_ss(d_tmem + h * 256, adesc, bdesc, idesc, ((kb - kb0) | k) != 0);
            }
          }
          mma_commit(empty_bar(s));  // smem slot reusable once these MMAs retire
        }
        mma_commit(tfull_bar(a));    // accumulator stage complete
      }
      if (!stream_mode) pdl_launch_dependents();   // all MMAs of this CTA are issued: the next kernel may start its prologue
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    if (stream_mode) pdl_wait();      // residual reads / output writes follow the previous kernel
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;  // 0 .. kEpiThreads-1
    uint8_t* box_gen = smem_gen + kOffStaging + (warp - 2) * 4096;
    const uint32_t box = smem_base + kOffStaging + (warp - 2) * 4096;
    const int mode = p.epi_mode;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const int tile = p.reverse ? p.total_tiles - 1 - t : t;
        const int split = tile % p.k_splits, tt = tile / p.k_splits;
      const int n_tile = tt % tiles_n, m_tile = (tt / tiles_n) % tiles_m, batch = tt / (tiles_n * tiles_m);
      const int m0 = m_tile * BM, n0 = n_tile * BN;
      const int a = lt % kAcc;
      const uint32_t aph = (lt / kAcc) & 1;
      const int row0 = m0 + q * 32;                 // first row of this warp inside the batch
      const int row_local = row0 + lane;
      const bool row_ok = row_local < p.rows_per_batch;
      const long long row = static_cast<long long>(batch) * p.rows_per_batch + row_local;
      if (!kLN) {
        // stage this tile's bias slice (double-buffered by accumulator stage)
        float* sb = s_params + a * BN;
        for (int i = et; i < BN; i += kEpiThreads)      // the bias rides on the first K split only
          sb[i] = (p.epi.bias && split == 0 && n0 + i < p.N) ? p.epi.bias[n0 + i] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      const LnFold fold = kLN ? LnFold() : ln_fold_row(p.epi, row, row_ok);   // row statistics: fetched ahead of the accumulator
      mbar_wait(tfull_bar(a), aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (kLN ? 0 : a * BN);
      if (!kLN) {
        epilogue_plain_tile<BN>(p, &mapC, mode, t_row, s_params + a * BN, half, lane, box_gen, box, n0, row0,
                                p.k_splits > 1 ? split : batch, row, row_ok, fold, n_tile);
      } else {
        // y = act(LayerNorm_512(acc + bias)).  Two warps share a row (256 columns each): one TMEM pass for
        // (sum, sum of squares), partials exchanged through smem, one pass to normalise / activate / store.
        const float* s_bias = s_params;
        const float* s_gamma = s_params + 512;
        const float* s_beta = s_params + 1024;
        float2* s_stat = reinterpret_cast<float2*>(s_params + 1536) + (lt & 1) * 256;   // [half][128 rows]
        const int c_begin = half * 256;
        float2 sum2 = make_float2(0.f, 0.f), ssq2 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = c_begin; c < c_begin + 256; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 v = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]);
            sum2 = add2(sum2, v);
            ssq2 = fma2(v, v, ssq2);
          }
        }
        const float sum = sum2.x + sum2.y, ssq = ssq2.x + ssq2.y;
        s_stat[half * 128 + q * 32 + lane] = make_float2(sum, ssq);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        const float2 other = s_stat[(half ^ 1) * 128 + q * 32 + lane];
        const float mean = (sum + other.x) * (1.0f / 512.0f);
        const float var = fmaxf((ssq + other.y) * (1.0f / 512.0f) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.epi.ln_eps);
        const float2 rstd2 = make_float2(rstd, rstd), nmean2 = make_float2(-mean, -mean);
#pragma unroll 1
        for (int c = c_begin; c < c_begin + 256; c += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + c, r0);
          tmem_ld32(t_row + c + 32, r1);
          tmem_ld_wait();
          __align__(8) float o[64];
          float2* o2 = reinterpret_cast<float2*>(o);
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c);
          const float2* g2 = reinterpret_cast<const float2*>(s_gamma + c);
          const float2* be2 = reinterpret_cast<const float2*>(s_beta + c);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t* r = i < 16 ? r0 : r1;
            const int j = i & 15;
            const float2 v = add2(add2(make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), b2[i]), nmean2);
            o2[i] = fma2(v, mul2(g2[i], rstd2), be2[i]);
          }
          if (p.epi.act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o2[i] = gelu2(o2[i]);
          } else if (p.epi.act == ACT_GELU_TANH) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o2[i] = gelu2_tanh(o2[i]);
          } else if (p.epi.act == ACT_GELU_AS) {
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = gelu_fast(o[i]);
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          stage_row_bf16(box_gen, lane, o);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.rows_per_batch) {
            tma_store_3d(&mapC, box, c, row0, batch);
            tma_store_commit();
          }
        }
      }
      // all TMEM reads of this stage are complete (tmem_ld_wait above): hand the stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(a));
      if (!kLN && p.epi.rowln_counters)
        rowln_after_tile<kEpiThreads>(p, m_tile, m0, et, lane, reinterpret_cast<volatile int*>(smem_gen + kOffBars + 240));
    }
    if (lane == 0) tma_store_wait_all<0>();   // outstanding TMA stores complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// =================================================================================================
// CTA-pair kernel (cta_group::2): a cluster of two CTAs on one TPC owns a 256 x 256 output tile.  Each CTA loads its
// own 128 A rows and ITS HALF of the W rows (32 KB per k-block instead of 48 KB for the same math per SM), the
// leader CTA issues one M = 256 tcgen05.mma per k-step for both SMs, and each CTA runs the epilogue of its own 128
// accumulator rows.  The 128 x 256 single-CTA tile is bound by the L2 -> shared-memory feed (profiles/r01_notes.md);
// the pair tile needs 1.5x fewer bytes per flop.
//   full[s]   (leader's copy)  : leader producer arrive.expect_tx(2 x stage) ; both CTAs' TMA loads complete_tx on it
//   empty[s]  (each CTA's copy): multicast tcgen05.commit from the leader's MMA thread
//   tfull[a]  (each CTA's copy): multicast tcgen05.commit ; tempty[a] (leader's copy): one arrive per epilogue warp of
//                                both CTAs (remote mbarrier.arrive for the peer)
// =================================================================================================
template <int BK>
struct Tc2Cfg {
  static constexpr int BN = 256;
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = 128 * BK * 2;               // this CTA's half of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kParamBytes = 2 * BN * 4;
  static constexpr int kFixed = kStagingBytes + 256 + kParamBytes;
  static constexpr int kStagesRaw = (kMaxSmem - 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed + 1024;
};

struct PairUnit { int n_tile, m_tile, batch, n0, width; };
__device__ __forceinline__ PairUnit pair_unit(const TcKernelParams& p, int u, int BN) {
  if (p.reverse) u = p.total_units - 1 - u;
  int tile = u, sub = 0, width = BN;
  if (u >= p.tail_start && p.tail_sub > 1) {
    const int v = u - p.tail_start;
    tile = p.tail_start + v / p.tail_sub;
    sub = v % p.tail_sub;
    width = BN / p.tail_sub;
  }
  PairUnit r;
  r.n_tile = tile % p.tiles_n;
  r.m_tile = (tile / p.tiles_n) % p.tiles_m;
  r.batch = tile / (p.tiles_n * p.tiles_m);
  r.n0 = r.n_tile * BN + sub * width;
  r.width = width;
  return r;
}

template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg<BK>::kThreads, 1)
tc_gemm_2sm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapBn,
                   const TcKernelParams p) {
  using Cfg = Tc2Cfg<BK>;
  constexpr int BN = Cfg::BN;
  constexpr int kStages = Cfg::kStages;
  constexpr int kEpiThreads = Cfg::kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kOffStaging = kStages * Cfg::kStageBytes;
  constexpr int kOffBars = kOffStaging + Cfg::kStagingBytes;
  constexpr int kOffParams = kOffBars + 256;
  const uint32_t bar_base = smem_base + kOffBars;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBars + 8 * (2 * kStages + 4));
  float* s_params = reinterpret_cast<float*>(smem_gen + kOffParams);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    if (p.epi_mode != EPI_DIRECT) prefetch_tmap(&mapC);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * Cfg::kEpiWarps);     // one elected arrive per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int tiles_n = p.tiles_n, tiles_m = p.tiles_m;       // tiles_m counts 256-row pair tiles
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs) =====
      const uint64_t a_policy = p.a_hint == 2 ? l2_policy_evict_last() : l2_policy_evict_first();
      uint32_t it = 0;
      for (int t = pair; t < p.total_units; t += n_pairs) {
        const PairUnit u = pair_unit(p, t, BN);
        const int m0 = u.m_tile * 256 + (int)rank * 128, batch = u.batch;
        const bool narrow = u.width != BN;                   // column slice of a tail tile: this CTA's half = width / 2 W rows
        const int nb0 = u.n0 + (int)rank * (u.width >> 1);
        const uint32_t unit_bytes = 2u * (Cfg::kABytes + (uint32_t)(u.width >> 1) * BK * 2);
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (p.debug & 1) {             // timing experiment: operands are whatever the shared memory holds
            if (leader) mbar_arrive(full_bar(s));
            continue;
          }
          if (leader) mbar_expect_tx(full_bar(s), unit_bytes);
          const uint32_t full_leader = mapa_shared(full_bar(s), 0);
          const uint32_t a_dst = smem_base + s * Cfg::kStageBytes;
          if (p.a_hint) tma_load_3d_2sm_hint(a_dst, &mapA, full_leader, kb * BK, m0, batch, a_policy);
          else tma_load_3d_2sm(a_dst, &mapA, full_leader, kb * BK, m0, batch);
          tma_load_2d_2sm(a_dst + Cfg::kABytes, narrow ? &mapBn : &mapB, full_leader, kb * BK, nb0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader CTA only) =====
      constexpr uint32_t idesc_full = umma_idesc_bf16(256, BN);
      const uint32_t idesc_narrow = umma_idesc_bf16(256, BN / (p.tail_sub > 1 ? p.tail_sub : 1));
      uint32_t it = 0, lt = 0;
      for (int t = pair; t < p.total_units; t += n_pairs, ++lt) {
        const uint32_t idesc = pair_unit(p, t, BN).width != BN ? idesc_narrow : idesc_full;
        const int a = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(tempty_bar(a), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
          if (!(p.debug & 2)) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              mma_bf16_ss_2sm(d_tmem, make_desc(a_addr + k * 32, BK), make_desc(b_addr + k * 32, BK), idesc, (kb | k) != 0);
          }
          mma_commit_2sm(empty_bar(s), 3);   // both CTAs' smem slots
        }
        mma_commit_2sm(tfull_bar(a), 3);     // both CTAs' accumulator halves
      }
      pdl_launch_dependents();
    }
  } else {
    // ===== epilogue warps (both CTAs): this CTA's 128 accumulator rows =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    uint8_t* box_gen = smem_gen + kOffStaging + (warp - 2) * 4096;
    const uint32_t box = smem_base + kOffStaging + (warp - 2) * 4096;
    const int mode = p.epi_mode;
    uint32_t lt = 0;
    for (int t = pair; t < p.total_units; t += n_pairs, ++lt) {
      const PairUnit u = pair_unit(p, t, BN);
      const int n_tile = u.n_tile, m_tile = u.m_tile, batch = u.batch;
      const int m0 = m_tile * 256 + (int)rank * 128, n0 = u.n0;
      const int a = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      const int row0 = m0 + q * 32;
      const int row_local = row0 + lane;
      const bool row_ok = row_local < p.rows_per_batch;
      const long long row = static_cast<long long>(batch) * p.rows_per_batch + row_local;
      float* sb = s_params + a * BN;
      for (int i = et; i < BN; i += kEpiThreads) sb[i] = (p.epi.bias && n0 + i < p.N) ? p.epi.bias[n0 + i] : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      const LnFold fold = ln_fold_row(p.epi, row, row_ok);      // row statistics: fetched ahead of the accumulator
      mbar_wait(tfull_bar(a), aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
      if (!(p.debug & 4)) {
        if (u.width == BN)
          epilogue_plain_tile<BN>(p, &mapC, mode, t_row, sb, half, lane, box_gen, box, n0, row0, batch, row, row_ok, fold, n_tile);
        else   // column slice of a tail tile (<= 64 columns): the warps of column half 0 take it, 64 columns each
          epilogue_plain_tile<128>(p, &mapC, mode, t_row, sb, half, lane, box_gen, box, n0, row0, batch, row, row_ok, fold, n_tile,
                                   n0 + u.width);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(a), 0));
      if (p.epi.rowln_counters)
        rowln_after_tile<kEpiThreads>(p, m_tile * 2 + (int)rank, m0, et, lane, reinterpret_cast<volatile int*>(smem_gen + kOffBars + 240));
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();            // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}


// =================================================================================================
// Full-row (N = 512) tile with fused bias + LayerNorm(512) + GELU, PIPELINED: the 512 accumulator columns are produced
// as two 256-column passes over K (h = 0, 1), each with its own tfull / tempty barrier.  The 16 epilogue warps take the
// row statistics of half 0 while the tensor pipe works on half 1, and the MMAs of the next tile's half 0 start as soon
// as half 0 has been normalised and stored -- the tensor pipe only waits for (stats of half 1 + pass 2 of half 0)
// instead of the whole epilogue (tc_gemm_kernel<512,*> runs it at 48 % tensor-active, profiles/r01_ncu_summary_v8.txt).
//   warp 0: TMA producer (A tile is re-loaded for the second pass; W half h per pass)   warp 1: MMA issuer
//   warps 2..17: epilogue; lane quarter q = warp % 4, column slice j = (warp - 2) / 4: columns [64 j, 64 j + 64) of
//   each half; partial (sum, sum of squares) exchanged through smem.
// =================================================================================================
template <int BK>
struct TcLnCfg {
  static constexpr int kEpiWarps = 16;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = 256 * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kParamBytes = 3 * 512 * 4 + 2 * 4 * 128 * 8;     // bias|gamma|beta + stats[2][4][128] float2
  static constexpr int kFixed = kStagingBytes + 256 + kParamBytes;
  static constexpr int kStagesRaw = (kMaxSmem - 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed + 1024;
};

template <int BK>
__global__ void __launch_bounds__(TcLnCfg<BK>::kThreads, 1)
tc_conv_ln_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapC, const TcKernelParams p) {
  using Cfg = TcLnCfg<BK>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kEpiThreads = Cfg::kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kOffStaging = kStages * Cfg::kStageBytes;
  constexpr int kOffBars = kOffStaging + Cfg::kStagingBytes;
  constexpr int kOffParams = kOffBars + 256;
  const uint32_t bar_base = smem_base + kOffBars;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int h) { return bar_base + 8u * (2 * kStages + h); };
  auto tempty_bar = [&](int h) { return bar_base + 8u * (2 * kStages + 2 + h); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBars + 8 * (2 * kStages + 4));
  float* s_params = reinterpret_cast<float*>(smem_gen + kOffParams);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapC);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(tfull_bar(h), 1);
      mbar_init(tempty_bar(h), Cfg::kEpiWarps);      // one elected arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 512; i += Cfg::kThreads) {
    s_params[i] = p.epi.bias ? p.epi.bias[i] : 0.f;
    s_params[512 + i] = p.epi.ln_gamma[i];
    s_params[1024 + i] = p.epi.ln_beta[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int total_m = p.tiles_m * (p.total_tiles / (p.tiles_m * p.tiles_n));   // m-tiles x batches

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_m; t += gridDim.x) {
        const int m_tile = t % p.tiles_m, batch = t / p.tiles_m;
        const int m0 = m_tile * BM;
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
            const uint32_t a_dst = smem_base + s * Cfg::kStageBytes;
            tma_load_3d(a_dst, &mapA, full_bar(s), kb * BK, m0, batch);
            tma_load_2d(a_dst + Cfg::kABytes, &mapB, full_bar(s), kb * BK, h * 256);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(BM, 256);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < total_m; t += gridDim.x, ++lt) {
        for (int h = 0; h < 2; ++h) {
          mbar_wait(tempty_bar(h), (lt & 1) ^ 1);     // half h of the previous tile has been normalised and stored
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + h * 256;
          for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_addr = smem_base + s * Cfg::kStageBytes;
            const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              mma_bf16_ss(d_tmem, make_desc(a_addr + k * 32, BK), make_desc(b_addr + k * 32, BK), idesc, (kb | k) != 0);
            mma_commit(empty_bar(s));
          }
          mma_commit(tfull_bar(h));
        }
      }
      pdl_launch_dependents();
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int j = (warp - 2) >> 2;
    uint8_t* box_gen = smem_gen + kOffStaging + (warp - 2) * 4096;
    const uint32_t box = smem_base + kOffStaging + (warp - 2) * 4096;
    const float* s_bias = s_params;
    const float* s_gamma = s_params + 512;
    const float* s_beta = s_params + 1024;
    float2* s_stat_all = reinterpret_cast<float2*>(s_params + 1536);      // [2][4][128]
    const int act = p.epi.act;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < total_m; t += gridDim.x, ++lt) {
      const int m_tile = t % p.tiles_m, batch = t / p.tiles_m;
      const int row0 = m_tile * BM + q * 32;
      const uint32_t tph = lt & 1;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      float2* s_stat = s_stat_all + (lt & 1) * 512;
      // ---- pass 1: row statistics of this warp's 64 + 64 columns (half 0 while the tensor pipe fills half 1) ----
      float2 sum2 = make_float2(0.f, 0.f), ssq2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(tfull_bar(h), tph);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int c = h * 256 + j * 64 + sub * 32;
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 v = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]);
            sum2 = add2(sum2, v);
            ssq2 = fma2(v, v, ssq2);
          }
        }
      }
      s_stat[j * 128 + q * 32 + lane] = make_float2(sum2.x + sum2.y, ssq2.x + ssq2.y);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      float sum = 0.f, ssq = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float2 o = s_stat[jj * 128 + q * 32 + lane];
        sum += o.x;
        ssq += o.y;
      }
      const float mean = sum * (1.0f / 512.0f);
      const float var = fmaxf(ssq * (1.0f / 512.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.epi.ln_eps);
      const float2 rstd2 = make_float2(rstd, rstd), nmean2 = make_float2(-mean, -mean);
      // ---- pass 2: normalise / GELU / store, half 0 first so that the next tile's MMAs can start on it ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int c0 = h * 256 + j * 64;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int c = c0 + sub * 32;
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          __align__(8) float o[32];
          float2* o2 = reinterpret_cast<float2*>(o);
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c);
          const float2* g2 = reinterpret_cast<const float2*>(s_gamma + c);
          const float2* be2 = reinterpret_cast<const float2*>(s_beta + c);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 v = add2(add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]), nmean2);
            o2[i] = fma2(v, mul2(g2[i], rstd2), be2[i]);
          }
          if (act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o2[i] = gelu2(o2[i]);
          } else if (act == ACT_GELU_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o2[i] = gelu2_tanh(o2[i]);
          } else if (act == ACT_GELU_AS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = gelu_fast(o[i]);
          }
          // 32 bf16 = 64 bytes = chunks 4*sub .. 4*sub+3 of this lane's 128-byte box row (SWIZZLE_128B)
          uint8_t* rowp = box_gen + lane * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(rowp + (((4 * sub + ch) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16x2(o[8 * ch], o[8 * ch + 1]), pack_bf16x2(o[8 * ch + 2], o[8 * ch + 3]),
                           pack_bf16x2(o[8 * ch + 4], o[8 * ch + 5]), pack_bf16x2(o[8 * ch + 6], o[8 * ch + 7]));
        }
        // all TMEM reads of half h by this warp are complete: hand it back to the MMA warp
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tempty_bar(h));
          if (row0 < p.rows_per_batch) {
            tma_store_3d(&mapC, box, c0, row0, batch);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================================
// CTA-pair (cta_group::2) version of the pipelined full-row LayerNorm tile: a 2-CTA cluster owns 256 rows x 512 columns.
// The single-CTA kernel streams the whole weight matrix (512 x K) plus the A tile twice per 128 rows -- 2.3 MB per tile
// for conv-1 (K = 1536), 7.4 GB per launch at B = 64, i.e. L2-bandwidth bound (615 us at ~12 TB/s; measured 600 us).
// Here every CTA loads its 128 A rows and only ITS HALF of the 256 W rows of a pass; the leader issues M = 256 MMAs, so
// the weight traffic per output row halves.  Barrier protocol as in tc_gemm_2sm_kernel; epilogue as in
// tc_conv_ln_kernel (each CTA normalises its own 128 rows).
// =================================================================================================
template <int BK>
struct TcLn2Cfg {
  static constexpr int kEpiWarps = 16;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = 128 * BK * 2;              // this CTA's half of the 256 W rows of a pass
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kParamBytes = 3 * 512 * 4 + 2 * 4 * 128 * 8;     // bias|gamma|beta + stats[2][4][128] float2
  static constexpr int kFixed = kStagingBytes + 256 + kParamBytes;
  static constexpr int kStagesRaw = (kMaxSmem - 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed + 1024;
};

template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TcLn2Cfg<BK>::kThreads, 1)
tc_conv_ln_2sm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapC, const TcKernelParams p) {
  using Cfg = TcLn2Cfg<BK>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kEpiThreads = Cfg::kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kOffStaging = kStages * Cfg::kStageBytes;
  constexpr int kOffBars = kOffStaging + Cfg::kStagingBytes;
  constexpr int kOffParams = kOffBars + 256;
  const uint32_t bar_base = smem_base + kOffBars;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int h) { return bar_base + 8u * (2 * kStages + h); };
  auto tempty_bar = [&](int h) { return bar_base + 8u * (2 * kStages + 2 + h); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBars + 8 * (2 * kStages + 4));
  float* s_params = reinterpret_cast<float*>(smem_gen + kOffParams);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapC);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(tfull_bar(h), 1);
      mbar_init(tempty_bar(h), 2 * Cfg::kEpiWarps);  // one elected arrive per epilogue warp of both CTAs (leader's copy)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  for (int i = threadIdx.x; i < 512; i += Cfg::kThreads) {
    s_params[i] = p.epi.bias ? p.epi.bias[i] : 0.f;
    s_params[512 + i] = p.epi.ln_gamma[i];
    s_params[1024 + i] = p.epi.ln_beta[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int total_m = p.total_tiles;                 // 256-row pair tiles x batches
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;
      for (int t = pair; t < total_m; t += n_pairs) {
        const int m_tile = t % p.tiles_m, batch = t / p.tiles_m;
        const int m0 = m_tile * 256 + (int)rank * 128;
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            if (leader) mbar_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);
            const uint32_t full_leader = mapa_shared(full_bar(s), 0);
            const uint32_t a_dst = smem_base + s * Cfg::kStageBytes;
            tma_load_3d_2sm(a_dst, &mapA, full_leader, kb * BK, m0, batch);
            tma_load_2d_2sm(a_dst + Cfg::kABytes, &mapB, full_leader, kb * BK, h * 256 + (int)rank * 128);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader CTA only): M = 256 over the pair =====
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
      uint32_t it = 0, lt = 0;
      for (int t = pair; t < total_m; t += n_pairs, ++lt) {
        for (int h = 0; h < 2; ++h) {
          mbar_wait(tempty_bar(h), (lt & 1) ^ 1);     // half h of the previous tile has been normalised and stored
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + h * 256;
          for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_addr = smem_base + s * Cfg::kStageBytes;
            const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              mma_bf16_ss_2sm(d_tmem, make_desc(a_addr + k * 32, BK), make_desc(b_addr + k * 32, BK), idesc, (kb | k) != 0);
            mma_commit_2sm(empty_bar(s), 3);
          }
          mma_commit_2sm(tfull_bar(h), 3);
        }
      }
      pdl_launch_dependents();
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int j = (warp - 2) >> 2;
    uint8_t* box_gen = smem_gen + kOffStaging + (warp - 2) * 4096;
    const uint32_t box = smem_base + kOffStaging + (warp - 2) * 4096;
    const float* s_bias = s_params;
    const float* s_gamma = s_params + 512;
    const float* s_beta = s_params + 1024;
    float2* s_stat_all = reinterpret_cast<float2*>(s_params + 1536);      // [2][4][128]
    const int act = p.epi.act;
    uint32_t lt = 0;
    for (int t = pair; t < total_m; t += n_pairs, ++lt) {
      const int m_tile = t % p.tiles_m, batch = t / p.tiles_m;
      const int row0 = m_tile * 256 + (int)rank * 128 + q * 32;
      const uint32_t tph = lt & 1;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      float2* s_stat = s_stat_all + (lt & 1) * 512;
      // ---- pass 1: row statistics of this warp's 64 + 64 columns (half 0 while the tensor pipe fills half 1) ----
      float2 sum2 = make_float2(0.f, 0.f), ssq2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(tfull_bar(h), tph);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int c = h * 256 + j * 64 + sub * 32;
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 v = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]);
            sum2 = add2(sum2, v);
            ssq2 = fma2(v, v, ssq2);
          }
        }
      }
      s_stat[j * 128 + q * 32 + lane] = make_float2(sum2.x + sum2.y, ssq2.x + ssq2.y);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      float sum = 0.f, ssq = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float2 o = s_stat[jj * 128 + q * 32 + lane];
        sum += o.x;
        ssq += o.y;
      }
      const float mean = sum * (1.0f / 512.0f);
      const float var = fmaxf(ssq * (1.0f / 512.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.epi.ln_eps);
      const float2 rstd2 = make_float2(rstd, rstd), nmean2 = make_float2(-mean, -mean);
      // ---- pass 2: normalise / GELU / store, half 0 first so that the next tile's MMAs can start on it ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int c0 = h * 256 + j * 64;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int c = c0 + sub * 32;
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
          __align__(8) float o[32];
          float2* o2 = reinterpret_cast<float2*>(o);
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c);
          const float2* g2 = reinterpret_cast<const float2*>(s_gamma + c);
          const float2* be2 = reinterpret_cast<const float2*>(s_beta + c);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 v = add2(add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]), nmean2);
            o2[i] = fma2(v, mul2(g2[i], rstd2), be2[i]);
          }
          if (act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o2[i] = gelu2(o2[i]);
          } else if (act == ACT_GELU_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o2[i] = gelu2_tanh(o2[i]);
          } else if (act == ACT_GELU_AS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = gelu_fast(o[i]);
          }
          // 32 bf16 = 64 bytes = chunks 4*sub .. 4*sub+3 of this lane's 128-byte box row (SWIZZLE_128B)
          uint8_t* rowp = box_gen + lane * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(rowp + (((4 * sub + ch) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16x2(o[8 * ch], o[8 * ch + 1]), pack_bf16x2(o[8 * ch + 2], o[8 * ch + 3]),
                           pack_bf16x2(o[8 * ch + 4], o[8 * ch + 5]), pack_bf16x2(o[8 * ch + 6], o[8 * ch + 7]));
        }
        // all TMEM reads of half h by this warp are complete: hand it back to the MMA warp
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_shared(tempty_bar(h), 0));
          if (row0 < p.rows_per_batch) {
            tma_store_3d(&mapC, box, c0, row0, batch);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();            // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int g_gelu_override = 0;
int tc_get_gelu_variant() { return g_gelu_override; }
void tc_set_gelu_variant(int act) { g_gelu_override = (act == ACT_GELU_TANH || act == ACT_GELU_AS) ? act : 0; }

static bool force_direct_epilogue() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTDF_TC_EPILOGUE");
    v = (e && e[0] == 'd') ? 1 : 0;   // RTDF_TC_EPILOGUE=direct disables the TMA-store epilogue (debug / A-B runs)
  }
  return v == 1;
}

// Output path: a single output tensor goes through TMA stores (bf16 / fp32) or the TMA reduce-add (in-place residual).
static int choose_epilogue_mode(TcKernelParams& p, CUtensorMap* mapC, const CUtensorMap* unused_map, const TcOperandA& A, int N,
                                const TcEpilogue& epi, bool ln_variant, int BN) {
  p.epi_mode = EPI_DIRECT;
  const bool one_f32 = epi.out_f32 && !epi.out_bf16 && (epi.ld_f32 % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(epi.out_f32) % 16 == 0);
  const bool one_bf16 = epi.out_bf16 && !epi.out_f32 && !epi.resid && (epi.ld_bf16 % 8 == 0) &&
                        (reinterpret_cast<uintptr_t>(epi.out_bf16) % 16 == 0);
  if (epi.xb_out) {
    RTDF_REQUIRE(!ln_variant && BN >= 64 && one_f32 && epi.resid == epi.out_f32 && epi.ldr == epi.ld_f32 && epi.ld_f32 == N &&
                 N % BN == 0 && N / BN * 2 <= 8 && A.batches == 1 && epi.stats_out && !epi.partials && !epi.rowln_counters &&
                 (reinterpret_cast<uintptr_t>(epi.xb_out) % 16 == 0),
                 "tc_gemm: the bf16-copy + row-statistics epilogue needs an in-place fp32 residual with dense rows, "
                 "N a multiple of the tile width and at most 8 (tile, half) slots per row");
    p.epi_mode = EPI_XRES;
  } else if (ln_variant) {
    RTDF_REQUIRE(one_bf16, "tc_gemm: the LayerNorm variant writes exactly one bf16 output (16-byte aligned rows)");
    p.epi_mode = EPI_TMA_BF16;
  } else if (!force_direct_epilogue()) {
    if (one_f32 && !epi.resid) p.epi_mode = EPI_TMA_F32;
    else if (one_f32 && epi.resid == epi.out_f32 && epi.ldr == epi.ld_f32) p.epi_mode = EPI_TMA_F32_ADD;
    else if (one_bf16 && BN >= 128) p.epi_mode = EPI_TMA_BF16;
  }
  if (epi.fold_stats)
    RTDF_REQUIRE(epi.fold_c && N % 32 == 0 && A.batches == 1 && !ln_variant && !epi.partials &&
                 (reinterpret_cast<uintptr_t>(epi.fold_c) % 16 == 0) && (reinterpret_cast<uintptr_t>(epi.fold_stats) % 16 == 0),
                 "tc_gemm: folded LayerNorm needs column sums, N a multiple of 32 and one batch");
  if (p.epi_mode == EPI_TMA_F32 || p.epi_mode == EPI_TMA_F32_ADD || p.epi_mode == EPI_XRES) {
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)A.rows_per_batch, (uint64_t)A.batches};
    uint64_t strides[2] = {(uint64_t)epi.ld_f32 * 4, (uint64_t)epi.ld_f32 * 4 * (uint64_t)A.rows_per_batch};
    uint32_t box[3] = {32, 32, 1};
    RTDF_TRY(make_tmap_f32(mapC, epi.out_f32, 3, dims, strides, box, TMAP_SW128));
  } else if (p.epi_mode == EPI_TMA_BF16) {
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)A.rows_per_batch, (uint64_t)A.batches};
    uint64_t strides[2] = {(uint64_t)epi.ld_bf16 * 2, (uint64_t)epi.ld_bf16 * 2 * (uint64_t)A.rows_per_batch};
    uint32_t box[3] = {64, 32, 1};
    RTDF_TRY(make_tmap_bf16(mapC, epi.out_bf16, 3, dims, strides, box, TMAP_SW128));
  } else {
    *mapC = *unused_map;
  }
  if (epi.rowln_counters) {
    RTDF_REQUIRE(!ln_variant && (BN == 256 || BN == 64) && (p.epi_mode == EPI_TMA_F32 || p.epi_mode == EPI_TMA_F32_ADD),
                 "tc_gemm: the fused row LayerNorm needs a 256- or 64-wide variant with a single fp32 output");
    RTDF_REQUIRE(A.batches == 1 && N % 128 == 0 && N <= 1024 && epi.ld_f32 == N && epi.rowln_gamma && epi.rowln_beta &&
                 (epi.rowln_out_bf16 || epi.rowln_out_f32),
                 "tc_gemm: fused row LayerNorm: N must be a multiple of 128 (<= 1024), dense rows, one batch");
  }
  return RTDF_OK;
}

template <int BN, int BK>
static int launch_variant(cudaStream_t stream, const TcOperandA& A, const bf16* W, int N, int Kw, int mode,
                          const TcEpilogue& epi) {
  using Cfg = TcCfg<BN, BK>;
  static_assert(Cfg::kStages >= 2, "need at least two stages");
  static_assert(8 * (2 * Cfg::kStages + 5) <= 256, "barrier block too small");
  CUtensorMap mapA, mapB, mapC;
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
  p.reverse = epi.reverse_tiles ? 1 : 0;
  p.a_hint = epi.a_cache_hint;
  if (p.epi.act == ACT_GELU && g_gelu_override) p.epi.act = g_gelu_override;
  RTDF_TRY(choose_epilogue_mode(p, &mapC, &mapA, A, N, epi, Cfg::kLN, BN));
  p.stream_mode = (pdl_enabled() && !epi.rowln_counters && stream_prefetch_enabled()) ? 1 : 0;
  p.k_splits = 1;
  p.kb_per_split = p.num_kb;
  if (epi.partials) {
    // Deterministic split-K for skinny GEMMs: split s writes its partial sum (bias on split 0) to partials[s] with plain
    // TMA stores; the consumer (layernorm_accum_rows) folds them into the residual stream in a fixed order.
    RTDF_REQUIRE(mode == TC_PLAIN && !Cfg::kLN && BN == 64 && A.batches == 1 && epi.act == ACT_NONE && !epi.resid &&
                 !epi.out_bf16 && !epi.rowln_counters && epi.scale == 1.0f,
                 "tc_gemm: split-K partials need a plain 64-wide GEMM without activation / residual / bf16 output");
    const int splits = tc_plan_splits(A.rows_per_batch, N, Kw);
    RTDF_REQUIRE(splits > 1, "tc_gemm: split-K requested for a shape that tc_plan_splits() does not split");
    p.kb_per_split = ceil_div(p.num_kb, splits);
    p.k_splits = ceil_div(p.num_kb, p.kb_per_split);
    RTDF_REQUIRE(p.k_splits == splits, "tc_gemm: inconsistent split plan");
    p.total_tiles *= p.k_splits;
    p.epi.out_f32 = epi.partials;
    p.epi.ld_f32 = N;
    p.epi_mode = EPI_TMA_F32;
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)A.rows_per_batch, (uint64_t)p.k_splits};
    uint64_t strides[2] = {(uint64_t)N * 4, (uint64_t)N * 4 * (uint64_t)A.rows_per_batch};
    uint32_t box[3] = {32, 32, 1};
    RTDF_TRY(make_tmap_f32(&mapC, epi.partials, 3, dims, strides, box, TMAP_SW128));
  }
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  ProfRec rec{};
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.a));
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.b));
    rec.flops = 2.0 * (double)A.rows_per_batch * (double)A.batches * (double)N * (double)Kw;
    rec.variant = epi.profile_as_wide ? 256 : BN + (BK == 32 ? 1 : 0);
    RTDF_CHECK_CUDA(cudaEventRecord(rec.a, stream));
  }
  RTDF_CHECK_CUDA(launch_pdl(tc_gemm_kernel<BN, BK>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, mapA, mapB,
                             mapC, p));
  RTDF_LAUNCH_CHECK();
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventRecord(rec.b, stream));
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back(rec);
  }
  return RTDF_OK;
}


template <int BK>
static int launch_2sm(cudaStream_t stream, const TcOperandA& A, const bf16* W, int N, int Kw, const TcEpilogue& epi) {
  using Cfg = Tc2Cfg<BK>;
  static_assert(Cfg::kStages >= 3, "need at least three stages");
  static_assert(8 * (2 * Cfg::kStages + 5) <= 256, "barrier block too small");
  CUtensorMap mapA, mapB, mapBn, mapC;
  {
    uint64_t dims[3] = {(uint64_t)A.k_extent, (uint64_t)A.rows_per_batch, (uint64_t)A.batches};
    uint64_t strides[2] = {(uint64_t)A.row_stride * 2, (uint64_t)(A.batches > 1 ? A.batch_stride : A.row_stride * A.rows_per_batch) * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    RTDF_TRY(make_tmap_bf16(&mapA, A.ptr, 3, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  {
    uint64_t dims[2] = {(uint64_t)Kw, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)Kw * 2};
    uint32_t box[2] = {(uint32_t)BK, 128};
    RTDF_TRY(make_tmap_bf16(&mapB, W, 2, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
    box[1] = 32;       // this CTA's half of a 64-column slice of a tail tile
    RTDF_TRY(make_tmap_bf16(&mapBn, W, 2, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  TcKernelParams p;
  p.rows_per_batch = (int)A.rows_per_batch;
  p.N = N;
  p.num_kb = ceil_div(Kw, BK);
  p.tiles_n = ceil_div(N, 256);
  p.tiles_m = ceil_div((int)A.rows_per_batch, 256);
  const long long total = (long long)p.tiles_n * p.tiles_m * A.batches;
  RTDF_REQUIRE(total < (1LL << 29), "tc_gemm: too many tiles");
  p.total_tiles = (int)total;
  p.a_kb_col_step = BK; p.a_kb_row_step = 0; p.a_row_off = 0; p.a_col_per_ntile = 0;
  p.epi = epi;
  p.reverse = epi.reverse_tiles ? 1 : 0;
  p.a_hint = epi.a_cache_hint;
  if (p.epi.act == ACT_GELU && g_gelu_override) p.epi.act = g_gelu_override;
  RTDF_TRY(choose_epilogue_mode(p, &mapC, &mapA, A, N, epi, false, 256));
  // Wave quantisation: with more tiles than CTA pairs the last round of the persistent loop is partly empty (out_proj /
  // fc2 at the timed batch: 200 tiles on 74 pairs = 2.7 rounds).  The tiles of that round are cut into four 64-column
  // slices each -- same kernel, narrower MMAs -- when the model says the round gets cheaper: a slice costs about 0.32 of
  // a tile (a quarter of the math, the A tile loaded again).
  const int n_pairs_h = p.total_tiles < kNumSMs / 2 ? p.total_tiles : kNumSMs / 2;
  p.tail_start = p.total_tiles;
  p.tail_sub = 1;
  p.total_units = p.total_tiles;
  {
    const int full_rounds = p.total_tiles / n_pairs_h, rem = p.total_tiles % n_pairs_h;
    const bool plain_epi = !epi.xb_out && !epi.rowln_counters && !epi.partials && N % 256 == 0;
    if (tail_slices_enabled() && plain_epi && full_rounds >= 1 && rem > 0) {
      const int slice_rounds = ceil_div(rem * 4, n_pairs_h);
      if (0.32 * slice_rounds < 1.0) {
        p.tail_start = full_rounds * n_pairs_h;
        p.tail_sub = 4;
        p.total_units = p.tail_start + rem * 4;
      }
    }
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("RTDF_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_2sm_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  int grid = 2 * (p.total_tiles < kNumSMs / 2 ? p.total_tiles : kNumSMs / 2);
  ProfRec rec{};
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.a));
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.b));
    rec.flops = 2.0 * (double)A.rows_per_batch * (double)A.batches * (double)N * (double)Kw;
    rec.variant = 256;
    RTDF_CHECK_CUDA(cudaEventRecord(rec.a, stream));
  }
  RTDF_CHECK_CUDA(launch_pdl(tc_gemm_2sm_kernel<BK>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, mapA, mapB,
                             mapC, mapBn, p));
  RTDF_LAUNCH_CHECK();
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventRecord(rec.b, stream));
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back(rec);
  }
  return RTDF_OK;
}

template <int BK>
static int launch_conv_ln(cudaStream_t stream, const TcOperandA& A, const bf16* W, int Kw, const TcEpilogue& epi) {
  using Cfg = TcLnCfg<BK>;
  static_assert(Cfg::kStages >= 3, "need at least three stages");
  static_assert(8 * (2 * Cfg::kStages + 5) <= 256, "barrier block too small");
  const int N = 512;
  CUtensorMap mapA, mapB, mapC;
  {
    uint64_t dims[3] = {(uint64_t)A.k_extent, (uint64_t)A.rows_per_batch, (uint64_t)A.batches};
    uint64_t strides[2] = {(uint64_t)A.row_stride * 2, (uint64_t)(A.batches > 1 ? A.batch_stride : A.row_stride * A.rows_per_batch) * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    RTDF_TRY(make_tmap_bf16(&mapA, A.ptr, 3, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  {
    uint64_t dims[2] = {(uint64_t)Kw, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)Kw * 2};
    uint32_t box[2] = {(uint32_t)BK, 256};
    RTDF_TRY(make_tmap_bf16(&mapB, W, 2, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  TcKernelParams p;
  p.rows_per_batch = (int)A.rows_per_batch;
  p.N = N;
  p.num_kb = ceil_div(Kw, BK);
  p.tiles_n = 1;
  p.tiles_m = ceil_div((int)A.rows_per_batch, BM);
  const long long total = (long long)p.tiles_m * A.batches;
  RTDF_REQUIRE(total < (1LL << 31), "tc_gemm: too many tiles");
  p.total_tiles = (int)total;
  p.a_kb_col_step = BK; p.a_kb_row_step = 0; p.a_row_off = 0; p.a_col_per_ntile = 0;
  p.epi = epi;
  p.reverse = epi.reverse_tiles ? 1 : 0;
  p.a_hint = epi.a_cache_hint;
  if (p.epi.act == ACT_GELU && g_gelu_override) p.epi.act = g_gelu_override;
  RTDF_TRY(choose_epilogue_mode(p, &mapC, &mapA, A, N, epi, true, 512));
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_ln_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  ProfRec rec{};
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.a));
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.b));
    rec.flops = 2.0 * (double)A.rows_per_batch * (double)A.batches * (double)N * (double)Kw;
    rec.variant = 514;
    RTDF_CHECK_CUDA(cudaEventRecord(rec.a, stream));
  }
  RTDF_CHECK_CUDA(launch_pdl(tc_conv_ln_kernel<BK>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, mapA, mapB,
                             mapC, p));
  RTDF_LAUNCH_CHECK();
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventRecord(rec.b, stream));
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back(rec);
  }
  return RTDF_OK;
}

template <int BK>
static int launch_conv_ln_2sm(cudaStream_t stream, const TcOperandA& A, const bf16* W, int Kw, const TcEpilogue& epi) {
  using Cfg = TcLn2Cfg<BK>;
  static_assert(Cfg::kStages >= 3, "need at least three stages");
  static_assert(8 * (2 * Cfg::kStages + 5) <= 256, "barrier block too small");
  const int N = 512;
  CUtensorMap mapA, mapB, mapC;
  {
    uint64_t dims[3] = {(uint64_t)A.k_extent, (uint64_t)A.rows_per_batch, (uint64_t)A.batches};
    uint64_t strides[2] = {(uint64_t)A.row_stride * 2, (uint64_t)(A.batches > 1 ? A.batch_stride : A.row_stride * A.rows_per_batch) * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    RTDF_TRY(make_tmap_bf16(&mapA, A.ptr, 3, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  {
    uint64_t dims[2] = {(uint64_t)Kw, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)Kw * 2};
    uint32_t box[2] = {(uint32_t)BK, 128};
    RTDF_TRY(make_tmap_bf16(&mapB, W, 2, dims, strides, box, BK == 64 ? TMAP_SW128 : TMAP_SW64));
  }
  TcKernelParams p;
  p.rows_per_batch = (int)A.rows_per_batch;
  p.N = N;
  p.num_kb = ceil_div(Kw, BK);
  p.tiles_n = 1;
  p.tiles_m = ceil_div((int)A.rows_per_batch, 256);
  const long long total = (long long)p.tiles_m * A.batches;
  RTDF_REQUIRE(total < (1LL << 31), "tc_gemm: too many tiles");
  p.total_tiles = (int)total;
  p.a_kb_col_step = BK; p.a_kb_row_step = 0; p.a_row_off = 0; p.a_col_per_ntile = 0;
  p.epi = epi;
  p.reverse = epi.reverse_tiles ? 1 : 0;
  p.a_hint = epi.a_cache_hint;
  if (p.epi.act == ACT_GELU && g_gelu_override) p.epi.act = g_gelu_override;
  RTDF_TRY(choose_epilogue_mode(p, &mapC, &mapA, A, N, epi, true, 512));
  RTDF_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_ln_2sm_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int grid = 2 * (p.total_tiles < kNumSMs / 2 ? p.total_tiles : kNumSMs / 2);
  ProfRec rec{};
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.a));
    RTDF_CHECK_CUDA(cudaEventCreate(&rec.b));
    rec.flops = 2.0 * (double)A.rows_per_batch * (double)A.batches * (double)N * (double)Kw;
    rec.variant = 516;
    RTDF_CHECK_CUDA(cudaEventRecord(rec.a, stream));
  }
  RTDF_CHECK_CUDA(launch_pdl(tc_conv_ln_2sm_kernel<BK>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, mapA, mapB,
                             mapC, p));
  RTDF_LAUNCH_CHECK();
  if (g_prof_on) {
    RTDF_CHECK_CUDA(cudaEventRecord(rec.b, stream));
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back(rec);
  }
  return RTDF_OK;
}

int tc_plan_splits(long long rows, int N, int Kw) {
  const long long tiles = ((rows + BM - 1) / BM) * ((N + 63) / 64);
  const int num_kb = (Kw + 63) / 64;
  long long want = tiles > 0 ? kNumSMs / tiles : 1;
  if (want > num_kb / 2) want = num_kb / 2;
  if (want > 8) want = 8;
  if (want < 2) return 1;
  const int per = (num_kb + (int)want - 1) / (int)want;
  return (num_kb + per - 1) / per;
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
    case 2256:   // CTA-pair (cta_group::2) 256 x 256 tiles
      RTDF_REQUIRE(mode == TC_PLAIN, "tc_gemm: the CTA-pair variant implements plain GEMMs only");
      return launch_2sm<64>(stream, A, W, N, Kw, epi);
    case 514:   // pipelined full-row LayerNorm tile (two 256-column passes, 16 epilogue warps)
    case 515:
      RTDF_REQUIRE(N == 512 && epi.ln_gamma && epi.ln_beta && mode == TC_PLAIN, "tc_gemm: LN variant needs N == 512 and LN parameters");
      return variant == 514 ? launch_conv_ln<32>(stream, A, W, Kw, epi) : launch_conv_ln<64>(stream, A, W, Kw, epi);
    case 516:   // CTA-pair version of the pipelined LayerNorm tile
      RTDF_REQUIRE(N == 512 && epi.ln_gamma && epi.ln_beta && mode == TC_PLAIN, "tc_gemm: LN variant needs N == 512 and LN parameters");
      return launch_conv_ln_2sm<64>(stream, A, W, Kw, epi);
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
