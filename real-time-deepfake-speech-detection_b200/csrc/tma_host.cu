#include "tma_host.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

namespace rtdf {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// Programmatic dependent launch: RTDF_PDL=0/1 forces it off/on; otherwise the forward turns it on for small
// problems (streaming chunks), where kernels last a few microseconds and overlapping the next kernel's prologue
// (barrier init, TMEM allocation, tensor-map prefetch) with the current kernel's tail is worth ~5 % of the latency.
// At batch 64 it measured 2.5 % slower inside the CUDA graph (r01), so large problems keep plain stream order.
// thread-local: each forward call sets it on its own thread right before it enqueues its launches, so two host threads
// driving two contexts never see each other's choice
static thread_local int g_pdl_auto = 0;
void pdl_set_auto(bool on) { g_pdl_auto = on ? 1 : 0; }
bool pdl_enabled() {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("RTDF_PDL");
    forced = (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
  }
  return forced >= 0 ? forced == 1 : g_pdl_auto == 1;
}

cudaError_t raise_max_dyn_smem(const void* func, size_t bytes) {
  static std::map<std::pair<int, const void*>, size_t> cur;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  size_t& have = cur[std::make_pair(dev, func)];
  if (bytes <= have) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return RTDF_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw == TMAP_SW128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw == TMAP_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu] box [%u,%u,%u]",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
              (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
              rank > 2 ? box[2] : 0);
    return RTDF_ERR_CUDA;
  }
  return RTDF_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle sw) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, sw);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle sw) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, sw);
}

}  // namespace rtdf
