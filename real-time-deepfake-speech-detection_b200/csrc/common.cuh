// Shared helpers for the rtdf sm_100a kernels: error plumbing, math, reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace rtdf {

// ---- error plumbing (thread-local message surfaced by rtdf_last_error) -----------------
void set_error(const char* fmt, ...);
const char* get_error();
long long launch_count();

#define RTDF_OK 0
#define RTDF_ERR_INVALID (-1)
#define RTDF_ERR_CUDA (-2)
#define RTDF_ERR_STATE (-3)
#define RTDF_ERR_UNSUPPORTED (-4)

#define RTDF_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      rtdf::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                      cudaGetErrorString(_e));                                             \
      return RTDF_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define RTDF_REQUIRE(cond, ...)         \
  do {                                  \
    if (!(cond)) {                      \
      rtdf::set_error(__VA_ARGS__);     \
      return RTDF_ERR_INVALID;          \
    }                                   \
  } while (0)

#define RTDF_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != RTDF_OK) return _r; \
  } while (0)

// every kernel launch of the library passes through here (rtdf_launch_count() reports the total)
void count_launch();
#define RTDF_LAUNCH_CHECK()              \
  do {                                   \
    rtdf::count_launch();                \
    RTDF_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

typedef __nv_bfloat16 bf16;

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device math ------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// GELU with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): 2 MUFU + ~12 FMA instead of the
// ~30-instruction branchy erff.  Used where the result is rounded to bf16 anyway (tensor-core epilogues).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_z = poly * t * __expf(-z * z);     // erfc(|x|/sqrt2)
  const float erf_x = copysignf(1.0f - erfc_z, x);
  return 0.5f * x * (1.0f + erf_x);
}
__device__ __forceinline__ float selu_f(float x) {
  const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
  return x > 0.f ? scale * x : scale * alpha * expm1f(x);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigmoid_f(x); }

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SWISH = 2, ACT_SELU = 3 };
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(x);
    case ACT_SWISH: return swish_f(x);
    case ACT_SELU: return selu_f(x);
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// typed load/store used by kernels templated on the activation type (float | bf16)
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace rtdf
