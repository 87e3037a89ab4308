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

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// Kernels of the per-layer chain are launched with the programmatic-stream-serialization attribute: their CTAs may
// become resident and run their prologue (barrier init, TMEM allocation, tensor-map prefetch) while the previous
// kernel drains; pdl_wait() then blocks until that kernel has completed and its writes are visible.  Every kernel
// launched through launch_pdl() MUST call pdl_wait() before its first global-memory access that depends on (or
// could overwrite the inputs of) the previous kernel.  On for small problems (B*T <= 1024 rows: streaming chunks, ~5 %
// lower latency), off for large ones (2.5 % slower inside the CUDA graph at batch 64); RTDF_PDL=0/1 forces it.
bool pdl_enabled();
void pdl_set_auto(bool on);   // set per forward call from the problem size (model.cu)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Opt-in dynamic shared memory of a kernel whose request varies from launch to launch: the per-function limit is only ever
// RAISED (per device).  Lowering it between launches is legal for plain launches but breaks tools that re-launch the kernel
// nodes of a captured CUDA graph with the function's current attributes (ncu: LaunchFailed on the node that needs more).
cudaError_t raise_max_dyn_smem(const void* func, size_t bytes);

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
// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two lanes per issue slot) -----------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GELU(x) = x * Phi(x) with Phi(x) ~= sigmoid(2 x (c1 + c3 x^2 + c5 x^4)); the odd polynomial is a minimax
// fit of the erf form (max |error| of the whole GELU 2.6e-5 over R, tools/fit_gelu.py), evaluated on two
// values per FFMA2 issue slot plus EX2 + RCP on the MUFU pipe.  -2*log2(e) is folded into the coefficients;
// x^2 is clamped to 64 (Phi(8) == 1 in fp32) so the quintic never turns over.
__device__ __forceinline__ float2 gelu2(float2 x) {
  const float k = -2.8853900817779268f;   // -2 / ln 2
  float2 s = mul2(x, x);
  s.x = fminf(s.x, 64.0f);
  s.y = fminf(s.y, 64.0f);
  float2 p = fma2(make_float2(k * -0.0003515189394188129f, k * -0.0003515189394188129f), s,
                  make_float2(k * 0.03700565997219061f, k * 0.03700565997219061f));
  p = fma2(p, s, make_float2(k * 0.7975078680535282f, k * 0.7975078680535282f));
  const float2 u = mul2(x, p);
  const float2 d = add2(make_float2(ex2_approx(u.x), ex2_approx(u.y)), make_float2(1.0f, 1.0f));
  return mul2(x, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
// tanh form on the hardware tanh (1 MUFU per value, |error| <= ~5e-4: only for A/B timing runs)
__device__ __forceinline__ float2 gelu2_tanh(float2 x) {
  const float2 s = mul2(x, x);
  const float2 p = fma2(make_float2(0.034701004943096476f, 0.034701004943096476f), s,
                        make_float2(0.8001568294135658f, 0.8001568294135658f));
  const float2 y = mul2(x, p);
  const float2 h = mul2(x, make_float2(0.5f, 0.5f));
  return fma2(h, make_float2(tanh_approx(y.x), tanh_approx(y.y)), h);
}
__device__ __forceinline__ float gelu_sig(float x) { return gelu2(make_float2(x, x)).x; }

__device__ __forceinline__ float selu_f(float x) {
  const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
  return x > 0.f ? scale * x : scale * alpha * expm1f(x);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigmoid_f(x); }

// ACT_GELU_TANH / ACT_GELU_AS select alternative GELU evaluations in the tensor-core epilogues (A/B timing);
// everywhere else they mean the exact erf GELU.
enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SWISH = 2, ACT_SELU = 3, ACT_GELU_TANH = 4, ACT_GELU_AS = 5 };
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU:
    case ACT_GELU_TANH:
    case ACT_GELU_AS: return gelu_erf(x);
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
