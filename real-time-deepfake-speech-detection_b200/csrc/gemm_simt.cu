// fp32 SIMT GEMM (FFMA, no tensor cores): the arithmetic of the fp32 verification mode
// (max |logit diff| <= 1e-4 vs the fp32 PyTorch forward) and of the small fp32 back-end
// projections.  Same operand view / epilogue contract as the tcgen05 kernel.
#include "gemm_simt.cuh"

namespace rtdf {

constexpr int SB = 64;   // tile M = N
constexpr int SK = 16;   // tile K

template <typename TA>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const TA* __restrict__ A, long long k_extent, int rows_per_batch, long long row_stride,
                 long long batch_stride, const float* __restrict__ W, int N, int K, TcEpilogue e) {
  __shared__ float sA[SK][SB + 4];
  __shared__ float sW[SK][SB + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB, batch = blockIdx.z;
  const TA* Ab = A + batch * batch_stride;
  float acc[4][4] = {};
  const int lr = threadIdx.x >> 2;        // 0..63 : tile row loaded by this thread
  const int lk = (threadIdx.x & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += SK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      const int m = m0 + lr, n = n0 + lr;
      sA[lk + i][lr] = (m < rows_per_batch && k < K && k < k_extent) ? to_f32(Ab[m * row_stride + k]) : 0.f;
      sW[lk + i][lr] = (n < N && k < K) ? W[(long long)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = sW[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= rows_per_batch) continue;
    const long long row = (long long)batch * rows_per_batch + m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float x = acc[i][j];
      if (e.bias) x += e.bias[n];
      x = apply_act(x, e.act) * e.scale;
      if (e.resid) x += e.resid[row * e.ldr + n];
      if (e.out_f32) e.out_f32[row * e.ld_f32 + n] = x;
      if (e.out_bf16) e.out_bf16[row * e.ld_bf16 + n] = __float2bfloat16_rn(x);
    }
  }
}

template <typename TA>
static int launch(cudaStream_t stream, const TA* A, long long k_extent, long long rows, long long batches,
                  long long row_stride, long long batch_stride, const float* W, int N, int K, const TcEpilogue& e) {
  RTDF_REQUIRE(A && W && N > 0 && K > 0 && rows > 0 && batches > 0 && batches <= 65535, "simt_gemm: bad arguments");
  dim3 grid(ceil_div(N, SB), ceil_div((int)rows, SB), (unsigned)batches);
  simt_gemm_kernel<TA><<<grid, 256, 0, stream>>>(A, k_extent, (int)rows, row_stride, batch_stride, W, N, K, e);
  RTDF_LAUNCH_CHECK();
  return RTDF_OK;
}

int simt_gemm_f32(cudaStream_t stream, const SimtOperandA& A, const float* W, int N, int K, const TcEpilogue& e) {
  return launch<float>(stream, A.ptr, A.k_extent, A.rows_per_batch, A.batches, A.row_stride, A.batch_stride, W, N, K, e);
}

}  // namespace rtdf
