#pragma once
#include "gemm_tc.cuh"

namespace rtdf {

struct SimtOperandA {
  const float* ptr = nullptr;
  long long k_extent = 0;
  long long rows_per_batch = 0;
  long long batches = 1;
  long long row_stride = 0;
  long long batch_stride = 0;
};

// D = epilogue(A * W^T), everything fp32, FFMA accumulation in k order within 16-wide blocks.
int simt_gemm_f32(cudaStream_t stream, const SimtOperandA& A, const float* W, int N, int K, const TcEpilogue& e);

}  // namespace rtdf
