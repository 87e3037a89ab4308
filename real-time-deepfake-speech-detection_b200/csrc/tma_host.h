// Host-side TMA tensor-map construction (driver entry point fetched through the runtime, so the
// library does not link libcuda).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace rtdf {

enum TmapSwizzle { TMAP_SW128 = 0, TMAP_SW64 = 1, TMAP_SW_NONE = 2 };

// bf16 tensor, rank 2 or 3.  dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride
// of dims[i+1] (multiple of 16).  Out-of-bound box elements are zero-filled (coordinates may be
// negative).  Rows of a box are box[0] elements = 128 B (SW128) or 64 B (SW64); any multiple of 16 B without swizzle.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle sw);
// same for fp32 tensors (TMA store / reduce-add targets)
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle sw);

}  // namespace rtdf
