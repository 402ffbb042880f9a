// hrv_host.h — host-side helpers shared by the translation units of libhrviton_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/hrviton_sm100.h"

namespace hrv {
// Records a thread-local error string (printf-style) and returns `code`.
int set_error(int code, const char* fmt, ...);
int sm_count();
// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time dependency on libcuda).
int encode_tensor_map(CUtensorMap* map, int rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);
}  // namespace hrv
