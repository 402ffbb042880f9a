// hrv_host.h — host-side helpers shared by the translation units of libhrviton_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#ifdef HRV_F16
#include "hrv_f16_rename.h"  // the fp16-storage flavour of every entry point: hrv_xxx -> hrv_xxx_f16 (same signatures)
#endif
#include "../../include/hrviton_sm100.h"

// Helpers defined once (capi.cu) and shared by both storage flavours of the kernel translation units.  The kernel TUs are
// compiled twice: as namespace hrv (bf16 storage) and, with -DHRV_F16 -Dhrv=hrv_f16, as namespace hrv_f16 (IEEE fp16 storage).
namespace hrv_host {
// Records a thread-local error string (printf-style) and returns `code`.
int set_error(int code, const char* fmt, ...);
int sm_count();
// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time dependency on libcuda).
int encode_tensor_map(CUtensorMap* map, int rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);
}  // namespace hrv_host
namespace hrv {
using namespace hrv_host;
}
