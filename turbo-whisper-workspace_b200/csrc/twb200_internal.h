// Internal glue shared by the .cu translation units of libtwb200.so.
#pragma once
#include "../../include/twb200.h"
#include <cuda.h>
#include <cuda_runtime.h>

namespace tw {
void set_error(const char* fmt, ...);
// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency, so the
// library loads on hosts without a driver for the symbol-export check).
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle);
int num_sms();
}  // namespace tw
