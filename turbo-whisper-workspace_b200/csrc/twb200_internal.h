// Internal glue shared by the .cu translation units of libtwb200.so.
#pragma once
#include "../../include/twb200.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>

namespace tw {
void set_error(const char* fmt, ...);
// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency, so the
// library loads on hosts without a driver for the symbol-export check).
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle);
int num_sms();
// Binds the device that owns `device_ptr` to the calling host thread (checked on every call: one host thread may serve
// several devices through the C ABI).  The library links
// its own static cudart: a host thread in which PyTorch has not yet made a context current would otherwise reach the
// driver without a context (cuTensorMapEncodeTiled -> CUDA_ERROR_INVALID_CONTEXT) or launch on device 0.
int ensure_device(const void* device_ptr);
// device currently bound to this host thread (0..63); kernel attributes (max dynamic smem) are per device
int current_device();
// per-device one-time setup (kernel attributes): an atomic bit per device, safe across the engine-context threads
bool device_needs_setup(std::atomic<unsigned long long>& done_mask);
void mark_device_done(std::atomic<unsigned long long>& done_mask);
// cross_attn.cu: persistent TMA-fed decoder cross-attention; 0 = launched, -1 = layout not supported (caller falls back
// to the scalar kernel of decode.cu), other = error (tw_last_error)
int cross_attn_stream_launch(const void* q, void* out, const void* k, const void* v, long long row_stride,
                             long long batch_stride, long long head_stride, const int* enc_row, int S, int B, int H,
                             int split_cap, float* part, unsigned int* counters, cudaStream_t stream, bool pdl);
}  // namespace tw
