// Error reporting, device queries and the TMA descriptor encoder used by every kernel file.
#include "common.cuh"
#include "twb200_internal.h"
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include <mutex>

namespace tw {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
    });
    return fn;
}

int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver / not an sm_100 host)");
        return 1;
    }
    cuuint64_t gdims[5];
    cuuint64_t gstr[5];
    cuuint32_t gbox[5];
    cuuint32_t estr[5];
    for (uint32_t i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUresult r = fn(map, dtype, rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %u, dims %llu/%llu/%llu, "
                  "stride0 %llu, box %u/%u)",
                  (int)r, rank, (unsigned long long)dims[0],
                  (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0],
                  rank > 1 ? box[1] : 0);
        return 1;
    }
    return 0;
}

// Binds the calling host thread to the device that owns `device_ptr` — on EVERY call, so that one host thread can serve
// several GPUs through the C ABI (a plain C caller has no torch.cuda.device() guard around it).  The pointer query is
// ~0.2 us; cudaSetDevice only happens when the current device differs.
int ensure_device(const void* device_ptr) {
    if (device_ptr == nullptr) return 0;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, device_ptr);
    if (e != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        set_error("expected a device pointer (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "host memory");
        return 1;
    }
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur == attr.device) return 0;
    e = cudaSetDevice(attr.device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed: %s", attr.device, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

// true exactly until `mark_device_done` was called for the current device (kernel attributes are per device); safe to race:
// cudaFuncSetAttribute is idempotent, the flag word is atomic
bool device_needs_setup(std::atomic<unsigned long long>& done_mask) {
    return !((done_mask.load(std::memory_order_acquire) >> current_device()) & 1ull);
}
void mark_device_done(std::atomic<unsigned long long>& done_mask) {
    done_mask.fetch_or(1ull << current_device(), std::memory_order_release);
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return dev & 63;
}

int num_sms() {
    static std::atomic<int> cache[64];   // per device (zero-initialised): a thread may serve several devices
    const int dev = current_device();
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace tw

extern "C" const char* tw_last_error(void) { return tw::g_err; }
extern "C" int tw_abi_version(void) { return TW_ABI_VERSION; }
