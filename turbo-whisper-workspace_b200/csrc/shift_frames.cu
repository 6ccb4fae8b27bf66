// Seek-shifted window gather for the short-form seek loop
// (WhisperGenerationMixin._get_input_segment, $TF/models/whisper/generation_whisper.py:1831-1850):
// the next 30 s window of a row starts at mel frame seek[b] and is zero-padded (feature value 0)
// to 3000 frames.  Operates on the time-major bf16 feature layout [*, rows, cols]; row_off skips
// the leading conv-padding row, which stays zero.
#include "common.cuh"
#include "twb200_internal.h"

namespace tw {
__global__ void __launch_bounds__(256) shift_frames_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                          const int* __restrict__ src_row,
                                                          const int* __restrict__ seek, int frames, int vec_per_row,
                                                          long long batch_stride_vec, int row_off) {
    const int b = blockIdx.y;
    const int sb = src_row ? src_row[b] : b;
    const int sk = seek[b];
    const long long total = (long long)frames * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / vec_per_row), c = (int)(i % vec_per_row);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t + sk < frames) v = src[sb * batch_stride_vec + (long long)(row_off + sk + t) * vec_per_row + c];
        dst[b * batch_stride_vec + (long long)(row_off + t) * vec_per_row + c] = v;
    }
}
}  // namespace tw

extern "C" int tw_shift_frames(const void* src_bf16, void* dst_bf16, const int32_t* src_row, const int32_t* seek,
                               int32_t batch, int32_t frames, int32_t cols, int64_t batch_stride, int32_t row_off,
                               void* stream) {
    using namespace tw;
    TW_REQUIRE(src_bf16 && dst_bf16 && seek, "tw_shift_frames: null argument");
    if (tw::ensure_device(src_bf16)) return 1;
    TW_REQUIRE(cols % 8 == 0 && batch_stride % 8 == 0, "tw_shift_frames: cols and batch_stride must be multiples of 8");
    TW_REQUIRE(src_bf16 != dst_bf16, "tw_shift_frames: in-place shift is not supported");
    TW_REQUIRE(batch >= 0 && batch <= 65535, "tw_shift_frames: bad batch");
    if (batch == 0) return 0;
    dim3 grid(148, batch);
    shift_frames_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src_bf16, (uint4*)dst_bf16, src_row, seek,
                                                               frames, cols / 8, batch_stride / 8, row_off);
    TW_CUDA_CHECK(cudaGetLastError());
    return 0;
}
