// Host: dynamic time warping of the token x frame alignment matrix -> one encoder frame per token.
// Restates _dynamic_time_warping and the jump extraction of _extract_token_timestamps
// ($TF/models/whisper/generation_whisper.py:64-112, 367-369), which HF runs as a Python double loop on the CPU
// (output_length x input_length iterations per window).  Same arithmetic: the cost table is float32, each cell is
// the float64 sum of the negated matrix entry and the float32 predecessor cost rounded back to float32, and the
// predecessor is chosen with the same strict comparisons (diagonal, then up, else left).
#include "twb200_internal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>
#include <thread>
#include <vector>

// one window; returns 0, or 3 on an unexpected trace value
static int dtw_one(const float* matrix, int64_t ld, int32_t n_tok, int32_t n_frames, int32_t* token_frame) {
    const int64_t W = (int64_t)n_frames + 1;
    std::vector<float> cost((size_t)(n_tok + 1) * W, std::numeric_limits<float>::infinity());
    std::vector<int8_t> trace((size_t)(n_tok + 1) * W, -1);
    cost[0] = 0.f;
    for (int32_t i = 1; i <= n_tok; ++i) {
        const float* mrow = matrix + (int64_t)(i - 1) * ld;
        const float* up = cost.data() + (int64_t)(i - 1) * W;
        float* cur = cost.data() + (int64_t)i * W;
        int8_t* tr = trace.data() + (int64_t)i * W;
        for (int32_t j = 1; j <= n_frames; ++j) {
            const float c0 = up[j - 1], c1 = up[j], c2 = cur[j - 1];
            float c;
            int8_t t;
            if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else { c = c2; t = 2; }
            cur[j] = (float)(-(double)mrow[j - 1] + (double)c);
            tr[j] = t;
        }
    }
    for (int64_t j = 0; j < W; ++j) trace[j] = 2;
    for (int32_t i = 0; i <= n_tok; ++i) trace[(int64_t)i * W] = 1;
    // backtrace; the path visits every token index, and the token's frame is the first (smallest) frame of its run
    int32_t i = n_tok, j = n_frames;
    for (int32_t t = 0; t < n_tok; ++t) token_frame[t] = -1;
    while (i > 0 || j > 0) {
        // overwritten while the run of token i-1 continues towards earlier frames; on the table's border (only reached
        // with NaN costs) HF records frame index -1 the same way
        if (i >= 1) token_frame[i - 1] = j - 1;
        const int8_t t = trace[(int64_t)i * W + j];
        if (t == 0) { --i; --j; }
        else if (t == 1) { --i; }
        else if (t == 2) { --j; }
        else return 3;
    }
    return 0;
}

extern "C" int tw_dtw_token_frames(const float* matrix, int64_t ld, int32_t n_tok, int32_t n_frames,
                                   int32_t* token_frame) {
    using tw::set_error;
    if (!matrix || !token_frame) { set_error("tw_dtw_token_frames: null argument"); return 2; }
    if (n_tok < 1 || n_frames < 1 || ld < n_frames) { set_error("tw_dtw_token_frames: bad shape %d x %d (ld %lld)", n_tok, n_frames, (long long)ld); return 2; }
    if (dtw_one(matrix, ld, n_tok, n_frames, token_frame)) { set_error("tw_dtw_token_frames: unexpected trace value"); return 3; }
    return 0;
}

// batch of windows on up to n_threads host threads (windows are independent); rows with n_frames[b] == 0 get frame -1
// for every token, which is what HF's backtrace yields on an empty frame axis
extern "C" int tw_dtw_token_frames_batch(const float* matrices, int64_t batch_stride, int64_t ld, int32_t batch,
                                         int32_t n_tok, const int32_t* n_frames, int32_t* token_frames,
                                         int32_t n_threads) {
    using tw::set_error;
    if (!matrices || !n_frames || !token_frames) { set_error("tw_dtw_token_frames_batch: null argument"); return 2; }
    if (batch < 0 || n_tok < 1) { set_error("tw_dtw_token_frames_batch: bad batch %d / tokens %d", batch, n_tok); return 2; }
    for (int32_t b = 0; b < batch; ++b)
        if (n_frames[b] < 0 || n_frames[b] > ld) { set_error("tw_dtw_token_frames_batch: row %d has %d frames (ld %lld)", b, n_frames[b], (long long)ld); return 2; }
    std::atomic<int32_t> next(0), failed(0);
    auto work = [&]() {
        for (;;) {
            const int32_t b = next.fetch_add(1);
            if (b >= batch) return;
            int32_t* out = token_frames + (int64_t)b * n_tok;
            if (n_frames[b] == 0) {
                for (int32_t t = 0; t < n_tok; ++t) out[t] = -1;
            } else if (dtw_one(matrices + (int64_t)b * batch_stride, ld, n_tok, n_frames[b], out)) {
                failed.store(1);
            }
        }
    };
    const int32_t nt = std::max(1, std::min(n_threads, batch));
    std::vector<std::thread> pool;
    for (int32_t i = 1; i < nt; ++i) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (failed.load()) { set_error("tw_dtw_token_frames_batch: unexpected trace value"); return 3; }
    return 0;
}
