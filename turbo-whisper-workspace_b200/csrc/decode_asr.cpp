// Host-side stitching of per-window Whisper token streams into timestamped chunks.
//
// Restates the behaviour of transformers' `_decode_asr` and `_find_longest_common_sequence`
// ($TF/models/whisper/tokenization_whisper.py:901-1150 and :1153-1270 in the survey's numbering; 5.5.0) for
// return_timestamps in {False, True}: the timestamp / stride state machine (timestamps inside a stride are skipped
// in pairs, times are offset by chunk_len - stride_right per window and by the length of earlier seek segments
// inside a window, chunks are closed on end timestamps), and the sliding best-overlap merge of the token runs of
// consecutive windows (score = matches / i + i / 10000, needs more than one match, split at the mid-points).
// Token ids in, token ids + times out: the byte-level text decode stays with the caller, which owns the vocabulary.
// Pure integer / double arithmetic in the same operation order as the Python original, and Python's round(x, 2) is
// reproduced through the correctly rounded "%.2f" conversion, so times are bit-identical.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "twb200_internal.h"

// same contract as the device files' TW_REQUIRE (common.cuh is CUDA-only): record the message, return status 2
#define TW_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            tw::set_error(__VA_ARGS__); \
            return 2;                  \
        }                              \
    } while (0)

namespace {

using Run = std::vector<int32_t>;

double py_round2(double x) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.2f", x);
    return std::strtod(buf, nullptr);
}

// Merge the token runs of consecutive windows.  For every window boundary all alignments "last i tokens region"
// are scored; the best one (if it has at least two equal tokens) decides where the left run stops and the right
// run starts (the mid-points of the overlap: the left window is trusted for the left half, the right one for the rest).
Run merge_runs(const std::vector<Run>& runs) {
    Run total;
    if (runs.empty()) return total;
    Run left = runs[0];
    for (size_t s = 1; s < runs.size(); ++s) {
        const Run& right = runs[s];
        const long ll = (long)left.size(), rl = (long)right.size();
        double best = 0.0;
        long b_ls = ll, b_le = ll, b_rs = 0, b_re = 0;
        for (long i = 1; i < ll + rl; ++i) {
            const long ls = std::max(0L, ll - i), le = std::min(ll, ll + rl - i);
            const long rs = std::max(0L, i - ll);
            long matches = 0;
            for (long k = 0; k < le - ls; ++k) matches += left[ls + k] == right[rs + k];
            const double score = (double)matches / (double)i + (double)i / 10000.0;
            if (matches > 1 && score > best) {
                best = score;
                b_ls = ls; b_le = le; b_rs = rs; b_re = std::min(rl, i);
            }
        }
        const long lmid = (b_le + b_ls) / 2, rmid = (b_re + b_rs) / 2;
        total.insert(total.end(), left.begin(), left.begin() + lmid);
        left.assign(right.begin() + rmid, right.end());
    }
    total.insert(total.end(), left.begin(), left.end());
    return total;
}

struct Chunk {
    bool has_t0 = false, has_t1 = false;
    double t0 = 0.0, t1 = 0.0;
    int32_t lang = -1;
    Run tokens;
};

}  // namespace

extern "C" int tw_decode_asr(const tw_asr_window* windows, int32_t n_windows, const tw_asr_config* cfg,
                             int32_t* out_tokens, int64_t out_tokens_cap, int64_t* chunk_offsets, double* chunk_t0,
                             double* chunk_t1, int32_t* chunk_lang, int32_t max_chunks, int32_t* n_chunks_out,
                             int32_t* flags_out) {
    using namespace tw;
    TW_REQUIRE(cfg && n_chunks_out && (n_windows == 0 || windows), "tw_decode_asr: null argument");
    TW_REQUIRE(cfg->n_special == 0 || (cfg->special_ids && cfg->special_lang), "tw_decode_asr: special id table missing");
    TW_REQUIRE(cfg->time_precision > 0.0, "tw_decode_asr: time_precision must be positive");
    const int32_t ts_begin = cfg->timestamp_begin;
    const double prec = cfg->time_precision;
    const bool with_ts = cfg->return_timestamps != 0;

    auto special_lang = [&](int32_t tok, int32_t* lang) -> bool {   // binary search in the sorted special id table
        int32_t lo = 0, hi = cfg->n_special;
        while (lo < hi) {
            const int32_t mid = (lo + hi) / 2;
            if (cfg->special_ids[mid] < tok) lo = mid + 1; else hi = mid;
        }
        if (lo < cfg->n_special && cfg->special_ids[lo] == tok) { *lang = cfg->special_lang[lo]; return true; }
        return false;
    };

    std::vector<Chunk> chunks;
    Chunk chunk;
    int32_t last_language = -1;
    auto fresh_chunk = [&]() { Chunk c; c.lang = last_language; return c; };
    double time_offset = 0.0;
    std::vector<Run> previous;   // token runs waiting to be merged into the open chunk
    bool skip = false;
    int32_t flags = 0;

    for (int32_t w = 0; w < n_windows; ++w) {
        const tw_asr_window& win = windows[w];
        TW_REQUIRE(win.n_tokens == 0 || win.tokens, "tw_decode_asr: window %d has no token pointer", w);
        const int32_t* ids = win.tokens;
        int32_t n = win.n_tokens;
        // a leading <|startofprev|> ... prompt is dropped up to <|startoftranscript|>
        if (n > 0 && ids[0] == cfg->prompt_token_id) {
            int32_t k = 0;
            while (k < n && ids[k] != cfg->decoder_start_token_id) ++k;
            ids += k;
            n -= k;   // k == n: nothing left
        }
        bool have_last_ts = false;
        int32_t last_ts = 0;
        double first_ts = (double)ts_begin;
        double cur_max = 0.0, prev_segments = 0.0, penultimate = 0.0;
        double right_stride_start = 0.0;
        if (win.has_stride) {
            time_offset -= win.stride_left;
            right_stride_start = win.chunk_len - win.stride_right;
            if (win.stride_left != 0.0) first_ts = win.stride_left / prec + (double)ts_begin;
            if (win.stride_right != 0.0) {
                // timestamps that fall into the right stride: the last timestamp of the window always does
                for (int32_t i = n - 1; i >= 0; --i) {
                    const int32_t t = ids[i];
                    if (t >= ts_begin) {
                        if (have_last_ts && (double)(t - ts_begin) * prec < right_stride_start) break;
                        last_ts = t;
                        have_last_ts = true;
                    }
                }
            }
        }
        Run current;
        for (int32_t i = 0; i < n; ++i) {
            const int32_t t = ids[i];
            int32_t lang = -1;
            if (special_lang(t, &lang)) {
                if (lang >= 0) {
                    if (last_language >= 0 && lang != last_language && !with_ts) {
                        // language switch without timestamps: close the chunk here
                        previous.push_back(current);
                        chunk.tokens = merge_runs(previous);
                        chunks.push_back(chunk);
                        previous.clear();
                        current.clear();
                        chunk = fresh_chunk();
                    }
                    chunk.lang = lang;
                    last_language = lang;
                }
                // every other special token is ignored
            } else if (t >= ts_begin) {
                const double stamp = (double)((double)(t - ts_begin) * prec);
                if (stamp < cur_max) {
                    // a new seek segment of the same window has started: its timestamps restart near zero
                    const bool single_ending = i >= 2 && !(ids[i - 1] >= ts_begin && ids[i - 2] >= ts_begin);
                    if (single_ending) {
                        prev_segments += prec * (double)cfg->segment_size;
                    } else {
                        cur_max = penultimate;
                        prev_segments += penultimate;
                    }
                }
                penultimate = cur_max;
                cur_max = stamp;
                const double time = py_round2((double)(t - ts_begin) * prec + time_offset + prev_segments);
                if (have_last_ts && t >= last_ts) {
                    skip = true;    // inside the right stride: skipped together with its partner
                } else if (skip || (!previous.empty() && (double)t < first_ts)) {
                    skip = false;
                } else if (!chunk.has_t0) {
                    chunk.t0 = time;
                    chunk.has_t0 = true;
                } else if (time == chunk.t0) {
                    // duplicated start timestamp: stays a start
                } else {
                    chunk.t1 = time;
                    chunk.has_t1 = true;
                    previous.push_back(current);
                    chunk.tokens = merge_runs(previous);
                    chunks.push_back(chunk);
                    previous.clear();
                    current.clear();
                    chunk = fresh_chunk();
                }
            } else {
                current.push_back(t);
            }
        }
        if (win.has_stride) time_offset += win.chunk_len - win.stride_right;
        if (!current.empty()) {
            previous.push_back(current);
        } else {
            bool any = false;
            for (const Run& r : previous) any = any || !r.empty();
            if (!any) {
                chunk = fresh_chunk();
                previous.clear();
            }
        }
    }
    if (!previous.empty()) {
        if (with_ts) flags |= 1;   // no closing timestamp predicted (audio cut mid-word, or timestamps not generated)
        chunk.tokens = merge_runs(previous);
        chunks.push_back(chunk);
    }

    *n_chunks_out = (int32_t)chunks.size();
    if (flags_out) *flags_out = flags;
    TW_REQUIRE((int32_t)chunks.size() <= max_chunks, "tw_decode_asr: %d chunks exceed the caller's capacity %d",
               (int)chunks.size(), (int)max_chunks);
    TW_REQUIRE(chunks.empty() || (out_tokens && chunk_offsets && chunk_t0 && chunk_t1 && chunk_lang),
               "tw_decode_asr: null output array");
    int64_t off = 0;
    for (size_t c = 0; c < chunks.size(); ++c) {
        const Chunk& ch = chunks[c];
        TW_REQUIRE(off + (int64_t)ch.tokens.size() <= out_tokens_cap, "tw_decode_asr: token capacity %lld too small",
                   (long long)out_tokens_cap);
        chunk_offsets[c] = off;
        for (int32_t t : ch.tokens) out_tokens[off++] = t;
        chunk_t0[c] = ch.has_t0 ? ch.t0 : std::nan("");
        chunk_t1[c] = ch.has_t1 ? ch.t1 : std::nan("");
        chunk_lang[c] = ch.lang;
    }
    if (chunk_offsets) chunk_offsets[chunks.size()] = off;
    return 0;
}
