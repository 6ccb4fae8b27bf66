// Host: native FLAC reader for the audio-ingest step ("GPU polyphase resampler + WAV/FLAC
// reader").  The reference hands file paths to the HF pipeline, whose `ffmpeg_read`
// ($TF/pipelines/audio_utils.py:9-45) spawns an `ffmpeg` subprocess per file; the reference's own example input is a
// FLAC file (ref:examples/Test1/ChrisAndAlexDiTest.flac).  This decodes the FLAC bit stream in-process to interleaved
// integer PCM, which the ingest kernel (resample.cu) converts, down-mixes and resamples on the GPU.
//
// Restates the published FLAC format (RFC 9639): STREAMINFO, frame header (UTF-8 coded number, CRC-8), subframes
// CONSTANT / VERBATIM / FIXED (order 0-4) / LPC (order 1-32) with wasted bits, partitioned Rice residuals (4- and 5-bit
// parameters, escape partitions), inter-channel decorrelation (left/side, right/side, mid/side), frame CRC-16.
// Integrity: every frame's CRC-8 / CRC-16 is checked here; the caller checks the MD5 of the decoded samples against
// STREAMINFO (Python hashlib).
#include "twb200_internal.h"

#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct BitReader {
    const uint8_t* p;
    int64_t n, pos;      // byte position of the next unread byte
    uint64_t acc;        // unread bits, left-aligned in the low `cnt` bits
    int cnt;
    bool bad;
    BitReader(const uint8_t* d, int64_t len, int64_t start) : p(d), n(len), pos(start), acc(0), cnt(0), bad(false) {}
    inline void refill() {
        while (cnt <= 56) {
            if (pos < n) acc = (acc << 8) | p[pos];
            else { acc <<= 8; if (pos >= n + 8) bad = true; }
            ++pos;
            cnt += 8;
        }
    }
    inline uint32_t bits(int k) {          // k in [0, 32]
        if (k == 0) return 0;
        if (cnt < k) refill();
        cnt -= k;
        return (uint32_t)((acc >> cnt) & ((k == 32) ? 0xffffffffull : ((1ull << k) - 1)));
    }
    inline int32_t sbits(int k) {
        if (k == 0) return 0;
        const uint32_t v = bits(k);
        const uint32_t m = 1u << (k - 1);
        return (int32_t)((v ^ m) - m);
    }
    inline uint32_t unary() {              // number of 0 bits before the next 1 bit
        uint32_t z = 0;
        for (;;) {
            if (cnt == 0) refill();
            const uint64_t window = acc & ((cnt == 64) ? ~0ull : ((1ull << cnt) - 1));
            if (window == 0) {
                z += cnt;
                cnt = 0;
                if (pos > n + 8) { bad = true; return z; }
                continue;
            }
            const int lead = __builtin_clzll(window) - (64 - cnt);
            z += lead;
            cnt -= lead + 1;
            return z;
        }
    }
    inline void align() { cnt -= cnt % 8; }
    inline int64_t byte_pos() const { return pos - cnt / 8; }   // valid when aligned
};

uint8_t crc8(const uint8_t* d, int64_t n) {
    uint8_t c = 0;
    for (int64_t i = 0; i < n; ++i) {
        c ^= d[i];
        for (int b = 0; b < 8; ++b) c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
    }
    return c;
}

uint16_t crc16_table[256];
bool crc16_ready = false;
void crc16_init() {
    for (int i = 0; i < 256; ++i) {
        uint16_t c = (uint16_t)(i << 8);
        for (int b = 0; b < 8; ++b) c = (uint16_t)((c & 0x8000) ? ((c << 1) ^ 0x8005) : (c << 1));
        crc16_table[i] = c;
    }
    crc16_ready = true;
}
uint16_t crc16(const uint8_t* d, int64_t n) {
    if (!crc16_ready) crc16_init();
    uint16_t c = 0;
    for (int64_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ crc16_table[(c >> 8) ^ d[i]]);
    return c;
}

struct Info {
    int32_t sample_rate, channels, bps, min_block, max_block;
    int64_t total_samples, first_frame;
    uint8_t md5[16];
};

int parse_header(const uint8_t* d, int64_t n, Info* info) {
    int64_t pos = 0;
    if (n >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3') {   // ID3v2 tag in front of the stream
        const int64_t sz = ((int64_t)(d[6] & 0x7f) << 21) | ((d[7] & 0x7f) << 14) | ((d[8] & 0x7f) << 7) | (d[9] & 0x7f);
        pos = 10 + sz;
    }
    if (pos + 4 > n || std::memcmp(d + pos, "fLaC", 4) != 0) { tw::set_error("flac: no fLaC marker"); return 2; }
    pos += 4;
    bool have_info = false;
    for (;;) {
        if (pos + 4 > n) { tw::set_error("flac: truncated metadata"); return 2; }
        const bool last = d[pos] & 0x80;
        const int type = d[pos] & 0x7f;
        const int64_t len = ((int64_t)d[pos + 1] << 16) | (d[pos + 2] << 8) | d[pos + 3];
        pos += 4;
        if (pos + len > n) { tw::set_error("flac: truncated metadata block"); return 2; }
        if (type == 0) {
            if (len < 34) { tw::set_error("flac: short STREAMINFO"); return 2; }
            const uint8_t* s = d + pos;
            info->min_block = (s[0] << 8) | s[1];
            info->max_block = (s[2] << 8) | s[3];
            info->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            info->channels = ((s[12] >> 1) & 7) + 1;
            info->bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            info->total_samples = ((int64_t)(s[13] & 0x0f) << 32) | ((int64_t)s[14] << 24) | (s[15] << 16) | (s[16] << 8) | s[17];
            std::memcpy(info->md5, s + 18, 16);
            have_info = true;
        }
        pos += len;
        if (last) break;
    }
    if (!have_info) { tw::set_error("flac: no STREAMINFO block"); return 2; }
    if (info->sample_rate <= 0 || info->bps < 4 || info->bps > 32) { tw::set_error("flac: bad STREAMINFO"); return 2; }
    info->first_frame = pos;
    return 0;
}

// residual of one subframe into res[order .. blocksize)
int read_residual(BitReader& br, int32_t* res, int blocksize, int order) {
    const int method = br.bits(2);
    if (method > 1) { tw::set_error("flac: reserved residual coding method"); return 3; }
    const int pbits = method == 0 ? 4 : 5;
    const uint32_t escape = method == 0 ? 15 : 31;
    const int porder = br.bits(4);
    const int parts = 1 << porder;
    if ((blocksize >> porder) << porder != blocksize && porder > 0) { tw::set_error("flac: partition order does not divide the block"); return 3; }
    int idx = order;
    for (int pt = 0; pt < parts; ++pt) {
        int count = (blocksize >> porder) - (pt == 0 ? order : 0);
        if (count < 0) { tw::set_error("flac: partition shorter than the predictor order"); return 3; }
        const uint32_t param = br.bits(pbits);
        if (param == escape) {
            const int raw = br.bits(5);
            for (int i = 0; i < count; ++i) res[idx++] = br.sbits(raw);
        } else {
            for (int i = 0; i < count; ++i) {
                const uint32_t q = br.unary();
                const uint32_t u = (q << param) | br.bits(param);
                res[idx++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
            }
        }
        if (br.bad) { tw::set_error("flac: truncated residual"); return 3; }
    }
    return 0;
}

int read_subframe(BitReader& br, int64_t* out, int32_t* scratch, int blocksize, int bps) {
    if (br.bits(1)) { tw::set_error("flac: subframe padding bit set"); return 3; }
    const int type = br.bits(6);
    int wasted = 0;
    if (br.bits(1)) wasted = (int)br.unary() + 1;
    bps -= wasted;
    if (bps < 1) { tw::set_error("flac: wasted bits exceed the sample size"); return 3; }
    auto sample = [&](int k) -> int64_t {       // k up to 33 bits (side channel of 32-bit audio)
        if (k <= 32) return br.sbits(k);
        const int64_t hi = br.sbits(k - 32);
        return (hi << 32) | br.bits(32);
    };
    if (type == 0) {
        const int64_t v = sample(bps);
        for (int i = 0; i < blocksize; ++i) out[i] = v;
    } else if (type == 1) {
        for (int i = 0; i < blocksize; ++i) out[i] = sample(bps);
    } else if (type >= 8 && type <= 12) {
        const int order = type - 8;
        if (order > blocksize) { tw::set_error("flac: fixed order exceeds the block"); return 3; }
        for (int i = 0; i < order; ++i) out[i] = sample(bps);
        if (int rc = read_residual(br, scratch, blocksize, order)) return rc;
        for (int i = order; i < blocksize; ++i) {
            int64_t pred = 0;
            switch (order) {
                case 1: pred = out[i - 1]; break;
                case 2: pred = 2 * out[i - 1] - out[i - 2]; break;
                case 3: pred = 3 * out[i - 1] - 3 * out[i - 2] + out[i - 3]; break;
                case 4: pred = 4 * out[i - 1] - 6 * out[i - 2] + 4 * out[i - 3] - out[i - 4]; break;
                default: break;
            }
            out[i] = pred + scratch[i];
        }
    } else if (type >= 32) {
        const int order = type - 31;
        if (order > blocksize) { tw::set_error("flac: LPC order exceeds the block"); return 3; }
        for (int i = 0; i < order; ++i) out[i] = sample(bps);
        const int prec = br.bits(4) + 1;
        if (prec == 16) { tw::set_error("flac: reserved LPC precision"); return 3; }
        const int shift = br.sbits(5);
        if (shift < 0) { tw::set_error("flac: negative LPC shift"); return 3; }
        int32_t coef[32];
        for (int i = 0; i < order; ++i) coef[i] = br.sbits(prec);
        if (int rc = read_residual(br, scratch, blocksize, order)) return rc;
        for (int i = order; i < blocksize; ++i) {
            int64_t acc = 0;
            for (int j = 0; j < order; ++j) acc += (int64_t)coef[j] * out[i - 1 - j];
            out[i] = (acc >> shift) + scratch[i];
        }
    } else {
        tw::set_error("flac: reserved subframe type %d", type);
        return 3;
    }
    if (wasted)
        for (int i = 0; i < blocksize; ++i) out[i] = out[i] * ((int64_t)1 << wasted);
    if (br.bad) { tw::set_error("flac: truncated subframe"); return 3; }
    return 0;
}

}  // namespace

extern "C" int tw_flac_info_read(const uint8_t* data, int64_t n, tw_flac_info* out) {
    if (!data || !out) { tw::set_error("tw_flac_info_read: null argument"); return 2; }
    Info info;
    if (int rc = parse_header(data, n, &info)) return rc;
    out->sample_rate = info.sample_rate;
    out->channels = info.channels;
    out->bits_per_sample = info.bps;
    out->max_block = info.max_block;
    out->total_samples = info.total_samples;
    std::memcpy(out->md5, info.md5, 16);
    return 0;
}

// Decodes every frame into out[sample][channel] (int32, interleaved; NULL = count only).  n_decoded receives the number
// of samples per channel found in the stream.
extern "C" int tw_flac_decode(const uint8_t* data, int64_t n, int32_t* out, int64_t out_cap_samples, int64_t* n_decoded) {
    if (!data || !n_decoded) { tw::set_error("tw_flac_decode: null argument"); return 2; }
    Info info;
    if (int rc = parse_header(data, n, &info)) return rc;
    const int C = info.channels;
    std::vector<int64_t> chan((size_t)C * 65536);
    std::vector<int32_t> scratch(65536);
    int64_t pos = info.first_frame, done = 0;
    while (pos + 2 <= n) {
        if (!(data[pos] == 0xff && (data[pos + 1] & 0xfe) == 0xf8)) {
            // trailing bytes that are not a frame (e.g. an ID3v1 tag) end the stream
            if (done > 0) break;
            tw::set_error("flac: lost frame sync at byte %lld", (long long)pos);
            return 3;
        }
        BitReader br(data, n, pos);
        br.bits(16);
        const int bs_code = br.bits(4), sr_code = br.bits(4), ch_code = br.bits(4), ss_code = br.bits(3);
        if (br.bits(1)) { tw::set_error("flac: reserved header bit set"); return 3; }
        // UTF-8 style coded frame / sample number (1-7 bytes)
        const int lead = br.bits(8);
        int extra = 0;
        if (lead & 0x80) {
            for (int m = 0x40; lead & m; m >>= 1) ++extra;
            if (extra == 0 || extra > 6) { tw::set_error("flac: bad coded frame number"); return 3; }
        }
        for (int i = 0; i < extra; ++i)
            if ((br.bits(8) & 0xc0) != 0x80) { tw::set_error("flac: bad coded frame number"); return 3; }
        int blocksize;
        if (bs_code == 0) { tw::set_error("flac: reserved block size code"); return 3; }
        else if (bs_code == 1) blocksize = 192;
        else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
        else if (bs_code == 6) blocksize = br.bits(8) + 1;
        else if (bs_code == 7) blocksize = br.bits(16) + 1;
        else blocksize = 256 << (bs_code - 8);
        if (sr_code == 12) br.bits(8);
        else if (sr_code == 13 || sr_code == 14) br.bits(16);
        else if (sr_code == 15) { tw::set_error("flac: invalid sample rate code"); return 3; }
        const int64_t hdr_end = br.byte_pos();
        if (hdr_end + 1 > n) { tw::set_error("flac: truncated frame header"); return 3; }
        const uint8_t want8 = (uint8_t)br.bits(8);
        if (crc8(data + pos, hdr_end - pos) != want8) { tw::set_error("flac: frame header CRC-8 mismatch at byte %lld", (long long)pos); return 3; }
        static const int ss_table[8] = {0, 8, 12, -1, 16, 20, 24, 32};
        int bps = ss_code == 0 ? info.bps : ss_table[ss_code];
        if (bps < 0) { tw::set_error("flac: reserved sample size code"); return 3; }
        int nch;
        if (ch_code < 8) nch = ch_code + 1;
        else if (ch_code <= 10) nch = 2;
        else { tw::set_error("flac: reserved channel assignment"); return 3; }
        if (nch != C) { tw::set_error("flac: frame has %d channels, STREAMINFO %d", nch, C); return 3; }
        for (int c = 0; c < C; ++c) {
            const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
            if (int rc = read_subframe(br, chan.data() + (size_t)c * 65536, scratch.data(), blocksize, bps + (side ? 1 : 0)))
                return rc;
        }
        br.align();
        const int64_t body_end = br.byte_pos();
        if (body_end + 2 > n) { tw::set_error("flac: truncated frame"); return 3; }
        const uint16_t want16 = (uint16_t)br.bits(16);
        if (crc16(data + pos, body_end - pos) != want16) { tw::set_error("flac: frame CRC-16 mismatch at byte %lld", (long long)pos); return 3; }
        int64_t* a = chan.data();
        int64_t* b = chan.data() + 65536;
        if (ch_code == 8) for (int i = 0; i < blocksize; ++i) b[i] = a[i] - b[i];
        else if (ch_code == 9) for (int i = 0; i < blocksize; ++i) a[i] = a[i] + b[i];
        else if (ch_code == 10)
            for (int i = 0; i < blocksize; ++i) {
                const int64_t side = b[i];
                const int64_t mid = (a[i] * 2) | (side & 1);
                a[i] = (mid + side) >> 1;
                b[i] = (mid - side) >> 1;
            }
        if (out) {
            if (done + blocksize > out_cap_samples) { tw::set_error("flac: output buffer too small (%lld samples)", (long long)out_cap_samples); return 2; }
            for (int c = 0; c < C; ++c) {
                const int64_t* src = chan.data() + (size_t)c * 65536;
                int32_t* dst = out + done * C + c;
                for (int i = 0; i < blocksize; ++i) dst[(int64_t)i * C] = (int32_t)src[i];
            }
        }
        done += blocksize;
        pos = body_end + 2;
    }
    *n_decoded = done;
    return 0;
}
