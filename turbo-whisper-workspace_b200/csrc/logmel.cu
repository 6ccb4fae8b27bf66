// K1 — fused log-mel front end.
//
// Replaces WhisperFeatureExtractor._torch_extract_fbank_features
// ($TF/models/whisper/feature_extraction_whisper.py:135-164): reflect-padded 400-point STFT at
// hop 160 with a periodic Hann window, power spectrum, 128-bin slaney mel filterbank, log10 clamp
// at 1e-10, per-clip `max - 8` floor and (x + 4) / 4.
//
// Two launches:
//   logmel_power_kernel : PCM -> log10(mel) into an fp32 scratch [B,128,3000] + per-clip max
//                         (ordered-int atomicMax).  One CTA owns FR consecutive frames of one
//                         clip: samples are staged once in shared memory (each sample is reused by
//                         2.5 frames), the 400-point real FFT runs as a 200-point complex Stockham
//                         FFT (radix 8,5,5) entirely in shared memory, then the sparse (<=9 taps)
//                         mel projection.
//   logmel_norm_kernel  : clamp to max-8, (x+4)/4; writes fp32 [B,128,3000] (reference layout)
//                         and/or the bf16 time-major padded layout [B, rows, 128] the conv stem's
//                         TMA im2col view consumes (row 1+t = frame t).
//
// The phases are written as __host__ __device__ functions over (tid, nthreads) so that the exact
// index arithmetic is also exercised on the CPU by tests/csrc/logmel_host_test.cu.
#include "common.cuh"
#include "twb200_internal.h"
#include <math.h>

namespace tw {
namespace logmel {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_BINS = 201;
constexpr int N_MEL = 128;
constexpr int N_FRAMES = 3000;
constexpr int N_SAMPLES = 480000;
constexpr int NC = 200;  // complex FFT length
constexpr int FR = 16;   // frames per CTA
constexpr int NT = 256;  // threads per CTA
constexpr int SX = FR * HOP + (N_FFT - HOP);  // 2800 staged samples
constexpr int POW_LD = 203;                   // padded row length of the power buffer
constexpr int MEL_MAX_TAPS = 12;

struct Tables {
    float2 tw200[NC];       // exp(-2 pi i m / 200)
    float2 tw400[N_BINS];   // exp(-2 pi i k / 400)
    float window[N_FFT];    // periodic Hann
    int mel_start[N_MEL];
    int mel_cnt[N_MEL];
    float mel_w[N_MEL][MEL_MAX_TAPS];
};

struct Smem {
    float x[SX];
    float2 a[FR][NC];
    float2 b[FR][NC];
    float2 tw200[NC];
    float2 tw400[N_BINS];
    float red[NT / 32];
};

#define HD __host__ __device__ __forceinline__

HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
HD float2 cmul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)

HD void dft4(float2 a, float2 b, float2 c, float2 d, float2& o0, float2& o1, float2& o2, float2& o3) {
    float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
    o0 = cadd(s0, s2);
    o1 = cadd(s1, cmul_mi(s3));
    o2 = csub(s0, s2);
    o3 = cadd(s1, cmul_pi(s3));
}
HD void dft8(float2* v) {
    const float h = 0.70710678118654752f;
    float2 e0, e1, e2, e3, o0, o1, o2, o3;
    dft4(v[0], v[2], v[4], v[6], e0, e1, e2, e3);
    dft4(v[1], v[3], v[5], v[7], o0, o1, o2, o3);
    // w8^1 = (1 - i)/sqrt2, w8^2 = -i, w8^3 = (-1 - i)/sqrt2
    float2 t1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));
    float2 t2 = cmul_mi(o2);
    float2 t3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}
HD void dft5(float2* v) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 r0 = cadd(v[0], cadd(a1, a2));
    float2 p1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    float2 p2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    v[0] = r0;
    v[1] = cadd(p1, cmul_mi(q1));
    v[4] = cadd(p1, cmul_pi(q1));
    v[2] = cadd(p2, cmul_mi(q2));
    v[3] = cadd(p2, cmul_pi(q2));
}

// phase 0: stage samples [160*f0 - 200, 160*f0 - 200 + SX) with reflect padding on the 480000-
// sample (zero-padded) clip; samples at or beyond n_valid read as 0 (HF pads with zeros first,
// then torch.stft(center=True) reflects the padded clip).
HD void phase_load(int tid, int nt, const float* pcm, int n_valid, int f0, float* sx) {
    const int base = HOP * f0 - N_FFT / 2;
    for (int i = tid; i < SX; i += nt) {
        int g = base + i;
        if (g < 0) g = -g;
        if (g >= N_SAMPLES) g = 2 * (N_SAMPLES - 1) - g;
        sx[i] = (g < n_valid) ? pcm[g] : 0.0f;
    }
}
// phase 1: window + pack to complex + radix-8 stage (Ns = 1, no twiddles).  25 items / frame.
HD void phase_fft_r8(int tid, int nt, const float* sx, const float* window, float2 (*out)[NC]) {
    constexpr int T = NC / 8;  // 25
    for (int w = tid; w < FR * T; w += nt) {
        const int fr = w / T, j = w - fr * T;
        const float* xf = sx + fr * HOP;
        float2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int n = j + t * T;
            v[t] = make_float2(xf[2 * n] * window[2 * n], xf[2 * n + 1] * window[2 * n + 1]);
        }
        dft8(v);
#pragma unroll
        for (int u = 0; u < 8; ++u) out[fr][j * 8 + u] = v[u];
    }
}
// phases 2/3: radix-5 Stockham stage with sub-transform length Ns (8, then 40).  40 items / frame.
HD void phase_fft_r5(int tid, int nt, int Ns, const float2* tw200, const float2 (*in)[NC],
                     float2 (*out)[NC]) {
    constexpr int T = NC / 5;  // 40
    const int twstep = NC / (Ns * 5);
    for (int w = tid; w < FR * T; w += nt) {
        const int fr = w / T, j = w - fr * T;
        const int k = j % Ns;
        float2 v[5];
        v[0] = in[fr][j];
#pragma unroll
        for (int t = 1; t < 5; ++t) v[t] = cmul(in[fr][j + t * T], tw200[t * k * twstep]);
        dft5(v);
        const int j0 = (j / Ns) * Ns * 5 + k;
#pragma unroll
        for (int u = 0; u < 5; ++u) out[fr][j0 + u * Ns] = v[u];
    }
}
// phase 4: real-FFT post-process + power.  X[k] = E[k] + W400^k O[k], k = 0..200.
HD void phase_power(int tid, int nt, const float2* tw400, const float2 (*Z)[NC], float* pw) {
    for (int w = tid; w < FR * N_BINS; w += nt) {
        const int fr = w / N_BINS, k = w - fr * N_BINS;
        const float2 zk = Z[fr][k == NC ? 0 : k];
        const float2 zr = Z[fr][(k == 0 || k == NC) ? 0 : NC - k];
        const float2 zc = make_float2(zr.x, -zr.y);
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
        const float2 o = cmul_mi(d);
        const float2 x = cadd(e, cmul(tw400[k], o));
        pw[fr * POW_LD + k] = x.x * x.x + x.y * x.y;
    }
}
// phase 5: sparse mel projection + log10 clamp.  Item = (mel, frame) with frame fastest so that
// the global store of one warp covers 2 mel rows x 16 consecutive frames (64 B segments).
HD float phase_mel(int tid, int nt, const int* mel_start, const int* mel_cnt,
                   const float (*mel_w)[MEL_MAX_TAPS], const float* pw, float* out_clip, int f0) {
    float vmax = -1e30f;
    for (int w = tid; w < FR * N_MEL; w += nt) {
        const int m = w / FR, fr = w - m * FR;
        const int f = f0 + fr;
        if (f >= N_FRAMES) continue;
        const int s = mel_start[m], c = mel_cnt[m];
        float acc = 0.0f;
        for (int i = 0; i < c; ++i) acc = fmaf(mel_w[m][i], pw[fr * POW_LD + s + i], acc);
        const float v = log10f(fmaxf(acc, 1e-10f));
        out_clip[(size_t)m * N_FRAMES + f] = v;
        vmax = fmaxf(vmax, v);
    }
    return vmax;
}

#ifndef TW_HOST_TEST
__global__ void __launch_bounds__(NT) logmel_power_kernel(const float* __restrict__ pcm,
                                                         long long pcm_stride,
                                                         const int* __restrict__ n_valid,
                                                         const Tables* __restrict__ tab,
                                                         float* __restrict__ scratch,
                                                         int* __restrict__ clip_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * FR;
    const int tid = threadIdx.x;
    const int nv = n_valid ? min(n_valid[b], N_SAMPLES) : N_SAMPLES;

    for (int i = tid; i < NC; i += NT) s.tw200[i] = tab->tw200[i];
    for (int i = tid; i < N_BINS; i += NT) s.tw400[i] = tab->tw400[i];
    phase_load(tid, NT, pcm + (size_t)b * pcm_stride, nv, f0, s.x);
    __syncthreads();
    phase_fft_r8(tid, NT, s.x, tab->window, s.a);
    __syncthreads();
    phase_fft_r5(tid, NT, 8, s.tw200, s.a, s.b);
    __syncthreads();
    phase_fft_r5(tid, NT, 40, s.tw200, s.b, s.a);
    __syncthreads();
    float* pw = reinterpret_cast<float*>(&s.b[0][0]);  // FR*POW_LD floats <= FR*NC*2
    phase_power(tid, NT, s.tw400, s.a, pw);
    __syncthreads();
    float vmax = phase_mel(tid, NT, tab->mel_start, tab->mel_cnt, tab->mel_w, pw,
                           scratch + (size_t)b * N_MEL * N_FRAMES, f0);
    vmax = warp_max(vmax);
    if ((tid & 31) == 0) s.red[tid >> 5] = vmax;
    __syncthreads();
    if (tid == 0) {
        float m = s.red[0];
#pragma unroll
        for (int i = 1; i < NT / 32; ++i) m = fmaxf(m, s.red[i]);
        atomicMax(clip_max + b, float_to_ordered(m));
    }
}

// normalise: y = (max(x, clipmax - 8) + 4) / 4.  Tile = 32 frames x 128 mels; the fp32 output
// keeps the reference [128,3000] layout, the bf16 output is transposed through shared memory
// into [rows,128] (row = row_off + frame), 256 B per frame, fully coalesced on both sides.
constexpr int NTILE_F = 32;
__global__ void __launch_bounds__(256) logmel_norm_kernel(const float* __restrict__ scratch,
                                                         const int* __restrict__ clip_max,
                                                         float* __restrict__ out_f32,
                                                         __nv_bfloat16* __restrict__ out_t,
                                                         long long out_t_bstride, int row_off) {
    __shared__ float tile[N_MEL][NTILE_F + 1];
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * NTILE_F;
    const float floor_v = ordered_to_float(clip_max[b]) - 8.0f;
    const float* src = scratch + (size_t)b * N_MEL * N_FRAMES;
    const int tid = threadIdx.x;
    for (int w = tid; w < N_MEL * NTILE_F; w += 256) {
        const int m = w / NTILE_F, fr = w % NTILE_F;
        const int f = f0 + fr;
        float v = 0.0f;
        if (f < N_FRAMES) {
            v = (fmaxf(src[(size_t)m * N_FRAMES + f], floor_v) + 4.0f) * 0.25f;
            if (out_f32) out_f32[((size_t)b * N_MEL + m) * N_FRAMES + f] = v;
        }
        tile[m][fr] = v;
    }
    if (!out_t) return;
    __syncthreads();
    for (int w = tid; w < NTILE_F * (N_MEL / 2); w += 256) {
        const int fr = w / (N_MEL / 2), m2 = w % (N_MEL / 2);
        const int f = f0 + fr;
        if (f >= N_FRAMES) continue;
        uint32_t p = pack_bf16x2(tile[2 * m2][fr], tile[2 * m2 + 1][fr]);
        reinterpret_cast<uint32_t*>(out_t + (size_t)b * out_t_bstride +
                                    (size_t)(row_off + f) * N_MEL)[m2] = p;
    }
}
#endif  // TW_HOST_TEST

// Host: build the constant tables from the dense [201,128] fp32 filterbank (HF mel_filters).
int build_tables(const float* mel_filters_201x128, Tables* t) {
    const double PI = 3.14159265358979323846;
    for (int m = 0; m < NC; ++m)
        t->tw200[m] = make_float2((float)cos(2.0 * PI * m / NC), (float)-sin(2.0 * PI * m / NC));
    for (int k = 0; k < N_BINS; ++k)
        t->tw400[k] = make_float2((float)cos(2.0 * PI * k / N_FFT), (float)-sin(2.0 * PI * k / N_FFT));
    for (int n = 0; n < N_FFT; ++n) t->window[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / N_FFT));
    for (int m = 0; m < N_MEL; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < N_BINS; ++k)
            if (mel_filters_201x128[k * N_MEL + m] != 0.0f) {
                if (first < 0) first = k;
                last = k;
            }
        if (first < 0) { first = 0; last = -1; }
        const int cnt = last - first + 1;
        if (cnt > MEL_MAX_TAPS) {
            set_error("mel filter %d has %d taps (max %d)", m, cnt, MEL_MAX_TAPS);
            return 2;
        }
        t->mel_start[m] = first;
        t->mel_cnt[m] = cnt;
        for (int i = 0; i < MEL_MAX_TAPS; ++i)
            t->mel_w[m][i] = (i < cnt) ? mel_filters_201x128[(first + i) * N_MEL + m] : 0.0f;
    }
    return 0;
}

}  // namespace logmel
}  // namespace tw

#ifndef TW_HOST_TEST
using namespace tw;
using namespace tw::logmel;

extern "C" size_t tw_logmel_tables_bytes(void) { return sizeof(Tables); }
extern "C" size_t tw_logmel_scratch_bytes(int batch) {
    return (size_t)batch * N_MEL * N_FRAMES * sizeof(float) + (size_t)batch * sizeof(int);
}

extern "C" int tw_logmel_init(void* tables_dev, const float* mel_filters_host_201x128) {
    TW_REQUIRE(tables_dev && mel_filters_host_201x128, "tw_logmel_init: null argument");
    Tables* h = new Tables;
    int rc = build_tables(mel_filters_host_201x128, h);
    if (rc == 0) {
        cudaError_t e = cudaMemcpy(tables_dev, h, sizeof(Tables), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("tw_logmel_init: cudaMemcpy failed: %s", cudaGetErrorString(e));
            rc = 1;
        }
    }
    delete h;
    return rc;
}

extern "C" int tw_logmel(const void* tables_dev, const float* pcm, int64_t pcm_stride,
                         const int32_t* n_valid, int batch, void* scratch, float* out_f32,
                         void* out_bf16_t, int64_t out_t_bstride, int32_t out_t_row_off,
                         void* stream) {
    TW_REQUIRE(tables_dev && pcm && scratch, "tw_logmel: null argument");
    if (tw::ensure_device(pcm)) return 1;
    TW_REQUIRE(batch >= 0 && batch <= 65535, "tw_logmel: batch %d out of range", batch);
    TW_REQUIRE(pcm_stride >= N_SAMPLES, "tw_logmel: pcm_stride %lld < %d", (long long)pcm_stride,
               N_SAMPLES);
    if (batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* scr = (float*)scratch;
    int* clip_max = (int*)(scr + (size_t)batch * N_MEL * N_FRAMES);
    static std::atomic<unsigned long long> attr_done{0};
    if (device_needs_setup(attr_done)) {
        TW_CUDA_CHECK(cudaFuncSetAttribute(logmel_power_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(Smem)));
        mark_device_done(attr_done);
    }
    // 0x80000000 is the ordered-int image of the most negative float
    TW_CUDA_CHECK(cudaMemsetAsync(clip_max, 0x80, (size_t)batch * sizeof(int), st));
    dim3 g1((N_FRAMES + FR - 1) / FR, batch);
    logmel_power_kernel<<<g1, NT, sizeof(Smem), st>>>(pcm, (long long)pcm_stride, n_valid,
                                                      (const Tables*)tables_dev, scr, clip_max);
    dim3 g2((N_FRAMES + NTILE_F - 1) / NTILE_F, batch);
    logmel_norm_kernel<<<g2, 256, 0, st>>>(scr, clip_max, out_f32, (__nv_bfloat16*)out_bf16_t,
                                           (long long)out_t_bstride, out_t_row_off);
    TW_CUDA_CHECK(cudaGetLastError());
    return 0;
}
#endif
