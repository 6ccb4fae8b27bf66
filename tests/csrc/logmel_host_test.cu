// CPU emulation of the log-mel CUDA kernel's phases (same __host__ __device__ code, threads run
// serially).  Usage: logmel_host_test pcm.f32 n_valid melfilters.f32 out.f32
// Lets the index arithmetic of logmel.cu be checked against the oracle without a GPU.
#define TW_HOST_TEST 1
#include "../../turbo-whisper-workspace_b200/csrc/logmel.cu"
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <vector>
namespace tw { void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); } }
using namespace tw::logmel;

static std::vector<float> read_f32(const char* path) {
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v(n / 4); if (fread(v.data(), 4, v.size(), f) != v.size()) exit(1); fclose(f); return v;
}
int main(int argc, char** argv) {
    if (argc != 5) return 2;
    std::vector<float> pcm = read_f32(argv[1]);
    int n_valid = atoi(argv[2]);
    std::vector<float> fb = read_f32(argv[3]);
    pcm.resize(N_SAMPLES, 0.f);
    Tables* tab = new Tables;
    if (build_tables(fb.data(), tab)) return 1;
    std::vector<float> scratch((size_t)N_MEL * N_FRAMES), out((size_t)N_MEL * N_FRAMES);
    Smem* s = new Smem;
    float vmax = -1e30f;
    for (int f0 = 0; f0 < N_FRAMES; f0 += FR) {
        for (int tid = 0; tid < NT; ++tid) phase_load(tid, NT, pcm.data(), n_valid, f0, s->x);
        for (int tid = 0; tid < NT; ++tid) phase_fft_r8(tid, NT, s->x, tab->window, s->a);
        for (int tid = 0; tid < NT; ++tid) phase_fft_r5(tid, NT, 8, tab->tw200, s->a, s->b);
        for (int tid = 0; tid < NT; ++tid) phase_fft_r5(tid, NT, 40, tab->tw200, s->b, s->a);
        float* pw = reinterpret_cast<float*>(&s->b[0][0]);
        for (int tid = 0; tid < NT; ++tid) phase_power(tid, NT, tab->tw400, s->a, pw);
        for (int tid = 0; tid < NT; ++tid)
            vmax = fmaxf(vmax, phase_mel(tid, NT, tab->mel_start, tab->mel_cnt, tab->mel_w, pw, scratch.data(), f0));
    }
    for (size_t i = 0; i < out.size(); ++i) out[i] = (fmaxf(scratch[i], vmax - 8.0f) + 4.0f) * 0.25f;
    FILE* f = fopen(argv[4], "wb"); fwrite(out.data(), 4, out.size(), f); fclose(f);
    return 0;
}
