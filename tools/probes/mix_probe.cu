// Probe: cycles per warp for the softmax instruction mix of one 128-key row (64 FFMA2, 128 MUFU.EX2, 64 FADD2, 64 F2FP)
// and for F2FP alone.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a mix_probe.cu -o mix_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
    float v[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) v[i] = -0.001f * (threadIdx.x + i);
    uint32_t acc = 0;
    uint64_t sum2[2] = {0, 0};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float mb = -0.5f * it;
        const uint64_t sc2 = pack2(1.4426950408889634f, 1.4426950408889634f), mb2 = pack2(mb, mb);
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            if (MODE == 0) {            // full mix
                float a0, a1;
                unpack2(fma2(pack2(v[2 * i], v[2 * i + 1]), sc2, mb2), a0, a1);
                const float p0 = ex2(a0), p1 = ex2(a1);
                sum2[i & 1] = add2(sum2[i & 1], pack2(p0, p1));
                acc ^= pack_bf16x2(p0, p1);
            } else if (MODE == 1) {     // F2FP only
                acc ^= pack_bf16x2(v[2 * i] + mb, v[2 * i + 1]);
            } else if (MODE == 2) {     // mix without F2FP
                float a0, a1;
                unpack2(fma2(pack2(v[2 * i], v[2 * i + 1]), sc2, mb2), a0, a1);
                const float p0 = ex2(a0), p1 = ex2(a1);
                sum2[i & 1] = add2(sum2[i & 1], pack2(p0, p1));
            } else {                    // scalar FFMA + MUFU + FADD, no packing
                const float p0 = ex2(fmaf(v[2 * i], 1.4426950408889634f, mb)), p1 = ex2(fmaf(v[2 * i + 1], 1.4426950408889634f, mb));
                float s0, s1; unpack2(sum2[i & 1], s0, s1);
                sum2[i & 1] = pack2(s0 + p0, s1 + p1);
            }
        }
    }
    long long t1 = clock64();
    float s0, s1, s2, s3; unpack2(sum2[0], s0, s1); unpack2(sum2[1], s2, s3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + __uint_as_float(acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 500;
    const char* names[4] = {"full mix (FFMA2+2 MUFU+FADD2+F2FP)", "F2FP only (+1 FADD)", "mix without F2FP", "scalar FFMA+MUFU+FADD"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps = 4; warps <= 8; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<148, warps * 32>>>(out, cyc, iters);
                else if (mode == 1) probe<1><<<148, warps * 32>>>(out, cyc, iters);
                else if (mode == 2) probe<2><<<148, warps * 32>>>(out, cyc, iters);
                else probe<3><<<148, warps * 32>>>(out, cyc, iters);
            }
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-40s warps/SMSP %d: %.0f cycles per 128-key row-step per warp\n", names[mode], warps / 4, (double)h / iters);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
