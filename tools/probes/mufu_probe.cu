// Probe: MUFU.EX2 issue rate per SM sub-partition (cycles per warp instruction) with 1, 2, 4 warps per SMSP,
// and the fp32x2 FMA polynomial alternative.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a mufu_probe.cu -o mufu_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (MODE == 0) v[i] = ex2(v[i]);
            else {  // degree-3 polynomial on the FMA pipe (Cody-Waite split done with the magic-number trick)
                float x = fmaxf(v[i], -126.0f);
                float xr = x + 12582912.0f;            // round to nearest integer
                float xi = xr - 12582912.0f;
                float f = x - xi;                       // [-0.5, 0.5]
                float p = fmaf(f, 0.0555041f, 0.2402265f);
                p = fmaf(p, f, 0.6931472f);
                p = fmaf(p, f, 1.0f);
                v[i] = __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 1000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 32; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<148, warps * 32>>>(out, cyc, iters); else probe<1><<<148, warps * 32>>>(out, cyc, iters);
            }
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double per_instr = (double)h / (iters * 32.0);              // cycles per warp-instruction slot of one warp
            double per_smsp = per_instr / (warps / 4.0);                 // cycles per warp-op on one SMSP
            printf("%s warps/SM %2d (per SMSP %d): %.2f cycles per op per warp, %.2f cycles per warp-op per SMSP, %.1f ops/clk/SM\n",
                   mode == 0 ? "MUFU.EX2" : "poly-exp2", warps, warps / 4, per_instr, per_smsp, 32.0 * 4 / per_smsp);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
