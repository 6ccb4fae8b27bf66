// K4 — LayerNorm over the last dimension: fp32 residual stream in, bf16 GEMM operand out.
// One warp per row, row cached in registers (cols <= 2048), two-pass mean / variance in fp32
// (same formulation as nn.LayerNorm: biased variance, eps inside the sqrt).
// $TF/models/whisper/modeling_whisper.py:380-414 (self_attn_layer_norm / final_layer_norm), :640.
#include "common.cuh"
#include "twb200_internal.h"

namespace tw {
namespace ln {

constexpr int MAX_VEC = 16;  // float4 per lane -> cols <= 16*4*32 = 2048

// Grid-stride over rows with the NEXT row's loads issued before the current row is reduced and stored: a warp
// always has a full row (5 KB at 1280 columns) in flight, instead of idling through two reductions and the stores
// of every row.  NV = float4 per lane actually used (ceil(cols / 128)), so the register buffers fit the row.
template <int NV>
__global__ void __launch_bounds__(256, 2) layernorm_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          __nv_bfloat16* __restrict__ out,
                                                          long long rows, int cols, float eps) {
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 2;  // cols % 4 == 0
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    float4 v[NV], nx[NV];
    auto load = [&](float4 (&dst)[NV], long long r) {
        const float4* xr = reinterpret_cast<const float4*>(x + r * cols);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) dst[i] = xr[c];
        }
    };
    load(v, row);
    for (; row < rows; row += warps_total) {
        const long long next = row + warps_total;
        if (next < rows) load(nx, next);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (lane + i * 32 < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(s) / (float)cols;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (lane + i * 32 < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (cc * cc + d * d);
            }
        const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
        uint2* orow = reinterpret_cast<uint2*>(out + row * cols);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) {
                const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
                uint2 pk;
                pk.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
                pk.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
                orow[c] = pk;
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = nx[i];
    }
}

}  // namespace ln
}  // namespace tw

extern "C" int tw_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16,
                            int64_t rows, int32_t cols, float eps, void* stream) {
    using namespace tw;
    TW_REQUIRE(x && gamma && beta && out_bf16, "tw_layernorm: null argument");
    if (tw::ensure_device(x)) return 1;
    TW_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= tw::ln::MAX_VEC * 128,
               "tw_layernorm: cols (%d) must be a multiple of 4 and <= %d", cols, tw::ln::MAX_VEC * 128);
    TW_REQUIRE(rows >= 0, "tw_layernorm: negative rows");
    if (rows == 0) return 0;
    const int warps = 8;
    const long long want = (rows + warps - 1) / warps;
    const int sms = tw::num_sms() > 0 ? tw::num_sms() : 148;
    const unsigned blocks = (unsigned)(want < (long long)sms * 2 ? want : (long long)sms * 2);   // 2 resident CTAs per SM, grid-stride
    const int nv = (cols + 127) / 128;
    auto* o = (__nv_bfloat16*)out_bf16;
    cudaStream_t st = (cudaStream_t)stream;
    if (nv <= 2) tw::ln::layernorm_kernel<2><<<blocks, warps * 32, 0, st>>>(x, gamma, beta, o, rows, cols, eps);
    else if (nv <= 4) tw::ln::layernorm_kernel<4><<<blocks, warps * 32, 0, st>>>(x, gamma, beta, o, rows, cols, eps);
    else if (nv <= 10) tw::ln::layernorm_kernel<10><<<blocks, warps * 32, 0, st>>>(x, gamma, beta, o, rows, cols, eps);
    else tw::ln::layernorm_kernel<16><<<blocks, warps * 32, 0, st>>>(x, gamma, beta, o, rows, cols, eps);
    TW_CUDA_CHECK(cudaGetLastError());
    return 0;
}
