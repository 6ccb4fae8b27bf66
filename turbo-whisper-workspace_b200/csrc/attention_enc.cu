// K6 — encoder self-attention, flash-style forward on tcgen05/TMEM (non-causal, no mask, head 64).
//
// Replaces WhisperAttention.forward's attention_interface call for the encoder
// ($TF/models/whisper/modeling_whisper.py:335-352, scaling = 1.0 because q was scaled by
// head_dim**-0.5 right after q_proj, :310 — here that factor is folded into Wq/bq at load time,
// which is exact for a power of two).
//
// Input is the fused QKV projection output [B*T, 3*D] bf16 (q | k | v, head h at columns 64h..).
// One CTA = one (batch, head, 128-query tile).  192 threads: warp 0 TMA producer, warp 1 MMA
// issuer, warps 2-5 softmax / accumulation (one query row per thread, no shuffles).
//   S  = Q K_j^T      tcgen05.mma 128x128x64  -> TMEM cols [0,128)
//   P  = exp2(S - m)  softmax threads, written to smem as the bf16 K-major SW128 A-operand
//   O_j = P V_j       tcgen05.mma 128x64x128 (V consumed MN-major straight from its TMA tile)
//                     -> TMEM cols [128,192); threads fold it into fp32 registers with the usual
//                     online-softmax rescale.
// 96 KB smem and 256 TMEM columns per CTA so two CTAs share an SM and overlap each other's
// softmax and MMA phases.
#include "common.cuh"
#include "twb200_internal.h"

namespace tw {
namespace attn {

constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per iteration
constexpr int DH = 64;
constexpr int NUM_THREADS = 192;
constexpr int TILE_BYTES = 128 * DH * 2;  // 16 KB: Q, K, V tiles and each half of P
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0;
constexpr int O_COL = 128;
constexpr int SMEM_BYTES = 6 * TILE_BYTES + 1024 + 128;  // Q, K0, K1, V, P_lo, P_hi
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    int T, H, D;        // sequence length, heads, model width (H*64)
    long long out_ld;   // elements
    __nv_bfloat16* out;
};

__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_enc_kernel(const __grid_constant__ CUtensorMap tmQKV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;       // 2 stages
    uint8_t* sV = smem + 3 * TILE_BYTES;
    uint8_t* sP = smem + 4 * TILE_BYTES;   // 2 halves (keys 0-63, 64-127)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* v_full = bars + 5;
    uint64_t* v_empty = bars + 6;
    uint64_t* s_full = bars + 7;
    uint64_t* p_full = bars + 8;
    uint64_t* o_full = bars + 9;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (p.T + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
        mbar_init(&k_empty[0], 1); mbar_init(&k_empty[1], 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(&tmQKV, q_full, sQ, h * DH, q0, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                const uint32_t kph = (j >> 1) & 1;
                mbar_wait(&k_empty[s], kph ^ 1);
                mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
                tma_load_3d(&tmQKV, &k_full[s], sK + s * TILE_BYTES, p.D + h * DH, j * BKV, b);
                mbar_wait(v_empty, (j & 1) ^ 1);
                mbar_arrive_expect_tx(v_full, TILE_BYTES);
                tma_load_3d(&tmQKV, v_full, sV, 2 * p.D + h * DH, j * BKV, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, DH, 0, 1);  // B (= V) is MN-major
            const uint32_t q_addr = smem_u32(sQ), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
            mbar_wait(q_full, 0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                const uint32_t ph = j & 1;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(sK + s * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    tcgen05_mma_f16(tmem_base + S_COL, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                                    umma_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0);
                tcgen05_commit(s_full);
                tcgen05_commit(&k_empty[s]);
                mbar_wait(p_full, ph);
                mbar_wait(v_full, ph);
                tcgen05_fence_after();
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    tcgen05_mma_f16(tmem_base + O_COL,
                                    umma_desc_sw128(p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024),
                                    umma_desc_sw128(v_addr + k * 2048, 16, 1024), idesc_o, k != 0);
                tcgen05_commit(o_full);
                tcgen05_commit(v_empty);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        float m = -INFINITY, l = 0.f;
        float acc[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) acc[i] = 0.f;
        uint8_t* p_row = sP + r * 128;
        const int sw = r & 7;

        for (int j = 0; j < nkv; ++j) {
            const uint32_t ph = j & 1;
            const int kvalid = p.T - j * BKV;  // keys >= kvalid are padding (TMA zero-fill)
            mbar_wait(s_full, ph);
            tcgen05_fence_after();
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + t_lane + S_COL + c * 32, v);
                tmem_ld_wait();
                if (c * 32 + 32 <= kvalid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            const float m_new = fmaxf(m, mx);
            const float alpha = exp2f((m - m_new) * LOG2E);  // m = -inf on the first tile -> 0
            const float mb = m_new * LOG2E;
            m = m_new;
            // pass 2: P = exp2(s*log2e - m*log2e), row sum, bf16 pack into the swizzled A tile
            float rs = 0.f;
#pragma unroll 1
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + t_lane + S_COL + c * 32, v);
                tmem_ld_wait();
                float e[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float x = exp2f(fmaf(__uint_as_float(v[i]), LOG2E, -mb));
                    e[i] = (c * 32 + i < kvalid) ? x : 0.f;
                    rs += e[i];
                }
                uint8_t* dst = p_row + (c >> 1) * TILE_BYTES;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = (c & 1) * 4 + g;  // 16-byte chunk within the 128-byte row
                    uint4 pk;
                    pk.x = pack_bf16x2(e[g * 8 + 0], e[g * 8 + 1]);
                    pk.y = pack_bf16x2(e[g * 8 + 2], e[g * 8 + 3]);
                    pk.z = pack_bf16x2(e[g * 8 + 4], e[g * 8 + 5]);
                    pk.w = pack_bf16x2(e[g * 8 + 6], e[g * 8 + 7]);
                    *reinterpret_cast<uint4*>(dst + ((chunk ^ sw) << 4)) = pk;
                }
            }
            l = l * alpha + rs;
            fence_proxy_async_smem();  // make the generic-proxy smem writes visible to tcgen05.mma
            tcgen05_fence_before();
            mbar_arrive(p_full);
            // fold O_j into the running output
            mbar_wait(o_full, ph);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + t_lane + O_COL + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], alpha, __uint_as_float(v[i]));
            }
            tcgen05_fence_before();
        }
        const int q = q0 + r;
        if (q < p.T) {
            const float inv = 1.0f / l;
            __nv_bfloat16* o = p.out + ((size_t)b * p.T + q) * p.out_ld + h * DH;
#pragma unroll
            for (int g = 0; g < DH / 8; ++g) {
                uint4 pk;
                pk.x = pack_bf16x2(acc[g * 8 + 0] * inv, acc[g * 8 + 1] * inv);
                pk.y = pack_bf16x2(acc[g * 8 + 2] * inv, acc[g * 8 + 3] * inv);
                pk.z = pack_bf16x2(acc[g * 8 + 4] * inv, acc[g * 8 + 5] * inv);
                pk.w = pack_bf16x2(acc[g * 8 + 6] * inv, acc[g * 8 + 7] * inv);
                reinterpret_cast<uint4*>(o)[g] = pk;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn
}  // namespace tw

extern "C" int tw_attention_enc(const void* qkv_bf16, void* out_bf16, int32_t batch, int32_t seq,
                                int32_t heads, int64_t out_ld, void* stream) {
    using namespace tw;
    using namespace tw::attn;
    TW_REQUIRE(qkv_bf16 && out_bf16, "tw_attention_enc: null argument");
    TW_REQUIRE(batch >= 0 && seq > 0 && heads > 0, "tw_attention_enc: bad shape");
    TW_REQUIRE(batch <= 65535 && heads <= 65535, "tw_attention_enc: grid dimension too large");
    TW_REQUIRE(out_ld % 8 == 0 && ((uintptr_t)qkv_bf16 & 15) == 0 && ((uintptr_t)out_bf16 & 15) == 0,
               "tw_attention_enc: alignment");
    if (batch == 0) return 0;
    const int D = heads * DH;
    CUtensorMap tm;
    const uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)seq, (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)seq * 3 * D * 2};
    const uint32_t box[3] = {DH, 128, 1};
    if (encode_tensor_map(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv_bf16, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B))
        return 1;
    Params p;
    p.T = seq; p.H = heads; p.D = D; p.out_ld = out_ld; p.out = (__nv_bfloat16*)out_bf16;
    static bool attr_set = false;
    if (!attr_set) {
        TW_CUDA_CHECK(cudaFuncSetAttribute(attention_enc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid((seq + BQ - 1) / BQ, heads, batch);
    attention_enc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tm, p);
    TW_CUDA_CHECK(cudaGetLastError());
    return 0;
}
