// K6 — encoder self-attention, flash-style forward on tcgen05/TMEM (non-causal, no mask, head 64).
//
// Replaces WhisperAttention.forward's attention_interface call for the encoder
// ($TF/models/whisper/modeling_whisper.py:335-352, scaling = 1.0 because q was scaled by
// head_dim**-0.5 right after q_proj, :310 — here that factor is folded into Wq/bq at load time,
// which is exact for a power of two).
//
// Input is the fused QKV projection output [B*T, 3*D] bf16 (q | k | v, head h at columns 64h..).
// One CTA = one (batch, head) and TWO 128-query tiles (A, B) that ping-pong on the tensor pipe.  320 threads:
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 / 6-9 softmax + accumulation of tile A / B (one query row
// per thread, no shuffles).  While the threads of A run the softmax of kv-tile j, the tensor core computes
// S_B(j), P_B V(j-1) ... : MMA order S_A(0) S_B(0) | PV_A(j) S_A(j+1) PV_B(j) S_B(j+1) | ...
//   S  = Q K_j^T      tcgen05.mma 128x128x64  -> TMEM cols [0,128)
//   P  = exp2(S - m)  softmax threads, written to smem as the bf16 K-major SW128 A-operand
//   O_j = P V_j       tcgen05.mma 128x64x128 (V consumed MN-major straight from its TMA tile)
//                     -> TMEM cols [128,192); threads fold it into fp32 registers with the usual
//                     online-softmax rescale.
// K/V tiles are shared by both query tiles (half the smem / L2 traffic per query); 160 KB smem and 384 of the 512
// TMEM columns per CTA (S_A, S_B, O_A, O_B).
#include "common.cuh"
#include "twb200_internal.h"

namespace tw {
namespace attn {

constexpr int BQ = 128;   // query rows per tile; a CTA owns two tiles (A, B) that ping-pong
constexpr int BKV = 128;  // keys per iteration
constexpr int DH = 64;
constexpr int NUM_THREADS = 320;           // warp 0 TMA, warp 1 MMA, warps 2-5 softmax A, warps 6-9 softmax B
constexpr int TILE_BYTES = 128 * DH * 2;   // 16 KB: Q, K, V tiles and each half of P
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;                   // S_A [0,128)   S_B [128,256)
constexpr int O_N = 80;                    // 64 output columns + 16 row-sum columns (V is extended by a block of ones)
constexpr int O_COL = 256;                 // O_A [256,336) O_B [336,416); row sums at O_COL + 64
constexpr int ONES_BYTES = 2048;           // 16 key rows x 128 B of bf16 1.0: second MN atom of the PV B operand
// smem tiles: Q_A Q_B | K0 K1 | V0 V1 | ones.  P never touches shared memory: the softmax threads store it (bf16,
// two keys per 32-bit column) into the first 64 TMEM columns of their own S tile and tcgen05.mma reads the A
// operand from tensor memory.
constexpr int SMEM_BYTES = 6 * TILE_BYTES + ONES_BYTES + 1024 + 256;
constexpr float LOG2E = 1.4426950408889634f;

// single-instruction exp2 (MUFU.EX2, flush-to-zero): arguments here are <= 0, so the slow path of exp2f()
// (denormal-input scaling, 4 extra instructions per element) is never needed
// two exponentials per MUFU op on bf16 pairs: the result is directly the packed bf16 P operand.  The argument
// is rounded to bf16 first (|error| <= 2^-9 |x| in the exponent, below the bf16 rounding of P itself for the
// entries that carry weight).
TW_DEVINL uint32_t ex2_bf16x2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
TW_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Params {
    int T, H, D;        // sequence length, heads, model width (H*64)
    long long out_ld;   // elements
    __nv_bfloat16* out;
    long long* dbg;   // optional clock64 trace of CTA (0,0,0): [0..63] MMA thread, [64..] softmax A row 0
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_enc_kernel(const __grid_constant__ CUtensorMap tmQKV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // 2 tiles
    uint8_t* sK = smem + 2 * TILE_BYTES;      // 2 stages
    uint8_t* sV = smem + 4 * TILE_BYTES;      // 2 stages
    uint8_t* sOnes = smem + 6 * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES + ONES_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* v_full = bars + 5;   // [2]
    uint64_t* v_empty = bars + 7;  // [2]
    uint64_t* s_full = bars + 9;   // [2] per group
    uint64_t* p_full = bars + 11;  // [2] per group
    uint64_t* o_full = bars + 13;  // [2] per group
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * BQ), h = blockIdx.y, b = blockIdx.z;
    const int nkv = (p.T + BKV - 1) / BKV;
    const bool trace = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    int ti = 0;
#define TRACE(base) do { if (trace && ti < 60) p.dbg[(base) + ti++] = clock64(); } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NUM_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;  // bf16 1.0 pairs
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
            tma_load_3d(&tmQKV, q_full, sQ, h * DH, q0, b);
            tma_load_3d(&tmQKV, q_full, sQ + TILE_BYTES, h * DH, q0 + BQ, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
                tma_load_3d(&tmQKV, &k_full[s], sK + s * TILE_BYTES, p.D + h * DH, j * BKV, b);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
                tma_load_3d(&tmQKV, &v_full[s], sV + s * TILE_BYTES, 2 * p.D + h * DH, j * BKV, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, O_N, 0, 1);  // B = [V | ones] is MN-major, N = 80
            // Every descriptor is loop-invariant: build them once.  The single issuing thread must spend ~2
            // instructions per tcgen05.mma, not ~30, because these MMAs only last 32-64 cycles each.
            uint64_t dq[2][DH / 16], dk[2][DH / 16], dv[2][BKV / 16];
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    dq[g][k] = umma_desc_sw128(smem_u32(sQ) + g * TILE_BYTES + k * 32, 16, 1024);
                    dk[g][k] = umma_desc_sw128(smem_u32(sK) + g * TILE_BYTES + k * 32, 16, 1024);
                }
            }
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) {
#pragma unroll
                for (int st = 0; st < 2; ++st) {
                    const uint32_t va = smem_u32(sV) + st * TILE_BYTES + k * 2048;
                    // MN atom 0 = the 64 head-dim columns of V (16 key rows x 128 B); atom 1 (LBO away) = ones
                    dv[st][k] = umma_desc_sw128(va, smem_u32(sOnes) - va, 1024);
                }
            }
            auto issue_s = [&](int g, int j) {   // S_g = Q_g K_j^T
                const int st = j & 1;
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    tcgen05_mma_f16(tmem_base + S_COL + g * BKV, dq[g][k], dk[st][k], idesc_s, k != 0);
                tcgen05_commit(&s_full[g]);
            };
            auto issue_pv = [&](int g, int j) {  // [O_g | rowsum_g] += P_g [V_j | 1]   (A = P from TMEM)
                const int st = j & 1;
                const uint32_t p_tmem = tmem_base + S_COL + g * BKV;   // P aliases the first 64 columns of S_g
                const uint32_t acc0 = j != 0;
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    tcgen05_mma_f16_ts(tmem_base + O_COL + g * O_N, p_tmem + k * 8, dv[st][k], idesc_o, k != 0 ? 1u : acc0);
                tcgen05_commit(&o_full[g]);
            };
            TRACE(0);
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            TRACE(0);
            tcgen05_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            tcgen05_commit(&k_empty[0]);
            // The two query tiles are served in whatever order their P tiles become ready (non-blocking polls), so a
            // slow softmax of one tile never delays the MMAs of the other.  K/V stages are released once both tiles
            // have issued the MMAs that read them; a tile can therefore run at most one kv tile ahead of the other.
            int jg[2] = {0, 0};          // next PV index per tile
            while (jg[0] < nkv || jg[1] < nkv) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int j = jg[g];
                    if (j >= nkv) continue;
                    if (!mbar_try_wait(&p_full[g], j & 1)) continue;
                    mbar_wait(&v_full[j & 1], (j >> 1) & 1);
                    tcgen05_fence_after();
                    issue_pv(g, j);
                    if (jg[g ^ 1] > j) tcgen05_commit(&v_empty[j & 1]);          // second reader of V_j
                    if (j + 1 < nkv) {
                        mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                        tcgen05_fence_after();
                        issue_s(g, j + 1);
                        if (jg[g ^ 1] > j) tcgen05_commit(&k_empty[(j + 1) & 1]);  // second reader of K_{j+1}
                    }
                    jg[g] = j + 1;
                }
            }
        }
    } else {
        const int grp = (warp - 2) >> 2;    // 0: tile A, 1: tile B
        const int quarter = warp & 3;       // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const uint32_t s_col = S_COL + grp * BKV, o_col = O_COL + grp * O_N, l_col = o_col + DH;
        // O and the row sums accumulate in TMEM across kv tiles (tcgen05.mma accumulate); S is read from TMEM
        // exactly once per tile (TMEM read bandwidth is the scarce resource).  The softmax reference point
        // m_used only moves when the running maximum grows by more than RESCALE_LOG2 (lazy rescaling: P <= 2^8,
        // mathematically identical after the final O / rowsum division); a move rescales O and rowsum in TMEM.
        constexpr float RESCALE_LOG2 = 8.0f;
        float m_used = -INFINITY;   // in log2 units (score * log2e)

        for (int j = 0; j < nkv; ++j) {
            const uint32_t ph = j & 1;
            const int kvalid = p.T - j * BKV;  // keys >= kvalid are padding (TMA zero-fill)
            const bool full = kvalid >= BKV;
            if (grp == 0 && r == 0) TRACE(64);
            mbar_wait(&s_full[grp], ph);
            if (j > 0) mbar_wait(&o_full[grp], (j - 1) & 1);
            if (grp == 0 && r == 0) TRACE(64);  // PV(j-1) done (implied by S(j) done; keeps phases in step)
            tcgen05_fence_after();
            uint32_t v[BKV];
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c)
                tmem_ld_32x32b_x32(tmem_base + t_lane + s_col + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
            tmem_ld_wait();
            if (grp == 0 && r == 0) TRACE(64);
            float mx = -INFINITY;
            if (full) {
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains
#pragma unroll
                for (int i = 0; i < BKV; i += 4) {
                    m4[0] = fmaxf(m4[0], __uint_as_float(v[i]));
                    m4[1] = fmaxf(m4[1], __uint_as_float(v[i + 1]));
                    m4[2] = fmaxf(m4[2], __uint_as_float(v[i + 2]));
                    m4[3] = fmaxf(m4[3], __uint_as_float(v[i + 3]));
                }
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
#pragma unroll
                for (int i = 0; i < BKV; ++i)
                    if (i < kvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float mx2 = mx * LOG2E;
            const bool move = mx2 > m_used + RESCALE_LOG2;   // always true on the first tile (m_used = -inf)
            const float alpha = move ? ex2_approx(m_used - mx2) : 1.0f;   // first tile: exp2(-inf) = 0, O is overwritten
            if (move) m_used = mx2;
            const float mb = m_used;
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t pk[16];
                if (full || c * 32 + 32 <= kvalid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        pk[i] = ex2_bf16x2(pack_bf16x2(fmaf(__uint_as_float(v[c * 32 + 2 * i]), LOG2E, -mb),
                                                       fmaf(__uint_as_float(v[c * 32 + 2 * i + 1]), LOG2E, -mb)));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float a0 = (c * 32 + 2 * i < kvalid) ? fmaf(__uint_as_float(v[c * 32 + 2 * i]), LOG2E, -mb) : -INFINITY;
                        const float a1 = (c * 32 + 2 * i + 1 < kvalid) ? fmaf(__uint_as_float(v[c * 32 + 2 * i + 1]), LOG2E, -mb) : -INFINITY;
                        pk[i] = ex2_bf16x2(pack_bf16x2(a0, a1));
                    }
                }
                // keys [32c, 32c+32) of this row -> TMEM columns [16c, 16c+16) of the P tile (aliases S_g, already in registers)
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                    :
                    : "r"(tmem_base + t_lane + s_col + c * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]),
                      "r"(pk[5]), "r"(pk[6]), "r"(pk[7]), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]),
                      "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                    : "memory");
            }
            // rescale the TMEM accumulators of this row when its reference point moved (warp-collective ld/st)
            if (j > 0 && __any_sync(0xffffffffu, move)) {
#pragma unroll
                for (int c = 0; c < DH / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32b_x32(tmem_base + t_lane + o_col + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32b_x32(tmem_base + t_lane + o_col + c * 32, o);
                }
                const uint32_t ls = tmem_ld_32x32b_x1(tmem_base + t_lane + l_col);
                tmem_ld_wait();
                tmem_st_32x32b_x1(tmem_base + t_lane + l_col, __float_as_uint(__uint_as_float(ls) * alpha));
            }
            tmem_st_wait();
            if (grp == 0 && r == 0) TRACE(64);
            tcgen05_fence_before();
            mbar_arrive(&p_full[grp]);
        }
        mbar_wait(&o_full[grp], (nkv - 1) & 1);
        tcgen05_fence_after();
        const int q = q0 + grp * BQ + r;
        const uint32_t ls = tmem_ld_32x32b_x1(tmem_base + t_lane + l_col);
        tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(ls);
        __nv_bfloat16* o = p.out + ((size_t)b * p.T + q) * p.out_ld + h * DH;
#pragma unroll
        for (int c = 0; c < DH / 32; ++c) {
            uint32_t acc[32];
            tmem_ld_32x32b_x32(tmem_base + t_lane + o_col + c * 32, acc);
            tmem_ld_wait();
            if (q < p.T) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 pk;
                    pk.x = pack_bf16x2(__uint_as_float(acc[g * 8 + 0]) * inv, __uint_as_float(acc[g * 8 + 1]) * inv);
                    pk.y = pack_bf16x2(__uint_as_float(acc[g * 8 + 2]) * inv, __uint_as_float(acc[g * 8 + 3]) * inv);
                    pk.z = pack_bf16x2(__uint_as_float(acc[g * 8 + 4]) * inv, __uint_as_float(acc[g * 8 + 5]) * inv);
                    pk.w = pack_bf16x2(__uint_as_float(acc[g * 8 + 6]) * inv, __uint_as_float(acc[g * 8 + 7]) * inv);
                    reinterpret_cast<uint4*>(o)[c * 4 + g] = pk;
                }
            }
        }
        tcgen05_fence_before();
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

static long long* g_attn_dbg = nullptr;
extern "C" int tw_attention_enc_set_trace(void* dev_buf_int64_x192) { g_attn_dbg = (long long*)dev_buf_int64_x192; return 0; }

extern "C" int tw_attention_enc(const void* qkv_bf16, void* out_bf16, int32_t batch, int32_t seq,
                                int32_t heads, int64_t out_ld, void* stream) {
    using namespace tw;
    using namespace tw::attn;
    TW_REQUIRE(qkv_bf16 && out_bf16, "tw_attention_enc: null argument");
    if (tw::ensure_device(qkv_bf16)) return 1;
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
    p.T = seq; p.H = heads; p.D = D; p.out_ld = out_ld; p.out = (__nv_bfloat16*)out_bf16; p.dbg = g_attn_dbg;
    static bool attr_set[64] = {};
    const int dev = current_device();
    if (!attr_set[dev]) {
        TW_CUDA_CHECK(cudaFuncSetAttribute(attention_enc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[dev] = true;
    }
    dim3 grid((seq + 2 * BQ - 1) / (2 * BQ), heads, batch);
    attention_enc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tm, p);
    TW_CUDA_CHECK(cudaGetLastError());
    return 0;
}
