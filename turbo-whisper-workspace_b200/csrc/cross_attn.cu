// Decoder cross-attention as a persistent K/V streaming kernel (sm_100a).
//
// Replaces the attention core of WhisperAttention for the decoder's encoder_attn with one query per (decode row, head)
// ($TF/models/whisper/modeling_whisper.py:263-352 with is_cross_attention: the keys / values are the 1500 encoder
// positions, cached once per window).  The work is pure HBM streaming: every decode step reads the whole head-major
// K and V block of every row (7.68 MB per row and decoder layer) for one query vector.
//
//   cross_attn_stream_kernel   persistent CTAs (2 per SM).  One producer lane walks the CTA's work items
//                              (row, head, key split) and keeps a 3-stage ring of 128-key K and V tiles (16 KB each,
//                              TMA, SWIZZLE_128B) in flight ACROSS item boundaries, so the HBM pipe never drains between
//                              items; four consumer warps each own 32 keys of a stage:
//                                scores  S = q K^T   mma.sync.m16n8k16, q as the (replicated) A rows, K via ldmatrix
//                                softmax online, per warp, fp32, exp2
//                                output  O += P V    P from the score fragments (no shuffles), split into bf16 hi + lo
//                                                    parts that ride in rows 0-7 / 8-15 of the same A fragment (P keeps
//                                                    16 mantissa bits), V via ldmatrix.trans
//                              and hand their partial (max, sum, o[64]) to an epilogue warp, which combines them in
//                              fixed order while the consumers already stream the next item; key splits are combined by
//                              the last CTA to finish a (row, head), in split order (deterministic).
//
// Measured at 96 decode rows (ncu, profiles/r2end_ncu_cross_*_96rows.csv): 19.6 M warp instructions and 18.6 % issue-active
// against 78.4 M and 61.9 % for the scalar kernel in decode.cu, 6.5 TB/s against 6.15 TB/s.  Opt-in: see tw_dec_cross_attn.
#include "common.cuh"
#include "twb200_internal.h"
#include <algorithm>

namespace tw {
namespace xattn {

#ifndef XATTN_KCH
#define XATTN_KCH 128
#endif
#ifndef XATTN_NST
#define XATTN_NST 3
#endif
#ifndef XATTN_CTAS
#define XATTN_CTAS 2
#endif
constexpr int KCH = XATTN_KCH;               // keys per stage (128 or 64)
constexpr int NST = XATTN_NST;               // stages in flight per CTA
constexpr int NCW = 4;                       // consumer warps; warp w owns keys [KPW w, KPW (w + 1)) of every stage
constexpr int KPW = KCH / NCW;               // 32 or 16 keys per warp and stage
constexpr int NT = KPW / 8;                  // 8-key score tiles per warp and stage
constexpr int NKB = KPW / 16;                // 16-key blocks of the P V product
constexpr int THREADS = (NCW + 2) * 32;      // + the producer warp + the epilogue warp
constexpr int TILE_BYTES = KCH * 128;        // 128 keys x 64 bf16
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // K tile + V tile
constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;
constexpr int CTAS_PER_SM = XATTN_CTAS;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    CUtensorMap kmap, vmap;   // rank 3: (64 dims, S keys, H * Bc blocks), box (64, 128, 1), SWIZZLE_128B
    const __nv_bfloat16* q;   // [B, D]
    __nv_bfloat16* out;       // [B, D]
    const int* enc_row;       // [B] or null
    int B, H, D, S, Bc;
    int splits, cps, n_items; // key splits per (row, head), 128-key chunks per split, B * H * splits
    float* part;              // [B][H][splits][66]
    unsigned int* counters;   // [B][H], zero between launches
};

TW_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
TW_DEVINL void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
TW_DEVINL void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
TW_DEVINL float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// bf16 pair of (x0, x1) and the bf16 pair of what the first rounding lost
TW_DEVINL void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(x0 - h0, x1 - h1);
}

struct Item { int b, h, split, c0, c1; };
TW_DEVINL Item decode_item(const Params& p, int item, int nchunks) {
    Item it;
    it.split = item % p.splits;
    const int bh = item / p.splits;
    it.h = bh % p.H;
    it.b = bh / p.H;
    it.c0 = it.split * p.cps;
    it.c1 = min(nchunks, it.c0 + p.cps);
    return it;
}

__global__ void __launch_bounds__(THREADS, CTAS_PER_SM) cross_attn_stream_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[NST], empty[NST], done[2], freeb[2];
    __shared__ __align__(16) float s_o[2][NCW][64];
    __shared__ float s_m[2][NCW], s_l[2][NCW];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (p.S + KCH - 1) / KCH;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        for (int s = 0; s < 2; ++s) { mbar_init(&done[s], NCW); mbar_init(&freeb[s], 1); }
        fence_barrier_init();
        tma_prefetch_desc(&p.kmap);
        tma_prefetch_desc(&p.vmap);
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();   // q comes from the projection launched before this kernel

    if (warp == NCW) {
        // ---------------- producer: one lane feeds the ring, across item boundaries
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const Item it = decode_item(p, item, nchunks);
                const int eb = p.enc_row ? p.enc_row[it.b] : it.b;
                const int z = it.h * p.Bc + eb;
                for (int c = it.c0; c < it.c1; ++c) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    uint8_t* dst = smem + stage * STAGE_BYTES;
                    tma_load_3d(&p.kmap, &full[stage], dst, 0, c * KCH, z);               // rows past S arrive as zeros
                    tma_load_3d(&p.vmap, &full[stage], dst + TILE_BYTES, 0, c * KCH, z);
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    if (warp == NCW + 1) {
        // ---------------- epilogue warp: combines the four consumer warps' partials of an item (fixed order), then the
        // key splits of a (row, head) — off the streaming warps' critical path (a fence + an atomic round trip per item)
        for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
            const Item it = decode_item(p, item, nchunks);
            const int par = k & 1;
            mbar_wait(&done[par], (uint32_t)(k >> 1) & 1u);
            float M = -INFINITY, L = 0.f, O0 = 0.f, O1 = 0.f;     // lane owns dims 2 lane, 2 lane + 1
#pragma unroll
            for (int w = 0; w < NCW; ++w) M = fmaxf(M, s_m[par][w]);
#pragma unroll
            for (int w = 0; w < NCW; ++w) {
                const float wgt = (s_m[par][w] == -INFINITY) ? 0.f : ex2f(s_m[par][w] - M);
                const float2 ov = *reinterpret_cast<const float2*>(&s_o[par][w][2 * lane]);
                L = fmaf(s_l[par][w], wgt, L);
                O0 = fmaf(ov.x, wgt, O0);
                O1 = fmaf(ov.y, wgt, O1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&freeb[par]);
            uint32_t* out2 = reinterpret_cast<uint32_t*>(p.out + (size_t)it.b * p.D + it.h * 64) + lane;
            if (p.splits == 1) {
                *out2 = pack_bf16x2(O0 / L, O1 / L);
                continue;
            }
            float* my = p.part + (((size_t)it.b * p.H + it.h) * p.splits + it.split) * 66;
            *reinterpret_cast<float2*>(my + 2 + 2 * lane) = make_float2(O0, O1);
            if (lane == 0) { my[0] = M; my[1] = L; }
            __threadfence();
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                const unsigned int prev = atomicAdd(&p.counters[it.b * p.H + it.h], 1u);
                last = (prev == (unsigned)p.splits - 1);
                if (last) p.counters[it.b * p.H + it.h] = 0;   // re-arm for the next launch
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (!last) continue;
            __threadfence();
            const float* base = p.part + ((size_t)it.b * p.H + it.h) * p.splits * 66;
            float Mg = -INFINITY;
            for (int sp = 0; sp < p.splits; ++sp) Mg = fmaxf(Mg, __ldcg(base + sp * 66));
            float Lg = 0.f, G0 = 0.f, G1 = 0.f;
            for (int sp = 0; sp < p.splits; ++sp) {
                const float ms = __ldcg(base + sp * 66);
                const float wgt = (ms == -INFINITY) ? 0.f : ex2f(ms - Mg);
                const float2 ov = __ldcg(reinterpret_cast<const float2*>(base + sp * 66 + 2 + 2 * lane));
                Lg = fmaf(__ldcg(base + sp * 66 + 1), wgt, Lg);
                G0 = fmaf(ov.x, wgt, G0);
                G1 = fmaf(ov.y, wgt, G1);
            }
            *out2 = pack_bf16x2(G0 / Lg, G1 / Lg);
        }
        return;
    }

    // ---------------- consumers
    const int g = lane >> 2, t = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    int par = 0;                         // item parity: the hand-over buffers are double-buffered
    // ldmatrix lane addressing inside a tile: row r at r * 128 B, 16-byte piece c at ((c ^ (r & 7)) << 4)
    const int k_row0 = warp * KPW + (lane & 7);                       // + nt * 8
    const int k_piece = lane >> 3;                                   // + 4 * half
    const int v_row0 = warp * KPW + (lane & 7) + ((lane >> 3) & 1) * 8;   // + kb * 16
    const int v_piece = lane >> 4;                                   // + 2 * dp

    uint32_t qa[4][2], qn[4][2];          // this item's query fragments, and the next item's (loaded one item ahead)
    auto load_q = [&](const Item& it, uint32_t (&qa)[4][2]) {
        const uint32_t* q32 = reinterpret_cast<const uint32_t*>(p.q + (size_t)it.b * p.D + it.h * 64);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qa[ks][0] = __ldg(q32 + ks * 8 + t);        // dims 16 ks + 2 t, + 1
            qa[ks][1] = __ldg(q32 + ks * 8 + 4 + t);    // dims 16 ks + 8 + 2 t, + 1
        }
    };

    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, par ^= 1, ++k) {
        const Item it = decode_item(p, item, nchunks);
        if (k == 0) load_q(it, qa);
        else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) { qa[ks][0] = qn[ks][0]; qa[ks][1] = qn[ks][1]; }
        }
        if (item + (int)gridDim.x < p.n_items) load_q(decode_item(p, item + gridDim.x, nchunks), qn);
        float m = -INFINITY, l = 0.f;      // running maximum in log2 units (score * log2 e), partial row sum of this lane
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

        for (int c = it.c0; c < it.c1; ++c) {
            mbar_wait(&full[stage], phase);
            const uint32_t kbase = smem_u32(smem + stage * STAGE_BYTES), vbase = kbase + TILE_BYTES;
            // scores of this warp's 32 keys: 4 n-tiles of 8 keys; lane (g, t) ends up with keys 8 nt + 2 t, + 1
            float sc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
                const int row = k_row0 + nt * 8;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r0, r1, r2, r3;
                    ldsm_x4(kbase + row * 128 + (((k_piece + 4 * half) ^ (row & 7)) << 4), r0, r1, r2, r3);
                    mma16816(sc[nt], qa[2 * half][0], 0u, qa[2 * half][1], 0u, r0, r1);
                    mma16816(sc[nt], qa[2 * half + 1][0], 0u, qa[2 * half + 1][1], 0u, r2, r3);
                }
            }
            float mx = -INFINITY;
            const int key0 = c * KCH + warp * KPW + 2 * t;
            if (c * KCH + KCH > p.S) {        // ragged last chunk: keys past S are masked
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (key0 + nt * 8 >= p.S) sc[nt][0] = -INFINITY;
                    if (key0 + nt * 8 + 1 >= p.S) sc[nt][1] = -INFINITY;
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                sc[nt][0] *= LOG2E;
                sc[nt][1] *= LOG2E;
                mx = fmaxf(mx, fmaxf(sc[nt][0], sc[nt][1]));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            if (mx != -INFINITY) {            // uniform over the warp (every lane holds the warp's maximum)
                if (mx > m) {
                    const float alpha = (m == -INFINITY) ? 0.f : ex2f(m - mx);
                    l *= alpha;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { o[i][0] *= alpha; o[i][1] *= alpha; o[i][2] *= alpha; o[i][3] *= alpha; }
                    m = mx;
                }
                float pr[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    pr[nt][0] = ex2f(sc[nt][0] - m);
                    pr[nt][1] = ex2f(sc[nt][1] - m);
                    l += pr[nt][0] + pr[nt][1];
                }
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    uint32_t a0, a1, a2, a3;          // rows 0-7: bf16(P); rows 8-15: what that rounding lost
                    split_bf16x2(pr[2 * kb][0], pr[2 * kb][1], a0, a1);
                    split_bf16x2(pr[2 * kb + 1][0], pr[2 * kb + 1][1], a2, a3);
                    const int row = v_row0 + kb * 16;
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        uint32_t r0, r1, r2, r3;
                        ldsm_x4_trans(vbase + row * 128 + (((v_piece + 2 * dp) ^ (row & 7)) << 4), r0, r1, r2, r3);
                        mma16816(o[2 * dp], a0, a1, a2, a3, r0, r1);
                        mma16816(o[2 * dp + 1], a0, a1, a2, a3, r2, r3);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == NST) { stage = 0; phase ^= 1u; }
        }

        // ---- end of the item: hand the warp's (max, sum, o[64]) to the epilogue warp and go on streaming
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        mbar_wait(&freeb[par], ((uint32_t)(k >> 1) & 1u) ^ 1u);     // the epilogue warp has read this buffer's last use
        if (g == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s_o[par][warp][8 * i + 2 * t] = o[i][0] + o[i][2];
                s_o[par][warp][8 * i + 2 * t + 1] = o[i][1] + o[i][3];
            }
            if (t == 0) { s_m[par][warp] = m; s_l[par][warp] = l; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[par]);
    }
}

// splits per (row, head): the SMs take the items round-robin (CTA c and c + #SM share an SM), so the cost of a choice is
// the items of the fullest SM times (chunks per item + a fixed per-item overhead); fewest splits among equals
static void choose_splits(int rows_heads, int nchunks, int cap, int sms, int& splits, int& cps) {
    double best = 1e30;
    splits = 1;
    cps = nchunks;
    for (int s = 1; s <= cap && s <= nchunks; ++s) {
        const int c = (nchunks + s - 1) / s, sp = (nchunks + c - 1) / c;
        if (sp != s) continue;
        const long long items = (long long)rows_heads * sp;
        const double cost = (double)((items + sms - 1) / sms) * (c + 0.35);
        if (cost < best - 1e-9) { best = cost; splits = sp; cps = c; }
    }
}

static std::atomic<unsigned long long> g_attr_done{0};

}  // namespace xattn

// Launches the streaming kernel when the K/V layout allows it (64-wide head-major rows, dense [head][window][key][64]
// blocks); returns -1 when it does not apply and the caller must use the scalar kernel.
int cross_attn_stream_launch(const void* q, void* out, const void* k, const void* v, long long row_stride,
                             long long batch_stride, long long head_stride, const int* enc_row, int S, int B, int H,
                             int split_cap, float* part, unsigned int* counters, cudaStream_t stream, bool pdl) {
    using namespace xattn;
    if (row_stride != 64 || batch_stride != (long long)S * 64 || head_stride % batch_stride != 0 || S < KCH) return -1;
    if ((reinterpret_cast<uintptr_t>(k) & 127) || (reinterpret_cast<uintptr_t>(v) & 127)) return -1;
    const int Bc = (int)(head_stride / batch_stride);
    Params p;
    const uint64_t dims[3] = {64, (uint64_t)S, (uint64_t)H * Bc};
    const uint64_t strides[2] = {128, (uint64_t)S * 128};
    const uint32_t box[3] = {64, KCH, 1};
    if (encode_tensor_map(&p.kmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (encode_tensor_map(&p.vmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    p.q = (const __nv_bfloat16*)q; p.out = (__nv_bfloat16*)out; p.enc_row = enc_row;
    p.B = B; p.H = H; p.D = H * 64; p.S = S; p.Bc = Bc;
    const int nchunks = (S + KCH - 1) / KCH, sms = num_sms();
    choose_splits(B * H, nchunks, (part && counters) ? split_cap : 1, sms, p.splits, p.cps);
    p.n_items = B * H * p.splits;
    p.part = part; p.counters = counters;
    if (device_needs_setup(g_attr_done)) {
        TW_CUDA_CHECK(cudaFuncSetAttribute(cross_attn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        mark_device_done(g_attr_done);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min(p.n_items, CTAS_PER_SM * sms));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    TW_CUDA_CHECK(cudaLaunchKernelEx(&cfg, cross_attn_stream_kernel, p));
    TW_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace tw

// host-only view of the split choice (tests/test_host_logic.py checks the balance it promises)
extern "C" int tw_cross_attn_plan(int32_t rows, int32_t heads, int32_t src_len, int32_t split_cap, int32_t sms,
                                  int32_t* splits, int32_t* chunks_per_split, int32_t* grid) {
    using namespace tw::xattn;
    TW_REQUIRE(rows > 0 && heads > 0 && src_len >= KCH && split_cap >= 1 && sms > 0 && splits && chunks_per_split && grid,
               "tw_cross_attn_plan: bad argument");
    int sp = 1, cps = 1;
    choose_splits(rows * heads, (src_len + KCH - 1) / KCH, split_cap, sms, sp, cps);
    *splits = sp;
    *chunks_per_split = cps;
    *grid = std::min(rows * heads * sp, CTAS_PER_SM * sms);
    return 0;
}
