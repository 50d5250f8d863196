// Self-attention on the 5th-generation tensor cores (sm_100a), head width 64:
//
//   out = softmax(q k^T / sqrt(d)) v        per (image, head), logits never leave the SM
//
// for AttentionBlock of the ADM U-Net (azula/plugins/adm/_src/unet.py:328-345,361-381; the reference
// materialises the T x T logits in fp32, :343) and MultiheadSelfAttention of the in-repo DiT
// (azula/nn/attention.py:110-116, F.scaled_dot_product_attention).
//
// One CTA = 128 queries of one (image, head).  Warp-specialised:
//   warp 8      TMA producer: Q tile once, then K (pass 1) and K + V (pass 2) tiles of 128 keys through a
//               3-stage mbarrier ring.  One 3-d tensor map (channels, tokens, images) serves q, k and v, which
//               live in the same qkv buffer at different channel offsets; rows beyond T are zero-filled.
//   warp 9      MMA issuer (one thread): S = Q K^T as tcgen05.mma M=128 N=128 K=16 x4 into one of two TMEM
//               accumulators; O += P V as M=128 N=64 K=16 x8 with V as an MN-major B operand straight from
//               its natural (keys x channels) layout -- no transpose pass.
//   warps 0-7   softmax: thread = (query row, half of the key columns); tcgen05.ld delivers one accumulator
//               row per lane, so row maxima and sums need no shuffles (the two halves meet in shared memory).  P is written as bf16 into a 128-byte-swizzled K-major
//               shared-memory tile that the P V MMA reads as its A operand.
//
// TWO passes over the keys instead of the online-softmax rescaling of the accumulator: pass 1 computes the
// exact row maximum (Q K^T only), pass 2 recomputes S, exponentiates against the final maximum and
// accumulates O in TMEM without ever touching it from registers.  The extra Q K^T costs tensor time that is
// otherwise idle (the kernel is bound by the exponentials), K tiles come from L2, and the result does not
// depend on the tile order.

#include "common.cuh"
#include "tc.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int BQ = 128;    // queries per CTA
constexpr int D = 64;      // head width
constexpr int STAGES = 3;
constexpr int SOFTMAX_WARPS = 8;  // warps w and w + 4 share TMEM lane quarter w % 4 and split the key columns
constexpr int THREADS = 32 * (SOFTMAX_WARPS + 2);
constexpr int Q_BYTES = BQ * D * 2;

// BK = keys per tile.  128: one CTA per SM (long sequences); 64: 192 TMEM columns and 97 KiB of shared memory,
// so two CTAs share an SM and overlap each other's prologue / epilogue (short sequences).
template <int BK>
struct Cfg {
    static constexpr int TILE_BYTES = BK * D * 2;    // one K or V tile
    static constexpr int P_BYTES = BQ * BK * 2;      // BK / 64 blocks of 128 rows x 128 bytes
    static constexpr int SMEM = Q_BYTES + STAGES * 2 * TILE_BYTES + 2 * P_BYTES + 1024;
    static constexpr uint32_t TMEM_COLS = BK == 128 ? 512 : 256;  // S0 [0,BK) S1 [BK,2BK) O [2BK,2BK+64)
    static constexpr int HALF = BK / 2;              // key columns per softmax warp
    static constexpr int CTAS_PER_SM = BK == 128 ? 1 : 2;
};

__device__ __forceinline__ float ex2_fast(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// Two positive fp32 values -> packed bf16 pair, round to nearest (ties away) with integer adds and one byte permute: the
// conversion instruction (F2FP.BF16.PACK_AB) shares the MUFU pipe with the exponentials (ncu: XU pipe 1.5 x the ex2 count).
__device__ __forceinline__ uint32_t pack_bf16_pos(float lo, float hi) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo) + 0x8000u), "r"(__float_as_uint(hi) + 0x8000u));
    return r;
}
__device__ __forceinline__ void softmax_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * SOFTMAX_WARPS) : "memory"); }

struct AttnTcParams {
    __nv_bfloat16* out;
    int64_t out_ld;
    int T, heads;
    int head_stride, k_delta, v_delta;
    float scale_log2e;
    int turns;         // short-sequence kernel: the two softmax groups exponentiate in turns (AZB_ATTN_TURNS, A/B switch)
    long long* trace;  // diagnosis only (AZB_ATTN_TRACE=<file>, short-sequence kernel): clock64 per (CTA, role, item, event)
};

template <int BK>
__global__ void __launch_bounds__(THREADS, Cfg<BK>::CTAS_PER_SM)
    attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const AttnTcParams p) {
    pdl_enter();
    using C = Cfg<BK>;
    constexpr int TILE_BYTES = C::TILE_BYTES, P_BYTES = C::P_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + Q_BYTES;
    const uint32_t p_smem = kv_smem + STAGES * 2 * TILE_BYTES;

    __shared__ __align__(8) uint64_t bar_q, bar_o;
    __shared__ __align__(8) uint64_t bar_kv_full[STAGES], bar_kv_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_s_full[2], bar_s_empty[2], bar_p_full[2], bar_p_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float red[2][BQ];  // partial row maxima, then partial row sums, of the two column halves

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, hd = blockIdx.y, img = blockIdx.z;
    const int n_tiles = (p.T + BK - 1) / BK;
    const int ch_q = hd * p.head_stride, ch_k = ch_q + p.k_delta, ch_v = ch_q + p.v_delta;

    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar_q), 1);
        tc::mbar_init(tc::smem_u32(&bar_o), 1);
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_kv_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_kv_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_s_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_s_empty[s]), SOFTMAX_WARPS);
            tc::mbar_init(tc::smem_u32(&bar_p_full[s]), SOFTMAX_WARPS);
            tc::mbar_init(tc::smem_u32(&bar_p_empty[s]), 1);
        }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_q);
        tc::prefetch_tmap(&tmap_kv);
    }
    if (warp == SOFTMAX_WARPS + 1) {
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), C::TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;

    if (warp == SOFTMAX_WARPS) {
        // ===== TMA producer (one ELECTED lane: operands go straight to uniform registers) =====
        if (tc::elect_one()) {
            tc::mbar_expect_tx(tc::smem_u32(&bar_q), Q_BYTES);
            tc::tma_load_3d(q_smem, &tmap_q, tc::smem_u32(&bar_q), ch_q, q0, img);
            for (int it = 0; it < 2 * n_tiles; ++it) {
                const int s = it % STAGES;
                tc::mbar_wait(tc::smem_u32(&bar_kv_empty[s]), ((it / STAGES) & 1) ^ 1);
                const bool second = it >= n_tiles;
                const int key0 = (second ? it - n_tiles : it) * BK;
                const uint32_t full = tc::smem_u32(&bar_kv_full[s]);
                const uint32_t k_dst = kv_smem + s * 2 * TILE_BYTES;
                tc::mbar_expect_tx(full, second ? 2 * TILE_BYTES : TILE_BYTES);
                tc::tma_load_3d(k_dst, &tmap_kv, full, ch_k, key0, img);
                if (second) tc::tma_load_3d(k_dst + TILE_BYTES, &tmap_kv, full, ch_v, key0, img);
            }
        }
    } else if (warp == SOFTMAX_WARPS + 1) {
        // ===== MMA issuer =====
        if (tc::elect_one()) {
            constexpr uint32_t idesc_s = tc::idesc_bf16_f32(BQ, BK);
            constexpr uint32_t idesc_o = tc::idesc_bf16_f32_b_mn(BQ, D);
            const uint64_t desc_q = tc::smem_desc_sw128(q_smem);
            const uint32_t tmem_o = tmem_base + 2 * BK;
            tc::mbar_wait(tc::smem_u32(&bar_q), 0);

            auto issue_s = [&](int g) {  // S tile number g (pass 1: g < n_tiles) from ring slot g
                const int s = g % STAGES;
                tc::mbar_wait(tc::smem_u32(&bar_kv_full[s]), (g / STAGES) & 1);
                tc::mbar_wait(tc::smem_u32(&bar_s_empty[g & 1]), ((g >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint64_t desc_k = tc::smem_desc_sw128(kv_smem + s * 2 * TILE_BYTES);
                const uint32_t acc = tmem_base + (uint32_t)((g & 1) * BK);
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    tc::mma_f16_ss(acc, desc_q + (uint64_t)(2 * k), desc_k + (uint64_t)(2 * k), idesc_s, k != 0);
                tc::mma_commit(tc::smem_u32(&bar_s_full[g & 1]));
            };

            for (int g = 0; g < n_tiles; ++g) {  // pass 1: logits only; the slot is free once the MMAs retire
                issue_s(g);
                tc::mma_commit(tc::smem_u32(&bar_kv_empty[g % STAGES]));
            }
            issue_s(n_tiles);
            for (int j = 0; j < n_tiles; ++j) {  // pass 2
                if (j + 1 < n_tiles) issue_s(n_tiles + j + 1);  // overlaps the softmax of tile j
                const int s = (n_tiles + j) % STAGES;
                tc::mbar_wait(tc::smem_u32(&bar_p_full[j & 1]), (j >> 1) & 1);
                tc::fence_after_sync();
                const uint32_t v_src = kv_smem + s * 2 * TILE_BYTES + TILE_BYTES;
                const uint32_t p_src = p_smem + (j & 1) * P_BYTES;
#pragma unroll
                for (int kb = 0; kb < BK / 64; ++kb) {
                    const uint64_t desc_p = tc::smem_desc_sw128(p_src + kb * (BQ * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // V tile: key rows of 128 bytes, 8-row groups every 1024 bytes; 16 keys = 2048 bytes
                        const uint64_t desc_v = tc::smem_desc_sw128(v_src + (kb * 4 + k) * 2048);
                        tc::mma_f16_ss(tmem_o, desc_p + (uint64_t)(2 * k), desc_v, idesc_o, (j | kb | k) != 0);
                    }
                }
                tc::mma_commit(tc::smem_u32(&bar_kv_empty[s]));
                tc::mma_commit(tc::smem_u32(&bar_p_empty[j & 1]));
            }
            tc::mma_commit(tc::smem_u32(&bar_o));
        }
    } else {
        // ===== softmax: thread = (query row, half of the key columns) =====
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;  // == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int col0 = half * C::HALF;         // first key column of this warp inside a tile
        float m = -INFINITY;
        for (int g = 0; g < n_tiles; ++g) {  // pass 1: exact row maximum
            tc::mbar_wait(tc::smem_u32(&bar_s_full[g & 1]), (g >> 1) & 1);
            tc::fence_after_sync();
#pragma unroll
            for (int c = 0; c < C::HALF / 32; ++c) {
                uint32_t acc[32];
                tc::tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)((g & 1) * BK + col0 + c * 32), acc);
                tc::tmem_ld_wait();
                const int key = g * BK + col0 + c * 32;
                if (key + 32 <= p.T) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(acc[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (key + i < p.T) m = fmaxf(m, __uint_as_float(acc[i]));
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_s_empty[g & 1]));
        }
        red[half][row] = m;
        softmax_sync();
        m = fmaxf(m, red[half ^ 1][row]);
        softmax_sync();  // red is reused for the sums
        const float c1 = p.scale_log2e, c0 = -m * p.scale_log2e;
        float sum = 0.f;
        for (int j = 0; j < n_tiles; ++j) {  // pass 2: P = exp2(c1 s + c0) -> shared memory
            const int g = n_tiles + j;
            tc::mbar_wait(tc::smem_u32(&bar_s_full[g & 1]), (g >> 1) & 1);
            tc::mbar_wait(tc::smem_u32(&bar_p_empty[j & 1]), ((j >> 1) & 1) ^ 1);
            tc::fence_after_sync();
            const uint32_t p_dst = p_smem + (j & 1) * P_BYTES + row * 128;
#pragma unroll
            for (int c = 0; c < C::HALF / 32; ++c) {
                uint32_t acc[32];
                tc::tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)((g & 1) * BK + col0 + c * 32), acc);
                tc::tmem_ld_wait();
                const int ko = col0 + c * 32;  // key offset inside the tile
                const int key = j * BK + ko;
                uint32_t packed[16];
                if (key + 32 <= p.T) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e0 = ex2_fast(fmaf(__uint_as_float(acc[2 * i]), c1, c0));
                        const float e1 = ex2_fast(fmaf(__uint_as_float(acc[2 * i + 1]), c1, c0));
                        sum += e0 + e1;
                        __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
                        packed[i] = *reinterpret_cast<uint32_t*>(&t);
                    }
                } else {  // the ragged end of the sequence: keys >= T contribute nothing
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float e0 = ex2_fast(fmaf(__uint_as_float(acc[2 * i]), c1, c0));
                        float e1 = ex2_fast(fmaf(__uint_as_float(acc[2 * i + 1]), c1, c0));
                        if (key + 2 * i >= p.T) e0 = 0.f;
                        if (key + 2 * i + 1 >= p.T) e1 = 0.f;
                        sum += e0 + e1;
                        __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
                        packed[i] = *reinterpret_cast<uint32_t*>(&t);
                    }
                }
                // 32 keys = four 16-byte slots of 64-key block ko / 64; slot q of row r lives at q ^ (r & 7)
                const uint32_t blk = p_dst + (ko >> 6) * (BQ * 128);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t slot = (uint32_t)((((ko & 63) >> 3) + q) ^ (row & 7));
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(blk + (slot << 4)), "r"(packed[4 * q]),
                                 "r"(packed[4 * q + 1]), "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                                 : "memory");
                }
            }
            tc::fence_before_sync();
            tc::fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(tc::smem_u32(&bar_s_empty[g & 1]));
                tc::mbar_arrive(tc::smem_u32(&bar_p_full[j & 1]));
            }
        }
        red[half][row] = sum;
        softmax_sync();
        sum += red[half ^ 1][row];
        // epilogue: O / sum -> bf16 -> global; warp half h stores channels [32 h, 32 h + 32) of its rows
        tc::mbar_wait(tc::smem_u32(&bar_o), 0);
        tc::fence_after_sync();
        const float inv = 1.0f / sum;
        const int q = q0 + row;
        __nv_bfloat16* dst = p.out + ((int64_t)img * p.T + q) * p.out_ld + hd * D + half * 32;
        uint32_t acc[32];
        tc::tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)(2 * BK + half * 32), acc);
        tc::tmem_ld_wait();
        if (q < p.T) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(acc[8 * v + 2 * i]) * inv,
                                                             __uint_as_float(acc[8 * v + 2 * i + 1]) * inv);
                    w[i] = *reinterpret_cast<uint32_t*>(&t);
                }
                *reinterpret_cast<uint4*>(dst + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == SOFTMAX_WARPS + 1) tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// Short sequences (T <= 256: DiT-B/2 tokens, the 16 x 16 and 8 x 8 levels of ADM): a PERSISTENT kernel, one CTA per SM,
// work item = one (image, head) = 2 tiles of 128 queries x all (<= 256) keys.  Everything a tile needs stays on chip:
//   tensor memory   group g owns columns [256 g, 256 g + 256): the fp32 logits S_g of its 128 queries, then P_g as bf16
//                   PAIRS written back by tcgen05.st over logits that are already consumed, and the output accumulator
//                   O_g (64 columns) -- P is the A operand of P V straight from tensor memory (tcgen05.mma with A in
//                   TMEM), so the probabilities never touch shared memory.  T > 224: P chunks 0-3 in [0, 64), O in
//                   [64, 128), P chunks 4-7 in [128, 192), and P V starts on the first half while the second is still
//                   being exponentiated; shorter sequences: P in [0, 16 chunks), O in [128, 192)
//   shared memory   two stages of {Q, K} (64 KiB) and two of V (32 KiB): the {Q, K} stage of item j + 2 is released as
//                   soon as the two logit MMAs of item j retire, a whole item ahead of its use
//   warps 0-3 / 4-7 softmax groups 0 / 1: thread = query row (the whole row of logits sits in the thread's TMEM lane: no
//                   shuffles, no exchange through shared memory); exact two-pass softmax out of tensor memory (row
//                   maximum, then P = exp2(c1 s + c0), the sum of the unrounded values), O / sum -> bf16 -> global
//   warp 8          TMA producer (Q, K, V of an item are three boxes of 256 tokens x 64 channels of one tensor map; rows
//                   beyond T are zero-filled)
//   warps 9, 10     MMA issuers, one per group (a single thread each, asleep on the group's barriers between issues):
//                   S_g of the next item once group g has drained O_g, P_g V once P_g is written.  Nothing couples the
//                   groups except the stage rings, so they drift apart and one exponentiates (the MUFU pipe, 16 / clk /
//                   SM, is the floor of this kernel) while the other waits for its MMAs and reads O.  (A single polling
//                   issuer was measured first: its spin loop took issue slots from the softmax warps of its scheduler.)
// QKNORM: the per-head RMS normalisation of q and k (azula/nn/attention.py:103, elementwise_affine = False) is applied
// to the landed Q / K rows IN PLACE in shared memory by the softmax threads (row t of both by thread t, one item ahead of
// the logits), with the arithmetic of azb_segment_rmsnorm_bf16 -- the separate pass over the qkv projection is gone.
constexpr int TQ = 256;                       // tokens per item
constexpr int T256_GROUP_WARPS = 8;           // softmax warps per group: warps w and w + 4 share TMEM lane quarter w % 4 and
                                              // split the key columns ([0, 128) / [128, 256))
constexpr int T256_SOFTMAX = 2 * T256_GROUP_WARPS;
constexpr int T256_THREADS = 32 * (T256_SOFTMAX + 3);  // + TMA producer + one MMA issuer per group
constexpr int T256_QK_STAGE = 2 * TQ * 128;   // Q + K
constexpr int T256_V_STAGE = TQ * 128;
constexpr int T256_SMEM = 2 * T256_QK_STAGE + 2 * T256_V_STAGE + 2 * BQ * 128 + 1024;  // + one output staging tile per group

template <bool QKNORM>
__global__ void __launch_bounds__(T256_THREADS, 1)
    attention_t256_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const AttnTcParams p,
                          const int items, const float qk_eps) {
    pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    auto q_smem = [&](int s) -> uint32_t { return smem_base + (uint32_t)s * T256_QK_STAGE; };
    auto k_smem = [&](int s) -> uint32_t { return smem_base + (uint32_t)s * T256_QK_STAGE + TQ * 128; };
    auto v_smem = [&](int s) -> uint32_t { return smem_base + 2 * T256_QK_STAGE + (uint32_t)s * T256_V_STAGE; };

    __shared__ __align__(8) uint64_t bar_qk_full[2], bar_qk_ready[2], bar_qk_free[2], bar_v_full[2], bar_v_free[2];
    __shared__ __align__(8) uint64_t bar_s[2], bar_p[2][2], bar_o[2], bar_sfree[2], bar_turn[2];
    __shared__ uint32_t tmem_slot;
    // [group][row]: the two key halves of a row meet here in two steps (half 1 deposits, half 0 combines and deposits the
    // result) -- one array per group is all the static shared memory that is left next to 225 KiB of stages
    __shared__ float red[2][BQ];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto mark = [&](int role, int j, int e) {  // role 0 / 1: softmax groups, 2 / 3: their MMA issuers
        if (p.trace && j < 8) p.trace[(((int64_t)blockIdx.x * 4 + role) * 8 + j) * 8 + e] = clock64();
    };
    const int n_my = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int tiles = p.T > BQ ? 2 : 1;       // query tiles per item (group 1 idles on sequences of <= 128 tokens)
    const int nkc = (p.T + 31) >> 5;          // 32-key chunks that hold valid keys
    // Tensor-memory columns of a group: logits chunk c in [32 c, 32 c + 32); P chunks 0-3 (key half 0) in [0, 64), O in
    // [64, 128), P chunks 4-7 (key half 1) in [128, 192): a half's P only ever overwrites logits of the SAME half that are
    // already consumed, and O overwrites logits of half 0, whose P is complete before the first P V step is issued.
    constexpr uint32_t O_COL = 64, P1_COL = 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_qk_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_qk_ready[s]), T256_SOFTMAX);
            tc::mbar_init(tc::smem_u32(&bar_qk_free[s]), tiles);
            tc::mbar_init(tc::smem_u32(&bar_v_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_v_free[s]), tiles);
            tc::mbar_init(tc::smem_u32(&bar_s[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_p[s][0]), 4);
            tc::mbar_init(tc::smem_u32(&bar_p[s][1]), 4);
            tc::mbar_init(tc::smem_u32(&bar_o[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_sfree[s]), T256_GROUP_WARPS);
            tc::mbar_init(tc::smem_u32(&bar_turn[s]), T256_GROUP_WARPS);
        }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap);
        tc::prefetch_tmap(&tmap_out);
    }
    if (warp == T256_SOFTMAX + 1) {
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), 512);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;

    if (warp == T256_SOFTMAX) {
        // ===== TMA producer =====
        if (tc::elect_one()) {
            for (int j = 0; j < n_my; ++j) {
                const int item = (int)blockIdx.x + j * (int)gridDim.x;
                const int img = item / p.heads, hd = item - img * p.heads;
                const int ch_q = hd * p.head_stride, s = j & 1;
                const uint32_t par = ((uint32_t)(j >> 1) & 1u) ^ 1u;
                tc::mbar_wait(tc::smem_u32(&bar_qk_free[s]), par);
                const uint32_t full = tc::smem_u32(&bar_qk_full[s]);
                tc::mbar_expect_tx(full, T256_QK_STAGE);
                tc::tma_load_3d(q_smem(s), &tmap, full, ch_q, 0, img);
                tc::tma_load_3d(k_smem(s), &tmap, full, ch_q + p.k_delta, 0, img);
                tc::mbar_wait(tc::smem_u32(&bar_v_free[s]), par);
                const uint32_t vfull = tc::smem_u32(&bar_v_full[s]);
                tc::mbar_expect_tx(vfull, T256_V_STAGE);
                tc::tma_load_3d(v_smem(s), &tmap, vfull, ch_q + p.v_delta, 0, img);
            }
        }
    } else if (warp > T256_SOFTMAX) {
        // ===== MMA issuers: one per group, in order, asleep on the group's barriers between issues =====
        const int g = warp - (T256_SOFTMAX + 1);
        if (g < tiles && tc::elect_one()) {
            constexpr uint32_t idesc_s = tc::idesc_bf16_f32(BQ, TQ);
            constexpr uint32_t idesc_o = tc::idesc_bf16_f32_b_mn(BQ, D);
            const uint32_t tmem_g = tmem_base + (uint32_t)(g * 256);
            const int ksteps = 2 * nkc;  // 16 keys per P V step
            for (int j = 0; j < n_my; ++j) {
                const int s = j & 1;
                const uint32_t par = (uint32_t)(j >> 1) & 1u;
                // S_g = Q_g K^T once the group has drained O_g of the previous item and Q / K have landed (and are normalised)
                tc::mbar_wait(tc::smem_u32(&bar_sfree[g]), ((uint32_t)j & 1u) ^ 1u);
                mark(2 + g, j, 0);
                tc::mbar_wait(tc::smem_u32(QKNORM ? &bar_qk_ready[s] : &bar_qk_full[s]), par);
                mark(2 + g, j, 1);
                tc::fence_after_sync();
                const uint64_t desc_q = tc::smem_desc_sw128(q_smem(s) + (uint32_t)g * (BQ * 128));
                const uint64_t desc_k = tc::smem_desc_sw128(k_smem(s));
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    tc::mma_f16_ss(tmem_g, desc_q + (uint64_t)(2 * k), desc_k + (uint64_t)(2 * k), idesc_s, k != 0);
                tc::mma_commit(tc::smem_u32(&bar_s[g]));
                if (j + 2 < n_my) tc::mma_commit(tc::smem_u32(&bar_qk_free[s]));  // (both groups' logit MMAs: count = tiles)
                // P_g V -> O_g: the steps of key half 0, then those of key half 1, each as soon as its P is written
                tc::mbar_wait(tc::smem_u32(&bar_v_full[s]), par);
                mark(2 + g, j, 2);
                const uint32_t v_src = v_smem(s);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    tc::mbar_wait(tc::smem_u32(&bar_p[g][half]), (uint32_t)j & 1u);
                    mark(2 + g, j, 3 + half);
                    tc::fence_after_sync();
                    const int k0 = half ? 8 : 0, k1 = half ? ksteps : (ksteps < 8 ? ksteps : 8);
                    for (int k = k0; k < k1; ++k)  // 8 packed columns of P, 2048 bytes of V per step
                        tc::mma_f16_ts(tmem_g + O_COL, tmem_g + (k < 8 ? 8u * (uint32_t)k : P1_COL + 8u * (uint32_t)(k - 8)),
                                       tc::smem_desc_sw128(v_src + (uint32_t)k * 2048u), idesc_o, k != 0);
                }
                tc::mma_commit(tc::smem_u32(&bar_o[g]));
                mark(2 + g, j, 5);
                if (j + 2 < n_my) tc::mma_commit(tc::smem_u32(&bar_v_free[s]));
            }
        }
    } else {
        // ===== softmax groups: thread = (query row, key half) =====
        const int g = warp >> 3;
        const int half = (warp >> 2) & 1;
        const int row = (warp & 3) * 32 + lane;  // == TMEM lane; query g * 128 + row
        const uint32_t tmem_g = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 256);
        const int t = (int)threadIdx.x;          // 0 .. 511: Q row t, or K row t - 256, is normalised by this thread
        const uint32_t o_stage = smem_base + 2 * T256_QK_STAGE + 2 * T256_V_STAGE + (uint32_t)g * (BQ * 128);
        const bool tracer_thread = (warp & 7) == 0 && lane == 0;  // the group's thread that talks to the TMA unit
        bool store_pending = false;
        auto group_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); };
        // this thread's chunks: [c_lo, c_hi) of the 32-key chunks (key half 0: 0-3, half 1: 4-7, clipped to the valid ones)
        const int c_lo = 4 * half, c_hi = nkc < 4 * half + 4 ? nkc : 4 * half + 4;
        auto normalise = [&](int j) {
            const int s = j & 1;
            tc::mbar_wait(tc::smem_u32(&bar_qk_full[s]), (uint32_t)(j >> 1) & 1u);
            {
                // the eight 16-byte slots of a row in an order that keeps a quarter warp on distinct banks (a sum of
                // squares does not care which channels a slot holds, and every slot is scaled alike)
                const uint32_t base = (t < TQ ? q_smem(s) : k_smem(s)) + (uint32_t)(t & (TQ - 1)) * 128u;
                uint32_t w[8][4];
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(w[i][0]), "=r"(w[i][1]), "=r"(w[i][2]), "=r"(w[i][3])
                                 : "r"(base + (uint32_t)(((i + t) & 7) << 4))
                                 : "memory");
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float a = bf16_bits_to_f32(w[i][c] & 0xffffu), b = __uint_as_float(w[i][c] & 0xffff0000u);
                        ss = fmaf(a, a, ss), ss = fmaf(b, b, ss);
                    }
                const float rstd = rsqrtf(ss / (float)D + qk_eps);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        __nv_bfloat162 r2 = __floats2bfloat162_rn(bf16_bits_to_f32(w[i][c] & 0xffffu) * rstd,
                                                                  __uint_as_float(w[i][c] & 0xffff0000u) * rstd);
                        w[i][c] = *reinterpret_cast<uint32_t*>(&r2);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + (uint32_t)(((i + t) & 7) << 4)), "r"(w[i][0]),
                                 "r"(w[i][1]), "r"(w[i][2]), "r"(w[i][3])
                                 : "memory");
                }
            }
            tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) reads
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_qk_ready[s]));
        };
        if (QKNORM && n_my > 0) normalise(0);
        for (int j = 0; j < n_my; ++j) {
            if (QKNORM && j + 1 < n_my) normalise(j + 1);  // a whole item ahead of its logits
            if (g >= tiles) continue;
            const int item = (int)blockIdx.x + j * (int)gridDim.x;
            const int img = item / p.heads, hd = item - img * p.heads;
            const uint32_t par = (uint32_t)j & 1u;
            const bool tracer = tracer_thread;
            if (tracer) mark(g, j, 0);
            tc::mbar_wait(tc::smem_u32(&bar_s[g]), par);
            if (tracer) mark(g, j, 1);
            tc::fence_after_sync();
            // Pass 1: the exact row maximum (the two key halves meet in shared memory).  Not needed with QKNORM: the rows of
            // q and k are RMS-normalised, |q| |k| <= 64, so every logit is <= 64 and the constant shift 64 keeps exp2 in
            // [2^-23, 1] -- softmax does not care which shift is used, bf16 / fp32 are floating point.
            float m = (float)D;
            if constexpr (!QKNORM) {
                m = -INFINITY;
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t acc[32];
                    tc::tmem_ld_32x32b_x32(tmem_g + (uint32_t)(c * 32), acc);
                    tc::tmem_ld_wait();
                    if (c * 32 + 32 <= p.T) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(acc[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < p.T) m = fmaxf(m, __uint_as_float(acc[i]));
                    }
                }
                if (half) red[g][row] = m;
                group_sync();
                if (!half) red[g][row] = m = fmaxf(m, red[g][row]);
                group_sync();
                m = red[g][row];
                group_sync();  // red is reused for the sums
            }
            if (tracer) mark(g, j, 2);
            const float c1 = p.scale_log2e, c0 = -m * p.scale_log2e;
            float sum0 = 0.f, sum1 = 0.f;
            // The exponentials of the two groups may take turns on the MUFU pipe (AZB_ATTN_TURNS, A/B switch)
            if (tiles == 2 && p.turns) tc::mbar_wait(tc::smem_u32(&bar_turn[g]), g == 0 ? (par ^ 1u) : par);
            // Pass 2: P = exp2(c1 s + c0) as bf16 pairs over consumed logits of this thread's key half
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t acc[32];
                tc::tmem_ld_32x32b_x32(tmem_g + (uint32_t)(c * 32), acc);
                tc::tmem_ld_wait();
                uint32_t packed[16];
                if (c * 32 + 32 <= p.T) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e0 = ex2_fast(fmaf(__uint_as_float(acc[2 * i]), c1, c0));
                        const float e1 = ex2_fast(fmaf(__uint_as_float(acc[2 * i + 1]), c1, c0));
                        sum0 += e0, sum1 += e1;
                        packed[i] = pack_bf16_pos(e0, e1);
                    }
                } else {  // the ragged end of the sequence: keys >= T contribute nothing
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float e0 = ex2_fast(fmaf(__uint_as_float(acc[2 * i]), c1, c0));
                        float e1 = ex2_fast(fmaf(__uint_as_float(acc[2 * i + 1]), c1, c0));
                        if (c * 32 + 2 * i >= p.T) e0 = 0.f;
                        if (c * 32 + 2 * i + 1 >= p.T) e1 = 0.f;
                        sum0 += e0, sum1 += e1;
                        packed[i] = pack_bf16_pos(e0, e1);
                    }
                }
                tc::tmem_st_32x32b_x16(tmem_g + (c < 4 ? (uint32_t)(c * 16) : P1_COL + (uint32_t)((c - 4) * 16)), packed);
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(tc::smem_u32(&bar_p[g][half]));  // this key half of P is in tensor memory
                if (tiles == 2 && p.turns) tc::mbar_arrive(tc::smem_u32(&bar_turn[g ^ 1]));
            }
            if (tracer) mark(g, j, 4);
            float sum = sum0 + sum1;
            if (half) red[g][row] = sum;
            // O_g / sum -> bf16 -> 128-byte-swizzled staging tile -> ONE TMA store per tile (rows beyond T are clipped by
            // the tensor map).  Per-thread row stores were measured at 2000 - 4000 clk per tile: 32 lanes x 8 partial-sector
            // writes to 32 different lines per warp.  Key half h of the threads converts channels [32 h, 32 h + 32).
            // (the sums meet and the staging block is checked WHILE P V runs: nothing of it is left for after the wait)
            if (tracer && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            group_sync();  // half 1's partial sums are visible; the TMA unit has read the previous tile out of the staging block
            if (!half) red[g][row] = sum = sum + red[g][row];
            group_sync();
            const float inv = 1.0f / red[g][row];
            tc::mbar_wait(tc::smem_u32(&bar_o[g]), par);
            if (tracer) mark(g, j, 5);
            tc::fence_after_sync();
            const uint32_t my_row = o_stage + (uint32_t)row * 128u;
            {
                uint32_t acc[32];
                tc::tmem_ld_32x32b_x32(tmem_g + O_COL + (uint32_t)(half * 32), acc);
                tc::tmem_ld_wait();
                // O is in registers: the MMA warp may overwrite the group's columns with the next logits
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_sfree[g]));
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(acc[8 * v + 2 * i]) * inv,
                                                                  __uint_as_float(acc[8 * v + 2 * i + 1]) * inv);
                        w[i] = *reinterpret_cast<uint32_t*>(&t2);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my_row + (uint32_t)(((half * 4 + v) ^ (row & 7)) << 4)), "r"(w[0]),
                                 "r"(w[1]), "r"(w[2]), "r"(w[3])
                                 : "memory");
                }
            }
            tc::fence_proxy_async();
            group_sync();
            if (tracer) {
                tc::tma_store_3d(&tmap_out, o_stage, hd * D, g * BQ, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                mark(g, j, 6);
            }
            store_pending = true;
        }
        if (tracer_thread && store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // complete before exit
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == T256_SOFTMAX + 1) tc::tmem_dealloc(tmem_base, 512);
}

template <bool QKNORM>
int launch_t256(const void* qkv, int64_t ld, int64_t n, int64_t t, const AttnTcParams& p, float qk_eps, cudaStream_t stream) {
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)t, (uint64_t)n};
    uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)t};
    uint32_t box[3] = {D, TQ, 1};
    int rc = tc::make_map_bf16(&tm, qkv, 3, dims, str, box);
    if (rc) return rc == -1 ? AZB_E_DRIVER : AZB_E_SHAPE;
    CUtensorMap tmo;  // (channels, tokens, images) of the output; a store box = 128 tokens x one head
    uint64_t odims[3] = {(uint64_t)p.heads * D, (uint64_t)t, (uint64_t)n};
    uint64_t ostr[2] = {(uint64_t)p.out_ld * 2, (uint64_t)p.out_ld * 2 * (uint64_t)t};
    uint32_t obox[3] = {D, BQ, 1};
    rc = tc::make_map_bf16(&tmo, p.out, 3, odims, ostr, obox);
    if (rc) return rc == -1 ? AZB_E_DRIVER : AZB_E_SHAPE;
    static AzbPerDevice<bool> configured_dev;
    bool& configured = configured_dev.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_t256_kernel<QKNORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, T256_SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int64_t items = n * p.heads;
    if (items > 0x7fffffffLL) return AZB_E_SHAPE;
    const int sms = azb_sm_count();
    const int grid = items < sms ? (int)items : sms;
    if (const char* path = getenv("AZB_ATTN_TRACE")) {  // diagnosis: one synchronous launch with event timestamps
        AttnTcParams pt = p;
        const size_t bytes = (size_t)grid * 4 * 8 * 8 * sizeof(long long);
        if (cudaMalloc(&pt.trace, bytes) != cudaSuccess) return AZB_E_DRIVER;
        cudaMemsetAsync(pt.trace, 0, bytes, stream);
        azb_launch(attention_t256_kernel<QKNORM>, dim3((unsigned)grid), dim3(T256_THREADS), T256_SMEM, stream, tm, tmo, pt, (int)items, qk_eps);
        cudaStreamSynchronize(stream);
        long long* host = (long long*)malloc(bytes);
        cudaMemcpy(host, pt.trace, bytes, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(path, "wb")) {
            fwrite(host, 1, bytes, f);
            fclose(f);
        }
        free(host);
        cudaFree(pt.trace);
        return azb_launch_status();
    }
    azb_launch(attention_t256_kernel<QKNORM>, dim3((unsigned)grid), dim3(T256_THREADS), T256_SMEM, stream, tm, tmo, p, (int)items, qk_eps);
    return azb_launch_status();
}

template <int BK>
int launch_tc(const void* qkv, int64_t ld, int64_t n, int64_t t, const AttnTcParams& p, cudaStream_t stream) {
    using C = Cfg<BK>;
    CUtensorMap tq, tkv;
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)t, (uint64_t)n};
    uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)t};
    uint32_t box_q[3] = {D, BQ, 1}, box_kv[3] = {D, BK, 1};
    int rc = tc::make_map_bf16(&tq, qkv, 3, dims, str, box_q);
    if (!rc) rc = tc::make_map_bf16(&tkv, qkv, 3, dims, str, box_kv);
    if (rc) return rc == -1 ? AZB_E_DRIVER : AZB_E_SHAPE;
    static AzbPerDevice<bool> configured_dev;
    bool& configured = configured_dev.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((unsigned)((t + BQ - 1) / BQ), (unsigned)p.heads, (unsigned)n);
    azb_launch(attention_tc_kernel<BK>, grid, dim3(THREADS), C::SMEM, stream, tq, tkv, p);
    return azb_launch_status();
}

}  // namespace

// Returns AZB_E_UNSUPPORTED when the shape is not the one this kernel serves (the caller then uses the
// mma.sync kernel of attn.cu).
int azb_attention_tc_launch(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t, int64_t heads,
                            int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta, int qk_norm, float qk_eps,
                            void* stream) {
    if (d != D || out_ld % 8 || !azb_aligned(out, 16) || ld % 8 || !azb_aligned(qkv, 16)) return AZB_E_UNSUPPORTED;
    if (head_stride % 8 || k_delta % 8 || v_delta % 8) return AZB_E_UNSUPPORTED;
    if (qk_norm && t > 256) return AZB_E_UNSUPPORTED;  // the fused normalisation lives in the short-sequence kernel
    AttnTcParams p{};
    p.out = reinterpret_cast<__nv_bfloat16*>(out), p.out_ld = out_ld;
    p.T = (int)t, p.heads = (int)heads;
    p.head_stride = (int)head_stride, p.k_delta = (int)k_delta, p.v_delta = (int)v_delta;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
    static int turns = -1;
    if (turns < 0) {
        const char* e = getenv("AZB_ATTN_TURNS");
        turns = e ? atoi(e) : 1;
    }
    p.turns = turns;
    static int forced = -1;  // AZB_ATTN_BK=64|128 pins the tile (experiments); default 64: two CTAs per SM
    if (forced < 0) {
        const char* e = getenv("AZB_ATTN_BK");
        forced = e ? atoi(e) : 0;
    }
    // measured on B200 (scripts/attn_bench.py): BK = 64 wins at every ADM / DiT sequence length (T = 1024: 90 vs
    // 100 us, T = 256: 21 vs 25 us) -- the second resident CTA hides the barrier round trips of the first
    const int bk = forced == 128 ? 128 : 64;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    static int small = -1;  // AZB_ATTN_SMALL=0 keeps short sequences on the ring kernel below (A/B measurements)
    if (small < 0) {
        const char* e = getenv("AZB_ATTN_SMALL");
        small = e ? atoi(e) : 1;
    }
    if (t <= TQ && (small || qk_norm))
        return qk_norm ? launch_t256<true>(qkv, ld, n, t, p, qk_eps, s) : launch_t256<false>(qkv, ld, n, t, p, 0.f, s);
    return bk == 128 ? launch_tc<128>(qkv, ld, n, t, p, s) : launch_tc<64>(qkv, ld, n, t, p, s);
}
