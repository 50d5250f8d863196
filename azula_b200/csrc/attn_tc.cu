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
__device__ __forceinline__ void softmax_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * SOFTMAX_WARPS) : "memory"); }

struct AttnTcParams {
    __nv_bfloat16* out;
    int64_t out_ld;
    int T, heads;
    int head_stride, k_delta, v_delta;
    float scale_log2e;
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
                            int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta, void* stream) {
    if (d != D || out_ld % 8 || !azb_aligned(out, 16) || ld % 8 || !azb_aligned(qkv, 16)) return AZB_E_UNSUPPORTED;
    if (head_stride % 8 || k_delta % 8 || v_delta % 8) return AZB_E_UNSUPPORTED;
    AttnTcParams p{};
    p.out = reinterpret_cast<__nv_bfloat16*>(out), p.out_ld = out_ld;
    p.T = (int)t, p.heads = (int)heads;
    p.head_stride = (int)head_stride, p.k_delta = (int)k_delta, p.v_delta = (int)v_delta;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
    static int forced = -1;  // AZB_ATTN_BK=64|128 pins the tile (experiments); default 64: two CTAs per SM
    if (forced < 0) {
        const char* e = getenv("AZB_ATTN_BK");
        forced = e ? atoi(e) : 0;
    }
    // measured on B200 (scripts/attn_bench.py): BK = 64 wins at every ADM / DiT sequence length (T = 1024: 90 vs
    // 100 us, T = 256: 21 vs 25 us) -- the second resident CTA hides the barrier round trips of the first
    const int bk = forced == 128 ? 128 : 64;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return bk == 128 ? launch_tc<128>(qkv, ld, n, t, p, s) : launch_tc<64>(qkv, ld, n, t, p, s);
}
