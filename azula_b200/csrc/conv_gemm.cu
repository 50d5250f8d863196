// Convolution / linear layers as implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   out[pixel, co] = bias[co] + sum_{tap, ci} act[pixel + offset(tap), ci] * w[co, tap, ci]  (+ residual)
//
// replaces the cuDNN / cuBLAS calls behind nn.Conv2d(3x3, pad 1), nn.Conv2d(1x1), nn.Conv1d(k=1)
// and nn.Linear on the ADM path (azula/plugins/adm/_src/unet.py:182,207,213-215,277,285,471,602).
//
// Data layout: activations NHWC bf16 (pixel stride `ld` elements), weights bf16 [C_out][taps][K_tap]
// with K_tap = C_in rounded up to 64 (zero padded), accumulation fp32 in tensor memory.
//
// Per CTA: one 128-pixel x BLOCK_N output tile.  The 128 pixels are a (BN images x BH rows x BW
// columns) patch, so the A operand of filter tap (kh, kw) is ONE 4-d TMA box load at spatial offset
// (kh-1, kw-1); out-of-image coordinates are zero-filled by the TMA unit, which implements the
// padding for free.  TMA writes both operands with the 128-byte swizzle the tensor core expects;
// a single elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into TMEM; mbarriers form a
// STAGES-deep producer/consumer ring; the epilogue reads TMEM with tcgen05.ld, adds bias and the
// residual, and stores bf16 NHWC (or fp32 NCHW for the network output).

#include "common.cuh"
#include "tc.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 128 bytes of bf16 = one swizzle row
constexpr int UMMA_K = 16;

struct ConvParams {
    int N, H, W;              // activation extent (pixels)
    int BW, BH, BN;           // patch shape of one M tile (BW*BH*BN == 128)
    int tiles_w, tiles_h;     // tiles per row / column of one image group
    int n_tiles;              // tiles along C_out
    int taps, ksize, pad;     // 9,3,1 or 1,1,0
    int kb_per_tap;           // K blocks (of 64) per tap
    int c_out;                // valid output channels
    const float* bias;        // [c_out] or null
    const __nv_bfloat16* res; // residual, NHWC bf16, or null
    int64_t res_ld;
    void* out;
    int64_t out_ld;           // NHWC pixel stride (mode 0)
    int out_mode;             // 0: bf16 NHWC, 1: fp32 NCHW
};

template <int BLOCK_N>
struct Smem {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(128) conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                         const __grid_constant__ CUtensorMap tmap_b,
                                                         const ConvParams p) {
    using S = Smem<BLOCK_N>;
    constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int CHUNK = BLOCK_N < 32 ? 16 : 32;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates: C_out tiles fastest so that neighbouring CTAs share the activation patch in L2
    const int n_tile = blockIdx.x % p.n_tiles;
    int m_tile = blockIdx.x / p.n_tiles;
    const int tw = m_tile % p.tiles_w;
    m_tile /= p.tiles_w;
    const int th = m_tile % p.tiles_h;
    const int tn = m_tile / p.tiles_h;
    const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1);
        }
        tc::mbar_init(tc::smem_u32(&bar_acc), 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        if (lane == 0) {
            tc::prefetch_tmap(&tmap_a);
            tc::prefetch_tmap(&tmap_b);
        }
        __syncwarp();
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_acc = tmem_slot;

    const int num_kb = p.taps * p.kb_per_tap;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t parity = ((kb / STAGES) & 1) ^ 1;
            tc::mbar_wait(tc::smem_u32(&bar_empty[s]), parity);
            const int tap = kb / p.kb_per_tap;
            const int cb = kb - tap * p.kb_per_tap;
            const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
            const uint32_t full = tc::smem_u32(&bar_full[s]);
            const uint32_t a_dst = smem_base + s * S::STAGE_BYTES;
            const uint32_t b_dst = a_dst + S::A_BYTES;
            tc::mbar_expect_tx(full, S::STAGE_BYTES);
            tc::tma_load_4d(a_dst, &tmap_a, full, cb * BLOCK_K, w0 + kw - p.pad, h0 + kh - p.pad, n0);
            tc::tma_load_2d(b_dst, &tmap_b, full, kb * BLOCK_K, n_tile * BLOCK_N);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = tc::idesc_bf16_f32(BLOCK_M, BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t parity = (kb / STAGES) & 1;
            tc::mbar_wait(tc::smem_u32(&bar_full[s]), parity);
            tc::fence_after_sync();
            const uint32_t a_src = smem_base + s * S::STAGE_BYTES;
            const uint64_t da = tc::smem_desc_sw128(a_src);
            const uint64_t db = tc::smem_desc_sw128(a_src + S::A_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >>4)
                tc::mma_f16_ss(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
            tc::mma_commit(tc::smem_u32(&bar_empty[s]));  // frees the smem slot when these MMAs retire
        }
        tc::mma_commit(tc::smem_u32(&bar_acc));  // accumulator complete
    }
    __syncwarp();

    // ===== epilogue: all four warps, warp w owns TMEM lanes [32w, 32w+32) =====
    tc::mbar_wait(tc::smem_u32(&bar_acc), 0);
    tc::fence_after_sync();

    const int row = warp * 32 + lane;  // row of the tile = TMEM lane
    const int bw = row % p.BW;
    const int bh = (row / p.BW) % p.BH;
    const int bn = row / (p.BW * p.BH);
    const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
    const bool row_ok = (n < p.N) && (h < p.H) && (w < p.W);
    const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
    const int col_base = n_tile * BLOCK_N;

#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += CHUNK) {
        uint32_t acc[CHUNK];
        __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the divergent stores
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        if constexpr (CHUNK == 32) {
            tc::tmem_ld_32x32b_x32(taddr, acc);
        } else {
            tc::tmem_ld_32x32b_x16(taddr, acc);
        }
        tc::tmem_ld_wait();
        const int col0 = col_base + c0;
        if (!row_ok || col0 >= p.c_out) continue;

        if (p.out_mode == 0) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + col0;
            const __nv_bfloat16* rsd = p.res ? p.res + pix * p.res_ld + col0 : nullptr;
#pragma unroll
            for (int v = 0; v < CHUNK / 8; ++v) {
                if (col0 + v * 8 >= p.c_out) break;
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(acc[v * 8 + j]);
                if (p.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + v * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + v * 8) + 1);
                    f[0] += b0.x, f[1] += b0.y, f[2] += b0.z, f[3] += b0.w;
                    f[4] += b1.x, f[5] += b1.y, f[6] += b1.z, f[7] += b1.w;
                }
                if (rsd) {
                    const uint4 r = __ldg(reinterpret_cast<const uint4*>(rsd + v * 8));
                    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f[2 * j] += bf16_bits_to_f32(rr[j] & 0xffffu);
                        f[2 * j + 1] += bf16_bits_to_f32(rr[j] >> 16);
                    }
                }
                uint4 o;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
                o.x = *reinterpret_cast<uint32_t*>(&t0), o.y = *reinterpret_cast<uint32_t*>(&t1);
                o.z = *reinterpret_cast<uint32_t*>(&t2), o.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(dst + v * 8) = o;
            }
        } else {
            // fp32 NCHW: out[n][c][h][w]; consecutive lanes are consecutive w => coalesced per channel
            float* dst = reinterpret_cast<float*>(p.out);
            const int64_t plane = (int64_t)p.H * p.W;
            const int64_t base = (int64_t)n * p.c_out * plane + (int64_t)h * p.W + w;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const int c = col0 + j;
                if (c < p.c_out) dst[base + c * plane] = __uint_as_float(acc[j]) + (p.bias ? __ldg(p.bias + c) : 0.0f);
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_acc, TMEM_COLS);
}

// ------------------------------------------------------------------ host: tensor maps + launch

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return AZB_E_DRIVER;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? AZB_OK : AZB_E_SHAPE;
}

template <int BLOCK_N, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const ConvParams& p, int64_t grid, cudaStream_t s) {
    constexpr int smem = STAGES * Smem<BLOCK_N>::STAGE_BYTES + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, STAGES>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    conv_gemm_kernel<BLOCK_N, STAGES><<<(unsigned)grid, 128, smem, s>>>(ta, tb, p);
    return azb_launch_status();
}

}  // namespace

extern "C" int azb_conv_gemm_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                                  const void* wpack, int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap,
                                  const float* bias, const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                                  int out_mode, void* stream) {
    AZB_CHECK_PTR(act);
    AZB_CHECK_PTR(wpack);
    AZB_CHECK_PTR(out);
    if (n <= 0 || h <= 0 || w <= 0 || c_in <= 0 || c_out <= 0) return AZB_E_SHAPE;
    if (taps != 1 && taps != 9) return AZB_E_SHAPE;
    if (k_per_tap % BLOCK_K || k_per_tap < c_in) return AZB_E_SHAPE;
    if (c_in % 8 || act_ld % 8 || act_ld < c_in) return AZB_E_ALIGN;
    if (!azb_aligned(act, 16) || !azb_aligned(wpack, 16) || !azb_aligned(out, 16)) return AZB_E_ALIGN;
    if (out_mode == 0 && (c_out % 8 || out_ld % 8 || (residual && (res_ld % 8 || !azb_aligned(residual, 16)))))
        return AZB_E_ALIGN;
    if (bias && !azb_aligned(bias, 16)) return AZB_E_ALIGN;
    if (out_mode != 0 && out_mode != 1) return AZB_E_SHAPE;

    int block_n = c_out_rows >= 128 ? 128 : c_out_rows >= 64 ? 64 : c_out_rows >= 32 ? 32 : 16;
    if (c_out_rows % block_n || c_out_rows < c_out) return AZB_E_SHAPE;

    ConvParams p{};
    p.N = (int)n, p.H = (int)h, p.W = (int)w;
    // patch: as wide as the image up to 16 columns, then rows, then images
    int bw = 1;
    while (bw < 16 && bw < w) bw <<= 1;
    int bh = 1;
    while (bw * bh < BLOCK_M && bh < h) bh <<= 1;
    int bn = BLOCK_M / (bw * bh);
    p.BW = bw, p.BH = bh, p.BN = bn;
    p.tiles_w = (int)((w + bw - 1) / bw);
    p.tiles_h = (int)((h + bh - 1) / bh);
    const int64_t tiles_n_img = (n + bn - 1) / bn;
    p.n_tiles = (int)(c_out_rows / block_n);
    p.taps = taps, p.ksize = taps == 9 ? 3 : 1, p.pad = taps == 9 ? 1 : 0;
    p.kb_per_tap = (int)(k_per_tap / BLOCK_K);
    p.c_out = (int)c_out;
    p.bias = bias;
    p.res = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.res_ld = res_ld;
    p.out = out, p.out_ld = out_ld, p.out_mode = out_mode;

    CUtensorMap ta, tb;
    {
        uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)act_ld * 2, (uint64_t)act_ld * 2 * w, (uint64_t)act_ld * 2 * w * h};
        uint32_t box[4] = {BLOCK_K, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
        int rc = make_map(&ta, act, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)(taps * k_per_tap), (uint64_t)c_out_rows};
        uint64_t str[1] = {(uint64_t)(taps * k_per_tap) * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)block_n};
        int rc = make_map(&tb, wpack, 2, dims, str, box);
        if (rc) return rc;
    }
    const int64_t grid = (int64_t)p.n_tiles * p.tiles_w * p.tiles_h * tiles_n_img;
    if (grid > 0x7fffffffLL) return AZB_E_SHAPE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (block_n) {
        case 128: return launch<128, 3>(ta, tb, p, grid, s);
        case 64: return launch<64, 4>(ta, tb, p, grid, s);
        case 32: return launch<32, 4>(ta, tb, p, grid, s);
        default: return launch<16, 4>(ta, tb, p, grid, s);
    }
}
