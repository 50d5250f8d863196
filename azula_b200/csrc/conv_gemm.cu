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
// Persistent, warp-specialised kernel, one CTA per SM, tiles handed out round-robin:
//   warp 0      TMA producer.  One 128-pixel x BLOCK_N output tile = (BN images x BH rows x BW columns)
//               of pixels, so the A operand of filter tap (kh, kw) is ONE 4-d TMA box load at spatial
//               offset (kh-1, kw-1); out-of-image coordinates are zero-filled by the TMA unit, which
//               implements the padding for free.  128-byte swizzle, STAGES-deep mbarrier ring.
//   warp 1      MMA issuer: a single elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into one
//               of TWO accumulator stages in tensor memory, so the epilogue of tile i overlaps the main
//               loop of tile i+1.  BLOCK_N = 256 keeps the shared-memory operand traffic of the SS-mode
//               MMA at 96 B/clk (128 B/clk is the SM limit that a 128x128 tile sits on).
//   warps 2..9  epilogue: tcgen05.ld (warp w reads TMEM lanes 32*(w%4)...), bias, residual, bf16 NHWC
//               store (or fp32 NCHW for the network output); optionally per-channel sums / sums of
//               squares of the stored values for the GroupNorm that consumes this tensor (see
//               azb_gn_finalize_f32), which removes a full read pass over the activation.
//
// CTA pairs (PAIR = true, large layers): two CTAs on the two SMs of a TPC form a cluster and share one
// 256-pixel x BLOCK_N tile through tcgen05.mma.cta_group::2.  Each CTA loads ITS 128 pixels of A and HALF of
// the weight tile (BLOCK_N / 2 rows); the leader's MMA thread drives both tensor cores, each of which reads the
// other half of B from the peer's shared memory.  Per CTA and k-block that is 32 KiB from L2 instead of 48
// (N = 256) -- the L2 -> SM path (~ 15 TB/s chip-wide at the single-CTA rate) is what caps the tensor pipe of
// the single-CTA kernel at ~ 75 % -- and 6 pipeline stages instead of 4 in the same shared memory.
//
// Halo tiles (HALO = true, 3 x 3 stride-1 layers on feature maps of at least 16 x 8 pixels with C_in % 64 == 0): an M tile
// is an 8-wide x 16-tall patch of ONE image and the A operand of a 64-channel block is ONE TMA box of (16 + 2) x (8 + 2)
// pixels.  The nine taps are shifted views of that tile (descriptor start + (kh * 10 + kw) * 128 bytes, 1280 bytes
// between 8-pixel row groups; the swizzle is a function of the absolute address, see tc::smem_desc_sw128_sbo), so the
// activation crosses the L2 -> SM path 1.4 times instead of 9 times.  Warp roles of a halo kernel (480 threads):
//   warp 0       stage ring: one weight tile per k-block; a k-block of the fused 1 x 1 operand (ResBlock skip
//                connection) takes two consecutive stages, plain A tile + weight tile, and these blocks are spread
//                between the halo items of a tile (ConvParams::item_mask)
//   warp 1       MMA issuer (as above; nine, or four, shifted views per halo item)
//   warps 2..9   epilogue (as above)
//   warps 10..13 input transform: rewrite each landed halo tile in place as  bf16(act(a[n, c] * x + b[n, c]))  -- the
//                GroupNorm (+ scale / shift + SiLU) that precedes the convolution in the reference
//                (_src/unet.py:177-181,203-207,238-243), coefficients from azb_gn_coef_f32 -- and reset out-of-image
//                pixels to zero (the reference pads the NORMALISED tensor).  No normalisation pass over HBM.
//   warp 14      halo-tile producer: keeps every free A slot loading, independently of the stage ring
// Upsampling inputs (the conv1 of an upsampling ResBlock, _src/unet.py:101-109,229-233) never exist in HBM either:
//   in_up = 1    the halo tile is loaded from the half-resolution tensor through a 5-d tensor map whose replication
//                dimensions have stride 0 (12 x 20 pixels of the virtual upsampled tensor, halo origin one row and one
//                column inside)
//   in_up = 2    phase decomposition: four 2 x 2 convolutions of the half-resolution tensor with pre-summed taps, each
//                (M tile, N tile, phase) a tile of this kernel (conv_impl)
// and the skip branch up(x) is read by the epilogue at pixel (h / 2, w / 2) (res_up).
//
// Short-reduction layers (the in-repo U-Net's 3 x 3 layers, K = 576 .. 2304 with 1 - 7 tiles per CTA; token GEMMs with
// K = 768): a tile lasts about a microsecond, nothing hides behind the main loop, and the kernel has these extras:
//   NORM         (template) the transform warps apply the per-pixel LayerNorm / RMSNorm + modulation that opens a UNetBlock
//                (in_norm): statistics from the per-pixel sums the PRODUCER's row-domain epilogue wrote (rowstat), requested a
//                tile ahead, shared among the eight lanes of a row group, applied with packed bf16 multiply-adds
//   64-column    halo tiles with the row-domain epilogue; the two warps of a TMEM lane quarter share a staging block (all
//   tiles        eight epilogue warps busy), two staging blocks per pair; one channel block + one N tile: the nine weight
//                tiles stay RESIDENT (requested before the programmatic-launch wait) and the 36 MMAs of a tile are issued as
//                a straight line with compile-time descriptor offsets
//   epilogue     bias / gate rows held one value per lane and broadcast with shuffles, the residual requested a chunk
//                ahead, all of it before the accumulator wait (tiles of <= 128 columns); out_up: every staged block stored
//                four times through the phase-scatter map (nn.Upsample after the block)
// Timelines behind these choices: scripts/conv_timeline.py (-DAZB_TIMELINE), scripts/graph_trace.py (azb_debug_trace).

#include "common.cuh"
#include "tc.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 128 bytes of bf16 = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int XF_WARPS = 4;                          // halo kernels: warps that transform the A tiles in place
constexpr int THREADS_HALO = THREADS + 32 * XF_WARPS + 32;  // + the A producer warp
constexpr int HALO_W = 8, HALO_H = 16;               // patch of one M tile (pixels)
constexpr int HALO_PITCH = HALO_W + 2;               // pixels per halo row
constexpr int HALO_ROWS = (HALO_H + 2) * HALO_PITCH; // 128-byte rows of one halo tile
constexpr int HALO_BYTES = HALO_ROWS * 128;
// Upsampling input (in_up): the tile is loaded from the HALF-resolution tensor through a tensor map whose x / y
// replication dimensions have stride 0 -- (2 x 6) x (2 x 10) = 12 x 20 pixels that cover the 10 x 18 halo region, whose
// origin sits one row and one column inside
constexpr int UP_PITCH = 12, UP_ROWS = 20 * UP_PITCH, UP_BYTES = UP_ROWS * 128, UP_ORIGIN = UP_PITCH + 1;
constexpr int SMEM_BUDGET = 192 * 1024;  // operand ring; + 32 KiB epilogue staging + alignment + ~1.5 KiB static <= 227 KiB

struct ConvParams {
    int N, H, W;              // activation extent (pixels)
    int BW, BH, BN;           // patch shape of one M tile (BW*BH*BN == 128)
    int tiles_w, tiles_h;     // tiles per row / column of one image group
    int n_tiles;              // tiles along C_out
    int total_tiles;
    int taps, ksize, pad;     // 9,3,1 or 1,1,0
    int kb_per_tap;           // K blocks (of 64) per tap
    int c_out;                // valid output channels
    const float* bias;        // [c_out] or null
    const __nv_bfloat16* res; // residual, NHWC bf16, or null
    int64_t res_ld;
    int res_up;               // 1: residual at (H / 2, W / 2), pixel (h, w) adds residual pixel (h / 2, w / 2)
    int out_up;               // 1 (EPI 2): the output is written through a nearest-neighbour 2x upsampling -- every staged block
                              // is stored four times, to (2 h + dy, 2 w + dx), through the 5-d map of the phase scatter
    void* out;
    int64_t out_ld;           // NHWC pixel stride (mode 0)
    int out_mode;             // 0: bf16 NHWC (TF32 mode: fp32 NHWC, or fp16 NHWC with out_f16), 1: fp32 NCHW
    int out_f16;              // TF32 mode, out_mode 0: store fp16 (the attention kernel's operands) instead of fp32
    float2* colsum;           // [m_tiles * 4][c_out / stat_gran] per-(32-row slab, channel block) {sum, sum of squares}
    int stat_gran;            // channels per colsum entry: 1 or 8
    int stride;               // 1 or 2: input pixel = stride * output pixel + tap - pad
    int act;                  // AZB_ACT_*: applied to acc + bias
    const float* gate;        // per-(sample, channel) multiplier of act(acc + bias), or null
    int64_t gate_ld;          // floats between the gate rows of consecutive samples (0 = one shared row)
    int gate_rows;            // output pixels per sample (sample = pixel index / gate_rows)
    int gate_uniform;         // every 32-row slab of an M tile lies in ONE sample (row-domain epilogue: gate row held per lane)
    int kb_extra;             // K blocks of the second (1x1, same resolution) operand appended after the taps
    unsigned long long* gn_acc;  // [N][c_out / stat_gran][4] exact fixed-point {sum hi, sum lo, sumsq hi, sumsq lo}
    int chunked;              // 1: CTA b owns the contiguous tile range [b * per, (b + 1) * per) instead of b, b + grid, ...
    int splits;               // split-K factor S (1 = off): tile t = split * tiles_out + output tile, single wave only
    int tiles_out;            // output tiles (= total_tiles / splits)
    float* ws_partial;        // [(S - 1) * tiles_out][128][BLOCK_N] fp32 partial accumulators of splits 1 .. S-1
    int* ws_flags;            // [tiles_out][EPI_WARPS] arrival counters, zero between launches
    int prefetch_kb;          // > 0: the producer pulls the weight tile of k-block kb + prefetch_kb into L2
    const float2* in_coef;    // halo kernels: [N][c_in] {a, b} of the input transform act(a x + b), or null (identity)
    int in_silu;              // the transform ends in SiLU (coefficients are halved, see azb_gn_coef_f32)
    int c_in;                 // row length of in_coef
    int in_up;                // halo kernels: the 3 x 3 operand is given at half resolution (nearest 2x upsampling on load)
    int phases;               // 4: phase-decomposed upsampling convolution (see conv_impl), tile / tiles_out = phase; else 1
    int a_slot;               // halo kernels: bytes per A slot
    int sa, sb;               // halo kernels: A slots and weight stages in the shared-memory budget
    int b_resident;           // halo kernels, one channel block, one N tile: the nine weight tiles are loaded ONCE per CTA and
                              // stay in stages 0 .. 8 (a 64 -> 64 layer streams 72 KiB of weights per 128-pixel tile otherwise:
                              // 8 KiB stages in flight against ~1 us of L2 latency paced its MMAs at a third of their rate)
    // halo kernels, per-pixel normalisation of the input (UNetBlock, azula/nn/unet.py:97-107): the convolution reads
    // bf16((1 + a[n][c]) norm_C(x)[pixel] + b[n][c]); statistics from the producer's epilogue (rowstat)
    int in_norm;                   // 0: none, 1: LayerNorm over the channels of a pixel, 2: RMSNorm
    float in_eps;
    const float2* in_rowstat;      // [pixels][c_in / 64] {sum, sum of squares} of each 64-channel block of the input
    const float* in_mod;           // [a(c_in) | b(c_in)] per sample (fp32), in_mod_ld floats between samples (0 = shared)
    int64_t in_mod_ld;
    float2* rowstat;               // EPI 2: [pixels][c_out / 64] {sum, sum of squares} of the stored values, or null
    unsigned long long* trace;     // azb_debug_trace buffer or null
    unsigned long long item_mask;  // halo kernels: bit i = item i of a tile is a halo item (else a 1 x 1 block): the blocks
                                   // of the fused 1 x 1 operand are spread between the halo items, so that every halo item
                                   // is preceded by >= 9 k-blocks of MMA time in which it can land and be transformed
};

// Adds v * 2^40 to a 96-bit fixed-point accumulator held as two int64 words (value = hi * 2^32 + lo, 0 <= lo < 2^32):
// integer additions commute, so the GroupNorm sums do not depend on the order in which tiles finish -- bit
// reproducible without a separate reduction pass.  Pure integer arithmetic on the bits of the fp32 partial sum (the
// double-precision pipe of this part is far too slow for an epilogue): |v| * 2^40 = mantissa << (exponent - 110);
// bits below 2^-40 are truncated, magnitudes above 2^30 saturate.
__device__ __forceinline__ void fixed_add(unsigned long long* acc, float v) {
    const uint32_t bits = __float_as_uint(v);
    const int e = (int)((bits >> 23) & 0xffu);
    if (e == 0) return;  // zero (or denormal: below 2^-126)
    const uint32_t man = (bits & 0x7fffffu) | 0x800000u;
    int sh = e - 110;
    unsigned long long hi, lo;
    if (sh >= 0) {
        sh = sh > 70 ? 70 : sh;
        const unsigned __int128 val = (unsigned __int128)man << sh;
        lo = (unsigned long long)(val & 0xffffffffu), hi = (unsigned long long)(val >> 32);
    } else {
        lo = sh > -24 ? (unsigned long long)(man >> (-sh)) : 0ull, hi = 0ull;
    }
    if (bits >> 31) {  // negative: -(hi * 2^32 + lo) = (-hi - borrow) * 2^32 + (2^32 - lo)
        hi = 0ull - hi - (lo != 0ull ? 1ull : 0ull);
        lo = (0x100000000ull - lo) & 0xffffffffull;
    }
    if (hi) atomicAdd(acc, hi);
    if (lo) atomicAdd(acc + 1, lo);
}

// The activation of N register values with the (runtime, CTA-uniform) selector tested ONCE, outside the element loop: with
// the switch inside, ptxas kept two compares and a branch per ELEMENT (ncu source page of the 768 -> 3072 token GEMM).
// SiLU: x sigmoid(x) = t + t tanh(t) with t = x / 2 -- ONE special-function instruction per element (tanh.approx, 2^-11
// relative) instead of an exponential and a reciprocal.
template <int N>
__device__ __forceinline__ void activate_all(float (&v)[N], int act) {
    if (act == AZB_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float t = 0.5f * v[j];
            v[j] = fmaf(t, tanh_approx(t), t);
        }
    } else if (act == AZB_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == AZB_ACT_RELU2) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float r = fmaxf(v[j], 0.f);
            v[j] = r * r;
        }
    }
}

template <int BLOCK_N, bool PAIR = false, bool HALO = false>
struct Cfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;  // weight rows this CTA stages per k-block
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
    // halo kernels: SA slots of one halo tile each (also used for the plain tiles of the fused 1x1 operand) and a
    // separate ring of weight tiles; otherwise one ring of {A tap tile, weight tile} stages
    static constexpr int A_SLOT = (HALO_BYTES + 1023) / 1024 * 1024;
    static constexpr int SA_MAX = 4;
    static constexpr int B_SLOT = B_BYTES < 1024 ? 1024 : B_BYTES;
    static constexpr int STAGE_BYTES = HALO ? B_SLOT : A_BYTES + B_BYTES;
    // halo kernels split the budget at run time (ConvParams::sa A slots, ::sb weight stages); STAGES is the most the
    // barrier arrays must hold
    // (halo kernels: up to 9, so that ALL nine weight tiles of a 64-channel layer can stay resident, ConvParams::b_resident)
    static constexpr int STAGES_CAP = HALO ? 9 : 8;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES > STAGES_CAP ? STAGES_CAP : SMEM_BUDGET / STAGE_BYTES;
    static constexpr int RING_BYTES = HALO ? SMEM_BUDGET : STAGES * STAGE_BYTES;
    static constexpr int weight_stages(int sa) {
        return (SMEM_BUDGET - sa * A_SLOT) / STAGE_BYTES > 8 ? 8 : (SMEM_BUDGET - sa * A_SLOT) / STAGE_BYTES;
    }
    static constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;  // columns of one accumulator stage
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
    static constexpr int COLS_PER_WARP = BLOCK_N >= 64 ? BLOCK_N / 2 : BLOCK_N;  // two warps share a lane quarter
    static constexpr int CHUNK = COLS_PER_WARP < 32 ? 16 : 32;
    static constexpr int STAGING_BYTES = EPI_WARPS * 32 * 128;  // per epilogue warp: 32 rows x 32 fp32, swizzled
    static constexpr int SMEM = RING_BYTES + STAGING_BYTES + 1024;
};

__device__ __forceinline__ void tile_coords(const ConvParams& p, int tile, int& n_tile, int& w0, int& h0, int& n0) {
    // C_out tiles fastest so that CTAs working at the same time share the activation patch in L2
    n_tile = tile % p.n_tiles;
    int m_tile = tile / p.n_tiles;
    const int tw = m_tile % p.tiles_w;
    m_tile /= p.tiles_w;
    const int th = m_tile % p.tiles_h;
    const int tn = m_tile / p.tiles_h;
    w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN;
}

// In-place input transform of one landed A tile (see the kernel's halo notes): thread (rg, chunk) owns the 16-byte chunk
// `chunk` (8 channels) of tile rows rg, rg + 16, ... (i & 7 == rg & 7: the chunk sits at the same swizzled position in
// each of them; a warp instruction touches 4 full 128-byte rows, conflict free).  Tile row i is pixel
// (h0 - OFF + i / PITCH, w0 - OFF + i % PITCH).  One warp per scheduler does this work, so instruction-level parallelism
// has to hide the shared-memory and SFU latencies: branch-free groups of four rows, all loads first.  Same arithmetic as
// gn_apply_kernel: f = fma(a, x, b), SiLU as f + f tanh(f) on halved coefficients, round to bf16.
template <int PITCH, int ROWS, int OFF>
__device__ __forceinline__ void transform_tile(uint32_t slot, int rg, int chunk, const float (&a)[8], const float (&b)[8],
                                               int silu, int h0, int w0, int H, int W) {
    const uint32_t base = slot + (uint32_t)rg * 128u + ((uint32_t)(chunk ^ (rg & 7)) << 4);
    constexpr int KS = (ROWS + 15) / 16;
#pragma unroll
    for (int k0 = 0; k0 < KS; k0 += 4) {
        uint32_t v[4][4];
        bool inside[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = rg + 16 * (k0 + u);
            const int y = i / PITCH, x = i - y * PITCH;
            // out-of-image pixels stay at (are reset to) zero: the convolution pads the NORMALISED tensor
            inside[u] = (unsigned)(h0 - OFF + y) < (unsigned)H && (unsigned)(w0 - OFF + x) < (unsigned)W;
            if (16 * (k0 + u) + 15 < ROWS || i < ROWS)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                             : "r"(base + (uint32_t)(k0 + u) * 2048u)
                             : "memory");
            else v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f0 = fmaf(a[2 * j], bf16_bits_to_f32(v[u][j] & 0xffffu), b[2 * j]);
                float f1 = fmaf(a[2 * j + 1], __uint_as_float(v[u][j] & 0xffff0000u), b[2 * j + 1]);
                if (silu) f0 = fmaf(f0, tanh_approx(f0), f0), f1 = fmaf(f1, tanh_approx(f1), f1);
                __nv_bfloat162 r = __floats2bfloat162_rn(f0, f1);
                v[u][j] = inside[u] ? *reinterpret_cast<uint32_t*>(&r) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = rg + 16 * (k0 + u);
            if (16 * (k0 + u) + 15 < ROWS || i < ROWS)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + (uint32_t)(k0 + u) * 2048u), "r"(v[u][0]),
                             "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3])
                             : "memory");
        }
    }
}

// The same pass for the per-pixel normalisation (ConvParams::in_norm): row k of this thread's rows is scaled with its own
// {rstd, -mean * rstd} (r2b[k], nm2b[k]: packed bf16 pairs; pixels outside the image stay zero), then modulated per channel:
// (1 + a) * norm(x) + b of UNetBlock._forward as two PACKED bf16 fused multiply-adds per channel pair,
//     t = fma(x, r, -m r),  y = fma(t, 1 + a[n][c], b[n][c])
// (HFMA2.BF16: one rounding each, so y is within ~1 bf16 ulp of the fp32 evaluation; the data never leaves its packed form --
// 1 instruction per element instead of 3.5, which is what the four transform warps of a microsecond-long tile can afford).
template <int PITCH, int ROWS, int OFF>
__device__ __forceinline__ void transform_tile_norm(uint32_t slot, int rg, int chunk, const float (&a)[8], const float (&b)[8],
                                                    const uint32_t (&r2b)[(ROWS + 15) / 16], const uint32_t (&nm2b)[(ROWS + 15) / 16],
                                                    int h0, int w0, int H, int W) {
    const uint32_t base = slot + (uint32_t)rg * 128u + ((uint32_t)(chunk ^ (rg & 7)) << 4);
    constexpr int KS = (ROWS + 15) / 16;
    __nv_bfloat162 a2[4], b2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a2[j] = __floats2bfloat162_rn(a[2 * j], a[2 * j + 1]), b2[j] = __floats2bfloat162_rn(b[2 * j], b[2 * j + 1]);
#pragma unroll
    for (int k0 = 0; k0 < KS; k0 += 4) {
        uint32_t v[4][4];
        bool inside[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = rg + 16 * (k0 + u);
            const int y = i / PITCH, x = i - y * PITCH;
            inside[u] = (unsigned)(h0 - OFF + y) < (unsigned)H && (unsigned)(w0 - OFF + x) < (unsigned)W;
            if (16 * (k0 + u) + 15 < ROWS || i < ROWS)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                             : "r"(base + (uint32_t)(k0 + u) * 2048u)
                             : "memory");
            else v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u < KS ? k0 + u : 0;
            const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&r2b[k]), nm2 = *reinterpret_cast<const __nv_bfloat162*>(&nm2b[k]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&v[u][j]);
                const __nv_bfloat162 y2 = __hfma2(__hfma2(x2, r2, nm2), a2[j], b2[j]);
                v[u][j] = inside[u] ? *reinterpret_cast<const uint32_t*>(&y2) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = rg + 16 * (k0 + u);
            if (16 * (k0 + u) + 15 < ROWS || i < ROWS)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + (uint32_t)(k0 + u) * 2048u), "r"(v[u][0]),
                             "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3])
                             : "memory");
        }
    }
}

// LEAN: the epilogue of the common case -- bf16 NHWC output, no activation, no gate, no split-K, GroupNorm sums (if
// any) as exact accumulators per 8-channel block -- with those switches resolved at compile time.
//
// EPI == 2: the ROW-DOMAIN epilogue with TMA stores.  tcgen05.ld delivers one accumulator row (= one output pixel) per
// lane; bias, activation, gate and residual are applied right there (the residual / gate rows of a lane are contiguous
// in memory: whole 32-byte sectors per lane), the 32 values are rounded to bf16 and written as four 16-byte slots into a
// 128-byte-swizzled staging block of 32 pixels x 64 channels (conflict free: a quarter warp covers eight distinct slots),
// and ONE elected lane hands the block to the TMA unit (cp.async.bulk.tensor store through a 4-d / 5-d tensor map of the
// output: partial tiles, channel tails and the (2 h + dy, 2 w + dx) scatter of a phase-decomposed upsampling convolution
// are the map's business).  ~60-200 instructions per 32-column chunk instead of ~455, no shared-memory transpose, no
// per-lane global stores.  The GroupNorm sums of the stored values fold 8 channels in the lane, then 32 rows with a
// transposed butterfly (12 shuffles per chunk).  Serves every bf16 NHWC layer without split-K / per-channel statistics.
//
// TF32 = true: the REFERENCE-NUMERICS mode.  Activations and weights are fp32 in HBM (what the reference's fp32 modules
// hold), a k-block is 32 fp32 channels (the same 128-byte swizzle rows, so tiles, ring and descriptors are unchanged),
// the tensor maps are CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 and the MMA is tcgen05.mma.kind::tf32 (K = 8 per instruction):
// operands with a 10-bit mantissa, fp32 accumulation -- what cuDNN does for the reference under PyTorch's default
// flags (torch.backends.cudnn.allow_tf32).  Tap-wise single-CTA kernels with EPI == 3 (row domain: one fp32 / fp16
// NHWC row segment per lane, direct 16-byte stores of whole 128-byte lines, fp32 residual) or the fp32 NCHW epilogue.
template <int BLOCK_N, bool PAIR, int EPI, bool HALO, bool TF32 = false, bool NORM = false>
__global__ void __launch_bounds__(HALO ? THREADS_HALO : THREADS, 1)
    conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_out,
                     const ConvParams p) {
    constexpr bool LEAN = EPI == 1;
    static_assert(!TF32 || !HALO, "the TF32 mode uses the tap-wise kernels");
    static_assert(!NORM || (HALO && EPI == 2), "the per-pixel normalisation lives in the halo kernels' input transform");
    static_assert((EPI == 3) == (TF32 && EPI != 0), "EPI 3 is the TF32 mode's NHWC epilogue");
    constexpr int KE = TF32 ? BLOCK_K / 2 : BLOCK_K;  // elements per k-block (128 bytes)
    using C = Cfg<BLOCK_N, PAIR, HALO>;
    constexpr int STAGES = C::STAGES;
    static_assert(!PAIR || BLOCK_N >= 128, "a CTA pair splits the weight tile in two halves of >= 64 rows");

    pdl_trigger();  // the next kernel of the stream may begin launching; it waits for this grid before touching memory
    const unsigned long long t_enter = p.trace ? azb_globaltimer() : 0ull;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[2];
    __shared__ __align__(8) uint64_t bar_acc_empty[2];
    // halo kernels: A slot landed (TMA) -> transformed (XF_WARPS warps, of both CTAs of a pair) -> consumed (MMA commit)
    __shared__ __align__(8) uint64_t bar_a_full[C::SA_MAX];
    __shared__ __align__(8) uint64_t bar_a_ready[C::SA_MAX];
    __shared__ __align__(8) uint64_t bar_a_empty[C::SA_MAX];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // CTA pair: rank 0 (the leader) issues the MMAs and owns the `full` and `accumulator drained` barriers
    const uint32_t cta_rank = PAIR ? (blockIdx.x & 1u) : 0u;  // cluster (2, 1, 1): rank in the pair = parity of the block index
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_acc_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&bar_acc_empty[s]), PAIR ? 2 * EPI_WARPS : EPI_WARPS);
        }
        if constexpr (HALO) {
            for (int s = 0; s < C::SA_MAX; ++s) {
                tc::mbar_init(tc::smem_u32(&bar_a_full[s]), 1);
                tc::mbar_init(tc::smem_u32(&bar_a_ready[s]), PAIR ? 2 * XF_WARPS : XF_WARPS);
                tc::mbar_init(tc::smem_u32(&bar_a_empty[s]), 1);
            }
        }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
        if (p.kb_extra) tc::prefetch_tmap(&tmap_a2);
        if constexpr (EPI == 2) tc::prefetch_tmap(&tmap_out);
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            tc::tmem_alloc_pair(tc::smem_u32(&tmem_slot), C::TMEM_COLS);
            tc::tmem_relinquish_pair();
        } else {
            tc::tmem_alloc(tc::smem_u32(&tmem_slot), C::TMEM_COLS);
            tc::tmem_relinquish();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if constexpr (PAIR) tc::cluster_sync();  // the peer's barriers are initialised before anything is sent to them
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    // Resident weights are constants of the network, not results of the previous kernel: their nine tiles are requested
    // BEFORE the programmatic-launch wait and land while the previous grid drains.
    if constexpr (HALO && !PAIR) {
        if (p.b_resident && warp == 0 && tc::elect_one()) {
            const uint32_t ring = smem_base + (uint32_t)(p.sa * p.a_slot);
            for (int t = 0; t < 9; ++t) {
                const uint32_t full = tc::smem_u32(&bar_full[t]);
                tc::mbar_expect_tx(full, C::B_BYTES);
                tc::tma_load_2d(ring + t * C::STAGE_BYTES, &tmap_b, full, t * p.kb_per_tap * BLOCK_K, 0);
            }
        }
    }
    // barrier setup, tensor-memory allocation and descriptor prefetch overlapped the tail of the previous kernel; its
    // results (activations, GroupNorm sums, coefficients) are read and this kernel's outputs written only from here on
    pdl_wait();
    // tile timeline of CTA 0 (diagnostics, scripts/conv_timeline.py; compiled in with AZB_NVCC_EXTRA=-DAZB_TIMELINE only):
    // words [1 << 16, ...) of the trace buffer, 16 stamps per (tile < 16)
    auto mark = [&](int local, int e) {
#ifdef AZB_TIMELINE
        if (p.trace && blockIdx.x == 0 && local < 16) p.trace[(1 << 16) + local * 16 + e] = azb_globaltimer();
#else
        (void)local, (void)e;
#endif
    };
    unsigned long long* trace_slot = nullptr;
    if (p.trace && threadIdx.x == 0) {
        trace_slot = p.trace + 8 + 4 * *reinterpret_cast<volatile unsigned long long*>(p.trace);
        atomicMin(trace_slot, t_enter);
        atomicMin(trace_slot + 1, azb_globaltimer());
    }

    const int num_kb_taps = p.taps * p.kb_per_tap;
    const int num_kb = num_kb_taps + p.kb_extra;
    const int kb_per_split = num_kb / p.splits;  // the host guarantees divisibility

    // this CTA's tiles: round robin, or (chunked) a contiguous range -- then consecutive tiles lie in the same image
    // and the GroupNorm sums can be carried across tiles instead of hitting the accumulators once per tile
    // A pair schedules like one CTA: unit u = (pair of M tiles 2 mp, 2 mp + 1) x N tile; this CTA works on M tile
    // 2 mp + rank.  p.total_tiles counts units.
    const int sched_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int sched_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    int tile_first, tile_step, tile_count;
    if (p.chunked) {
        const int per = (p.total_tiles + sched_n - 1) / sched_n;
        tile_first = sched_id * per, tile_step = 1;
        tile_count = max(0, min(per, p.total_tiles - tile_first));
    } else {
        tile_first = sched_id, tile_step = sched_n;
        tile_count = sched_id < p.total_tiles ? (p.total_tiles - sched_id + tile_step - 1) / tile_step : 0;
    }
    auto unit_to_tile = [&](int unit) -> int {
        if constexpr (PAIR) return (2 * (unit / p.n_tiles) + (int)cta_rank) * p.n_tiles + unit % p.n_tiles;
        else return unit;
    };

    // halo kernels: A items per tile = 64-channel blocks of the 3 x 3 operand (one halo tile, nine weight tiles each)
    // followed by the 64-channel blocks of the fused 1 x 1 operand (one plain 128-pixel tile, one weight tile each)
    const int items = p.kb_per_tap + p.kb_extra;
    const int halo_taps = HALO && p.phases > 1 ? 4 : 9;  // k-blocks per halo item
    const int SA = p.sa, SB = p.sb;
    const uint32_t b_ring = smem_base + (uint32_t)(SA * p.a_slot);

    if (HALO && warp == 0) {
        // ===== TMA producer (halo): the stage ring =====
        // One stage per k-block of a halo item (its weight tile).  A k-block of the fused 1 x 1 operand takes TWO
        // consecutive stages, the plain 128-pixel A tile and the weight tile -- the tap-wise kernel's scheme, so these
        // blocks need no A slot, no pass through the transform warps and no extra handshake, and are requested a full
        // ring ahead of the MMA like everything else in the ring.
        if (tc::elect_one()) {
            int sb = 0;
            uint32_t pb = 1;  // parity of the `empty` barriers: the first pass finds every slot free
            const uint32_t full_b0 = PAIR ? tc::mapa(tc::smem_u32(&bar_full[0]), 0) : tc::smem_u32(&bar_full[0]);
            // (resident weights were requested before the programmatic-launch wait, see above)
            for (int local = 0; local < (p.b_resident ? 0 : tile_count); ++local) {
                const int tile = unit_to_tile(tile_first + local * tile_step);
                int n_tile, w0, h0, n0;
                tile_coords(p, tile % p.tiles_out, n_tile, w0, h0, n0);
                const int b_row0 = n_tile * BLOCK_N + (int)cta_rank * C::B_ROWS;
                const int phase = tile / p.tiles_out;  // 0 unless phase-decomposed
                for (int it = 0, hi = 0, pi = 0; it < items; ++it) {
                    const bool halo = (p.item_mask >> it) & 1ull;
                    const int nb = halo ? halo_taps : 1;
                    if (!halo) {
                        tc::mbar_wait(tc::smem_u32(&bar_empty[sb]), pb);
                        const uint32_t a_dst = b_ring + sb * C::STAGE_BYTES;
                        const uint32_t full = full_b0 + 8u * (uint32_t)sb;
                        if constexpr (PAIR) {
                            if (leader) tc::mbar_expect_tx(tc::smem_u32(&bar_full[sb]), 2 * C::A_BYTES);
                            tc::tma_load_4d_pair(a_dst, &tmap_a2, full, pi * BLOCK_K, w0, h0, n0);
                        } else {
                            tc::mbar_expect_tx(full, C::A_BYTES);
                            tc::tma_load_4d(a_dst, &tmap_a2, full, pi * BLOCK_K, w0, h0, n0);
                        }
                        if (++sb == SB) sb = 0, pb ^= 1u;
                    }
                    // weights are packed [tap][channel block]: tap t of channel block hi is k-block t * kb_per_tap + hi
                    // (phase-decomposed: [phase][2 x 2 taps][channel block])
                    int kbx = halo ? phase * halo_taps * p.kb_per_tap + hi++ : num_kb_taps + pi++;
                    for (int t = 0; t < nb; ++t, kbx += p.kb_per_tap) {
                        tc::mbar_wait(tc::smem_u32(&bar_empty[sb]), pb);
                        const uint32_t b_dst = b_ring + sb * C::STAGE_BYTES;
                        const uint32_t full = full_b0 + 8u * (uint32_t)sb;
                        if constexpr (PAIR) {
                            if (leader) tc::mbar_expect_tx(tc::smem_u32(&bar_full[sb]), 2 * C::B_BYTES);
                            tc::tma_load_2d_pair(b_dst, &tmap_b, full, kbx * BLOCK_K, b_row0);
                        } else {
                            tc::mbar_expect_tx(full, C::B_BYTES);
                            tc::tma_load_2d(b_dst, &tmap_b, full, kbx * BLOCK_K, b_row0);
                        }
                        if (++sb == SB) sb = 0, pb ^= 1u;
                    }
                }
            }
        }
    } else if (HALO && warp == 2 + EPI_WARPS + XF_WARPS) {
        // ===== TMA producer (halo): the halo tiles, a lane of its own =====
        // Every free A slot is put to work at once, independently of the stage ring: a halo tile needs TMA latency plus
        // the transform pass before the MMA reaches it.
        if (tc::elect_one()) {
            int sa = 0;
            uint32_t pa = 1;
            for (int local = 0; local < tile_count; ++local) {
                const int tile = unit_to_tile(tile_first + local * tile_step);
                int n_tile, w0, h0, n0;
                tile_coords(p, tile % p.tiles_out, n_tile, w0, h0, n0);
                for (int hi = 0; hi < p.kb_per_tap; ++hi) {
                    tc::mbar_wait(tc::smem_u32(&bar_a_empty[sa]), pa);
                    const uint32_t full = tc::smem_u32(&bar_a_full[sa]);
                    const uint32_t dst = smem_base + (uint32_t)(sa * p.a_slot);
                    if (p.in_up) {  // (channels, x replica, x / 2, y replica, image rows / 2): see UP_PITCH
                        tc::mbar_expect_tx(full, UP_BYTES);
                        tc::tma_load_5d(dst, &tmap_a, full, hi * BLOCK_K, 0, (w0 >> 1) - 1, 0, n0 * (p.H >> 1) + (h0 >> 1) - 1);
                    } else {
                        tc::mbar_expect_tx(full, HALO_BYTES);
                        tc::tma_load_4d(dst, &tmap_a, full, hi * BLOCK_K, w0 - 1, h0 - 1, n0);
                    }
                    if (++sa == SA) sa = 0, pa ^= 1u;
                }
            }
        }
    } else if (HALO && warp == 1) {
        // ===== MMA issuer (halo) =====
        if (leader && tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_bf16_f32(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
            int sb = 0, sa = 0;
            uint32_t pb = 0, pa = 0;
            auto mma4 = [&](uint64_t da, uint64_t db, uint32_t tmem_acc, bool first) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    // +32 bytes per K = 16 step inside the 128-byte swizzle row (address field is >> 4)
                    if constexpr (PAIR)
                        tc::mma_f16_ss_pair(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, !first || k != 0);
                    else
                        tc::mma_f16_ss(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, !first || k != 0);
                }
            };
            auto release = [&](uint64_t* bar) {  // the barrier (of both CTAs of a pair) arrives when the MMAs so far retire
                if constexpr (PAIR) tc::mma_commit_pair(tc::smem_u32(bar), 0b11);
                else tc::mma_commit(tc::smem_u32(bar));
            };
            for (int local = 0; local < tile_count; ++local) {
                const int as = local & 1;
                mark(local, 0);
                tc::mbar_wait(tc::smem_u32(&bar_acc_empty[as]), ((local >> 1) & 1) ^ 1);
                mark(local, 1);
                tc::fence_after_sync();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * C::ACC_COLS);
                const int phase = HALO && p.phases > 1 ? unit_to_tile(tile_first + local * tile_step) / p.tiles_out : 0;
                for (int it = 0; it < items; ++it) {
                    if ((p.item_mask >> it) & 1ull) {
                        // the halo tile has landed AND been transformed (by the transform warps of both CTAs of a pair;
                        // they arrive with release.cluster after fence.proxy.async, and what they wrote is read by each
                        // CTA's own tensor core, never by this thread: the CTA-scope acquire of try_wait suffices)
                        tc::mbar_wait(tc::smem_u32(&bar_a_ready[sa]), pa);
                        if (it == 0) mark(local, 2);
                        tc::fence_after_sync();
                        // tap (0, 0): the view that starts at halo pixel (0, 0)
                        uint32_t a_src = smem_base + (uint32_t)(sa * p.a_slot) + (p.in_up ? UP_ORIGIN * 128u : 0u);
                        const uint32_t pitch = p.in_up ? UP_PITCH : HALO_PITCH;
                        // taps: the 3 x 3 window, or (phase (dy, dx) of an upsampling convolution) the 2 x 2 window of
                        // half-resolution pixels at rows dy, dy + 1 and columns dx, dx + 1 of it
                        const int kw_end = halo_taps == 9 ? 3 : 2;
                        if (halo_taps != 9) a_src += ((uint32_t)(phase >> 1) * pitch + (uint32_t)(phase & 1)) * 128u;
                        if (p.b_resident) {
                            // Resident weights (one channel block, nine taps, stage = tap, loaded once): a straight line of 36
                            // MMAs whose descriptors differ from two base descriptors by COMPILE-TIME offsets -- the generic
                            // loop below spends ~70 clk of this one thread per MMA on waits, descriptor arithmetic and commits,
                            // against 48 clk of tensor time for a 64-column MMA.
                            if (local == 0) {
                                for (int t = 0; t < 9; ++t) tc::mbar_wait(tc::smem_u32(&bar_full[t]), 0u);
                                tc::fence_after_sync();
                            }
                            const uint64_t da0 = tc::smem_desc_sw128_sbo(a_src, HALO_PITCH * 128u);
                            const uint64_t db0 = tc::smem_desc_sw128(b_ring);
#pragma unroll
                            for (int t = 0; t < 9; ++t) {
                                const uint32_t a_off = (uint32_t)(((t / 3) * HALO_PITCH + (t % 3)) * 128);
                                const uint32_t b_off = (uint32_t)(t * C::STAGE_BYTES);
#pragma unroll
                                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                    tc::mma_f16_ss(tmem_acc, da0 + (uint64_t)((a_off >> 4) + 2 * k), db0 + (uint64_t)((b_off >> 4) + 2 * k),
                                                   idesc, (t | k) != 0);
                            }
                            release(&bar_a_empty[sa]);
                            if (++sa == SA) sa = 0, pa ^= 1u;
                            continue;
                        }
                        for (int t = 0, kw = 0; t < halo_taps; ++t) {
                            const int st = p.b_resident ? t : sb;
                            if (!p.b_resident) {
                                tc::mbar_wait(tc::smem_u32(&bar_full[st]), pb);
                                tc::fence_after_sync();
                            }
                            mma4(tc::smem_desc_sw128_sbo(a_src, pitch * 128u), tc::smem_desc_sw128(b_ring + st * C::STAGE_BYTES),
                                 tmem_acc, (it | t) == 0);
                            if (!p.b_resident) {
                                release(&bar_empty[sb]);
                                if (++sb == SB) sb = 0, pb ^= 1u;
                            }
                            // next tap: one pixel to the right, or back to column 0 of the next halo row
                            if (++kw == kw_end) kw = 0, a_src += (pitch - (uint32_t)kw_end + 1u) * 128u;
                            else a_src += 128u;
                        }
                        release(&bar_a_empty[sa]);
                        if (++sa == SA) sa = 0, pa ^= 1u;
                    } else {
                        // a k-block of the fused 1 x 1 operand: plain A tile in this stage, its weights in the next
                        const int s_a = sb;
                        tc::mbar_wait(tc::smem_u32(&bar_full[sb]), pb);
                        if (++sb == SB) sb = 0, pb ^= 1u;
                        tc::mbar_wait(tc::smem_u32(&bar_full[sb]), pb);
                        tc::fence_after_sync();
                        mma4(tc::smem_desc_sw128(b_ring + s_a * C::STAGE_BYTES), tc::smem_desc_sw128(b_ring + sb * C::STAGE_BYTES),
                             tmem_acc, it == 0);
                        release(&bar_empty[s_a]);
                        release(&bar_empty[sb]);
                        if (++sb == SB) sb = 0, pb ^= 1u;
                    }
                }
                release(&bar_acc_full[as]);
                mark(local, 3);
            }
        }
    } else if (HALO && warp >= 2 + EPI_WARPS && warp < 2 + EPI_WARPS + XF_WARPS) {
        // ===== input transform (halo) =====
        // Thread (rg, chunk) owns the 16-byte chunk `chunk` (8 channels) of halo rows rg, rg + 16, ...: a warp
        // instruction touches 4 full 128-byte rows (conflict free for any swizzle phase).  Same arithmetic as
        // gn_apply_kernel: f = fma(a, x, b), SiLU as f + f tanh(f) on halved coefficients, round to bf16.
        const int tx = (int)threadIdx.x - 32 * (2 + EPI_WARPS);
        const int chunk = tx & 7, rg = tx >> 3;
        int sa = 0;
        uint32_t pa = 0;
        const uint32_t ready0 = PAIR ? tc::mapa(tc::smem_u32(&bar_a_ready[0]), 0) : tc::smem_u32(&bar_a_ready[0]);
        const bool xf = p.in_coef != nullptr;
        constexpr int KS = (HALO_ROWS + 15) / 16;
        // Per-pixel normalisation: the producer's per-block sums, requested ONE TILE AHEAD (all loads of a tile are independent
        // and in flight together: a single L2 round trip, hidden behind the previous tile).  The eight lanes that share the
        // rows rg, rg + 16, ... (one per 16-byte chunk) split them: lane `chunk` owns rows k = chunk and chunk + 8, turns their
        // sums into packed {rstd, -mean rstd} and the group exchanges the twelve results with shuffles.
        constexpr int KO = (KS + 7) / 8;  // rows per lane
        float ps1[NORM ? KO : 1], ps2[NORM ? KO : 1];
        auto request_stats = [&](int local) {
#pragma unroll
            for (int o = 0; o < (NORM ? KO : 1); ++o) ps1[o] = ps2[o] = 0.f;
            if (!NORM || local >= tile_count) return;
            const int tile = unit_to_tile(tile_first + local * tile_step);
            int n_tile, w0, h0, n0;
            tile_coords(p, tile % p.tiles_out, n_tile, w0, h0, n0);
            const int nblk = p.c_in >> 6;
            const float2* sp[KO];
#pragma unroll
            for (int o = 0; o < KO; ++o) {
                const int k = chunk + 8 * o;
                const int i = rg + 16 * k;
                const int y = i / HALO_PITCH, x = i - y * HALO_PITCH;
                const int hh = h0 - 1 + y, ww = w0 - 1 + x;
                const bool in = k < KS && i < HALO_ROWS && (unsigned)hh < (unsigned)p.H && (unsigned)ww < (unsigned)p.W;
                sp[o] = in ? p.in_rowstat + (((int64_t)n0 * p.H + hh) * p.W + ww) * nblk : nullptr;
            }
            for (int bk = 0; bk < nblk; ++bk) {
                float2 v[KO];
#pragma unroll
                for (int o = 0; o < KO; ++o) v[o] = sp[o] ? __ldg(sp[o] + bk) : make_float2(0.f, 0.f);
#pragma unroll
                for (int o = 0; o < KO; ++o) ps1[o] += v[o].x, ps2[o] += v[o].y;
            }
        };
        if constexpr (NORM) request_stats(0);
        for (int local = 0; local < tile_count; ++local) {
            const int tile = unit_to_tile(tile_first + local * tile_step);
            int n_tile, w0, h0, n0;
            tile_coords(p, tile % p.tiles_out, n_tile, w0, h0, n0);
            uint32_t nr[NORM ? KS : 1], nmr[NORM ? KS : 1];
            if constexpr (NORM) {
                const float inv_c = 1.0f / (float)p.c_in, inv_cm1 = 1.0f / (float)(p.c_in - 1);
                uint32_t own_r[KO], own_m[KO];
#pragma unroll
                for (int o = 0; o < KO; ++o) {
                    // LayerNorm: torch.var_mean's UNBIASED variance (azula/nn/layers.py:152-155); RMSNorm: the mean square
                    const float s1 = ps1[o], s2 = ps2[o];
                    const float mean = p.in_norm == 1 ? s1 * inv_c : 0.f;
                    const float var = p.in_norm == 1 ? fmaxf(fmaf(-mean, s1, s2), 0.f) * inv_cm1 : s2 * inv_c;
                    const float r = rsqrtf(var + p.in_eps);
                    const __nv_bfloat162 r2 = __float2bfloat162_rn(r), m2 = __float2bfloat162_rn(-mean * r);
                    own_r[o] = *reinterpret_cast<const uint32_t*>(&r2), own_m[o] = *reinterpret_cast<const uint32_t*>(&m2);
                }
#pragma unroll
                for (int k = 0; k < KS; ++k) {  // row k lives in lane (k & 7) of this thread's group of eight
                    nr[k] = __shfl_sync(0xffffffffu, own_r[k >> 3], (lane & 24) | (k & 7));
                    nmr[k] = __shfl_sync(0xffffffffu, own_m[k >> 3], (lane & 24) | (k & 7));
                }
                request_stats(local + 1);
            }
            for (int hi = 0; hi < p.kb_per_tap; ++hi) {
                float a[8], b[8];
                if (NORM) {
                    const float* mp = p.in_mod + (int64_t)n0 * p.in_mod_ld + hi * BLOCK_K + chunk * 8;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float4 va = __ldg(reinterpret_cast<const float4*>(mp) + j);
                        const float4 vb = __ldg(reinterpret_cast<const float4*>(mp + p.c_in) + j);
                        a[4 * j] = 1.f + va.x, a[4 * j + 1] = 1.f + va.y, a[4 * j + 2] = 1.f + va.z, a[4 * j + 3] = 1.f + va.w;
                        b[4 * j] = vb.x, b[4 * j + 1] = vb.y, b[4 * j + 2] = vb.z, b[4 * j + 3] = vb.w;
                    }
                } else if (xf) {
                    const float4* cp = reinterpret_cast<const float4*>(p.in_coef + (int64_t)n0 * p.c_in + hi * BLOCK_K + chunk * 8);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 v = __ldg(cp + j);
                        a[2 * j] = v.x, b[2 * j] = v.y, a[2 * j + 1] = v.z, b[2 * j + 1] = v.w;
                    }
                }
                tc::mbar_wait(tc::smem_u32(&bar_a_full[sa]), pa);
                if constexpr (NORM) {
                    const uint32_t slot = smem_base + (uint32_t)(sa * p.a_slot);
                    transform_tile_norm<HALO_PITCH, HALO_ROWS, 1>(slot, rg, chunk, a, b, nr, nmr, h0, w0, p.H, p.W);
                    tc::fence_proxy_async();
                } else if (xf) {
                    const uint32_t slot = smem_base + (uint32_t)(sa * p.a_slot);
                    if (p.in_up) transform_tile<UP_PITCH, UP_ROWS, 2>(slot, rg, chunk, a, b, p.in_silu, h0, w0, p.H, p.W);
                    else transform_tile<HALO_PITCH, HALO_ROWS, 1>(slot, rg, chunk, a, b, p.in_silu, h0, w0, p.H, p.W);
                    tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) reads
                }
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && !leader) tc::mbar_arrive_cluster(ready0 + 8u * (uint32_t)sa);
                    else tc::mbar_arrive(tc::smem_u32(&bar_a_ready[sa]));
                }
                if (++sa == SA) sa = 0, pa ^= 1u;
            }
        }
    } else if (warp == 0) {
        // ===== TMA producer =====
        // One ELECTED lane (elect.sync tells ptxas that exactly one lane is active, so descriptor and coordinate
        // operands go straight to uniform registers instead of through per-value waterfall loops).  The loop is the
        // pacemaker of the whole kernel -- one k-block must be issued every 512 cycles -- so it carries its ring slot,
        // parity and (tap, channel block) position incrementally: no division, no modulo.
        if (tc::elect_one()) {
            int s = 0;
            uint32_t parity = 1;  // of the `empty` barriers: the first pass over the ring finds every slot free
            const uint32_t full0 = PAIR ? tc::mapa(tc::smem_u32(&bar_full[0]), 0) : tc::smem_u32(&bar_full[0]);
            for (int local = 0; local < tile_count; ++local) {
                const int tile = unit_to_tile(tile_first + local * tile_step);
                int n_tile, w0, h0, n0;
                tile_coords(p, tile % p.tiles_out, n_tile, w0, h0, n0);
                const int kb0 = (tile / p.tiles_out) * kb_per_split;
                const int b_row0 = n_tile * BLOCK_N + (int)cta_rank * C::B_ROWS;
                for (int j = 0; j < p.prefetch_kb && j < kb_per_split; ++j)
                    tc::tma_prefetch_l2_2d(&tmap_b, (kb0 + j) * KE, n_tile * BLOCK_N);
                // position of k-block kb0 inside the taps: (kh, kw, channel block); one division per tile
                int tap = kb0 / p.kb_per_tap, cb = kb0 - tap * p.kb_per_tap;
                int kh = tap / p.ksize, kw = tap - kh * p.ksize;
                const int wbase = w0 * p.stride - p.pad, hbase = h0 * p.stride - p.pad;
                for (int kb = kb0; kb < kb0 + kb_per_split; ++kb) {
                    tc::mbar_wait(tc::smem_u32(&bar_empty[s]), parity);
                    const uint32_t a_dst = smem_base + s * C::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + C::A_BYTES;
                    const uint32_t full = full0 + 8u * (uint32_t)s;
                    const bool in_taps = kb < num_kb_taps;
                    // the fused 1x1 operand (ResBlock skip connection) follows the taps: same pixels, no offset
                    const CUtensorMap* ma = in_taps ? &tmap_a : &tmap_a2;
                    const int c0 = (in_taps ? cb : kb - num_kb_taps) * KE;
                    const int cw = in_taps ? wbase + kw : w0, ch = in_taps ? hbase + kh : h0;
                    if constexpr (PAIR) {
                        // both CTAs' bytes are credited to the LEADER's barrier, which its producer arms for the pair
                        if (leader) tc::mbar_expect_tx(tc::smem_u32(&bar_full[s]), 2 * C::STAGE_BYTES);
                        tc::tma_load_4d_pair(a_dst, ma, full, c0, cw, ch, n0);
                        tc::tma_load_2d_pair(b_dst, &tmap_b, full, kb * KE, b_row0);
                    } else {
                        tc::mbar_expect_tx(full, C::STAGE_BYTES);
                        tc::tma_load_4d(a_dst, ma, full, c0, cw, ch, n0);
                        tc::tma_load_2d(b_dst, &tmap_b, full, kb * KE, b_row0);
                    }
                    if (p.prefetch_kb && kb + p.prefetch_kb < kb0 + kb_per_split)
                        tc::tma_prefetch_l2_2d(&tmap_b, (kb + p.prefetch_kb) * KE, n_tile * BLOCK_N);
                    if (++cb == p.kb_per_tap) {
                        cb = 0;
                        if (++kw == p.ksize) kw = 0, ++kh;
                    }
                    if (++s == STAGES) s = 0, parity ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (leader && tc::elect_one()) {
            constexpr uint32_t idesc = TF32 ? tc::idesc_tf32_f32(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N)
                                            : tc::idesc_bf16_f32(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
            int s = 0;
            uint32_t parity = 0;
            for (int local = 0; local < tile_count; ++local) {
                const int as = local & 1;
                // wait until the epilogue has drained this accumulator stage (first use passes immediately)
                tc::mbar_wait(tc::smem_u32(&bar_acc_empty[as]), ((local >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * C::ACC_COLS);
                for (int kb = 0; kb < kb_per_split; ++kb) {
                    tc::mbar_wait(tc::smem_u32(&bar_full[s]), parity);
                    tc::fence_after_sync();
                    const uint32_t a_src = smem_base + s * C::STAGE_BYTES;
                    const uint64_t da = tc::smem_desc_sw128(a_src);
                    const uint64_t db = tc::smem_desc_sw128(a_src + C::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >>4)
                        if constexpr (PAIR && TF32)
                            tc::mma_tf32_ss_pair(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        else if constexpr (PAIR)
                            tc::mma_f16_ss_pair(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        else if constexpr (TF32)  // K = 8 fp32 words = the same 32 bytes per operand row
                            tc::mma_tf32_ss(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        else
                            tc::mma_f16_ss(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    // frees the smem slot (of both CTAs) when these MMAs retire
                    if constexpr (PAIR) tc::mma_commit_pair(tc::smem_u32(&bar_empty[s]), 0b11);
                    else tc::mma_commit(tc::smem_u32(&bar_empty[s]));
                    if (++s == STAGES) s = 0, parity ^= 1u;
                }
                // accumulator complete (in both CTAs' tensor memory)
                if constexpr (PAIR) tc::mma_commit_pair(tc::smem_u32(&bar_acc_full[as]), 0b11);
                else tc::mma_commit(tc::smem_u32(&bar_acc_full[as]));
            }
        }
    } else if constexpr (EPI == 3) {
        // ===== epilogue of the TF32 mode: row domain, fp32 (or fp16) NHWC output, direct stores =====
        // lane = output pixel; a 32-column chunk of its row is 128 contiguous bytes of fp32 (one full line per lane, eight
        // 16-byte stores) or 64 bytes of fp16 (the qkv projection feeding the attention kernel).  bias -> activation ->
        // fp32 residual, no rounding for fp32 outputs.
        const int e = warp - 2;
        const int quarter = warp & 3;
        constexpr int CPW = C::COLS_PER_WARP;
        constexpr int CHUNK = C::CHUNK;
        const int half = (BLOCK_N >= 64) ? (e >> 2) : 0;
        const bool active = (BLOCK_N >= 64) || (e < 4);
        for (int local = 0; local < tile_count; ++local) {
            const int tile = unit_to_tile(tile_first + local * tile_step);
            const int as = local & 1;
            int n_tile, w0, h0, n0;
            tile_coords(p, tile % p.tiles_out, n_tile, w0, h0, n0);
            const int col_base = n_tile * BLOCK_N + half * CPW;
            tc::mbar_wait_backoff(tc::smem_u32(&bar_acc_full[as]), (local >> 1) & 1);
            tc::fence_after_sync();
            if (active) {
                const int row = quarter * 32 + lane;
                const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
                const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
                const bool ok = (n < p.N) && (h < p.H) && (w < p.W);
                const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
                const float* resp = p.res ? reinterpret_cast<const float*>(p.res) + pix * p.res_ld + col_base : nullptr;
#pragma unroll 1
                for (int c0 = 0; c0 < CPW; c0 += CHUNK) {
                    uint32_t acc[CHUNK];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * C::ACC_COLS + half * CPW + c0);
                    if constexpr (CHUNK == 32) tc::tmem_ld_32x32b_x32(taddr, acc);
                    else tc::tmem_ld_32x32b_x16(taddr, acc);
                    const int col = col_base + c0;
                    float4 rsd[CHUNK / 4];
                    const bool use_res = resp != nullptr && ok;
                    if (use_res) {
#pragma unroll
                        for (int q = 0; q < CHUNK / 4; ++q)
                            rsd[q] = col + 4 * q < p.c_out ? __ldg(reinterpret_cast<const float4*>(resp + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    tc::tmem_ld_wait();
                    float v[CHUNK];
#pragma unroll
                    for (int q = 0; q < CHUNK / 4; ++q) {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias && col + 4 * q < p.c_out) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + q);
                        v[4 * q] = __uint_as_float(acc[4 * q]) + b4.x, v[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + b4.y;
                        v[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + b4.z, v[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + b4.w;
                    }
                    activate_all(v, p.act);
                    if (use_res) {
#pragma unroll
                        for (int q = 0; q < CHUNK / 4; ++q)
                            v[4 * q] += rsd[q].x, v[4 * q + 1] += rsd[q].y, v[4 * q + 2] += rsd[q].z, v[4 * q + 3] += rsd[q].w;
                    }
                    if (ok) {
                        if (p.out_f16) {
                            __half* dst = reinterpret_cast<__half*>(p.out) + pix * p.out_ld + col;
#pragma unroll
                            for (int q = 0; q < CHUNK / 8; ++q) {
                                if (col + 8 * q < p.c_out) {
                                    uint32_t w4[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        __half2 t = __floats2half2_rn(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
                                        w4[j] = *reinterpret_cast<uint32_t*>(&t);
                                    }
                                    *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                                }
                            }
                        } else {
                            float* dst = reinterpret_cast<float*>(p.out) + pix * p.out_ld + col;
#pragma unroll
                            for (int q = 0; q < CHUNK / 4; ++q)
                                if (col + 4 * q < p.c_out)
                                    *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && !leader) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bar_acc_empty[as]), 0));
                else tc::mbar_arrive(tc::smem_u32(&bar_acc_empty[as]));
            }
        }
    } else if constexpr (EPI == 2) {
        // ===== epilogue, row domain + TMA store (see the kernel's header) =====
        const int e = warp - 2;
        const int quarter = warp & 3;
        // columns per warp: whole 64-column store blocks -- or, for 64-column tiles, HALF a block: the two warps of a TMEM lane
        // quarter (e and e + 4) fill one staging block together and meet at a 64-thread named barrier, so that all eight
        // epilogue warps work on the short tiles of the narrow layers (with four, the epilogue of a 64 -> 64 layer of the
        // in-repo U-Net took 1.7 - 2.3 us per tile against 1.3 us of MMA issue)
        constexpr bool SPLIT64 = BLOCK_N == 64;
        constexpr int CPW = BLOCK_N >= 128 ? BLOCK_N / 2 : SPLIT64 ? 32 : BLOCK_N;
        static_assert(CPW % 64 == 0 || SPLIT64, "row-domain epilogue: BLOCK_N >= 64");
        const int half = e >> 2;
        const bool active = true;
        __shared__ float2 pair_sums[4][32];  // SPLIT64 + rowstat: the upper half's per-pixel sums on their way to the lower half
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };
        // 32 pixels x 128 bytes per warp, 1024-byte aligned.  With 64-column tiles only four warps are active: each uses
        // the idle partner's block as a SECOND staging buffer, so that a tile never waits for the TMA unit to finish reading
        // the previous one (measured on the 64-channel U-Net layers: that wait was most of a 1.4 us epilogue)
        constexpr bool TWO_STAGING = BLOCK_N == 64;
        const uint32_t stage0 = smem_base + C::RING_BYTES + (SPLIT64 ? quarter : e) * (32 * 128);
        const int sw = lane & 7;
        // this lane's pixel inside the tile (fixed for the whole kernel: the divisions are done once)
        const int row = quarter * 32 + lane;
        const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
        const int r0 = quarter * 32;  // first pixel of the 32 this warp stores: tile rows [32 quarter, 32 quarter + 32)
        const int r0h = (r0 / p.BW) % p.BH, r0n = r0 / (p.BW * p.BH);
        // tiles of a microsecond: no sleeping between polls of the accumulator barrier (only there: eight spinning warps cost
        // the power-capped 36-k-block layers of ADM 1.3 % of the sustained step)
        const bool short_k = num_kb < 24;
        // the lanes with (lane & 7) == 0 own 8-channel block (lane >> 3) of every chunk: GroupNorm sums carried across
        // the tiles of an image (chunked mode), slot = n_tile * 4 + chunk
        float carry_s[8], carry_q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) carry_s[i] = carry_q[i] = 0.f;
        int carry_img = -1;
        const bool owner = (lane & 7) == 0;
        const int own_blk = lane >> 3;
        auto flush_carry = [&](int img) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int col = (i >> 2) * BLOCK_N + half * CPW + (i & 3) * 32 + own_blk * 8;
                if (owner && img >= 0 && img < p.N && (i & 3) < CPW / 32 && (i >> 2) < p.n_tiles && col < p.c_out) {
                    unsigned long long* dst = p.gn_acc + ((int64_t)img * (p.c_out >> 3) + (col >> 3)) * 4;
                    fixed_add(dst, carry_s[i]);
                    fixed_add(dst + 2, carry_q[i]);
                }
                carry_s[i] = carry_q[i] = 0.f;
            }
        };
        bool store_pending = false;
        for (int local = 0; local < tile_count; ++local) {
            const int tile = unit_to_tile(tile_first + local * tile_step);
            const int as = local & 1;
            int n_tile, w0, h0, n0;
            const int out_tile = tile % p.tiles_out, phase = tile / p.tiles_out;
            tile_coords(p, out_tile, n_tile, w0, h0, n0);
            const int col_base = n_tile * BLOCK_N + half * CPW;
            if (p.chunked && n0 != carry_img) {
                flush_carry(carry_img);
                carry_img = n0;
            }
            auto wait_accumulator = [&]() {
                if (e == 0 && lane == 0) mark(local, 4);
                if (short_k) tc::mbar_wait(tc::smem_u32(&bar_acc_full[as]), (local >> 1) & 1);
                else tc::mbar_wait_backoff(tc::smem_u32(&bar_acc_full[as]), (local >> 1) & 1);
                if (e == 0 && lane == 0) mark(local, 5);
                tc::fence_after_sync();
            };
            // Narrow tiles (<= 128 columns): everything that does not depend on the accumulator happens BEFORE the wait for
            // it -- this lane's pixel, its residual / gate rows and the residual of the first chunk (an L2 round trip on
            // layers whose whole reduction lasts a microsecond).  The N = 256 layers (long reductions; the power-capped ADM
            // step) keep the order wait -> addresses: nothing to hide there, and fewer values live across the wait.
            constexpr bool PREP_FIRST = BLOCK_N <= 128;
            if constexpr (!PREP_FIRST) wait_accumulator();
            const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
            const bool ok = active && (n < p.N) && (h < p.H) && (w < p.W);
            const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
            const int64_t rpix = p.res_up ? ((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : pix;
            const __nv_bfloat16* resp = p.res ? p.res + rpix * p.res_ld + col_base : nullptr;
            const float* gatep = p.gate ? p.gate + (int64_t)((uint32_t)pix / (uint32_t)p.gate_rows) * p.gate_ld + col_base : nullptr;
            const bool use_res = resp != nullptr && ok;
            // (narrow tiles only: with N = 256 the extra 16 registers spill, and those layers have long reductions)
            constexpr bool RES_AHEAD = BLOCK_N <= 128;
            uint4 rsd_next[RES_AHEAD ? 4 : 1];
            if (RES_AHEAD && use_res) {
#pragma unroll
                for (int q = 0; q < (RES_AHEAD ? 4 : 1); ++q)
                    rsd_next[q] = col_base + 8 * q < p.c_out ? __ldg(reinterpret_cast<const uint4*>(resp) + q) : make_uint4(0, 0, 0, 0);
            }
            // bias and gate of this warp's columns, ONE value per lane and 32-column chunk, broadcast with shuffles when the
            // chunk is processed: these kernels leave the SM no L1 cache (the shared-memory carve-out is the whole array), so
            // every per-chunk __ldg of a bias / gate row was a full L2 round trip on the critical path of a short tile
            // (narrow tiles, like RES_AHEAD: the N = 256 layers have long reductions that hide the round trip, and on the
            // power-capped ADM step the extra shuffles cost more than they save -- same-box A/B, 8.47 vs 8.35 images/s)
            constexpr bool LANE_ROWS = BLOCK_N <= 128;
            float bias_l[LANE_ROWS ? CPW / 32 : 1], gate_l[LANE_ROWS ? CPW / 32 : 1];
            int gate_sample = -1;
            if constexpr (LANE_ROWS) {
                if (p.gate && p.gate_uniform) gate_sample = __reduce_max_sync(0xffffffffu, ok ? (int)((uint32_t)pix / (uint32_t)p.gate_rows) : -1);
#pragma unroll
                for (int c = 0; c < CPW / 32; ++c) {
                    const int cl = col_base + 32 * c + lane;
                    bias_l[c] = (p.bias && cl < p.c_out) ? __ldg(p.bias + cl) : 0.f;
                    gate_l[c] = (gate_sample >= 0 && cl < p.c_out) ? __ldg(p.gate + (int64_t)gate_sample * p.gate_ld + cl) : 1.f;
                }
            }
            const uint32_t stage = TWO_STAGING ? stage0 + 4 * (local & 1) * (32 * 128) : stage0;
            const uint32_t my_row = stage + (uint32_t)lane * 128u;
            if constexpr (PREP_FIRST) wait_accumulator();
            if (active) {
                const int sh = h0 + r0h, sn = n0 + r0n;
                float rs_sum = 0.f, rs_sq = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < CPW; c0 += 32) {
                    uint32_t acc[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * C::ACC_COLS + half * CPW + c0);
                    tc::tmem_ld_32x32b_x32(taddr, acc);
                    const int col = col_base + c0;
                    // this chunk's residual was requested one chunk (or one barrier wait) ago; request the next one
                    uint4 rsd[4];
                    if constexpr (RES_AHEAD) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) rsd[q] = rsd_next[q];
                        if (use_res && c0 + 32 < CPW) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                rsd_next[q] = col + 32 + 8 * q < p.c_out ? __ldg(reinterpret_cast<const uint4*>(resp + c0 + 32) + q) : make_uint4(0, 0, 0, 0);
                        }
                    } else if (use_res) {  // loads that do not depend on the accumulator go first
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            rsd[q] = col + 8 * q < p.c_out ? __ldg(reinterpret_cast<const uint4*>(resp + c0) + q) : make_uint4(0, 0, 0, 0);
                    }
                    tc::tmem_ld_wait();
                    if (e == 0 && lane == 0) mark(local, c0 == 0 ? 7 : 11);
                    float v[32];
                    float gate_c = 1.f;
                    if constexpr (LANE_ROWS) {
                        float bias_c = bias_l[0];
                        gate_c = gate_l[0];
#pragma unroll
                        for (int c = 1; c < CPW / 32; ++c)
                            if (c0 == 32 * c) bias_c = bias_l[c], gate_c = gate_l[c];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + __shfl_sync(0xffffffffu, bias_c, j);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.bias && col + 4 * q < p.c_out) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + q);
                            v[4 * q] = __uint_as_float(acc[4 * q]) + b4.x, v[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + b4.y;
                            v[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + b4.z, v[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + b4.w;
                        }
                    }
                    activate_all(v, p.act);
                    if (e == 0 && lane == 0) mark(local, c0 == 0 ? 8 : 12);
                    if (gate_sample >= 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= __shfl_sync(0xffffffffu, gate_c, j);
                    } else if (gatep && ok) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (col + 4 * q < p.c_out) {
                                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gatep + c0) + q);
                                v[4 * q] *= g4.x, v[4 * q + 1] *= g4.y, v[4 * q + 2] *= g4.z, v[4 * q + 3] *= g4.w;
                            }
                        }
                    }
                    if (use_res) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t rr[4] = {rsd[q].x, rsd[q].y, rsd[q].z, rsd[q].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                v[8 * q + 2 * j] += bf16_bits_to_f32(rr[j] & 0xffffu);
                                v[8 * q + 2 * j + 1] += bf16_bits_to_f32(rr[j] >> 16);
                            }
                        }
                    }
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        packed[j] = *reinterpret_cast<uint32_t*>(&t);
                    }
                    if (e == 0 && lane == 0) mark(local, c0 == 0 ? 9 : 13);
                    if (p.rowstat) {
                        // {sum, sum of squares} of the STORED values of this pixel's 64-channel block, for the per-pixel
                        // normalisation in the consumer's input transform (ConvParams::in_norm)
                        float s_ = 0.f, q_ = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float lo = bf16_bits_to_f32(packed[j] & 0xffffu), hi = __uint_as_float(packed[j] & 0xffff0000u);
                            s_ += lo + hi, q_ = fmaf(lo, lo, fmaf(hi, hi, q_));
                        }
                        if constexpr (SPLIT64) rs_sum = s_, rs_sq = q_;
                        else if ((c0 & 32) == 0) rs_sum = s_, rs_sq = q_;
                        else if (ok && col < p.c_out + 32)
                            p.rowstat[pix * (p.c_out >> 6) + ((col_base + (c0 & ~63)) >> 6)] = make_float2(rs_sum + s_, rs_sq + q_);
                    }
                    if (e == 0 && lane == 0) mark(local, c0 == 0 ? 10 : 14);
                    if constexpr (SPLIT64) {
                        // the pair's staging block is free once the TMA unit has read the store of two tiles ago out of it
                        if (store_pending) {
                            if (half == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            pair_sync();
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t slot = (uint32_t)((4 * half + q) ^ sw);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my_row + (slot << 4)), "r"(packed[4 * q]),
                                         "r"(packed[4 * q + 1]), "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                                         : "memory");
                        }
                        if (p.rowstat && half == 1) pair_sums[quarter][lane] = make_float2(rs_sum, rs_sq);
                        tc::fence_proxy_async();
                        pair_sync();
                        if (half == 0) {
                            if (lane == 0) {
                                const int cblk = n_tile * BLOCK_N;
                                if (cblk < p.c_out) tc::tma_store_4d(&tmap_out, stage, cblk, w0, sh, sn);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                            if (p.rowstat && ok) {
                                const float2 up = pair_sums[quarter][lane];
                                p.rowstat[pix * (p.c_out >> 6) + n_tile] = make_float2(rs_sum + up.x, rs_sq + up.y);
                            }
                        }
                        store_pending = true;
                    } else {
                    // the staging block is free once the TMA unit has read the previous store out of it
                    if ((c0 & 32) == 0 && store_pending) {
                        if (lane == 0) {
                            if (TWO_STAGING) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                        __syncwarp();
                    }
                    // 32 channels = slots (c0 & 63) / 8 + 0..3 of this pixel's 128-byte row; slot s lives at s ^ (row & 7)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t slot = (uint32_t)((((c0 & 63) >> 3) + q) ^ sw);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my_row + (slot << 4)), "r"(packed[4 * q]),
                                     "r"(packed[4 * q + 1]), "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                                     : "memory");
                    }
                    if ((c0 & 32) != 0) {  // a 64-channel block is complete: hand it to the TMA unit
                        tc::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            const int cblk = col_base + (c0 & ~63);
                            if (cblk < p.c_out) {
                                if (HALO && p.phases > 1) {  // (channel + dx * ld, w, dy, h, n) of the full-resolution tensor
                                    tc::tma_store_5d(&tmap_out, stage, cblk + (phase & 1) * (int)p.out_ld, w0, phase >> 1, sh, sn);
                                } else if (p.out_up) {  // nn.Upsample(2, nearest) of the output: the same map, all four positions
#pragma unroll
                                    for (int ph = 0; ph < 4; ++ph)
                                        tc::tma_store_5d(&tmap_out, stage, cblk + (ph & 1) * (int)p.out_ld, w0, ph >> 1, sh, sn);
                                } else {
                                    tc::tma_store_4d(&tmap_out, stage, cblk, w0, sh, sn);
                                }
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        store_pending = true;
                    }
                    }
                    if (p.gn_acc) {
                        // sums of the STORED (rounded) values: 8 channels in the lane, then the 32 rows of the slab
                        float s4[4], q4[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            float s_ = 0.f, q_ = 0.f;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float lo = bf16_bits_to_f32(packed[4 * b + j] & 0xffffu), hi = bf16_bits_to_f32(packed[4 * b + j] >> 16);
                                s_ += lo, q_ = fmaf(lo, lo, q_);
                                s_ += hi, q_ = fmaf(hi, hi, q_);
                            }
                            const bool valid = ok && col + 8 * b < p.c_out;
                            s4[b] = valid ? s_ : 0.f, q4[b] = valid ? q_ : 0.f;
                        }
                        // transposed butterfly: after offsets 16 and 8 the lane holds block 2 * bit4 + bit3, then a plain fold
                        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
                        float s0 = (b4 ? s4[2] : s4[0]) + __shfl_xor_sync(0xffffffffu, b4 ? s4[0] : s4[2], 16);
                        float s1 = (b4 ? s4[3] : s4[1]) + __shfl_xor_sync(0xffffffffu, b4 ? s4[1] : s4[3], 16);
                        float q0 = (b4 ? q4[2] : q4[0]) + __shfl_xor_sync(0xffffffffu, b4 ? q4[0] : q4[2], 16);
                        float q1 = (b4 ? q4[3] : q4[1]) + __shfl_xor_sync(0xffffffffu, b4 ? q4[1] : q4[3], 16);
                        float sa = (b3 ? s1 : s0) + __shfl_xor_sync(0xffffffffu, b3 ? s0 : s1, 8);
                        float qa = (b3 ? q1 : q0) + __shfl_xor_sync(0xffffffffu, b3 ? q0 : q1, 8);
#pragma unroll
                        for (int off = 4; off >= 1; off >>= 1) {
                            sa += __shfl_xor_sync(0xffffffffu, sa, off);
                            qa += __shfl_xor_sync(0xffffffffu, qa, off);
                        }
                        const int bcol = col + own_blk * 8;
                        if (owner && bcol < p.c_out) {
                            if (p.chunked) {
                                const int slot = n_tile * 4 + (c0 >> 5);
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i == slot) carry_s[i] += sa, carry_q[i] += qa;
                            } else {
                                const int img = n0 + (quarter * 32) / (p.BW * p.BH);  // the image this 32-row slab lies in
                                if (img < p.N) {
                                    unsigned long long* dst = p.gn_acc + ((int64_t)img * (p.c_out >> 3) + (bcol >> 3)) * 4;
                                    fixed_add(dst, sa);
                                    fixed_add(dst + 2, qa);
                                }
                            }
                        }
                    }
                }
            }
            // release the accumulator stage to the MMA warp
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && !leader) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bar_acc_empty[as]), 0));
                else tc::mbar_arrive(tc::smem_u32(&bar_acc_empty[as]));
            }
            if (e == 0 && lane == 0) mark(local, 6);
        }
        if (p.chunked) flush_carry(carry_img);
        if (lane == 0 && store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
    } else {
        // ===== epilogue: warp e reads TMEM lanes [32*(warp%4), +32) and one half of the columns =====
        // Row domain (lane = tile row, as tcgen05.ld delivers it) -> swizzled fp32 staging in shared memory
        // -> "coalesced domain": 4 lanes cover 32 consecutive channels of one pixel (64 contiguous bytes), a
        // warp instruction covers 8 pixels.  Bias, residual, rounding, the store and the GroupNorm sums all
        // happen in the coalesced domain (8x fewer memory wavefronts than one-row-per-lane stores).
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int act = LEAN ? 0 : p.act;
        const float* const gate = LEAN ? nullptr : p.gate;
        float2* const colsum = LEAN ? nullptr : p.colsum;
        const int stat_gran = LEAN ? 8 : p.stat_gran;
        const int out_mode = LEAN ? 0 : p.out_mode;
        const int splits = LEAN ? 1 : p.splits;
        const int half = (BLOCK_N >= 64) ? (e >> 2) : 0;
        const bool active = (BLOCK_N >= 64) || (e < 4);
        constexpr int CHUNK = C::CHUNK;
        const uint32_t stage_base = smem_base + C::RING_BYTES + e * (32 * 128);
        const int cg4 = lane & 3;       // which 8-channel group of the 32-column chunk
        const int rl = lane >> 2;       // row within a group of 8 rows
        // chunked mode: the lanes that own an 8-channel block (rl == 0) carry its GroupNorm sums across the tiles of
        // an image in registers, slot = n_tile * 4 + chunk (n_tiles <= 2, <= 4 chunks of 32 columns per warp)
        float carry_s[8], carry_q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) carry_s[i] = carry_q[i] = 0.f;
        int carry_img = -1;
        auto flush_carry = [&](int img) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int col = (i >> 2) * BLOCK_N + half * C::COLS_PER_WARP + (i & 3) * 32 + cg4 * 8;
                if (rl == 0 && img >= 0 && img < p.N && (i & 3) < C::COLS_PER_WARP / 32 && (i >> 2) < p.n_tiles &&
                    col < p.c_out) {
                    unsigned long long* dst = p.gn_acc + ((int64_t)img * (p.c_out >> 3) + (col >> 3)) * 4;
                    fixed_add(dst, carry_s[i]);
                    fixed_add(dst + 2, carry_q[i]);
                }
                carry_s[i] = carry_q[i] = 0.f;
            }
        };
        for (int local = 0; local < tile_count; ++local) {
            const int tile = unit_to_tile(tile_first + local * tile_step);
            const int as = local & 1;
            int n_tile, w0, h0, n0;
            const int out_tile = tile % p.tiles_out, split = tile / p.tiles_out;
            tile_coords(p, out_tile, n_tile, w0, h0, n0);
            const int col_base = n_tile * BLOCK_N + half * C::COLS_PER_WARP;
            const int m_tile = out_tile / p.n_tiles;
            if (p.chunked && n0 != carry_img) {  // a new image begins: hand the finished one to the accumulators
                flush_carry(carry_img);
                carry_img = n0;
            }

            // the epilogue is ahead of the main loop most of the time: wait with back-off instead of a hot spin that
            // would take issue slots (and power) from the producer and MMA lanes
            tc::mbar_wait_backoff(tc::smem_u32(&bar_acc_full[as]), (local >> 1) & 1);
            tc::fence_after_sync();

            if (!LEAN && active && split > 0) {
                // split-K helper: this CTA accumulated K range `split`; its raw fp32 accumulator goes to the workspace
                // (lane = tile row: 128 contiguous bytes per lane and chunk), then the owner (split 0) is signalled
                float* dst = p.ws_partial + ((int64_t)(split - 1) * p.tiles_out + out_tile) * (BLOCK_M * BLOCK_N) +
                             (int64_t)(quarter * 32 + lane) * BLOCK_N + half * C::COLS_PER_WARP;
#pragma unroll 1
                for (int c0 = 0; c0 < C::COLS_PER_WARP; c0 += CHUNK) {
                    uint32_t acc[CHUNK];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                           (uint32_t)(as * C::ACC_COLS + half * C::COLS_PER_WARP + c0);
                    if constexpr (CHUNK == 32) tc::tmem_ld_32x32b_x32(taddr, acc);
                    else tc::tmem_ld_32x32b_x16(taddr, acc);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < CHUNK; j += 4)
                        __stcg(reinterpret_cast<uint4*>(dst + c0 + j), make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(p.ws_flags + out_tile * EPI_WARPS + e, 1);
            } else if (active && out_mode == 0) {
                if (splits > 1) {  // owner: wait for the S - 1 helpers of this warp's part of the tile
                    int* flag = p.ws_flags + out_tile * EPI_WARPS + e;
                    if (lane == 0) {
                        int seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                        } while (seen < splits - 1);
                        *flag = 0;  // clean for the next launch (stream order)
                    }
                    __syncwarp();
                }
                {
                    // pixels of the 4 rows this lane handles in the coalesced domain; their output / residual addresses
                    // at this lane's first channel are computed once per tile, a chunk only adds its column offset
                    int64_t pixc[4];
                    bool okc[4];
                    __nv_bfloat16* outp[4];
                    const __nv_bfloat16* resp[4];
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = quarter * 32 + it * 8 + rl;
                        const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
                        const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
                        okc[it] = (n < p.N) && (h < p.H) && (w < p.W);
                        pixc[it] = ((int64_t)n * p.H + h) * p.W + w;
                        // phase (dy, dx) of an upsampling convolution: tile pixel (h, w) of the half-resolution grid is
                        // output pixel (2 h + dy, 2 w + dx)
                        const int64_t opix = HALO && p.phases > 1
                                                 ? ((int64_t)n * (2 * p.H) + 2 * h + (split >> 1)) * (2 * p.W) + 2 * w + (split & 1)
                                                 : pixc[it];
                        outp[it] = reinterpret_cast<__nv_bfloat16*>(p.out) + opix * p.out_ld + col_base + cg4 * 8;
                        // res_up: the residual lives at half the resolution and is read through a nearest-neighbour 2x
                        // upsampling (Upsample of the ResBlock's skip branch, _src/unet.py:101-109,231-233), never stored
                        const int64_t rpix = p.res_up ? ((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : pixc[it];
                        resp[it] = p.res + rpix * p.res_ld + col_base + cg4 * 8;
                    }
#pragma unroll 1
                    for (int c0 = 0; c0 < C::COLS_PER_WARP; c0 += 32) {
                        uint32_t acc[32];
                        __syncwarp();  // staging buffer free again; tcgen05.ld is warp-collective
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                               (uint32_t)(as * C::ACC_COLS + half * C::COLS_PER_WARP + c0);
                        if constexpr (CHUNK == 32) {
                            tc::tmem_ld_32x32b_x32(taddr, acc);
                        } else {  // BLOCK_N = 16: the upper half of the chunk does not exist
                            uint32_t lo[16];
                            tc::tmem_ld_32x32b_x16(taddr, lo);
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc[j] = lo[j], acc[16 + j] = 0u;
                        }
                        const int col = col_base + c0 + cg4 * 8;
                        const bool col_ok = col < p.c_out;
                        // residual and bias loads are issued before the TMEM wait so the latencies overlap
                        uint4 rsd[4];
                        const bool use_res = p.res != nullptr && col_ok;
                        if (use_res) {
#pragma unroll
                            for (int it = 0; it < 4; ++it)
                                rsd[it] = okc[it] ? __ldg(reinterpret_cast<const uint4*>(resp[it] + c0)) : make_uint4(0, 0, 0, 0);
                        }
                        float bias8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (p.bias && col_ok) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + 1);
                            bias8[0] = b0.x, bias8[1] = b0.y, bias8[2] = b0.z, bias8[3] = b0.w;
                            bias8[4] = b1.x, bias8[5] = b1.y, bias8[6] = b1.z, bias8[7] = b1.w;
                        }
                        tc::tmem_ld_wait();
                        if (splits > 1) {  // fold the helpers' partial accumulators (row domain, L2-resident)
                            for (int sp = 1; sp < splits; ++sp) {
                                const float* src = p.ws_partial + ((int64_t)(sp - 1) * p.tiles_out + out_tile) * (BLOCK_M * BLOCK_N) +
                                                   (int64_t)(quarter * 32 + lane) * BLOCK_N + half * C::COLS_PER_WARP + c0;
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + j));
                                    acc[j] = __float_as_uint(__uint_as_float(acc[j]) + v.x);
                                    acc[j + 1] = __float_as_uint(__uint_as_float(acc[j + 1]) + v.y);
                                    acc[j + 2] = __float_as_uint(__uint_as_float(acc[j + 2]) + v.z);
                                    acc[j + 3] = __float_as_uint(__uint_as_float(acc[j + 3]) + v.w);
                                }
                            }
                        }
                        // row domain -> staging: 16-byte slot j of row `lane` lives at slot j ^ (lane & 7)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t addr = stage_base + lane * 128 + ((j ^ (lane & 7)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(acc[4 * j]),
                                         "r"(acc[4 * j + 1]), "r"(acc[4 * j + 2]), "r"(acc[4 * j + 3])
                                         : "memory");
                        }
                        __syncwarp();
                        float s1[8], s2[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int r = it * 8 + rl;
                            float f[8];
#pragma unroll
                            for (int hslot = 0; hslot < 2; ++hslot) {
                                const uint32_t addr = stage_base + r * 128 + (((2 * cg4 + hslot) ^ (r & 7)) << 4);
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                             : "=f"(f[4 * hslot]), "=f"(f[4 * hslot + 1]), "=f"(f[4 * hslot + 2]),
                                               "=f"(f[4 * hslot + 3])
                                             : "r"(addr)
                                             : "memory");
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] += bias8[j];
                            activate_all(f, act);
                            if (gate && col_ok && okc[it]) {
                                const float* gp = gate + (pixc[it] / p.gate_rows) * p.gate_ld + col;
                                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp));
                                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
                                f[0] *= g0.x, f[1] *= g0.y, f[2] *= g0.z, f[3] *= g0.w;
                                f[4] *= g1.x, f[5] *= g1.y, f[6] *= g1.z, f[7] *= g1.w;
                            }
                            if (use_res) {
                                const uint32_t rr[4] = {rsd[it].x, rsd[it].y, rsd[it].z, rsd[it].w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    f[2 * j] += bf16_bits_to_f32(rr[j] & 0xffffu);
                                    f[2 * j + 1] += bf16_bits_to_f32(rr[j] >> 16);
                                }
                            }
                            uint32_t packed[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                                packed[j] = *reinterpret_cast<uint32_t*>(&t);
                            }
                            if (okc[it] && col_ok) {
                                *reinterpret_cast<uint4*>(outp[it] + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                                // statistics are taken of the STORED (rounded) values
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float a = bf16_bits_to_f32(packed[j] & 0xffffu), b = bf16_bits_to_f32(packed[j] >> 16);
                                    s1[2 * j] += a, s2[2 * j] += a * a;
                                    s1[2 * j + 1] += b, s2[2 * j + 1] += b * b;
                                }
                            }
                        }
                        if (colsum || p.gn_acc) {
                            const int64_t slab = (int64_t)m_tile * 4 + quarter;
                            const int img = n0 + (quarter * 32) / (p.BW * p.BH);  // the image this 32-row slab lies in
                            if (stat_gran == 8) {
                                // one {sum, sumsq} per 8-channel block: fold the lane's 8 channels, then the 8 row lanes
                                float a = ((s1[0] + s1[1]) + (s1[2] + s1[3])) + ((s1[4] + s1[5]) + (s1[6] + s1[7]));
                                float b = ((s2[0] + s2[1]) + (s2[2] + s2[3])) + ((s2[4] + s2[5]) + (s2[6] + s2[7]));
#pragma unroll
                                for (int off = 4; off <= 16; off <<= 1) {
                                    a += __shfl_xor_sync(0xffffffffu, a, off);
                                    b += __shfl_xor_sync(0xffffffffu, b, off);
                                }
                                if (rl == 0 && col_ok) {
                                    if (p.chunked) {  // BN == 1, n_tiles <= 2, stat_gran == 8: carried, flushed per image
                                        const int slot = n_tile * 4 + (c0 >> 5);
#pragma unroll
                                        for (int i = 0; i < 8; ++i)
                                            if (i == slot) carry_s[i] += a, carry_q[i] += b;
                                    } else if (p.gn_acc) {
                                        if (img < p.N) {
                                            unsigned long long* dst =
                                                p.gn_acc + ((int64_t)img * (p.c_out >> 3) + (col >> 3)) * 4;
                                            fixed_add(dst, a);
                                            fixed_add(dst + 2, b);
                                        }
                                    } else {
                                        colsum[slab * (p.c_out >> 3) + (col >> 3)] = make_float2(a, b);
                                    }
                                }
                            } else {
                                // per channel: transpose-reduce 8 values over the 8 row lanes (bits 2..4 of the lane id);
                                // afterwards the lane with row-lane id k holds channel col + k
#pragma unroll
                                for (int off = 16, nn = 4; off >= 4; off >>= 1, nn >>= 1) {
                                    const bool hi = (lane & off) != 0;
#pragma unroll
                                    for (int i = 0; i < nn; ++i) {
                                        const float send1 = hi ? s1[i] : s1[i + nn], keep1 = hi ? s1[i + nn] : s1[i];
                                        const float send2 = hi ? s2[i] : s2[i + nn], keep2 = hi ? s2[i + nn] : s2[i];
                                        s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
                                        s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
                                    }
                                }
                                const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                                if (col + k < p.c_out) {
                                    if (p.gn_acc) {
                                        if (img < p.N) {
                                            unsigned long long* dst = p.gn_acc + ((int64_t)img * p.c_out + col + k) * 4;
                                            fixed_add(dst, s1[0]);
                                            fixed_add(dst + 2, s2[0]);
                                        }
                                    } else {
                                        colsum[slab * p.c_out + col + k] = make_float2(s1[0], s2[0]);
                                    }
                                }
                            }
                        }
                    }
                }
            } else if (!LEAN && active) {
                // fp32 NCHW network output: out[n][c][h][w]; consecutive lanes are consecutive w => coalesced per channel
                const int row = quarter * 32 + lane;
                const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
                const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
                const bool row_ok = (n < p.N) && (h < p.H) && (w < p.W);
#pragma unroll 1
                for (int c0 = 0; c0 < C::COLS_PER_WARP; c0 += CHUNK) {
                    uint32_t acc[CHUNK];
                    __syncwarp();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                           (uint32_t)(as * C::ACC_COLS + half * C::COLS_PER_WARP + c0);
                    if constexpr (CHUNK == 32) {
                        tc::tmem_ld_32x32b_x32(taddr, acc);
                    } else {
                        tc::tmem_ld_32x32b_x16(taddr, acc);
                    }
                    tc::tmem_ld_wait();
                    const int col0 = col_base + c0;
                    if (!row_ok || col0 >= p.c_out) continue;
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
            // release the accumulator stage to the MMA warp
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && !leader) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bar_acc_empty[as]), 0));
                else tc::mbar_arrive(tc::smem_u32(&bar_acc_empty[as]));
            }
        }
        if (p.chunked) flush_carry(carry_img);
    }

    tc::fence_before_sync();
    __syncthreads();
    if (trace_slot) {
        atomicMax(trace_slot + 2, azb_globaltimer());
        if (blockIdx.x == 0) atomicAdd(p.trace, 1ull);
    }
    if constexpr (PAIR) {
        tc::cluster_sync();  // neither CTA retires (shared memory, barriers, tensor memory) while its peer still works
        if (warp == 1) tc::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    } else {
        if (warp == 1) tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host: tensor maps + launch

int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, const uint32_t* elem_strides = nullptr) {
    const int rc = tc::make_map_bf16(m, base, rank, dims, strides_bytes, box, elem_strides);
    return rc == 0 ? AZB_OK : rc == -1 ? AZB_E_DRIVER : AZB_E_SHAPE;
}

int sm_count() { return azb_sm_count(); }

#define g_knob azb_knob

template <int BLOCK_N, bool PAIR = false, int EPI = 0, bool HALO = false, bool TF32 = false, bool NORM = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const ConvParams& p, cudaStream_t s,
           const CUtensorMap* tout = nullptr) {
    constexpr int smem = Cfg<BLOCK_N, PAIR, HALO>::SMEM;
    constexpr int threads = HALO ? THREADS_HALO : THREADS;
    auto kernel = conv_gemm_kernel<BLOCK_N, PAIR, EPI, HALO, TF32, NORM>;
    const CUtensorMap& to = tout ? *tout : ta;
    static AzbPerDevice<bool> configured_dev;
    bool& configured = configured_dev.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if constexpr (PAIR) {
        // one cluster of two CTAs (the two SMs of a TPC) per scheduling unit, at most one cluster per TPC
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)sm_count() & ~1u), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see pdl_trigger / pdl_wait
        attr[1].val.programmaticStreamSerializationAllowed = g_knob[AZB_KNOB_PDL] != 0 ? 1 : 0;
        cfg.attrs = attr, cfg.numAttrs = 2;
        static AzbPerDevice<int> resident_dev;  // clusters that fit on the device at once: the persistent grid is one wave
        int& resident = resident_dev.get();
        if (!resident) {
            if (cudaOccupancyMaxActiveClusters(&resident, kernel, &cfg) != cudaSuccess || resident < 1) {
                cudaGetLastError();
                resident = sm_count() / 2;
            }
        }
        const int pairs = p.total_tiles < resident ? p.total_tiles : resident;
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ta, tb, ta2, to, p);
        if (e != cudaSuccess) return (int)e;
    } else {
        const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
        return azb_launch(kernel, dim3((unsigned)grid), dim3(threads), smem, s, ta, tb, ta2, to, p);
    }
    return azb_launch_status();
}

void patch_shape(int64_t h, int64_t w, int& bw, int& bh, int& bn) {
    // patch: as wide as the image up to 16 columns, then rows, then images
    bw = 1;
    while (bw < 16 && bw < w) bw <<= 1;
    bh = 1;
    while (bw * bh < BLOCK_M && bh < h) bh <<= 1;
    bn = BLOCK_M / (bw * bh);
}

// Temporarily forces the tap-wise kernel (restores the knob on scope exit)
struct ExtraGuard {
    int& slot;
    int saved;
    explicit ExtraGuard(int& k) : slot(k), saved(k) { slot = 0; }
    ~ExtraGuard() { slot = saved; }
};

struct ConvExtra {
    int stride = 1;
    int act = AZB_ACT_NONE;
    const float* gate = nullptr;
    int64_t gate_ld = 0;
    int64_t gate_rows = 0;
    const void* act2 = nullptr;  // second operand: NHWC bf16 at the OUTPUT resolution, 1x1, weights appended along K
    int64_t c_in2 = 0, act2_ld = 0, k2 = 0;
    int64_t* gn_acc = nullptr;   // exact per-(image, channel block) sums of `out` for the GroupNorms that consume it
    void* workspace = nullptr;   // split-K scratch: flags (zero between launches) + fp32 partial tiles
    int64_t workspace_bytes = 0;
    int res_up = 0;              // residual given at half resolution (nearest 2x upsampling on the fly)
    int out_up = 0;              // output written through a nearest 2x upsampling: `out` is (n, 2 h, 2 w, ld)
    int in_up = 0;               // act given at half resolution: the convolution reads its nearest 2x upsampling (halo + in_coef)
    const float* in_coef = nullptr;  // [N][c_in] {a, b}: the input is act(a x + b), applied on the fly (halo kernels only)
    int in_silu = 0;
    int in_norm = 0;             // per-pixel LayerNorm (1) / RMSNorm (2) + modulation of the input (halo kernels)
    float in_eps = 1e-5f;
    const float* in_rowstat = nullptr;
    const float* in_mod = nullptr;
    int64_t in_mod_ld = 0;
    float* rowstat = nullptr;    // per-(pixel, 64-channel block) sums of the output (row-domain epilogue only)
    AzbConvChoice* choice = nullptr;  // dry run: report the launcher's choice instead of launching
};

// (h, w) are the INPUT extents; the output is ceil(h / stride) x ceil(w / stride).
int conv_impl(const void* act, int64_t n, int64_t h_in, int64_t w_in, int64_t c_in, int64_t act_ld, const void* wpack,
              int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap, const float* bias, const void* residual,
              int64_t res_ld, void* out, int64_t out_ld, int out_mode, float* colsum, int stat_gran, void* stream,
              const ConvExtra& ex = ConvExtra()) {
    AZB_CHECK_PTR(act);
    AZB_CHECK_PTR(wpack);
    AZB_CHECK_PTR(out);
    if (n <= 0 || h_in <= 0 || w_in <= 0 || c_in <= 0 || c_out <= 0) return AZB_E_SHAPE;
    // Phase-decomposed upsampling convolution (in_up = 2): conv3x3(up2(z)) at output pixel (2 i + dy, 2 j + dx) only sees
    // the 2 x 2 half-resolution pixels at rows i - 1 + dy, i + dy and columns j - 1 + dx, j + dx, with the 3 x 3 taps that
    // fall on the same pixel SUMMED (done once, in fp32, when the weights are packed: [C_out][phase][2 x 2][C_in]): four
    // 2 x 2 convolutions of the half-resolution tensor, 16 instead of 36 tap-GEMMs per half-resolution pixel (2.25 x fewer
    // FLOPs).  Each (M tile, N tile, phase) is a tile of the halo kernel; its taps are views of the same halo tile.
    const int phases = ex.in_up == 2 ? 4 : 1;
    const int taps_w = taps;  // taps in the packed weights
    if (phases > 1) {
        if (taps != 16 || !ex.in_coef || (h_in & 1) || (w_in & 1) || ex.stride != 1 || ex.act2 || residual || out_mode != 0)
            return AZB_E_SHAPE;
        h_in >>= 1, w_in >>= 1;  // tiles, halo loads and the transform work on the half-resolution grid
        taps = 9;                // ... with the halo geometry of a 3 x 3 layer
    }
    const int taps_k = phases > 1 ? 4 : taps;  // k-blocks per channel block in one tile's reduction
    if (taps != 1 && taps != 9) return AZB_E_SHAPE;
    if (ex.stride != 1 && ex.stride != 2) return AZB_E_SHAPE;
    if (ex.act < AZB_ACT_NONE || ex.act > AZB_ACT_RELU2) return AZB_E_SHAPE;
    if (k_per_tap % BLOCK_K || k_per_tap < c_in) return AZB_E_SHAPE;
    if (c_in % 8 || act_ld % 8 || act_ld < c_in) return AZB_E_ALIGN;
    if (!azb_aligned(act, 16) || !azb_aligned(wpack, 16) || !azb_aligned(out, 16)) return AZB_E_ALIGN;
    if (out_mode == 0 && (c_out % 8 || out_ld % 8 || (residual && (res_ld % 8 || !azb_aligned(residual, 16)))))
        return AZB_E_ALIGN;
    if (bias && !azb_aligned(bias, 16)) return AZB_E_ALIGN;
    if (out_mode != 0 && out_mode != 1) return AZB_E_SHAPE;
    if ((colsum || ex.gn_acc) && (out_mode != 0 || (stat_gran != 1 && stat_gran != 8))) return AZB_E_SHAPE;
    if (colsum && (ex.gn_acc || !azb_aligned(colsum, 8))) return AZB_E_SHAPE;
    if (ex.gn_acc && !azb_aligned(ex.gn_acc, 8)) return AZB_E_ALIGN;
    if (ex.gate && n * ((h_in + ex.stride - 1) / ex.stride) * ((w_in + ex.stride - 1) / ex.stride) >= 0x7fffffffLL) return AZB_E_SHAPE;
    if (ex.gate && (out_mode != 0 || ex.gate_rows <= 0 || ex.gate_ld % 4 || !azb_aligned(ex.gate, 16))) return AZB_E_ALIGN;
    if (out_mode == 1 && ex.act != AZB_ACT_NONE) return AZB_E_UNSUPPORTED;
    if (ex.act2 && (ex.c_in2 <= 0 || ex.c_in2 % 8 || ex.act2_ld % 8 || ex.act2_ld < ex.c_in2 || ex.k2 % BLOCK_K ||
                    ex.k2 < ex.c_in2 || !azb_aligned(ex.act2, 16)))
        return AZB_E_ALIGN;
    const int64_t h = (h_in + ex.stride - 1) / ex.stride, w = (w_in + ex.stride - 1) / ex.stride;

    // The row-domain epilogue (EPI == 2) brings activation and gate to the halo kernels.  Measured: neutral to +3 % on the
    // ADM layers (no activation / gate, long reductions); +25 % on the token GEMMs with activation / gate epilogues
    // (768 -> 3072 + SiLU: 87 -> 70 us, scripts/gemm_one.py) and 2 - 7 us per launch on the in-repo U-Net's 3 x 3 layers
    // (scripts/unet_conv_ab.py) once the activation selector is tested outside the element loop, bias / gate rows are held
    // per lane instead of re-read through an L1 these kernels do not have, and 64-column tiles use all eight warps.
    const bool rowepi_ok = g_knob[AZB_CONV_KNOB_ROWEPI] != 0 && out_mode == 0 && !colsum && (!ex.gn_acc || stat_gran == 8) &&
                           (phases == 1 || (c_out % 64 == 0 && out_ld % 8 == 0));
    if (ex.in_norm < 0 || ex.in_norm > 2) return AZB_E_SHAPE;
    if (ex.in_norm && (!ex.in_rowstat || !ex.in_mod || ex.in_coef || ex.in_up || taps != 9 || ex.stride != 1 || c_in % BLOCK_K ||
                       ex.in_mod_ld % 4 || !azb_aligned(ex.in_mod, 16) || !azb_aligned(ex.in_rowstat, 8)))
        return AZB_E_SHAPE;
    if (ex.rowstat && (out_mode != 0 || c_out % 64 || phases != 1 || !azb_aligned(ex.rowstat, 8))) return AZB_E_SHAPE;
    // Halo tiles: 3 x 3, stride 1, maps of at least one 8 x 16 patch, whole 64-channel blocks (see the kernel's header)
    // ... and, unless the input transform / upsampling on load needs them, only on feature maps of at least 32 patches
    // (64 x 64 pixels): on smaller maps a CTA gets one or two tiles and the tap-wise kernel's plain load -> MMA stream starts
    // up faster than load -> transform slot -> MMA (in-repo U-Net, 128 / 256 channels at 32 x 32 / 16 x 16 x batch 32: 14 - 15
    // us tap-wise against 18 us).  The rule looks at ONE image, not at the batch: a shard of a batch must choose the same
    // kernels as the whole batch (the two classes sum K in different orders).
    const bool halo_needed = ex.in_coef || ex.in_norm || ex.in_up;
    const int64_t halo_tiles_per_image = ((w + HALO_W - 1) / HALO_W) * ((h + HALO_H - 1) / HALO_H);
    bool halo = g_knob[AZB_CONV_KNOB_HALO] != 0 && taps == 9 &&
                (halo_needed || g_knob[AZB_CONV_KNOB_HALO] == 1 || halo_tiles_per_image >= 32) &&
                ex.stride == 1 && h >= HALO_H && w >= HALO_W &&
                c_in % BLOCK_K == 0 && k_per_tap == c_in && (!ex.act2 || (ex.c_in2 % BLOCK_K == 0 && ex.k2 == ex.c_in2)) &&
                !colsum && ((ex.act == AZB_ACT_NONE && !ex.gate) || rowepi_ok) && (!ex.gn_acc || stat_gran == 8) &&
                c_in / BLOCK_K + (ex.act2 ? ex.c_in2 / BLOCK_K : 0) <= 64 && !(ex.act2 && out_mode == 1);
    if (ex.in_coef && !azb_aligned(ex.in_coef, 16)) return AZB_E_ALIGN;
    // in_up: (h_in, w_in) are the UPSAMPLED extents; the zero padding is restored by the input transform, so it needs one
    if (ex.in_up == 1 && (!ex.in_coef || (h_in & 1) || (w_in & 1) || ex.stride != 1 || taps != 9)) return AZB_E_SHAPE;
    if (ex.in_up < 0 || ex.in_up > 2) return AZB_E_SHAPE;

    ConvParams p{};
    p.N = (int)n, p.H = (int)h, p.W = (int)w;
    if (halo) p.BW = HALO_W, p.BH = HALO_H, p.BN = 1;
    else patch_shape(h, w, p.BW, p.BH, p.BN);
    p.tiles_w = (int)((w + p.BW - 1) / p.BW);
    p.tiles_h = (int)((h + p.BH - 1) / p.BH);
    const int64_t tiles_n_img = (n + p.BN - 1) / p.BN;
    const int64_t m_tiles = (int64_t)p.tiles_w * p.tiles_h * tiles_n_img;

    // N tile: the widest one that divides the (padded) row count and still gives every SM a tile
    int block_n = 16;
    const int sms = sm_count();
    const int cand[5] = {256, 128, 64, 32, 16};
    for (int i = 0; i < 5; ++i) {
        if (c_out_rows % cand[i]) continue;
        block_n = cand[i];
        if (m_tiles * (c_out_rows / cand[i]) >= (sms * 3) / 4 || cand[i] <= 64) break;
    }
    // forced CTA pairs (tuning knob): take the wide tile even when it leaves SMs idle
    if (g_knob[AZB_CONV_KNOB_PAIR] == 1 && m_tiles % 2 == 0 && c_out_rows % 128 == 0) block_n = c_out_rows % 256 ? 128 : 256;
    if (g_knob[AZB_CONV_KNOB_BLOCKN] >= 16 && c_out_rows % g_knob[AZB_CONV_KNOB_BLOCKN] == 0) block_n = g_knob[AZB_CONV_KNOB_BLOCKN];
    if (c_out_rows % block_n || c_out_rows < c_out) return AZB_E_SHAPE;

    // Split-K for small feature maps with long reductions (8 x 8 layers: 8 M tiles, K = 9216 .. 18432): a wide N tile
    // alone leaves most SMs idle and the narrow one (N = 64) re-reads the A operand at half MMA rate.  S CTAs share an
    // output tile, each reduces 1/S of K; the helpers park their fp32 accumulators in the workspace and the owner
    // folds them before its epilogue.  Single wave only (all CTAs co-resident), so the owner's wait cannot deadlock.
    int splits = 1;
    const int64_t num_kb_total = (int64_t)taps_k * (k_per_tap / BLOCK_K) + (ex.act2 ? ex.k2 / BLOCK_K : 0);
    // halo kernels exist for the wide lean tiles and for the narrowest generic one (the network's output convolution)
    // ... and, with the row-domain epilogue, for 64-column tiles (the 64-channel level of the in-repo U-Net)
    // (one channel block: resident weights.  With more, the layer is bound by the MMA issuer's per-k-block loop -- wait, fence,
    // four 48-clk MMAs, commit: ~190 ns against 100 ns of tensor work -- in both kernel classes, and the halo one adds its
    // slot handshakes: 192 -> 64 took 57 us with 8 weight stages and 56 us with 15, against 40 us tap-wise)
    const bool halo64 = out_mode == 0 && block_n == 64 && rowepi_ok && !ex.act2 && phases == 1 && !ex.gn_acc && splits == 1 &&
                        !ex.in_coef && !ex.in_up && (k_per_tap == BLOCK_K || ex.in_norm);
    if (halo && !((out_mode == 0 && block_n >= 128) || (out_mode == 1 && block_n == 16) || halo64)) {
        if (ex.in_coef || ex.in_norm) return AZB_E_UNSUPPORTED;
        // recompute the tiling for the tap-wise kernel
        ExtraGuard guard(g_knob[AZB_CONV_KNOB_HALO]);
        return conv_impl(act, n, h_in, w_in, c_in, act_ld, wpack, c_out, c_out_rows, taps, k_per_tap, bias, residual, res_ld,
                         out, out_ld, out_mode, colsum, stat_gran, stream, ex);
    }
    if (!halo && (ex.in_coef || ex.in_norm)) return AZB_E_UNSUPPORTED;
    if (!halo && !ex.res_up && g_knob[AZB_CONV_KNOB_SPLITK] != 0 && ex.workspace && out_mode == 0 && m_tiles * (c_out_rows / block_n) <= sms && (block_n <= 64 || m_tiles * (c_out_rows / block_n) < sms / 2)) {
        const int try_n[2] = {128, 256}, try_s[2] = {2, 4};
        for (int i = 0; i < 2 && splits == 1; ++i) {
            const int bn = try_n[i], sp = try_s[i];
            if (c_out_rows % bn || num_kb_total % sp || num_kb_total / sp < 16) continue;
            const int64_t tiles = m_tiles * (c_out_rows / bn);
            if (tiles * sp > sms || tiles * sp < sms / 2) continue;
            const int64_t need = 256 + ((tiles * EPI_WARPS * 4 + 255) / 256) * 256 + (int64_t)(sp - 1) * tiles * BLOCK_M * bn * 4;
            if (need > ex.workspace_bytes) continue;
            block_n = bn, splits = sp;
        }
    }

    p.n_tiles = (int)(c_out_rows / block_n);
    if (m_tiles * p.n_tiles * splits > 0x7fffffffLL) return AZB_E_SHAPE;
    p.tiles_out = (int)(m_tiles * p.n_tiles);
    // CTA pairs: wide tiles, an even number of M tiles, and enough of them that every TPC gets several units
    bool pair = block_n >= 128 && splits == 1 && m_tiles % 2 == 0 && g_knob[AZB_CONV_KNOB_PAIR] != 0;
    // measured on the ADM shapes (scripts/conv_ab.py): pairs win 5 - 20 % whenever the reduction is at least 16 k-blocks;
    // with short reductions (1x1 layers, K <= 512) the epilogue dominates and the pair's extra handshakes cost 5 - 8 %
    if (pair && g_knob[AZB_CONV_KNOB_PAIR] < 0) pair = num_kb_total >= 16;
    // halo kernels stage the plain tiles of a fused 1 x 1 operand in the weight ring: a stage must hold 128 x 64 bf16
    if (halo && ex.act2 && block_n == 128) pair = false;
    p.total_tiles = (pair ? p.tiles_out / 2 : p.tiles_out * splits) * phases;
    p.phases = phases;
    if (phases > 1 && !halo) return AZB_E_UNSUPPORTED;
    p.splits = splits;
    if (splits > 1) {
        if (!azb_aligned(ex.workspace, 256)) return AZB_E_ALIGN;
        p.ws_flags = reinterpret_cast<int*>(ex.workspace);
        p.ws_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ex.workspace) +
                                                 (((int64_t)p.tiles_out * EPI_WARPS * 4 + 255) / 256) * 256);
    }
    p.taps = taps_k, p.ksize = taps == 9 ? 3 : 1, p.pad = taps == 9 ? 1 : 0;
    p.kb_per_tap = (int)(k_per_tap / BLOCK_K);
    p.c_out = (int)c_out;
    p.bias = bias;
    p.res = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.res_ld = res_ld;
    if (ex.res_up && (!residual || (h & 1) || (w & 1) || out_mode != 0 || splits > 1)) return AZB_E_SHAPE;
    p.res_up = ex.res_up;
    p.out_up = ex.out_up;
    p.out = out, p.out_ld = out_ld, p.out_mode = out_mode;
    p.colsum = reinterpret_cast<float2*>(colsum);
    p.stat_gran = stat_gran;
    p.stride = ex.stride, p.act = ex.act;
    p.gate = ex.gate, p.gate_ld = ex.gate_ld, p.gate_rows = (int)ex.gate_rows;
    // a slab = 32 consecutive tile rows: inside one image when the per-image patch has >= 32 pixels; a sample is then a
    // whole number of images (gate_rows = k H W), or -- one image, 16-column token grids -- a whole number of slabs
    p.gate_uniform = (p.BW * p.BH >= 32 && (ex.gate_rows % (h * w) == 0 || (n == 1 && p.BW == w && ex.gate_rows % 32 == 0))) ? 1 : 0;
    p.kb_extra = ex.act2 ? (int)(ex.k2 / BLOCK_K) : 0;
    if (ex.gn_acc && (p.BW * p.BH) % 32) return AZB_E_SHAPE;  // a 32-row slab would straddle two images
    p.gn_acc = reinterpret_cast<unsigned long long*>(ex.gn_acc);
    // Large feature maps (one image per M tile, at most two N tiles): contiguous tile ranges per CTA, sums carried
    // across the tiles of an image.  With round-robin tiles every CTA works on the same image at the same time and the ~4 M
    // same-address atomics of a 256 x 256 layer serialise in L2 (measured: +15 % on the K = 2304 layers).
    // few M tiles => every weight tile is used by a handful of CTAs right after its first (HBM) read: with only
    // STAGES loads in flight the main loop would run at HBM latency; prefetch the weight stream into L2 ahead of use
    p.prefetch_kb = halo ? 0 : g_knob[AZB_CONV_KNOB_PREFETCH] >= 0 ? g_knob[AZB_CONV_KNOB_PREFETCH] : (m_tiles <= 32 ? 24 : 0);
    p.in_coef = reinterpret_cast<const float2*>(ex.in_coef), p.in_silu = ex.in_silu, p.c_in = (int)c_in;
    p.trace = reinterpret_cast<unsigned long long*>(azb_trace_buf);
    p.in_norm = ex.in_norm, p.in_eps = ex.in_eps, p.in_rowstat = reinterpret_cast<const float2*>(ex.in_rowstat);
    p.in_mod = ex.in_mod, p.in_mod_ld = ex.in_mod_ld, p.rowstat = reinterpret_cast<float2*>(ex.rowstat);
    p.in_up = ex.in_up == 1;
    p.a_slot = p.in_up ? UP_BYTES : Cfg<256, true, true>::A_SLOT;
    p.sa = g_knob[AZB_CONV_KNOB_HALO_SA] >= 2 && g_knob[AZB_CONV_KNOB_HALO_SA] <= 4 ? g_knob[AZB_CONV_KNOB_HALO_SA] : 3;
    if (halo) {
        // item order of a tile: halo item h is followed by the 1 x 1 blocks [h P / H, (h + 1) P / H)
        const int H = p.kb_per_tap, P = p.kb_extra, spread = g_knob[AZB_CONV_KNOB_HALO_SPREAD] != 0;
        p.item_mask = 0;
        int pos = 0;
        for (int hh = 0, done = 0; hh < H; ++hh) {
            p.item_mask |= 1ull << pos++;
            const int until = spread ? ((hh + 1) * P) / H : (hh + 1 == H ? P : 0);
            for (; done < until; ++done) ++pos;
        }
    }
    p.chunked = (ex.gn_acc && p.n_tiles <= 2 && p.BN == 1 && stat_gran == 8 && block_n >= 64 && splits == 1) ? 1 : 0;
    const int64_t k_total = taps_w * k_per_tap + (ex.act2 ? ex.k2 : 0);

    CUtensorMap ta, tb, ta2;
    if (p.in_up) {
        // virtual nearest-neighbour upsampling of the (n, h / 2, w / 2, c_in) tensor: (channel, x replica [stride 0], x / 2,
        // y replica [stride 0], image row / 2 over all images); out-of-range columns are zero-filled, out-of-image rows
        // read the neighbouring image -- either way the input transform resets out-of-image pixels to zero
        uint64_t dims[5] = {(uint64_t)c_in, 2, (uint64_t)(w_in / 2), 2, (uint64_t)(n * (h_in / 2))};
        uint64_t str[4] = {0, (uint64_t)act_ld * 2, 0, (uint64_t)act_ld * 2 * (uint64_t)(w_in / 2)};
        uint32_t box[5] = {BLOCK_K, 2, UP_PITCH / 2, 2, UP_ROWS / UP_PITCH / 2};
        int rc = make_map(&ta, act, 5, dims, str, box);
        if (rc) return rc;
    } else {
        // strided convolutions traverse the input with element strides (2, 2): the box spans stride * B pixels
        // of the input and delivers B of them; coordinates stay in input pixels
        const uint32_t st = (uint32_t)ex.stride;
        uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)act_ld * 2, (uint64_t)act_ld * 2 * w_in, (uint64_t)act_ld * 2 * w_in * h_in};
        uint32_t box[4] = {BLOCK_K, (uint32_t)p.BW * st, (uint32_t)p.BH * st, (uint32_t)p.BN};
        if (halo) box[1] = HALO_PITCH, box[2] = HALO_H + 2;
        uint32_t es[4] = {1, st, st, 1};
        if (box[1] > 256 || box[2] > 256) return AZB_E_SHAPE;
        int rc = make_map(&ta, act, 4, dims, str, box, es);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)c_out_rows};
        uint64_t str[1] = {(uint64_t)k_total * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)(pair ? block_n / 2 : block_n)};  // a pair member stages half the rows
        int rc = make_map(&tb, wpack, 2, dims, str, box);
        if (rc) return rc;
    }
    if (ex.act2) {
        uint64_t dims[4] = {(uint64_t)ex.c_in2, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)ex.act2_ld * 2, (uint64_t)ex.act2_ld * 2 * w, (uint64_t)ex.act2_ld * 2 * w * h};
        uint32_t box[4] = {BLOCK_K, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BN};
        int rc = make_map(&ta2, ex.act2, 4, dims, str, box);
        if (rc) return rc;
    } else {
        ta2 = ta;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool lean = (halo || g_knob[AZB_CONV_KNOB_LEAN] != 0) && out_mode == 0 && ex.act == AZB_ACT_NONE && !ex.gate && !colsum &&
                      splits == 1 && (!ex.gn_acc || stat_gran == 8) && block_n >= 128;
    // Row-domain epilogue with TMA stores (EPI == 2): every bf16 NHWC layer with an N tile of whole 64-channel store
    // blocks, without split-K / per-channel statistics; activation, gate, residual and exact GroupNorm sums included.
    // A halo kernel needs a wide tile unless it is this epilogue's 64-column instantiation.
    const bool rowepi = rowepi_ok && splits == 1 && block_n >= 64 && !(block_n == 64 && ex.gn_acc);
    if (ex.rowstat && !rowepi) return AZB_E_UNSUPPORTED;  // the per-pixel sums come from the row-domain epilogue only
    if (ex.out_up < 0 || ex.out_up > 1) return AZB_E_SHAPE;
    if (ex.out_up && (out_mode != 0 || c_out % 64 || out_ld % 8 || phases != 1 || ex.rowstat || ex.gn_acc)) return AZB_E_SHAPE;
    if (ex.out_up && (!rowepi || block_n < 128 || p.BW * p.BH < 32)) return AZB_E_UNSUPPORTED;
    if (halo && (ex.act != AZB_ACT_NONE || ex.gate) && !rowepi) return AZB_E_UNSUPPORTED;  // (unreachable: wide halo tiles)
    if (ex.choice) {
        ex.choice->halo = halo, ex.choice->pair = pair, ex.choice->lean = lean || rowepi, ex.choice->block_n = block_n, ex.choice->splits = splits;
        ex.choice->tiles = p.total_tiles;
        ex.choice->epi = rowepi ? 2 : lean ? 1 : 0;
        return AZB_OK;
    }
    if (rowepi) {
        // the 32 pixels a warp stores: tile rows [32 q, 32 q + 32) = (BW, 32 / BW rows) of one image, or whole small images
        const int rows_h = p.BW * p.BH >= 32 ? 32 / p.BW : p.BH;
        const int imgs = p.BW * p.BH >= 32 ? 1 : 32 / (p.BW * p.BH);
        CUtensorMap tout;
        int rc;
        if (phases > 1 || ex.out_up) {
            // full-resolution output (n, 2 H, 2 W, ld) seen as (channel + dx * ld, w, dy, h, n): one map serves all phases
            const uint64_t W2 = 2 * (uint64_t)w, H2 = 2 * (uint64_t)h;
            uint64_t dims[5] = {(uint64_t)out_ld + (uint64_t)c_out, (uint64_t)w, 2, (uint64_t)h, (uint64_t)n};
            uint64_t str[4] = {(uint64_t)out_ld * 4, W2 * out_ld * 2, 2 * W2 * out_ld * 2, H2 * W2 * out_ld * 2};
            uint32_t box[5] = {64, (uint32_t)p.BW, 1, (uint32_t)rows_h, 1};
            rc = make_map(&tout, out, 5, dims, str, box);
        } else {
            uint64_t dims[4] = {(uint64_t)c_out, (uint64_t)w, (uint64_t)h, (uint64_t)n};
            uint64_t str[3] = {(uint64_t)out_ld * 2, (uint64_t)out_ld * 2 * w, (uint64_t)out_ld * 2 * w * h};
            uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)rows_h, (uint32_t)imgs};
            rc = make_map(&tout, out, 4, dims, str, box);
        }
        if (rc) return rc;
        if (halo) {
            const int b_stage = pair ? Cfg<256, true, true>::STAGE_BYTES * block_n / 256 : Cfg<256, false, true>::STAGE_BYTES * block_n / 256;
            p.sb = (SMEM_BUDGET - p.sa * p.a_slot) / b_stage;
            // one channel block, one N tile, all nine weight tiles fit next to the A slots: resident weights
            p.b_resident = (!pair && p.kb_per_tap == 1 && p.kb_extra == 0 && phases == 1 && p.n_tiles == 1 && p.sb >= 9 &&
                            g_knob[AZB_CONV_KNOB_HALO_SB] < 0) ? 1 : 0;
            if (p.sb > 8) p.sb = 8;
            if (g_knob[AZB_CONV_KNOB_HALO_SB] >= 2 && g_knob[AZB_CONV_KNOB_HALO_SB] < p.sb) p.sb = g_knob[AZB_CONV_KNOB_HALO_SB];
            if (p.sb < 2) return AZB_E_SHAPE;
            if (p.in_norm) {  // the instantiations whose transform warps normalise per pixel
                if (pair) return block_n == 256 ? launch<256, true, 2, true, false, true>(ta, tb, ta2, p, s, &tout)
                                                : launch<128, true, 2, true, false, true>(ta, tb, ta2, p, s, &tout);
                if (block_n == 64) return launch<64, false, 2, true, false, true>(ta, tb, ta2, p, s, &tout);
                return block_n == 256 ? launch<256, false, 2, true, false, true>(ta, tb, ta2, p, s, &tout)
                                      : launch<128, false, 2, true, false, true>(ta, tb, ta2, p, s, &tout);
            }
            if (pair) return block_n == 256 ? launch<256, true, 2, true>(ta, tb, ta2, p, s, &tout) : launch<128, true, 2, true>(ta, tb, ta2, p, s, &tout);
            if (block_n == 64) return launch<64, false, 2, true>(ta, tb, ta2, p, s, &tout);
            return block_n == 256 ? launch<256, false, 2, true>(ta, tb, ta2, p, s, &tout) : launch<128, false, 2, true>(ta, tb, ta2, p, s, &tout);
        }
        if (pair) return block_n == 256 ? launch<256, true, 2>(ta, tb, ta2, p, s, &tout) : launch<128, true, 2>(ta, tb, ta2, p, s, &tout);
        switch (block_n) {
            case 256: return launch<256, false, 2>(ta, tb, ta2, p, s, &tout);
            case 128: return launch<128, false, 2>(ta, tb, ta2, p, s, &tout);
            default: return launch<64, false, 2>(ta, tb, ta2, p, s, &tout);
        }
    }
    if (halo) {
        const int b_stage = out_mode == 1 ? Cfg<16, false, true>::STAGE_BYTES
                                          : pair ? Cfg<256, true, true>::STAGE_BYTES * block_n / 256 : Cfg<256, false, true>::STAGE_BYTES * block_n / 256;
        p.sb = (SMEM_BUDGET - p.sa * p.a_slot) / b_stage;
        // (one channel block, one N tile: resident weights, as in the row-domain branch -- the 64 -> 3 output convolution of
        // the in-repo U-Net streamed nine 2 KiB weight tiles per 128-pixel tile through the ring)
        p.b_resident = (!pair && p.kb_per_tap == 1 && p.kb_extra == 0 && phases == 1 && p.n_tiles == 1 && p.sb >= 9 &&
                        g_knob[AZB_CONV_KNOB_HALO_SB] < 0) ? 1 : 0;
        if (p.sb > 8) p.sb = 8;
        if (g_knob[AZB_CONV_KNOB_HALO_SB] >= 2 && g_knob[AZB_CONV_KNOB_HALO_SB] < p.sb) p.sb = g_knob[AZB_CONV_KNOB_HALO_SB];
        if (p.sb < 2) return AZB_E_SHAPE;
        if (out_mode == 1) return launch<16, false, false, true>(ta, tb, ta2, p, s);
        if (pair) return block_n == 256 ? launch<256, true, true, true>(ta, tb, ta2, p, s) : launch<128, true, true, true>(ta, tb, ta2, p, s);
        return block_n == 256 ? launch<256, false, true, true>(ta, tb, ta2, p, s) : launch<128, false, true, true>(ta, tb, ta2, p, s);
    }
    if (pair && lean) return block_n == 256 ? launch<256, true, true>(ta, tb, ta2, p, s) : launch<128, true, true>(ta, tb, ta2, p, s);
    if (pair) return block_n == 256 ? launch<256, true>(ta, tb, ta2, p, s) : launch<128, true>(ta, tb, ta2, p, s);
    if (lean) return block_n == 256 ? launch<256, false, true>(ta, tb, ta2, p, s) : launch<128, false, true>(ta, tb, ta2, p, s);
    switch (block_n) {
        case 256: return launch<256>(ta, tb, ta2, p, s);
        case 128: return launch<128>(ta, tb, ta2, p, s);
        case 64: return launch<64>(ta, tb, ta2, p, s);
        case 32: return launch<32>(ta, tb, ta2, p, s);
        default: return launch<16>(ta, tb, ta2, p, s);
    }
}

// The reference-numerics mode (see the kernel's TF32 notes): fp32 NHWC activations (pixel stride act_ld floats), fp32
// weights [c_out_rows][taps][k_per_tap] with k_per_tap = c_in rounded up to 32, fp32 bias / residual; output fp32 NHWC
// (out_mode 0, or fp16 NHWC with out_f16) or fp32 NCHW (out_mode 1).  (h_in, w_in) are the input extents.
int conv_tf32_impl(const void* act, int64_t n, int64_t h_in, int64_t w_in, int64_t c_in, int64_t act_ld, const void* wpack,
                   int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap, int stride, const float* bias, int act_fn,
                   const void* residual, int64_t res_ld, void* out, int64_t out_ld, int out_mode, int out_f16, void* stream) {
    constexpr int KE = BLOCK_K / 2;
    AZB_CHECK_PTR(act);
    AZB_CHECK_PTR(wpack);
    AZB_CHECK_PTR(out);
    if (n <= 0 || h_in <= 0 || w_in <= 0 || c_in <= 0 || c_out <= 0) return AZB_E_SHAPE;
    if ((taps != 1 && taps != 9) || (stride != 1 && stride != 2)) return AZB_E_SHAPE;
    if (act_fn < AZB_ACT_NONE || act_fn > AZB_ACT_RELU2) return AZB_E_SHAPE;
    if (k_per_tap % KE || k_per_tap < c_in) return AZB_E_SHAPE;
    if (c_in % 4 || act_ld % 4 || act_ld < c_in) return AZB_E_ALIGN;
    if (!azb_aligned(act, 16) || !azb_aligned(wpack, 16) || !azb_aligned(out, 16)) return AZB_E_ALIGN;
    if (out_mode != 0 && out_mode != 1) return AZB_E_SHAPE;
    if (out_mode == 0 && (c_out % 8 || out_ld % (out_f16 ? 8 : 4) || out_ld < c_out)) return AZB_E_ALIGN;
    if (residual && (out_mode != 0 || res_ld % 4 || res_ld < c_out || !azb_aligned(residual, 16))) return AZB_E_ALIGN;
    if (bias && !azb_aligned(bias, 16)) return AZB_E_ALIGN;
    if (out_mode == 1 && (act_fn != AZB_ACT_NONE || out_f16)) return AZB_E_UNSUPPORTED;
    const int64_t h = (h_in + stride - 1) / stride, w = (w_in + stride - 1) / stride;

    ConvParams p{};
    p.N = (int)n, p.H = (int)h, p.W = (int)w;
    patch_shape(h, w, p.BW, p.BH, p.BN);
    p.tiles_w = (int)((w + p.BW - 1) / p.BW);
    p.tiles_h = (int)((h + p.BH - 1) / p.BH);
    const int64_t m_tiles = (int64_t)p.tiles_w * p.tiles_h * ((n + p.BN - 1) / p.BN);
    int block_n = 16;
    const int sms = sm_count();
    const int cand[5] = {256, 128, 64, 32, 16};
    for (int i = 0; i < 5; ++i) {
        if (c_out_rows % cand[i]) continue;
        block_n = cand[i];
        if (m_tiles * (c_out_rows / cand[i]) >= (sms * 3) / 4 || cand[i] <= 64) break;
    }
    if (g_knob[AZB_CONV_KNOB_BLOCKN] >= 16 && c_out_rows % g_knob[AZB_CONV_KNOB_BLOCKN] == 0 && out_mode == 0) block_n = g_knob[AZB_CONV_KNOB_BLOCKN];
    if (c_out_rows % block_n || c_out_rows < c_out) return AZB_E_SHAPE;
    p.n_tiles = (int)(c_out_rows / block_n);
    if (m_tiles * p.n_tiles > 0x7fffffffLL) return AZB_E_SHAPE;
    p.tiles_out = (int)(m_tiles * p.n_tiles);
    // CTA pairs (two SMs share a 256-pixel x N tile, each stages half of the weight tile): as in the bf16 kernels, for wide
    // tiles, an even number of M tiles and reductions of at least 16 k-blocks
    const int64_t num_kb_total = (int64_t)taps * (k_per_tap / KE);
    bool pair = out_mode == 0 && block_n >= 128 && m_tiles % 2 == 0 && g_knob[AZB_CONV_KNOB_PAIR] != 0;
    if (pair && g_knob[AZB_CONV_KNOB_PAIR] < 0) pair = num_kb_total >= 16;
    p.total_tiles = pair ? p.tiles_out / 2 : p.tiles_out;
    p.phases = 1, p.splits = 1;
    p.taps = taps, p.ksize = taps == 9 ? 3 : 1, p.pad = taps == 9 ? 1 : 0;
    p.kb_per_tap = (int)(k_per_tap / KE);
    p.c_out = (int)c_out;
    p.bias = bias;
    p.res = reinterpret_cast<const __nv_bfloat16*>(residual);  // (fp32 in this mode: the EPI 3 epilogue casts it back)
    p.res_ld = res_ld;
    p.out = out, p.out_ld = out_ld, p.out_mode = out_mode, p.out_f16 = out_f16;
    p.stat_gran = 1, p.stride = stride, p.act = act_fn;
    p.prefetch_kb = g_knob[AZB_CONV_KNOB_PREFETCH] >= 0 ? g_knob[AZB_CONV_KNOB_PREFETCH] : (m_tiles <= 32 ? 24 : 0);
    p.c_in = (int)c_in;

    CUtensorMap ta, tb;
    {
        const uint32_t st = (uint32_t)stride;
        uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)act_ld * 4, (uint64_t)act_ld * 4 * w_in, (uint64_t)act_ld * 4 * w_in * h_in};
        uint32_t box[4] = {KE, (uint32_t)p.BW * st, (uint32_t)p.BH * st, (uint32_t)p.BN};
        uint32_t es[4] = {1, st, st, 1};
        if (box[1] > 256 || box[2] > 256) return AZB_E_SHAPE;
        const int rc = tc::make_map_typed(&ta, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, act, 4, dims, str, box, es);
        if (rc) return rc == -1 ? AZB_E_DRIVER : AZB_E_SHAPE;
    }
    {
        const int64_t k_total = (int64_t)taps * k_per_tap;
        uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)c_out_rows};
        uint64_t str[1] = {(uint64_t)k_total * 4};
        uint32_t box[2] = {KE, (uint32_t)(pair ? block_n / 2 : block_n)};  // a pair member stages half the rows
        const int rc = tc::make_map_typed(&tb, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, wpack, 2, dims, str, box);
        if (rc) return rc == -1 ? AZB_E_DRIVER : AZB_E_SHAPE;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (out_mode == 1) {
        switch (block_n) {
            case 16: return launch<16, false, 0, false, true>(ta, tb, ta, p, s);
            case 32: return launch<32, false, 0, false, true>(ta, tb, ta, p, s);
            default: return AZB_E_UNSUPPORTED;  // (the network's output convolution has 3 or 6 channels)
        }
    }
    if (pair) return block_n == 256 ? launch<256, true, 3, false, true>(ta, tb, ta, p, s) : launch<128, true, 3, false, true>(ta, tb, ta, p, s);
    switch (block_n) {
        case 256: return launch<256, false, 3, false, true>(ta, tb, ta, p, s);
        case 128: return launch<128, false, 3, false, true>(ta, tb, ta, p, s);
        case 64: return launch<64, false, 3, false, true>(ta, tb, ta, p, s);
        case 32: return launch<32, false, 3, false, true>(ta, tb, ta, p, s);
        default: return launch<16, false, 3, false, true>(ta, tb, ta, p, s);
    }
}

}  // namespace

extern "C" int azb_conv_tf32(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld, const void* wpack,
                             int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap, int stride, const float* bias,
                             int act_fn, const void* residual, int64_t res_ld, void* out, int64_t out_ld, int out_mode,
                             int out_f16, void* stream) {
    return conv_tf32_impl(act, n, h, w, c_in, act_ld, wpack, c_out, c_out_rows, taps, k_per_tap, stride, bias, act_fn, residual,
                          res_ld, out, out_ld, out_mode, out_f16, stream);
}

extern "C" int azb_conv_gemm_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                                  const void* wpack, int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap,
                                  const float* bias, const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                                  int out_mode, void* stream) {
    return conv_impl(act, n, h, w, c_in, act_ld, wpack, c_out, c_out_rows, taps, k_per_tap, bias, residual, res_ld, out,
                     out_ld, out_mode, nullptr, 1, stream);
}

extern "C" int azb_conv_colsum_rows(int64_t n, int64_t h, int64_t w, int64_t* rows, int64_t* rows_per_image) {
    if (n <= 0 || h <= 0 || w <= 0 || !rows) return AZB_E_SHAPE;
    int bw, bh, bn;
    patch_shape(h, w, bw, bh, bn);
    const int64_t tiles = ((w + bw - 1) / bw) * ((h + bh - 1) / bh) * ((n + bn - 1) / bn);
    *rows = tiles * 4;
    // a 32-row slab lies inside one image iff the per-image part of the patch is a multiple of 32 rows
    if (rows_per_image) *rows_per_image = ((bw * bh) % 32 == 0) ? 1 : 0;
    return AZB_OK;
}

extern "C" int azb_conv_gemm_stats_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                                        const void* wpack, int64_t c_out, int64_t c_out_rows, int taps,
                                        int64_t k_per_tap, const float* bias, const void* residual, int64_t res_ld,
                                        void* out, int64_t out_ld, float* colsum, int stat_gran, void* stream) {
    AZB_CHECK_PTR(colsum);
    return conv_impl(act, n, h, w, c_in, act_ld, wpack, c_out, c_out_rows, taps, k_per_tap, bias, residual, res_ld, out,
                     out_ld, 0, colsum, stat_gran, stream);
}

extern "C" int azb_conv2d_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                               const void* wpack, int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap,
                               int stride, const float* bias, int act_fn, const float* gate, int64_t gate_ld,
                               int64_t gate_rows, const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                               int out_mode, float* colsum, int stat_gran, void* stream) {
    ConvExtra ex;
    ex.stride = stride, ex.act = act_fn, ex.gate = gate, ex.gate_ld = gate_ld, ex.gate_rows = gate_rows;
    return conv_impl(act, n, h, w, c_in, act_ld, wpack, c_out, c_out_rows, taps, k_per_tap, bias, residual, res_ld, out,
                     out_ld, out_mode, colsum, colsum ? stat_gran : 1, stream, ex);
}

extern "C" int azb_conv_skip_stats_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                                        const void* act2, int64_t c_in2, int64_t act2_ld, const void* wpack,
                                        int64_t c_out, int64_t c_out_rows, int64_t k_per_tap, int64_t k2,
                                        const float* bias, void* out, int64_t out_ld, float* colsum, int stat_gran,
                                        void* stream) {
    AZB_CHECK_PTR(act2);
    ConvExtra ex;
    ex.act2 = act2, ex.c_in2 = c_in2, ex.act2_ld = act2_ld, ex.k2 = k2;
    return conv_impl(act, n, h, w, c_in, act_ld, wpack, c_out, c_out_rows, 9, k_per_tap, bias, nullptr, 0, out, out_ld, 0,
                     colsum, colsum ? stat_gran : 1, stream, ex);
}

extern "C" int azb_conv_bf16(const AzbConv* d, void* stream) {
    AZB_CHECK_PTR(d);
    ConvExtra ex;
    ex.stride = d->stride ? d->stride : 1, ex.act = d->act_fn;
    ex.gate = d->gate, ex.gate_ld = d->gate_ld, ex.gate_rows = d->gate_rows;
    ex.act2 = d->act2, ex.c_in2 = d->c_in2, ex.act2_ld = d->act2_ld, ex.k2 = d->k2;
    ex.gn_acc = d->gn_acc;
    ex.workspace = d->workspace, ex.workspace_bytes = d->workspace_bytes;
    ex.in_coef = d->in_coef, ex.in_silu = d->in_silu;
    ex.res_up = d->res_up, ex.in_up = d->in_up;
    ex.in_norm = d->in_norm, ex.in_eps = d->in_eps, ex.in_rowstat = d->in_rowstat, ex.in_mod = d->in_mod, ex.in_mod_ld = d->in_mod_ld;
    ex.rowstat = d->rowstat, ex.out_up = d->out_up;
    return conv_impl(d->act, d->n, d->h, d->w, d->c_in, d->act_ld, d->wpack, d->c_out, d->c_out_rows, d->taps, d->k_per_tap,
                     d->bias, d->residual, d->res_ld, d->out, d->out_ld, d->out_mode, d->colsum,
                     (d->colsum || d->gn_acc) ? d->stat_gran : 1, stream, ex);
}

extern "C" int azb_conv_choice(const AzbConv* d, AzbConvChoice* choice) {
    AZB_CHECK_PTR(d);
    AZB_CHECK_PTR(choice);
    ConvExtra ex;
    ex.stride = d->stride ? d->stride : 1, ex.act = d->act_fn;
    ex.gate = d->gate, ex.gate_ld = d->gate_ld, ex.gate_rows = d->gate_rows;
    ex.act2 = d->act2, ex.c_in2 = d->c_in2, ex.act2_ld = d->act2_ld, ex.k2 = d->k2;
    ex.gn_acc = d->gn_acc;
    ex.workspace = d->workspace, ex.workspace_bytes = d->workspace_bytes;
    ex.in_coef = d->in_coef, ex.in_silu = d->in_silu;
    ex.res_up = d->res_up, ex.in_up = d->in_up;
    ex.in_norm = d->in_norm, ex.in_eps = d->in_eps, ex.in_rowstat = d->in_rowstat, ex.in_mod = d->in_mod, ex.in_mod_ld = d->in_mod_ld;
    ex.rowstat = d->rowstat, ex.out_up = d->out_up;
    ex.choice = choice;
    return conv_impl(d->act, d->n, d->h, d->w, d->c_in, d->act_ld, d->wpack, d->c_out, d->c_out_rows, d->taps, d->k_per_tap,
                     d->bias, d->residual, d->res_ld, d->out, d->out_ld, d->out_mode, d->colsum,
                     (d->colsum || d->gn_acc) ? d->stat_gran : 1, nullptr, ex);
}
