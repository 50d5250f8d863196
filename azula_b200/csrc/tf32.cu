// Element-wise / reduction kernels of the REFERENCE-NUMERICS mode (fp32 NHWC activations, TF32 tensor-core
// contractions: csrc/conv_gemm.cu TF32 = true).  They restate, in fp32 and one pass each, what the reference does around
// its convolutions in eager PyTorch with fp32 modules:
//
//   azb_gn_stats_f32      GroupNorm32 statistics (azula/plugins/adm/_src/nn.py:80-87 -> native_group_norm): mean and
//                         rstd per (image, group), two-level fp32 / fp64 accumulation
//   azb_gn_apply_f32      y = act((x - mean) rstd gamma + beta [(1 + scale) . + shift]) with optional nearest 2x
//                         upsampling or 2 x 2 average pooling of the result (_src/unet.py:101-109,135-137,177-181,
//                         203-207,229-243); stats NULL = resampling only (the skip branch's x_upd)
//   azb_nchw_to_nhwc_f32  network input (N, C, H, W) -> (N, H, W, C_pad), zero padded channels
//
// All HBM-bound streaming kernels; 128-bit accesses, grid sized in multiples of the SM count.

#include "common.cuh"

namespace {

constexpr int THREADS = 256;

// CTA (g, n, s) reduces pixel range s of S of group g of image n: the group's channels are contiguous per pixel
// (cg floats), pixels ld apart.  fp32 sums of <= 64 values are folded in fp64; with S > 1 the partials meet in a
// workspace and the LAST CTA of an (image, group) to arrive folds them in a fixed order (deterministic) and resets the
// arrival counter for the next launch.
__global__ void __launch_bounds__(THREADS) gn_stats_f32_kernel(const float* __restrict__ x, int64_t ld, int64_t hw, int c,
                                                               int groups, float eps, float* __restrict__ stats,
                                                               double* __restrict__ partial, int* __restrict__ counters) {
    pdl_enter();
    const int g = blockIdx.x, n = blockIdx.y, S = gridDim.z, sp = blockIdx.z;
    const int cg = c / groups, v4 = cg >> 2;  // float4 per pixel and group
    const int64_t p_lo = hw * sp / S, p_hi = hw * (sp + 1) / S;
    const float* base = x + ((int64_t)n * hw + p_lo) * ld + (int64_t)g * cg;
    const uint32_t items = (uint32_t)((p_hi - p_lo) * v4);  // (the host keeps a CTA's share below 2^31 items)
    double s = 0.0, q = 0.0;
    for (uint32_t i0 = threadIdx.x; i0 < items; i0 += THREADS * 16u) {
        float fs = 0.f, fq = 0.f;
#pragma unroll 4
        for (int u = 0; u < 16; ++u) {
            const uint32_t i = i0 + (uint32_t)u * THREADS;
            if (i < items) {
                const uint32_t pix = i / (uint32_t)v4;  // 32-bit: a 64-bit division per 16 bytes made this pass ALU-bound
                const uint32_t j = i - pix * (uint32_t)v4;
                const float4 v = ldg_stream4(base + (int64_t)pix * ld + 4 * j);
                fs += (v.x + v.y) + (v.z + v.w);
                fq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, fq))));
            }
        }
        s += (double)fs, q += (double)fq;
    }
    __shared__ double sh[2][THREADS / 32];
    __shared__ int last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s, sh[1][threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) ts += sh[0][w], tq += sh[1][w];
        const int64_t slot = (int64_t)n * groups + g;
        last = 1;
        if (S > 1) {
            double* mine = partial + (slot * S + sp) * 2;
            mine[0] = ts, mine[1] = tq;
            __threadfence();
            last = atomicAdd(counters + slot, 1) == S - 1;
            if (last) {
                __threadfence();
                ts = tq = 0.0;
                for (int k = 0; k < S; ++k) {
                    const volatile double* pk = partial + (slot * S + k) * 2;
                    ts += pk[0], tq += pk[1];
                }
                counters[slot] = 0;
            }
        }
        if (last) {
            const double cnt = (double)hw * cg;
            const double mean = ts / cnt;
            double var = tq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            float* o = stats + slot * 2;
            o[0] = (float)mean;
            o[1] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
}

struct ApplyParams {
    const float* x;
    int64_t x_ld;
    float* y;
    int64_t y_ld;
    int n, h, w, c, groups;
    const float* stats;  // [n][groups][2] or null (resampling only)
    const float* gamma;
    const float* beta;
    const float* ss;     // [scale(c) | shift(c)] per sample (stride ss_stride; 0 = shared) or null
    int64_t ss_stride;
    int silu, mode;      // mode 0: same size, 1: nearest 2x upsampling of the result, 2: 2 x 2 average pooling of the result
};

__device__ __forceinline__ float silu_exact(float v) { return __fdividef(v, 1.0f + __expf(-v)); }  // ~2e-6 relative

// Grid (x, output row, image); a thread walks the row's (pixel, float4-of-channels) items with a stride that is a
// multiple of the channel vectors for the common widths, so its channel vector -- and with it the affine coefficients
// (statistics, gamma / beta, scale / shift: five loads) -- stays the same and is computed once.  The transform is
// applied per INPUT pixel; pooling averages the four transformed values (the reference pools act(norm(x)),
// _src/unet.py:229-233).
__global__ void __launch_bounds__(THREADS) gn_apply_f32_kernel(const ApplyParams p) {
    pdl_enter();
    const int v4 = p.c >> 2;
    const int wo = p.mode == 1 ? 2 * p.w : p.mode == 2 ? p.w / 2 : p.w;
    const int ho = p.mode == 1 ? 2 * p.h : p.mode == 2 ? p.h / 2 : p.h;
    const int oh = blockIdx.y, n = blockIdx.z;
    const int cg = p.c / max(p.groups, 1);
    const int items = wo * v4;
    const float* xin = p.x + (int64_t)n * p.h * p.w * p.x_ld;
    float* yout = p.y + ((int64_t)n * ho + oh) * wo * p.y_ld;
    int jc = -1;
    float a[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    for (int it = blockIdx.x * THREADS + threadIdx.x; it < items; it += gridDim.x * THREADS) {
        const int ow = it / v4, j = it - ow * v4;
        const int ch = j * 4;
        if (p.stats && j != jc) {
            jc = j;
            const float* st = p.stats + ((int64_t)n * p.groups + ch / cg) * 2;  // (4 consecutive channels share a group: cg % 4 == 0)
            const float mean = __ldg(st), rstd = __ldg(st + 1);
            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + ch)), be = __ldg(reinterpret_cast<const float4*>(p.beta + ch));
            const float g4[4] = {ga.x, ga.y, ga.z, ga.w}, b4[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) a[e] = rstd * g4[e], b[e] = fmaf(-mean, a[e], b4[e]);
            if (p.ss) {
                const float* ss = p.ss + (int64_t)n * p.ss_stride;
                const float4 sc = __ldg(reinterpret_cast<const float4*>(ss + ch)), sf = __ldg(reinterpret_cast<const float4*>(ss + p.c + ch));
                const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, f4[4] = {sf.x, sf.y, sf.z, sf.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) a[e] *= 1.f + s4[e], b[e] = fmaf(b[e], 1.f + s4[e], f4[e]);
            }
        }
        auto load = [&](int ih, int iw) {
            const float4 v = ld_stream4_coherent(xin + ((int64_t)ih * p.w + iw) * p.x_ld + ch);  // (y may alias x)
            float f[4] = {fmaf(a[0], v.x, b[0]), fmaf(a[1], v.y, b[1]), fmaf(a[2], v.z, b[2]), fmaf(a[3], v.w, b[3])};
            if (p.silu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) f[e] = silu_exact(f[e]);
            }
            return make_float4(f[0], f[1], f[2], f[3]);
        };
        float4 o;
        if (p.mode == 2) {
            const float4 v0 = load(2 * oh, 2 * ow), v1 = load(2 * oh, 2 * ow + 1), v2 = load(2 * oh + 1, 2 * ow), v3 = load(2 * oh + 1, 2 * ow + 1);
            o = make_float4(0.25f * ((v0.x + v1.x) + (v2.x + v3.x)), 0.25f * ((v0.y + v1.y) + (v2.y + v3.y)),
                            0.25f * ((v0.z + v1.z) + (v2.z + v3.z)), 0.25f * ((v0.w + v1.w) + (v2.w + v3.w)));
        } else if (p.mode == 1) {
            o = load(oh >> 1, ow >> 1);
        } else {
            o = load(oh, ow);
        }
        stg_stream4(yout + (int64_t)ow * p.y_ld + ch, o);
    }
}

__global__ void __launch_bounds__(THREADS) nchw_to_nhwc_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int c,
                                                                   int h, int w, int c_pad) {
    pdl_enter();
    const int64_t items = (int64_t)n * h * w;
    for (int64_t pix = (int64_t)blockIdx.x * THREADS + threadIdx.x; pix < items; pix += (int64_t)gridDim.x * THREADS) {
        const int64_t hw = (int64_t)h * w;
        const int64_t img = pix / hw, sp = pix - img * hw;
        for (int z = 0; z < c_pad; ++z) y[pix * c_pad + z] = z < c ? __ldg(x + (img * c + z) * hw + sp) : 0.f;
    }
}

int grid_of(int64_t items) {
    int64_t blocks = (items + THREADS - 1) / THREADS;
    const int64_t cap = (int64_t)azb_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int azb_gn_stats_f32(const float* x, int64_t ld, int64_t n, int64_t hw, int64_t c, int64_t groups, float eps,
                                float* stats, void* workspace, int64_t workspace_bytes, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(stats);
    if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || c % groups || (c / groups) % 4 || n > 65535) return AZB_E_SHAPE;
    if (ld % 4 || ld < c || !azb_aligned(x, 16)) return AZB_E_ALIGN;
    // pixel ranges per (image, group): enough CTAs to fill the machine on large maps, one on small ones
    int64_t S = hw / 2048;
    if (S > 16) S = 16;
    if (S < 1) S = 1;
    if ((hw / S + 1) * (c / groups / 4) > 0x7fffffffLL) return AZB_E_SHAPE;
    const int64_t slots = n * groups;
    const int64_t need = ((slots * 4 + 255) / 256) * 256 + slots * S * 16;
    if (S > 1 && (!workspace || workspace_bytes < need || !azb_aligned(workspace, 256))) S = 1;  // no scratch: one CTA per slot
    int* counters = reinterpret_cast<int*>(workspace);
    double* partial = S > 1 ? reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + ((slots * 4 + 255) / 256) * 256) : nullptr;
    return azb_launch(gn_stats_f32_kernel, dim3((unsigned)groups, (unsigned)n, (unsigned)S), dim3(THREADS), 0,
                      reinterpret_cast<cudaStream_t>(stream), x, ld, hw, (int)c, (int)groups, eps, stats, partial, counters);
}

extern "C" int azb_gn_apply_f32(const float* x, int64_t x_ld, float* y, int64_t y_ld, int64_t n, int64_t h, int64_t w, int64_t c,
                                int64_t groups, const float* stats, const float* gamma, const float* beta,
                                const float* scale_shift, int64_t ss_stride, int silu, int mode, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(y);
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 4 || mode < 0 || mode > 2 || (mode == 2 && ((h | w) & 1))) return AZB_E_SHAPE;
    if (stats && (!gamma || !beta || groups <= 0 || c % groups || (c / groups) % 4)) return AZB_E_SHAPE;
    if (x_ld % 4 || y_ld % 4 || x_ld < c || y_ld < c || !azb_aligned(x, 16) || !azb_aligned(y, 16)) return AZB_E_ALIGN;
    if ((gamma && !azb_aligned(gamma, 16)) || (beta && !azb_aligned(beta, 16)) || (scale_shift && (!azb_aligned(scale_shift, 16) || ss_stride % 4)))
        return AZB_E_ALIGN;
    ApplyParams p{x, x_ld, y, y_ld, (int)n, (int)h, (int)w, (int)c, (int)(stats ? groups : 1), stats, gamma, beta, stats ? scale_shift : nullptr,
                  ss_stride, silu, mode};
    const int64_t ho = mode == 1 ? 2 * h : mode == 2 ? h / 2 : h, wo = mode == 1 ? 2 * w : mode == 2 ? w / 2 : w;
    if (ho > 65535 || n > 65535 || wo * (c / 4) > 0x7fffffffLL) return AZB_E_SHAPE;
    // a row of items per (output row, image); up to 16 items per thread (the stride keeps the channel vector fixed for
    // the common widths, so the coefficient set-up is amortised over them)
    int64_t gx = (wo * (c / 4) + 16 * THREADS - 1) / (16 * THREADS);
    if (gx < 1) gx = 1;
    return azb_launch(gn_apply_f32_kernel, dim3((unsigned)gx, (unsigned)ho, (unsigned)n), dim3(THREADS), 0,
                      reinterpret_cast<cudaStream_t>(stream), p);
}

extern "C" int azb_nchw_to_nhwc_f32(const float* x, float* y, int64_t n, int64_t c, int64_t h, int64_t w, int64_t c_pad, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(y);
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || c_pad < c || c_pad % 4) return AZB_E_SHAPE;
    return azb_launch(nchw_to_nhwc_f32_kernel, dim3((unsigned)grid_of(n * h * w)), dim3(THREADS), 0, reinterpret_cast<cudaStream_t>(stream),
                      x, y, (int)n, (int)c, (int)h, (int)w, (int)c_pad);
}
