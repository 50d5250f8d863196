// GroupNorm statistics and the fused normalise / modulate / SiLU / resample pass (sm_100a).
//
// Replaces, per normalisation site of the ADM UNet (azula/plugins/adm/_src/unet.py:177-181,
// 203-207,236-243,276,599-601 via _src/nn.py:80-87): native_group_norm + SiLU + (1+scale)*h+shift
// + nearest-upsample / 2x2 average pool -- four to six eager HBM round trips -- by ONE reduction
// pass (read x) and ONE apply pass (read x, write the conv input).  HBM-bound: 16-byte vector
// loads along the contiguous channel dimension of NHWC bf16, fp32 accumulation, deterministic
// (atomic-free) reductions.

#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int MAX_C = 4096;

struct StatsParams {
    const __nv_bfloat16* x;  // (N, HW, C) with pixel stride ld
    int64_t ld;
    int hw, c, groups, chunks, pix_per_chunk;
    float eps;
    float* partial;   // (N, chunks, groups, 2) sum / sum of squares
    float* stats;     // (N, groups, 2) mean / rstd
    int* counters;    // (N) zero-initialised; self-resetting
};

// One CTA = one (image, pixel chunk).  Thread (r, v): vector column v (8 channels), pixel rows
// r, r+R, ...  Per-channel partials go through shared memory and are reduced in a fixed order.
__global__ void __launch_bounds__(THREADS) gn_stats_kernel(const StatsParams p) {
    extern __shared__ float sm[];  // [R][C] sums, [R][C] squares, then [C] x2 channel totals
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int V = p.c >> 3;                       // vector columns
    const int R = THREADS / V > 0 ? THREADS / V : 1;
    const int v = threadIdx.x % V, r = threadIdx.x / V;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;

    const int p0 = chunk * p.pix_per_chunk;
    const int p1 = min(p0 + p.pix_per_chunk, p.hw);
    if (r < R) {
        const __nv_bfloat16* base = p.x + ((int64_t)n * p.hw) * p.ld + v * 8;
        for (int pix = p0 + r; pix < p1; pix += 4 * R) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                u[k] = pix + k * R < p1 ? __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(pix + k * R) * p.ld))
                                        : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = bf16_bits_to_f32(w[j] & 0xffffu), b = bf16_bits_to_f32(w[j] >> 16);
                    s[2 * j] += a, q[2 * j] += a * a;
                    s[2 * j + 1] += b, q[2 * j + 1] += b * b;
                }
            }
        }
    }
    float* ssum = sm;
    float* ssq = sm + R * p.c;
    float* csum = sm + 2 * R * p.c;
    float* csq = csum + p.c;
    if (r < R) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ssum[r * p.c + v * 8 + j] = s[j];
            ssq[r * p.c + v * 8 + j] = q[j];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.c; c += THREADS) {
        float a = 0.f, b = 0.f;
        for (int rr = 0; rr < R; ++rr) a += ssum[rr * p.c + c], b += ssq[rr * p.c + c];
        csum[c] = a, csq[c] = b;
    }
    __syncthreads();
    const int cg = p.c / p.groups;
    if (threadIdx.x < p.groups) {
        const int g = threadIdx.x;
        float a = 0.f, b = 0.f;
        for (int c = g * cg; c < (g + 1) * cg; ++c) a += csum[c], b += csq[c];
        float* dst = p.partial + (((int64_t)n * p.chunks + chunk) * p.groups + g) * 2;
        dst[0] = a, dst[1] = b;
    }
    // last CTA of this image folds the chunk partials (in chunk order, double precision)
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(p.counters + n, 1);
        is_last = (done == p.chunks - 1);
        if (is_last) p.counters[n] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order two-level fold: SL slices of the chunk range per group, then the slices in order
    const int SL = THREADS / p.groups;
    double* fold = reinterpret_cast<double*>(sm);  // [SL][groups][2], reuses the (finished) staging area
    {
        const int g = threadIdx.x % p.groups, sl = threadIdx.x / p.groups;
        if (sl < SL) {
            double a = 0.0, b = 0.0;
            for (int ch = sl; ch < p.chunks; ch += SL) {
                const float* src = p.partial + (((int64_t)n * p.chunks + ch) * p.groups + g) * 2;
                a += (double)__ldcg(src), b += (double)__ldcg(src + 1);
            }
            fold[(sl * p.groups + g) * 2] = a, fold[(sl * p.groups + g) * 2 + 1] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < p.groups) {
        const int g = threadIdx.x;
        double a = 0.0, b = 0.0;
        for (int sl = 0; sl < SL; ++sl) a += fold[(sl * p.groups + g) * 2], b += fold[(sl * p.groups + g) * 2 + 1];
        const double cnt = (double)p.hw * cg;
        const double mean = a / cnt;
        double var = b / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        p.stats[((int64_t)n * p.groups + g) * 2 + 0] = (float)mean;
        p.stats[((int64_t)n * p.groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
}

struct ApplyParams {
    const __nv_bfloat16* x;  // (N, H, W, C), pixel stride x_ld
    int64_t x_ld;
    __nv_bfloat16* y;        // (N, Ho, Wo, C), pixel stride y_ld
    int64_t y_ld;
    __nv_bfloat16* y2;       // DUAL (mode 2): the 2x2 average of the RAW input, same shape as y (the skip branch's x_upd)
    int64_t y2_ld;
    int n, h, w, c, groups;
    const float* stats;      // (N, groups, 2) or null (identity transform)
    const float* gamma;      // (C)
    const float* beta;       // (C)
    const float* scale_shift;  // (rows, 2C): [scale | shift], or null
    int64_t ss_stride;         // row stride between images (0 = broadcast one row)
    const int32_t* ss_step;    // optional device step index selecting a row block of `ss_step_stride`
    int64_t ss_step_stride;
    int silu;
    int pix_per_cta;         // output pixels per CTA
    const long long* acc[2];  // exact fixed-point sums (azb_conv_bf16 gn_acc) of channel ranges [0, c_a), [c_a, c): (N, c_x / gran, 4)
    int acc_c[2];
    int acc_gran;
    float eps;
    double acc_scale;  // 2^-40 / (h * w * channels per group)
};

// With SILU the caller passes HALVED coefficients, h = (A x + B) / 2, and SiLU(2h) = h + h tanh(h): one FMA, one
// SFU operation (tanh.approx, absolute error ~2^-11 on tanh, i.e. <= |h| 2^-11 on the result: a tenth of the bf16
// rounding that follows) and one FMA per element.  The exp + reciprocal form (2 SFU operations and ~10 issue slots
// per element) kept this HBM-bound pass at 65 % of the DRAM rate (ncu: issue 69 %, XU 48 %).
template <bool SILU, bool ADD = true>
__device__ __forceinline__ void affine8(const uint4 u, const float (&a)[8], const float (&b)[8], float (&acc)[8]) {
    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float f0 = fmaf(a[2 * j], bf16_bits_to_f32(wv[j] & 0xffffu), b[2 * j]);
        float f1 = fmaf(a[2 * j + 1], __uint_as_float(wv[j] & 0xffff0000u), b[2 * j + 1]);
        if (SILU) f0 = fmaf(f0, tanh_approx(f0), f0), f1 = fmaf(f1, tanh_approx(f1), f1);
        if (ADD) acc[2 * j] += f0, acc[2 * j + 1] += f1;
        else acc[2 * j] = f0, acc[2 * j + 1] = f1;
    }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8], float mul) {
    __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0] * mul, f[1] * mul), t1 = __floats2bfloat162_rn(f[2] * mul, f[3] * mul);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4] * mul, f[5] * mul), t3 = __floats2bfloat162_rn(f[6] * mul, f[7] * mul);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&t0), o.y = *reinterpret_cast<uint32_t*>(&t1);
    o.z = *reinterpret_cast<uint32_t*>(&t2), o.w = *reinterpret_cast<uint32_t*>(&t3);
    return o;
}

// (mean, rstd) of group g of image n: copied from `stats`, or folded from the producers' exact accumulators (integer
// fold, then a handful of double-precision operations)
__device__ __forceinline__ float2 group_stat(const ApplyParams& p, int n, int g, int cg) {
    if (p.stats)
        return make_float2(__ldg(p.stats + ((int64_t)n * p.groups + g) * 2), __ldg(p.stats + ((int64_t)n * p.groups + g) * 2 + 1));
    long long s_hi = 0, s_lo = 0, q_hi = 0, q_lo = 0;
    for (int ch = g * cg; ch < (g + 1) * cg; ch += p.acc_gran) {
        const int which = ch >= p.acc_c[0];
        const int local = which ? ch - p.acc_c[0] : ch;
        const long long* src = p.acc[which] + ((int64_t)n * (p.acc_c[which] / p.acc_gran) + local / p.acc_gran) * 4;
        s_hi += __ldg(src), s_lo += __ldg(src + 1), q_hi += __ldg(src + 2), q_lo += __ldg(src + 3);
    }
    // double precision only where the cancellation is (E[x^2] - mean^2); the slow DP pipe never sees a
    // division or a square root (acc_scale = 2^-40 / count comes from the host)
    const double m = ((double)s_hi * 4294967296.0 + (double)s_lo) * p.acc_scale;
    const double var = fma(-m, m, ((double)q_hi * 4294967296.0 + (double)q_lo) * p.acc_scale);
    return make_float2((float)m, rsqrtf(fmaxf((float)var, 0.f) + p.eps));
}

// {A, B} of channel c: y = act(A x + B) with A = rstd gamma (1 + scale), B = (beta - mean rstd gamma)(1 + scale) + shift;
// halved when the activation is SiLU (evaluated as h + h tanh(h), h = (A x + B) / 2)
__device__ __forceinline__ float2 channel_coef(const ApplyParams& p, const float2* s_stat, const float* ss, int c, int cg,
                                               bool silu, int g_first = 0) {
    float aa = 1.f, bb = 0.f;
    if (p.stats || p.acc[0]) {
        const float2 st = s_stat[c / cg - g_first];
        aa = __fmul_rn(st.y, __ldg(p.gamma + c));
        bb = __fsub_rn(__ldg(p.beta + c), __fmul_rn(st.x, aa));
    }
    if (ss) {
        const float sc = __fadd_rn(1.0f, __ldg(ss + c));
        aa = __fmul_rn(aa, sc);
        bb = __fadd_rn(__fmul_rn(bb, sc), __ldg(ss + p.c + c));
    }
    return silu ? make_float2(0.5f * aa, 0.5f * bb) : make_float2(aa, bb);
}

// coef[n][c] = {A, B} for the convolution kernels that apply the normalisation to their input on the fly
// (AzbConv::in_coef): one CTA per (256 channels, image); the first thread of each group folds its statistics.
__global__ void __launch_bounds__(THREADS) gn_coef_kernel(const ApplyParams p, float2* coef) {
    pdl_enter();
    const int n = blockIdx.y;
    const int cg = p.c / p.groups;
    const int c = blockIdx.x * THREADS + threadIdx.x;
    const float* ss = p.scale_shift ? p.scale_shift + (int64_t)n * p.ss_stride : nullptr;
    __shared__ float2 s_stat[THREADS];  // indexed by group - first group of this CTA
    const int g0 = (blockIdx.x * THREADS) / cg;
    if (c < p.c && (c % cg == 0 || threadIdx.x == 0)) s_stat[c / cg - g0] = group_stat(p, n, c / cg, cg);
    __syncthreads();
    if (c < p.c) coef[(int64_t)n * p.c + c] = channel_coef(p, s_stat, ss, c, cg, p.silu != 0, g0);
}

// y = act(A[c]*x + B[c]) with A = rstd*gamma*(1+scale), B = (beta - mean*rstd*gamma)*(1+scale) + shift.
// Thread (r, v) owns vector column v (8 channels, its A/B live in registers) and walks output pixels
// r, r+R, ... of the CTA's pixel range with UNROLL independent 16-byte loads in flight.
// MODE: 0 same size, 1 nearest x2 upsample, 2 2x2 average pool (of the activated values).
template <int MODE, bool SILU, bool DUAL = false>
__global__ void __launch_bounds__(THREADS) gn_apply_kernel(const ApplyParams p) {
    pdl_enter();
    const int n = blockIdx.y;
    const int V = p.c >> 3;
    const int VT = V < THREADS ? V : THREADS;  // vector columns handled per pass
    const int R = THREADS / VT;
    const int r = threadIdx.x / VT;
    const int wo = MODE == 1 ? p.w * 2 : MODE == 2 ? p.w / 2 : p.w;
    const int ho = MODE == 1 ? p.h * 2 : MODE == 2 ? p.h / 2 : p.h;
    const int out_pix = ho * wo;
    const int q0 = blockIdx.x * p.pix_per_cta;
    const int q1 = min(q0 + p.pix_per_cta, out_pix);
    const __nv_bfloat16* xin = p.x + (int64_t)n * p.h * p.w * p.x_ld;
    __nv_bfloat16* yout = p.y + (int64_t)n * out_pix * p.y_ld;
    __nv_bfloat16* yout2 = DUAL ? p.y2 + (int64_t)n * out_pix * p.y2_ld : nullptr;
    const int cg = p.c / p.groups;
    const float* ss = nullptr;
    if (p.scale_shift) {
        ss = p.scale_shift + (int64_t)n * p.ss_stride;
        if (p.ss_step) ss += (int64_t)(*p.ss_step) * p.ss_step_stride;
    }
    // (mean, rstd) of every group of image n, once per CTA: copied from `stats`, or folded from the producers' exact
    // accumulators (integer fold, then a handful of double-precision operations by <= 256 threads)
    __shared__ float2 s_stat[THREADS];
    if ((p.stats || p.acc[0]) && threadIdx.x < p.groups) s_stat[threadIdx.x] = group_stat(p, n, threadIdx.x, cg);
    __syncthreads();
    if (r >= R) return;

    for (int v = threadIdx.x % VT; v < V; v += VT) {
        float a[8], b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = v * 8 + j;
            const float2 ab = channel_coef(p, s_stat, ss, c, cg, SILU);
            a[j] = ab.x, b[j] = ab.y;
        }
        const __nv_bfloat16* xc = xin + v * 8;
        __nv_bfloat16* yc = yout + v * 8;
        constexpr int UNROLL = MODE == 2 ? 2 : 4;
        for (int q = q0 + r; q < q1; q += UNROLL * R) {
            if (MODE == 2) {
                uint4 u[UNROLL][4];
#pragma unroll
                for (int k = 0; k < UNROLL; ++k) {
                    const int qq = q + k * R;
                    if (qq < q1) {
                        const int oh = qq / wo, ow = qq - oh * wo;
                        const __nv_bfloat16* src = xc + ((int64_t)(2 * oh) * p.w + 2 * ow) * p.x_ld;
                        u[k][0] = ldg_stream16(src);
                        u[k][1] = ldg_stream16(src + p.x_ld);
                        u[k][2] = ldg_stream16(src + (int64_t)p.w * p.x_ld);
                        u[k][3] = ldg_stream16(src + (int64_t)(p.w + 1) * p.x_ld);
                    }
                }
#pragma unroll
                for (int k = 0; k < UNROLL; ++k) {
                    const int qq = q + k * R;
                    if (qq < q1) {
                        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int t = 0; t < 4; ++t) affine8<SILU>(u[k][t], a, b, acc);
                        *reinterpret_cast<uint4*>(yc + (int64_t)qq * p.y_ld) = pack8(acc, 0.25f);
                        if (DUAL) {  // the same 2x2 window of the raw values (what the identity transform would produce)
                            float raw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t wv[4] = {u[k][t].x, u[k][t].y, u[k][t].z, u[k][t].w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    raw[2 * j] += bf16_bits_to_f32(wv[j] & 0xffffu), raw[2 * j + 1] += __uint_as_float(wv[j] & 0xffff0000u);
                            }
                            *reinterpret_cast<uint4*>(yout2 + v * 8 + (int64_t)qq * p.y2_ld) = pack8(raw, 0.25f);
                        }
                    }
                }
            } else {
                uint4 u[UNROLL];
#pragma unroll
                for (int k = 0; k < UNROLL; ++k) {
                    const int qq = q + k * R;
                    if (qq < q1) {
                        int src = qq;
                        if (MODE == 1) {
                            const int oh = qq / wo, ow = qq - oh * wo;
                            src = (oh >> 1) * p.w + (ow >> 1);
                        }
                        u[k] = MODE == 1 ? __ldg(reinterpret_cast<const uint4*>(xc + (int64_t)src * p.x_ld))
                                         : ldg_stream16(xc + (int64_t)src * p.x_ld);
                    }
                }
#pragma unroll
                for (int k = 0; k < UNROLL; ++k) {
                    const int qq = q + k * R;
                    if (qq < q1) {
                        float acc[8];
                        affine8<SILU, false>(u[k], a, b, acc);
                        *reinterpret_cast<uint4*>(yc + (int64_t)qq * p.y_ld) = pack8(acc, 1.0f);
                    }
                }
            }
        }
    }
}

template <int MODE>
void launch_apply(const ApplyParams& p, dim3 grid, cudaStream_t s) {
    if (MODE == 2 && p.y2) {
        if (p.silu) azb_launch(gn_apply_kernel<2, true, true>, grid, dim3(THREADS), 0, s, p);
        else azb_launch(gn_apply_kernel<2, false, true>, grid, dim3(THREADS), 0, s, p);
    } else if (p.silu)
        azb_launch(gn_apply_kernel<MODE, true>, grid, dim3(THREADS), 0, s, p);
    else
        azb_launch(gn_apply_kernel<MODE, false>, grid, dim3(THREADS), 0, s, p);
}

struct FinalizeParams {
    const float2* src[2];  // per-(32-row slab, channel block) {sum, sum of squares} of up to two channel ranges
    int c[2];              // channels of each range
    int gran[2];           // channels per entry of each range (1 or 8)
    int n, hw, groups;
    int tiles_per_image;   // M tiles that hold rows of one image
    int images_per_tile;   // BN
    int slabs_per_image;   // 32-row slabs of one tile that belong to one image
    float eps;
    float* stats;          // (N, groups, 2) mean / rstd
};

// One CTA per (image, group): folds the column sums written by the convolution epilogues in a fixed
// order (thread-strided partials in double, then a shared-memory tree), so results are deterministic.
__global__ void __launch_bounds__(THREADS) gn_finalize_kernel(const FinalizeParams p) {
    __shared__ double red[2][THREADS];
    const int g = blockIdx.x, n = blockIdx.y;
    const int ctot = p.c[0] + p.c[1];
    const int cg = ctot / p.groups;
    const int tn = n / p.images_per_tile, bn = n - tn * p.images_per_tile;
    const int slabs = p.tiles_per_image * p.slabs_per_image;
    double a = 0.0, b = 0.0;
    // the group's channels [g*cg, (g+1)*cg) split into the part inside range 0 and the part inside range 1
    for (int which = 0; which < 2; ++which) {
        const int lo = max(g * cg, which ? p.c[0] : 0) - (which ? p.c[0] : 0);
        const int hi = min((g + 1) * cg, which ? ctot : p.c[0]) - (which ? p.c[0] : 0);
        if (hi <= lo) continue;
        const int gr = p.gran[which];
        const int e0 = lo / gr, ne = (hi - lo) / gr;  // entries (host guarantees alignment)
        const int epr = p.c[which] / gr;              // entries per slab row
        for (int item = threadIdx.x; item < slabs * ne; item += THREADS) {
            const int sl = item / ne, en = e0 + (item - sl * ne);
            const int t = sl / p.slabs_per_image, q = sl - t * p.slabs_per_image;
            const int64_t row = ((int64_t)tn * p.tiles_per_image + t) * 4 + bn * p.slabs_per_image + q;
            const float2 v = __ldcg(p.src[which] + row * epr + en);
            a += (double)v.x, b += (double)v.y;
        }
    }
    red[0][threadIdx.x] = a, red[1][threadIdx.x] = b;
    __syncthreads();
    for (int off = THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            red[0][threadIdx.x] += red[0][threadIdx.x + off];
            red[1][threadIdx.x] += red[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double cnt = (double)p.hw * cg;
        const double mean = red[0][0] / cnt;
        double var = red[1][0] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        p.stats[((int64_t)n * p.groups + g) * 2 + 0] = (float)mean;
        p.stats[((int64_t)n * p.groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
}

}  // namespace

extern "C" int azb_gn_stats_workspace(int64_t n, int64_t hw, int64_t c, int64_t groups, int64_t* partial_floats) {
    if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || !partial_floats) return AZB_E_SHAPE;
    int64_t chunks = (hw + 63) / 64;
    if (chunks > 1024) chunks = 1024;
    *partial_floats = n * chunks * groups * 2;
    return AZB_OK;
}

extern "C" int azb_gn_stats_bf16(const void* x, int64_t ld, int64_t n, int64_t hw, int64_t c, int64_t groups, float eps,
                                 float* partial, float* stats, int32_t* counters, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(partial);
    AZB_CHECK_PTR(stats);
    AZB_CHECK_PTR(counters);
    if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || c % groups || c % 8 || c > MAX_C) return AZB_E_SHAPE;
    if (groups > THREADS || (c >> 3) > THREADS) return AZB_E_SHAPE;
    if (ld % 8 || ld < c || !azb_aligned(x, 16)) return AZB_E_ALIGN;
    StatsParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.ld = ld, p.hw = (int)hw, p.c = (int)c, p.groups = (int)groups, p.eps = eps;
    int64_t chunks = (hw + 63) / 64;
    if (chunks > 1024) chunks = 1024;
    p.chunks = (int)chunks;
    p.pix_per_chunk = (int)((hw + chunks - 1) / chunks);
    p.partial = partial, p.stats = stats, p.counters = counters;
    const int V = (int)(c >> 3);
    const int R = THREADS / V > 0 ? THREADS / V : 1;
    size_t smem = (size_t)(2 * R * c + 2 * c) * sizeof(float);
    const size_t fold = (size_t)(THREADS / groups) * groups * 2 * sizeof(double);
    if (smem < fold) smem = fold;
    static AzbPerDevice<bool> configured_dev;
    bool& configured = configured_dev.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (smem > 64 * 1024) return AZB_E_SHAPE;
    gn_stats_kernel<<<dim3((unsigned)chunks, (unsigned)n), THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return azb_launch_status();
}

static int gn_apply_impl(const void* x, int64_t x_ld, void* y, int64_t y_ld, void* y2, int64_t y2_ld, int64_t n, int64_t h,
                         int64_t w, int64_t c,
                         int64_t groups, const float* stats, const int64_t* acc, int64_t c_a, const int64_t* acc_b,
                         int64_t c_b, int64_t gran, float eps, const float* gamma,
                         const float* beta, const float* scale_shift, int64_t ss_stride, const int32_t* ss_step,
                         int64_t ss_step_stride, int silu, int mode, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(y);
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 || c > MAX_C) return AZB_E_SHAPE;
    if ((stats || acc) && (!gamma || !beta || groups <= 0 || c % groups)) return AZB_E_NULL;
    if ((stats || acc) && groups > THREADS) return AZB_E_SHAPE;
    if (mode < 0 || mode > 2 || (mode == 2 && ((h | w) & 1))) return AZB_E_SHAPE;
    if (x_ld % 8 || y_ld % 8 || x_ld < c || y_ld < c || !azb_aligned(x, 16) || !azb_aligned(y, 16)) return AZB_E_ALIGN;
    ApplyParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x), p.x_ld = x_ld;
    p.y = reinterpret_cast<__nv_bfloat16*>(y), p.y_ld = y_ld;
    if (y2 && (mode != 2 || y2_ld % 8 || y2_ld < c || !azb_aligned(y2, 16))) return AZB_E_SHAPE;
    p.y2 = reinterpret_cast<__nv_bfloat16*>(y2), p.y2_ld = y2_ld;
    p.n = (int)n, p.h = (int)h, p.w = (int)w, p.c = (int)c, p.groups = (stats || acc) ? (int)groups : 1;
    p.stats = stats, p.gamma = gamma, p.beta = beta;
    if (acc) {
        if (c_a <= 0 || c_b < 0 || c_a + c_b != c || (c_b > 0 && !acc_b) || (gran != 1 && gran != 8)) return AZB_E_SHAPE;
        if (c_a % gran || c_b % gran || (c / groups) % gran) return AZB_E_SHAPE;
    }
    p.acc[0] = reinterpret_cast<const long long*>(acc), p.acc[1] = reinterpret_cast<const long long*>(acc_b);
    p.acc_c[0] = (int)c_a, p.acc_c[1] = (int)c_b, p.acc_gran = (int)gran, p.eps = eps;
    p.acc_scale = acc ? 1.0 / (1099511627776.0 * (double)h * (double)w * (double)(c / groups)) : 0.0;
    p.scale_shift = scale_shift, p.ss_stride = ss_stride, p.ss_step = ss_step, p.ss_step_stride = ss_step_stride;
    p.silu = silu;
    const int64_t ho = mode == 1 ? h * 2 : mode == 2 ? h / 2 : h, wo = mode == 1 ? w * 2 : mode == 2 ? w / 2 : w;
    const int64_t out_pix = ho * wo;
    if (out_pix > 0x7fffffffLL || h * w > 0x7fffffffLL) return AZB_E_SHAPE;
    // 16 vectors per thread and CTA (enough loads in flight, enough CTAs for every SM)
    const int64_t V = c >> 3;
    int64_t ppc = (16 * THREADS) / V;
    if (ppc < 1) ppc = 1;
    // keep at least ~4 CTAs per SM when the tensor is large enough
    while (ppc > 8 && ((out_pix + ppc - 1) / ppc) * n < 4 * 148) ppc >>= 1;
    // Large tensors: ONE wave of long-lived CTAs (`wave` per SM, all resident at once) instead of thousands of short
    // ones.  Every CTA pays two dependent global round trips (group statistics, then per-channel coefficients) before
    // its first data load; with 128-pixel CTAs that prologue was 10 - 15 % of a CTA's life and of the kernel's time.
    // Measured (scripts/gn_ab.py, B200): 15 - 30 % faster for 8 .. 270 MB inputs (13 vs 19 us at 32 x 32 x 512, 27 vs 32
    // at 64 x 64 x 512, 167 vs 196 for the 128 -> 256 upsample); the 0.5 - 1 GB tensors already stream at 5.7 - 5.9 TB/s
    // with short CTAs and lose 3 % to the imbalance of a single wave, and the pooling mode is indifferent.
    const bool mid = mode != 2 && (double)n * h * w * c * 2.0 <= 3.0e8;
    const int wave = azb_knob[AZB_GN_KNOB_WAVE] >= 0 ? azb_knob[AZB_GN_KNOB_WAVE] : (mid ? 4 : 0);
    if (wave > 0) {
        const int64_t per_image = (148 * (int64_t)wave) / n;
        if (per_image >= 1) {
            const int64_t VT = V < THREADS ? V : THREADS, step = 4 * (THREADS / VT);  // pixels per unrolled iteration
            int64_t big = (out_pix + per_image - 1) / per_image;
            big = ((big + step - 1) / step) * step;
            if (big > ppc) ppc = big;
        }
    }
    p.pix_per_cta = (int)ppc;
    const int64_t ctas = (out_pix + ppc - 1) / ppc;
    const dim3 grid((unsigned)ctas, (unsigned)n);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (mode == 0) launch_apply<0>(p, grid, s);
    else if (mode == 1) launch_apply<1>(p, grid, s);
    else launch_apply<2>(p, grid, s);
    return azb_launch_status();
}

extern "C" int azb_gn_apply_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t n, int64_t h, int64_t w,
                                 int64_t c, int64_t groups, const float* stats, const float* gamma, const float* beta,
                                 const float* scale_shift, int64_t ss_stride, const int32_t* ss_step,
                                 int64_t ss_step_stride, int silu, int mode, void* stream) {
    return gn_apply_impl(x, x_ld, y, y_ld, nullptr, 0, n, h, w, c, groups, stats, nullptr, 0, nullptr, 0, 1, 0.f, gamma, beta,
                         scale_shift, ss_stride, ss_step, ss_step_stride, silu, mode, stream);
}

extern "C" int azb_gn_apply_acc_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t n, int64_t h, int64_t w,
                                     int64_t c, int64_t groups, const int64_t* acc_a, int64_t c_a, const int64_t* acc_b,
                                     int64_t c_b, int64_t gran, float eps, const float* gamma, const float* beta,
                                     const float* scale_shift, int64_t ss_stride, int silu, int mode, void* stream) {
    AZB_CHECK_PTR(acc_a);
    return gn_apply_impl(x, x_ld, y, y_ld, nullptr, 0, n, h, w, c, groups, nullptr, acc_a, c_a, acc_b, c_b, gran, eps, gamma, beta,
                         scale_shift, ss_stride, nullptr, 0, silu, mode, stream);
}

extern "C" int azb_gn_pool_acc_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, void* y_raw, int64_t y_raw_ld, int64_t n,
                                    int64_t h, int64_t w, int64_t c, int64_t groups, const int64_t* acc_a, int64_t c_a,
                                    const int64_t* acc_b, int64_t c_b, int64_t gran, float eps, const float* gamma,
                                    const float* beta, const float* scale_shift, int64_t ss_stride, int silu, void* stream) {
    AZB_CHECK_PTR(acc_a);
    AZB_CHECK_PTR(y_raw);
    return gn_apply_impl(x, x_ld, y, y_ld, y_raw, y_raw_ld, n, h, w, c, groups, nullptr, acc_a, c_a, acc_b, c_b, gran, eps,
                         gamma, beta, scale_shift, ss_stride, nullptr, 0, silu, 2, stream);
}

extern "C" int azb_gn_coef_f32(int64_t n, int64_t h, int64_t w, int64_t c, int64_t groups, const int64_t* acc_a, int64_t c_a,
                               const int64_t* acc_b, int64_t c_b, int64_t gran, float eps, const float* gamma,
                               const float* beta, const float* scale_shift, int64_t ss_stride, int silu, float* coef,
                               void* stream) {
    AZB_CHECK_PTR(acc_a);
    AZB_CHECK_PTR(gamma);
    AZB_CHECK_PTR(beta);
    AZB_CHECK_PTR(coef);
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || groups <= 0 || groups > THREADS || c % groups || c > MAX_C) return AZB_E_SHAPE;
    if (c_a <= 0 || c_b < 0 || c_a + c_b != c || (c_b > 0 && !acc_b) || (gran != 1 && gran != 8)) return AZB_E_SHAPE;
    if (c_a % gran || c_b % gran || (c / groups) % gran) return AZB_E_SHAPE;
    if (!azb_aligned(coef, 16)) return AZB_E_ALIGN;
    ApplyParams p{};
    p.n = (int)n, p.h = (int)h, p.w = (int)w, p.c = (int)c, p.groups = (int)groups;
    p.gamma = gamma, p.beta = beta, p.scale_shift = scale_shift, p.ss_stride = ss_stride, p.silu = silu;
    p.acc[0] = reinterpret_cast<const long long*>(acc_a), p.acc[1] = reinterpret_cast<const long long*>(acc_b);
    p.acc_c[0] = (int)c_a, p.acc_c[1] = (int)c_b, p.acc_gran = (int)gran, p.eps = eps;
    p.acc_scale = 1.0 / (1099511627776.0 * (double)h * (double)w * (double)(c / groups));
    return azb_launch(gn_coef_kernel, dim3((unsigned)((c + THREADS - 1) / THREADS), (unsigned)n), dim3(THREADS), 0,
                      reinterpret_cast<cudaStream_t>(stream), p, reinterpret_cast<float2*>(coef));
}

extern "C" int azb_gn_finalize_f32(const float* colsum_a, int64_t c_a, int gran_a, const float* colsum_b, int64_t c_b,
                                   int gran_b, int64_t n, int64_t h, int64_t w, int64_t groups, float eps, float* stats,
                                   void* stream) {
    AZB_CHECK_PTR(colsum_a);
    AZB_CHECK_PTR(stats);
    if (n <= 0 || h <= 0 || w <= 0 || c_a <= 0 || c_b < 0 || groups <= 0 || (c_a + c_b) % groups) return AZB_E_SHAPE;
    if (c_b > 0 && !colsum_b) return AZB_E_NULL;
    if ((gran_a != 1 && gran_a != 8) || (c_b > 0 && gran_b != 1 && gran_b != 8)) return AZB_E_SHAPE;
    {
        // group boundaries must fall on entry boundaries of both ranges
        const int64_t cg = (c_a + c_b) / groups;
        if (gran_a == 8 && (cg % 8 || c_a % 8)) return AZB_E_SHAPE;
        if (c_b > 0 && gran_b == 8 && (cg % 8 || c_a % 8 || c_b % 8)) return AZB_E_SHAPE;
    }
    // geometry of the convolution's M tiles (same rule as azb_conv_gemm_*: 128 = BN x BH x BW pixels)
    int bw = 1;
    while (bw < 16 && bw < w) bw <<= 1;
    int bh = 1;
    while (bw * bh < 128 && bh < h) bh <<= 1;
    const int bn = 128 / (bw * bh);
    if ((bw * bh) % 32) return AZB_E_SHAPE;  // a 32-row slab would straddle two images
    FinalizeParams p{};
    p.src[0] = reinterpret_cast<const float2*>(colsum_a), p.src[1] = reinterpret_cast<const float2*>(colsum_b);
    p.c[0] = (int)c_a, p.c[1] = (int)c_b;
    p.gran[0] = gran_a, p.gran[1] = c_b > 0 ? gran_b : 1;
    p.n = (int)n, p.hw = (int)(h * w), p.groups = (int)groups;
    p.tiles_per_image = (int)(((w + bw - 1) / bw) * ((h + bh - 1) / bh));
    p.images_per_tile = bn;
    p.slabs_per_image = bn == 1 ? 4 : (bw * bh) / 32;
    p.eps = eps, p.stats = stats;
    gn_finalize_kernel<<<dim3((unsigned)groups, (unsigned)n), THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return azb_launch_status();
}
