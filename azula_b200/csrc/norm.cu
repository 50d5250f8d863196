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
        for (int pix = p0 + r; pix < p1; pix += R) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)pix * p.ld));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16_bits_to_f32(w[j] & 0xffffu), b = bf16_bits_to_f32(w[j] >> 16);
                s[2 * j] += a, q[2 * j] += a * a;
                s[2 * j + 1] += b, q[2 * j + 1] += b * b;
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
    if (threadIdx.x < p.groups) {
        const int g = threadIdx.x;
        double a = 0.0, b = 0.0;
        for (int ch = 0; ch < p.chunks; ++ch) {
            const float* src = p.partial + (((int64_t)n * p.chunks + ch) * p.groups + g) * 2;
            a += (double)__ldcg(src), b += (double)__ldcg(src + 1);
        }
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
    int n, h, w, c, groups;
    const float* stats;      // (N, groups, 2) or null (identity transform)
    const float* gamma;      // (C)
    const float* beta;       // (C)
    const float* scale_shift;  // (rows, 2C): [scale | shift], or null
    int64_t ss_stride;         // row stride between images (0 = broadcast one row)
    const int32_t* ss_step;    // optional device step index selecting a row block of `ss_step_stride`
    int64_t ss_step_stride;
    int silu;
    int mode;                // 0 same, 1 nearest x2 up, 2 2x2 average pool
    int pix_per_cta;         // output pixels per CTA
};

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// y = act(A[c]*x + B[c]) with A = rstd*gamma*(1+scale), B = (beta - mean*rstd*gamma)*(1+scale) + shift
// folded per (image, channel) into shared memory once per CTA; then a pure streaming pass.
__global__ void __launch_bounds__(THREADS) gn_apply_kernel(const ApplyParams p) {
    extern __shared__ float sm[];  // A[C], B[C]
    float* A = sm;
    float* B = sm + p.c;
    const int n = blockIdx.y;
    const int cg = p.c / p.groups;
    const float* ss = nullptr;
    if (p.scale_shift) {
        ss = p.scale_shift + (int64_t)n * p.ss_stride;
        if (p.ss_step) ss += (int64_t)(*p.ss_step) * p.ss_step_stride;
    }
    for (int c = threadIdx.x; c < p.c; c += THREADS) {
        float a = 1.f, b = 0.f;
        if (p.stats) {
            const int g = c / cg;
            const float mean = p.stats[((int64_t)n * p.groups + g) * 2];
            const float rstd = p.stats[((int64_t)n * p.groups + g) * 2 + 1];
            a = rstd * __ldg(p.gamma + c);
            b = __ldg(p.beta + c) - mean * a;
        }
        if (ss) {
            const float sc = 1.0f + __ldg(ss + c);
            a *= sc;
            b = b * sc + __ldg(ss + p.c + c);
        }
        A[c] = a, B[c] = b;
    }
    __syncthreads();

    const int V = p.c >> 3;
    const int ho = p.mode == 1 ? p.h * 2 : p.mode == 2 ? p.h / 2 : p.h;
    const int wo = p.mode == 1 ? p.w * 2 : p.mode == 2 ? p.w / 2 : p.w;
    const int64_t out_pix = (int64_t)ho * wo;
    const int64_t q0 = (int64_t)blockIdx.x * p.pix_per_cta;
    const int64_t q1 = min(q0 + (int64_t)p.pix_per_cta, out_pix);
    const __nv_bfloat16* xin = p.x + (int64_t)n * p.h * p.w * p.x_ld;
    __nv_bfloat16* yout = p.y + (int64_t)n * out_pix * p.y_ld;

    for (int64_t item = q0 * V + threadIdx.x; item < q1 * V; item += THREADS) {
        const int64_t q = item / V;
        const int v = (int)(item - q * V);
        const int oh = (int)(q / wo), ow = (int)(q - (int64_t)oh * wo);
        float a[8], b[8], acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = A[v * 8 + j], b[j] = B[v * 8 + j], acc[j] = 0.f;
        const int taps = p.mode == 2 ? 4 : 1;
        for (int t = 0; t < taps; ++t) {
            int ih, iw;
            if (p.mode == 1) ih = oh >> 1, iw = ow >> 1;
            else if (p.mode == 2) ih = oh * 2 + (t >> 1), iw = ow * 2 + (t & 1);
            else ih = oh, iw = ow;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(xin + ((int64_t)ih * p.w + iw) * p.x_ld + v * 8));
            const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f0 = fmaf(a[2 * j], bf16_bits_to_f32(wv[j] & 0xffffu), b[2 * j]);
                float f1 = fmaf(a[2 * j + 1], bf16_bits_to_f32(wv[j] >> 16), b[2 * j + 1]);
                if (p.silu) f0 = silu_f(f0), f1 = silu_f(f1);
                acc[2 * j] += f0, acc[2 * j + 1] += f1;
            }
        }
        if (p.mode == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
        }
        uint4 o;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(acc[0], acc[1]), t1 = __floats2bfloat162_rn(acc[2], acc[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(acc[4], acc[5]), t3 = __floats2bfloat162_rn(acc[6], acc[7]);
        o.x = *reinterpret_cast<uint32_t*>(&t0), o.y = *reinterpret_cast<uint32_t*>(&t1);
        o.z = *reinterpret_cast<uint32_t*>(&t2), o.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(yout + q * p.y_ld + v * 8) = o;
    }
}

}  // namespace

extern "C" int azb_gn_stats_workspace(int64_t n, int64_t hw, int64_t c, int64_t groups, int64_t* partial_floats) {
    if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || !partial_floats) return AZB_E_SHAPE;
    int64_t chunks = (hw + 255) / 256;
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
    int64_t chunks = (hw + 255) / 256;
    if (chunks > 1024) chunks = 1024;
    p.chunks = (int)chunks;
    p.pix_per_chunk = (int)((hw + chunks - 1) / chunks);
    p.partial = partial, p.stats = stats, p.counters = counters;
    const int V = (int)(c >> 3);
    const int R = THREADS / V > 0 ? THREADS / V : 1;
    const size_t smem = (size_t)(2 * R * c + 2 * c) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (smem > 64 * 1024) return AZB_E_SHAPE;
    gn_stats_kernel<<<dim3((unsigned)chunks, (unsigned)n), THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return azb_launch_status();
}

extern "C" int azb_gn_apply_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t n, int64_t h, int64_t w,
                                 int64_t c, int64_t groups, const float* stats, const float* gamma, const float* beta,
                                 const float* scale_shift, int64_t ss_stride, const int32_t* ss_step,
                                 int64_t ss_step_stride, int silu, int mode, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(y);
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 || c > MAX_C) return AZB_E_SHAPE;
    if (stats && (!gamma || !beta || groups <= 0 || c % groups)) return AZB_E_NULL;
    if (mode < 0 || mode > 2 || (mode == 2 && ((h | w) & 1))) return AZB_E_SHAPE;
    if (x_ld % 8 || y_ld % 8 || x_ld < c || y_ld < c || !azb_aligned(x, 16) || !azb_aligned(y, 16)) return AZB_E_ALIGN;
    ApplyParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x), p.x_ld = x_ld;
    p.y = reinterpret_cast<__nv_bfloat16*>(y), p.y_ld = y_ld;
    p.n = (int)n, p.h = (int)h, p.w = (int)w, p.c = (int)c, p.groups = stats ? (int)groups : 1;
    p.stats = stats, p.gamma = gamma, p.beta = beta;
    p.scale_shift = scale_shift, p.ss_stride = ss_stride, p.ss_step = ss_step, p.ss_step_stride = ss_step_stride;
    p.silu = silu, p.mode = mode;
    const int64_t ho = mode == 1 ? h * 2 : mode == 2 ? h / 2 : h, wo = mode == 1 ? w * 2 : mode == 2 ? w / 2 : w;
    const int64_t out_pix = ho * wo;
    // ~8 vectors per thread and CTA: pixels per CTA = 8*256 / V, at least 1
    int64_t ppc = (8 * THREADS) / (c >> 3);
    if (ppc < 1) ppc = 1;
    p.pix_per_cta = (int)ppc;
    const int64_t ctas = (out_pix + ppc - 1) / ppc;
    const size_t smem = (size_t)2 * c * sizeof(float);
    gn_apply_kernel<<<dim3((unsigned)ctas, (unsigned)n), THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return azb_launch_status();
}
