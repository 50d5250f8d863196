// Row-wise kernels of the in-repo backbones (azula/nn/unet.py, azula/nn/dit.py, azula/nn/vit.py,
// azula/nn/attention.py) on sm_100a.  "Row" = one pixel of an NHWC activation or one token: C contiguous
// bf16 channels.  All of them are HBM-bound streaming passes: 16-byte vector accesses along the channel
// dimension, fp32 arithmetic, warp-shuffle reductions, no shared memory.
//
//   azb_rownorm_mod_bf16      Ada-Norm-Zero prologue: (1 + a) * LayerNorm_C(x) + b   (nn/unet.py:99-104,
//                             nn/layers.py:152-155)  or  (1 + a) * RMSNorm_C(x) + b  (nn/dit.py:102-103)
//   azb_segment_rmsnorm_bf16  per-head query/key RMS normalisation, in place (nn/attention.py:103)
//   azb_patchify_f32 / azb_unpatchify_f32   pixels <-> tokens (nn/vit.py:97,105; nn/layers.py:198-246)
//   azb_linear_gather_f32     all second Ada-Norm-Zero linears of a network in one launch

#include "common.cuh"

namespace {

constexpr int THREADS = 256;

struct RowNormParams {
    const __nv_bfloat16* x;
    int64_t x_ld;
    __nv_bfloat16* y;
    int64_t y_ld;
    int64_t rows;
    int c, nvec, lpr;  // channels, 8-channel vectors per row, lanes per row (power of two <= 32)
    int kind;          // AZB_NORM_*
    float eps;
    const float* mod;  // [a(C) | b(C) | ...] per sample, or null
    int64_t mod_ld;
    int64_t rows_per_sample;
};

__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) f[2 * j] = bf16_bits_to_f32(w[j] & 0xffffu), f[2 * j + 1] = bf16_bits_to_f32(w[j] >> 16);
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&t0), o.y = *reinterpret_cast<uint32_t*>(&t1);
    o.z = *reinterpret_cast<uint32_t*>(&t2), o.w = *reinterpret_cast<uint32_t*>(&t3);
    return o;
}

__device__ __forceinline__ float group_sum(float v, int lanes) {
    for (int off = lanes >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// A group of `lpr` lanes owns one row; lane l of the group holds vectors l, l + lpr, ... (VPL of them).
// Two passes over registers: mean, then the centred sum of squares (as torch.var_mean does).
template <int VPL>
__global__ void __launch_bounds__(THREADS, VPL <= 4 ? 4 : 2) rownorm_mod_kernel(const RowNormParams p) {
    pdl_enter();  // (programmatic dependent launch: the CTAs are resident when the producer of x retires)
    const int lane = threadIdx.x & 31;
    const int sub = lane & (p.lpr - 1);
    const int rows_per_warp = 32 / p.lpr;
    const int64_t warp = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t row_groups = (p.rows + rows_per_warp - 1) / rows_per_warp;
    for (int64_t rg = warp; rg < row_groups; rg += warps) {
        const int64_t row = rg * rows_per_warp + lane / p.lpr;
        const bool row_ok = row < p.rows;
        float f[VPL][8];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int v = sub + k * p.lpr;
            uint4 u = make_uint4(0, 0, 0, 0);
            if (row_ok && v < p.nvec) u = ldg_stream16(p.x + row * p.x_ld + v * 8);
            unpack8(u, f[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[k][j];
        }
        float mean = 0.f, rstd;
        if (p.kind == AZB_NORM_LAYER) {
            mean = group_sum(s, p.lpr) / (float)p.c;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                if (sub + k * p.lpr < p.nvec) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float d = f[k][j] - mean;
                        q = fmaf(d, d, q);
                    }
                }
            }
            q = group_sum(q, p.lpr);
            rstd = rsqrtf(q / (float)(p.c - 1) + p.eps);  // torch.var_mean: unbiased
        } else {
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
#pragma unroll
                for (int j = 0; j < 8; ++j) q = fmaf(f[k][j], f[k][j], q);
            }
            q = group_sum(q, p.lpr);
            rstd = rsqrtf(q / (float)p.c + p.eps);
        }
        if (!row_ok) continue;
        const float* mod = p.mod ? p.mod + (row / p.rows_per_sample) * p.mod_ld : nullptr;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int v = sub + k * p.lpr;
            if (v >= p.nvec) continue;
            float o[8];
            if (mod) {
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(mod + v * 8));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(mod + v * 8) + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(mod + p.c + v * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(mod + p.c + v * 8) + 1);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(a[j] + 1.0f, (f[k][j] - mean) * rstd, b[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (f[k][j] - mean) * rstd;
            }
            *reinterpret_cast<uint4*>(p.y + row * p.y_ld + v * 8) = pack8(o);
        }
    }
}

// item = (row, segment); d / 8 lanes per item
__global__ void __launch_bounds__(THREADS) segment_rmsnorm_kernel(__nv_bfloat16* x, int64_t ld, int64_t rows, int segs,
                                                                  int d, float eps) {
    const int lps = d >> 3;  // lanes per segment (power of two <= 32)
    const int per_warp = 32 / lps;
    const int lane = threadIdx.x & 31;
    const int sub = lane & (lps - 1);
    const int64_t warp = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t items = rows * segs;
    const int64_t groups = (items + per_warp - 1) / per_warp;
    for (int64_t g = warp; g < groups; g += warps) {
        const int64_t item = g * per_warp + lane / lps;
        const bool ok = item < items;
        const int64_t row = ok ? item / segs : 0;
        const int seg = ok ? (int)(item - row * segs) : 0;
        __nv_bfloat16* ptr = x + row * ld + (int64_t)seg * d + sub * 8;
        float f[8];
        unpack8(ok ? *reinterpret_cast<const uint4*>(ptr) : make_uint4(0, 0, 0, 0), f);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) q = fmaf(f[j], f[j], q);
        q = group_sum(q, lps);
        const float rstd = rsqrtf(q / (float)d + eps);
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= rstd;
        *reinterpret_cast<uint4*>(ptr) = pack8(f);
    }
}

// The same item decomposition, with the rotary positional embedding applied after the (optional) normalisation:
// channel pair (2 i, 2 i + 1) of head h at token l is rotated by the angle whose {cos, sin} sits at
// rot[(l mod rows_per_sample) * (heads * d / 2) + h * d / 2 + i] (azula/nn/attention.py:105-108,124-156; the same
// angles for queries and keys).  Arithmetic in fp32, one rounding to bf16 at the end.
__global__ void __launch_bounds__(THREADS) qk_norm_rope_kernel(__nv_bfloat16* x, int64_t ld, int64_t rows, int heads, int d,
                                                               int norm, float eps, const float2* __restrict__ rot,
                                                               int64_t rows_per_sample) {
    const int segs = 2 * heads;
    const int lps = d >> 3;
    const int per_warp = 32 / lps;
    const int lane = threadIdx.x & 31;
    const int sub = lane & (lps - 1);
    const int64_t warp = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t items = rows * segs;
    const int64_t groups = (items + per_warp - 1) / per_warp;
    for (int64_t g = warp; g < groups; g += warps) {
        const int64_t item = g * per_warp + lane / lps;
        const bool ok = item < items;
        const int64_t row = ok ? item / segs : 0;
        const int seg = ok ? (int)(item - row * segs) : 0;
        __nv_bfloat16* ptr = x + row * ld + (int64_t)seg * d + sub * 8;
        float f[8];
        unpack8(ok ? *reinterpret_cast<const uint4*>(ptr) : make_uint4(0, 0, 0, 0), f);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) q = fmaf(f[j], f[j], q);
        q = group_sum(q, lps);
        const float rstd = norm ? rsqrtf(q / (float)d + eps) : 1.0f;
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= rstd;
        if (rot) {
            const int head = seg < heads ? seg : seg - heads;
            const float2* r = rot + (row % rows_per_sample) * (int64_t)(heads * (d >> 1)) + head * (d >> 1) + sub * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 cs = __ldg(r + j);
                const float re = f[2 * j], im = f[2 * j + 1];
                f[2 * j] = re * cs.x - im * cs.y;
                f[2 * j + 1] = re * cs.y + im * cs.x;
            }
        }
        *reinterpret_cast<uint4*>(ptr) = pack8(f);
    }
}

__global__ void __launch_bounds__(THREADS) patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ tok,
                                                           int n, int c, int hp, int wp, int p, int q, int k_pad) {
    const int vecs = k_pad >> 3;
    const int64_t total = (int64_t)n * hp * wp * vecs;
    const int H = hp * p, W = wp * q, K = c * p * q;
    for (int64_t item = (int64_t)blockIdx.x * THREADS + threadIdx.x; item < total; item += (int64_t)gridDim.x * THREADS) {
        const int v = (int)(item % vecs);
        const int64_t t = item / vecs;
        const int j = (int)(t % wp), i = (int)((t / wp) % hp), img = (int)(t / ((int64_t)wp * hp));
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = v * 8 + e;
            float val = 0.f;
            if (k < K) {
                const int b = k % q, a = (k / q) % p, z = k / (q * p);
                val = __ldg(x + (((int64_t)img * c + z) * H + i * p + a) * W + j * q + b);
            }
            f[e] = val;
        }
        *reinterpret_cast<uint4*>(tok + t * k_pad + v * 8) = pack8(f);
    }
}

__global__ void __launch_bounds__(THREADS) unpatchify_kernel(const float* __restrict__ yt, float* __restrict__ out, int n,
                                                             int c, int hp, int wp, int p, int q) {
    const int H = hp * p, W = wp * q;
    const int64_t tokens = (int64_t)n * hp * wp;
    const int64_t total = (int64_t)n * c * H * W;
    for (int64_t e = (int64_t)blockIdx.x * THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
        const int xx = (int)(e % W), yy = (int)((e / W) % H);
        const int z = (int)((e / ((int64_t)W * H)) % c), img = (int)(e / ((int64_t)W * H * c));
        const int64_t t = ((int64_t)img * hp + yy / p) * wp + xx / q;
        const int k = (z * p + yy % p) * q + xx % q;
        out[e] = __ldg(yt + (int64_t)k * tokens + t);
    }
}

// one warp per output column, rows over blockIdx.y
__global__ void __launch_bounds__(THREADS) linear_gather_kernel(const float* __restrict__ x, int64_t x_ld,
                                                                const int32_t* __restrict__ xoff,
                                                                const float* __restrict__ W, const float* __restrict__ b,
                                                                float* __restrict__ y, int M, int N, int K, int silu_in) {
    const int col = (blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (col >= N) return;
    const float* wrow = W + (int64_t)col * K;
    const int off = xoff ? __ldg(xoff + col) : 0;
    for (int m = blockIdx.y; m < M; m += gridDim.y) {
        const float* xrow = x + (int64_t)m * x_ld + off;
        float acc = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + k));
            float4 xv = __ldg(reinterpret_cast<const float4*>(xrow + k));
            if (silu_in) {
                xv.x = xv.x / (1.f + expf(-xv.x)), xv.y = xv.y / (1.f + expf(-xv.y));
                xv.z = xv.z / (1.f + expf(-xv.z)), xv.w = xv.w / (1.f + expf(-xv.w));
            }
            acc = fmaf(wv.x, xv.x, acc), acc = fmaf(wv.y, xv.y, acc);
            acc = fmaf(wv.z, xv.z, acc), acc = fmaf(wv.w, xv.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) y[(int64_t)m * N + col] = acc + (b ? __ldg(b + col) : 0.f);
    }
}

unsigned stream_grid(int64_t work_items, int per_cta) {
    int64_t blocks = (work_items + per_cta - 1) / per_cta;
    const int64_t cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int azb_rownorm_mod_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t rows, int64_t c,
                                    int kind, float eps, const float* mod, int64_t mod_ld, int64_t rows_per_sample,
                                    void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(y);
    if (rows <= 0 || c < 8 || c % 8 || c > 2048) return AZB_E_SHAPE;
    if (kind != AZB_NORM_LAYER && kind != AZB_NORM_RMS) return AZB_E_SHAPE;
    if (x_ld % 8 || y_ld % 8 || x_ld < c || y_ld < c || !azb_aligned(x, 16) || !azb_aligned(y, 16)) return AZB_E_ALIGN;
    if (mod && (rows_per_sample <= 0 || mod_ld % 4 || !azb_aligned(mod, 16))) return AZB_E_ALIGN;
    RowNormParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x), p.x_ld = x_ld;
    p.y = reinterpret_cast<__nv_bfloat16*>(y), p.y_ld = y_ld;
    p.rows = rows, p.c = (int)c, p.nvec = (int)(c >> 3);
    int lpr = 1;
    while (lpr < 32 && lpr < p.nvec) lpr <<= 1;
    p.lpr = lpr;
    p.kind = kind, p.eps = eps, p.mod = mod, p.mod_ld = mod_ld, p.rows_per_sample = mod ? rows_per_sample : 1;
    const int vpl = (p.nvec + lpr - 1) / lpr;
    const int rows_per_cta = (THREADS / 32) * (32 / lpr);
    const unsigned grid = stream_grid(rows, rows_per_cta);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (vpl) {
        case 1: azb_launch(rownorm_mod_kernel<1>, dim3(grid), dim3(THREADS), 0, s, p); break;
        case 2: azb_launch(rownorm_mod_kernel<2>, dim3(grid), dim3(THREADS), 0, s, p); break;
        case 3: azb_launch(rownorm_mod_kernel<3>, dim3(grid), dim3(THREADS), 0, s, p); break;
        case 4: azb_launch(rownorm_mod_kernel<4>, dim3(grid), dim3(THREADS), 0, s, p); break;
        case 5:
        case 6: azb_launch(rownorm_mod_kernel<6>, dim3(grid), dim3(THREADS), 0, s, p); break;
        default: azb_launch(rownorm_mod_kernel<8>, dim3(grid), dim3(THREADS), 0, s, p); break;
    }
    return azb_launch_status();
}

extern "C" int azb_segment_rmsnorm_bf16(void* x, int64_t ld, int64_t rows, int64_t segs, int64_t d, float eps,
                                        void* stream) {
    AZB_CHECK_PTR(x);
    if (rows <= 0 || segs <= 0 || (d != 8 && d != 16 && d != 32 && d != 64 && d != 128 && d != 256)) return AZB_E_SHAPE;
    if (ld % 8 || ld < segs * d || !azb_aligned(x, 16)) return AZB_E_ALIGN;
    const int per_cta = (THREADS / 32) * (32 / (int)(d >> 3));
    segment_rmsnorm_kernel<<<stream_grid(rows * segs, per_cta), THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(x), ld, rows, (int)segs, (int)d, eps);
    return azb_launch_status();
}

extern "C" int azb_qk_norm_rope_bf16(void* x, int64_t ld, int64_t rows, int64_t heads, int64_t d, int norm, float eps,
                                     const float* rot, int64_t rows_per_sample, void* stream) {
    AZB_CHECK_PTR(x);
    if (rows <= 0 || heads <= 0 || (d != 8 && d != 16 && d != 32 && d != 64 && d != 128 && d != 256)) return AZB_E_SHAPE;
    if (ld % 8 || ld < 2 * heads * d || !azb_aligned(x, 16)) return AZB_E_ALIGN;
    if (rot && (rows_per_sample <= 0 || !azb_aligned(rot, 8))) return AZB_E_SHAPE;
    const int per_cta = (THREADS / 32) * (32 / (int)(d >> 3));
    qk_norm_rope_kernel<<<stream_grid(rows * 2 * heads, per_cta), THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(x), ld, rows, (int)heads, (int)d, norm, eps, reinterpret_cast<const float2*>(rot),
        rows_per_sample > 0 ? rows_per_sample : 1);
    return azb_launch_status();
}

extern "C" int azb_patchify_f32(const float* x, void* tokens, int64_t n, int64_t c, int64_t hp, int64_t wp, int64_t p,
                                int64_t q, int64_t k_pad, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(tokens);
    if (n <= 0 || c <= 0 || hp <= 0 || wp <= 0 || p <= 0 || q <= 0 || k_pad % 8 || c * p * q > k_pad) return AZB_E_SHAPE;
    if (!azb_aligned(tokens, 16)) return AZB_E_ALIGN;
    patchify_kernel<<<stream_grid(n * hp * wp * (k_pad / 8), THREADS), THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, reinterpret_cast<__nv_bfloat16*>(tokens), (int)n, (int)c, (int)hp, (int)wp, (int)p, (int)q, (int)k_pad);
    return azb_launch_status();
}

extern "C" int azb_unpatchify_f32(const float* yt, float* out, int64_t n, int64_t c, int64_t hp, int64_t wp, int64_t p,
                                  int64_t q, void* stream) {
    AZB_CHECK_PTR(yt);
    AZB_CHECK_PTR(out);
    if (n <= 0 || c <= 0 || hp <= 0 || wp <= 0 || p <= 0 || q <= 0) return AZB_E_SHAPE;
    unpatchify_kernel<<<stream_grid(n * c * hp * p * wp * q, THREADS), THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        yt, out, (int)n, (int)c, (int)hp, (int)wp, (int)p, (int)q);
    return azb_launch_status();
}

extern "C" int azb_linear_gather_f32(const float* x, int64_t x_ld, const int32_t* xoff, const float* w, const float* b,
                                     float* y, int64_t m, int64_t n, int64_t k, int silu_in, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(w);
    AZB_CHECK_PTR(y);
    if (m <= 0 || n <= 0 || k <= 0 || k % 4 || x_ld % 4) return AZB_E_SHAPE;
    if (!azb_aligned(x, 16) || !azb_aligned(w, 16)) return AZB_E_ALIGN;
    dim3 grid((unsigned)((n + 7) / 8), (unsigned)(m < 64 ? m : 64));
    linear_gather_kernel<<<grid, THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_ld, xoff, w, b, y, (int)m, (int)n,
                                                                                      (int)k, silu_in);
    return azb_launch_status();
}
