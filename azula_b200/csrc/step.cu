// Fused reverse-diffusion transition q(X_s | X_t) for sm_100a.
//
// One launch replaces what the reference issues per step outside the backbone
// (azula/sample.py:204-216,248-261 + azula/denoise.py:322 / plugins/adm/__init__.py:125-134):
// ~86 ATen dispatches, 15 of them full-tensor passes (~132 B/element of HBM traffic), become
// 12 B/element: read x_t, read F, write x_s.  The N(0,1) draw is produced in registers with a
// Philox4x32-10 stream laid out like ATen's randn kernel, so equal seeds give equal bits.
//
// HBM-bound: 128-bit coalesced streaming loads/stores (L1 no-allocate), all loads of a work
// item issued before the arithmetic, grid sized in multiples of the SM count.

#include "common.cuh"

namespace {

struct Row {
    float c_skip, c_out, alpha_s, k, alpha_t, n, c_in_next, clip;
};

// ------------------------------------------------------------------------------- Philox

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// Box-Muller exactly as cuRAND's curand_normal4 evaluates it on the device
// (same constants, fused multiply-add for the uniforms, accurate logf/sqrtf, fast sincos).
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
    constexpr float INV32 = 2.3283064e-10f;
    constexpr float INV32_2PI = 2.3283064e-10f * 6.2831855f;
    float u = __fmaf_rn((float)x, INV32, INV32 / 2);
    float v = __fmaf_rn((float)y, INV32_2PI, INV32_2PI / 2);
    float s = sqrtf(-2.0f * logf(u));
    float sn, cs;
    __sincosf(v, &sn, &cs);
    return make_float2(sn * s, cs * s);
}

// Four normals of Philox stream `idx` at counter `ctr`: what curand_normal4 returns on the
// (ctr - offset/4)-th call after curand_init(seed, idx, offset).
__device__ __forceinline__ float4 normal4(uint64_t ctr, uint64_t idx, uint2 key) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)idx, (uint32_t)(idx >> 32)), key);
    float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

// Global element li of a randn over T threads: stream idx = li % T, call j = (li / T) / 4,
// lane ii = (li / T) % 4   (DistributionTemplates.h:65-88).
__device__ __forceinline__ float normal_at(int64_t li, int64_t T, uint64_t ctr0, uint2 key) {
    int64_t q = li / T;
    int64_t idx = li - q * T;
    float4 v = normal4(ctr0 + (uint64_t)(q >> 2), (uint64_t)idx, key);
    int ii = (int)(q & 3);
    return ii == 0 ? v.x : ii == 1 ? v.y : ii == 2 ? v.z : v.w;
}

// ------------------------------------------------------------------------- dtype helpers

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    static __device__ __forceinline__ float4 load(const void* p, int64_t i) {
        return ldg_stream4(reinterpret_cast<const float*>(p) + i);
    }
    static __device__ __forceinline__ float load1(const void* p, int64_t i) {
        return __ldg(reinterpret_cast<const float*>(p) + i);
    }
    static __device__ __forceinline__ void store(void* p, int64_t i, float4 v) {
        stg_stream4(reinterpret_cast<float*>(p) + i, v);
    }
    static __device__ __forceinline__ void store1(void* p, int64_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const void* p, int64_t i) {
        uint2 u = ldg_stream2u(reinterpret_cast<const __nv_bfloat16*>(p) + i);
        return make_float4(bf16_bits_to_f32(u.x & 0xffffu), bf16_bits_to_f32(u.x >> 16), bf16_bits_to_f32(u.y & 0xffffu),
                           bf16_bits_to_f32(u.y >> 16));
    }
    static __device__ __forceinline__ float load1(const void* p, int64_t i) {
        return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    }
    static __device__ __forceinline__ void store(void* p, int64_t i, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = u;
    }
    static __device__ __forceinline__ void store1(void* p, int64_t i, float v) {
        reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    }
};
template <>
struct Vec4<__half> {
    static __device__ __forceinline__ float4 load(const void* p, int64_t i) {
        uint2 u = ldg_stream2u(reinterpret_cast<const __half*>(p) + i);
        __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ float load1(const void* p, int64_t i) {
        return __half2float(reinterpret_cast<const __half*>(p)[i]);
    }
    static __device__ __forceinline__ void store(void* p, int64_t i, float4 v) {
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p) + i) = u;
    }
    static __device__ __forceinline__ void store1(void* p, int64_t i, float v) {
        reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    }
};
struct NoOut {};
template <>
struct Vec4<NoOut> {
    static __device__ __forceinline__ void store(void*, int64_t, float4) {}
    static __device__ __forceinline__ void store1(void*, int64_t, float) {}
};

// Coherent streaming load of the sampler state: the fused loop updates x in place (src == dst), so the
// read-only (.nc) path is not allowed for it.
__device__ __forceinline__ float4 ld_state4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ float ld_state1(const float* p) {
    float r;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}

// Extended row (include/azb.h AZB_R_*): selectors, stored-quantity / update coefficients, history weights.
struct RowX {
    float p, q, r;
    uint32_t flags;
    int32_t draw;
    float w[AZB_STEP_MAX_SLOTS];
};

struct StepArgs {
    const float* src[2];
    float* dst[2];
    const void* f;
    const void* fneg;
    const float* guidance;
    const float* eps;
    void* xin;
    float* hist;
    int64_t n_per_sample, numel, f_bstride, hist_stride;
    const float* table;
    const int32_t* step_idx;
    uint64_t seed;
    const int64_t* off_dev;
    int64_t off_host, off_inc, T, elem_off;
    int row_floats, xin_copies, noise_hint;
    unsigned long long np_magic;  // floor(2^64 / n_per_sample) + 1: exact i / n_per_sample for i < 2^32 by one multiply-high
};

__device__ __forceinline__ Row load_row_n(const StepArgs& a, RowX& rx) {
    const int32_t idx = *a.step_idx;
    const float* base = a.table + (int64_t)idx * a.row_floats;
    const float4* r = reinterpret_cast<const float4*>(base);
    float4 u = __ldg(r), v = __ldg(r + 1);
    rx.p = rx.q = rx.r = 0.f;
    rx.flags = 0u, rx.draw = 0;
#pragma unroll
    for (int j = 0; j < AZB_STEP_MAX_SLOTS; ++j) rx.w[j] = 0.f;
    if (a.row_floats >= AZB_ROW_COLS) {
        float4 c = __ldg(r + 2), d = __ldg(r + 3), w0 = __ldg(r + 4), w1 = __ldg(r + 5);
        rx.p = c.x, rx.q = c.y, rx.r = c.z;
        rx.flags = __float_as_uint(d.x), rx.draw = __float_as_int(d.y);
        rx.w[0] = w0.x, rx.w[1] = w0.y, rx.w[2] = w0.z, rx.w[3] = w0.w;
        rx.w[4] = w1.x, rx.w[5] = w1.y, rx.w[6] = w1.z, rx.w[7] = w1.w;
    }
    return Row{u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
}

// Index of element i of the state inside F when F carries more channels per sample than the state (learned
// variance: the first C of 2C): sample b = i / n_per_sample by one multiply-high when the tensor has < 2^32 elements
// (np_magic != 0), a 64-bit division otherwise.
__device__ __forceinline__ int64_t f_index(const StepArgs& a, int64_t i) {
    if (a.f_bstride == a.n_per_sample) return i;
    const int64_t b = a.np_magic ? (int64_t)__umul64hi((unsigned long long)i, a.np_magic) : i / a.n_per_sample;
    return b * a.f_bstride + (i - b * a.n_per_sample);
}

// Posterior mean of ONE element: mean = clip(c_skip*x + c_out*F) (denoise.py:322, adm/__init__.py:125-134); under
// classifier-free guidance each branch is clipped on its own and mu = m+ + w*(m+ - m-) (guidance/cfg.py:62-64),
// every operation rounded separately as eager does.
__device__ __forceinline__ float mean_of(const Row& r, float x, float f) {
    float m = __fadd_rn(__fmul_rn(r.c_skip, x), __fmul_rn(r.c_out, f));
    if (m == m) m = fminf(fmaxf(m, -r.clip), r.clip);
    return m;
}
template <bool CFG>
__device__ __forceinline__ float mean_cfg(const Row& r, float x, float f, float fn, float w) {
    float m = mean_of(r, x, f);
    if constexpr (CFG) {
        float mn = mean_of(r, x, fn);
        m = __fadd_rn(m, __fmul_rn(w, __fsub_rn(m, mn)));
    }
    return m;
}
//   x_s  = alpha_s*mean; x_s += k*(x - alpha_t*mean); x_s += n*eps        sample.py:212-214,257-259
__device__ __forceinline__ float affine_of(const Row& r, float x, float m, float eps) {
    float xs = __fmul_rn(r.alpha_s, m);
    xs = __fadd_rn(xs, __fmul_rn(r.k, __fsub_rn(x, __fmul_rn(r.alpha_t, m))));
    xs = __fadd_rn(xs, __fmul_rn(r.n, eps));
    return xs;
}

template <typename IT>
__device__ __forceinline__ void store_out4(const StepArgs& a, const Row& r, float* out, int64_t i, float4 o) {
    stg_stream4(out + i, o);
    float4 s = make_float4(__fmul_rn(r.c_in_next, o.x), __fmul_rn(r.c_in_next, o.y), __fmul_rn(r.c_in_next, o.z),
                           __fmul_rn(r.c_in_next, o.w));
    Vec4<IT>::store(a.xin, i, s);
    if (a.xin_copies > 1) Vec4<IT>::store(a.xin, i + a.numel, s);
}

template <typename FT, typename IT, bool CFG>
__device__ __forceinline__ void finish4(const StepArgs& a, const Row& r, float* out, float w, int64_t i, float4 x, float4 f,
                                        float4 fn, float4 e) {
    float4 o;
    o.x = affine_of(r, x.x, mean_cfg<CFG>(r, x.x, f.x, fn.x, w), e.x);
    o.y = affine_of(r, x.y, mean_cfg<CFG>(r, x.y, f.y, fn.y, w), e.y);
    o.z = affine_of(r, x.z, mean_cfg<CFG>(r, x.z, f.z, fn.z, w), e.z);
    o.w = affine_of(r, x.w, mean_cfg<CFG>(r, x.w, f.w, fn.w, w), e.w);
    store_out4<IT>(a, r, out, i, o);
}

#define AZB_AXPY4(acc, s, v) \
    (acc).x = fmaf((s), (v).x, (acc).x), (acc).y = fmaf((s), (v).y, (acc).y), (acc).z = fmaf((s), (v).z, (acc).z), \
    (acc).w = fmaf((s), (v).w, (acc).w)

// Vector path: numel, n_per_sample, f_bstride, elem_off multiples of 4; 16-byte aligned pointers.
// NOISE = false: the caller guarantees that no row draws noise (AzbStep::noise_hint < 0) -- the Philox / Box-Muller code
// is not compiled in, which halves the register count and lets the streaming path run at the HBM rate; a row that asks
// for noise nevertheless poisons the output with NaN instead of silently dropping the term.
template <typename FT, typename IT, bool CFG, bool NOISE>
__global__ void __launch_bounds__(256) step_vec4_kernel(const StepArgs a) {
    pdl_enter();
    RowX rx;
    const Row r = load_row_n(a, rx);
    // (the grid may be two-dimensional, see the noise path; the other paths use it as a flat one)
    const int64_t tid = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * blockDim.x;
    const float* xe_p = a.src[rx.flags & 1u];
    const float* xb_p = a.src[(rx.flags >> 1) & 1u];
    float* out_p = a.dst[(rx.flags >> 2) & 1u];
    const float gw = CFG ? __ldg(a.guidance) : 0.f;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    if (rx.flags & 8u) {
        // History mode (Heun, Adams-Bashforth family): h = p*x_e + q*m,  x_out = r*x_b + sum_j W[j]*H[j].
        const int wslot = (int)((rx.flags >> 4) & 15u), nslots = (int)((rx.flags >> 12) & 15u);
        const bool store_h = (rx.flags >> 8) & 1u, same = xe_p == xb_p;
        const int64_t n4 = a.numel >> 2;
        for (int64_t i4 = tid; i4 < n4; i4 += nthreads) {
            const int64_t i = i4 << 2;
            const float4 xe = ld_state4(xe_p + i);
            const float4 xb = same ? xe : ld_state4(xb_p + i);
            const int64_t fi = f_index(a, i);
            const float4 f = Vec4<FT>::load(a.f, fi);
            const float4 fn = CFG ? Vec4<FT>::load(a.fneg, fi) : zero4;
            float4 hs[AZB_STEP_MAX_SLOTS];
#pragma unroll
            for (int j = 0; j < AZB_STEP_MAX_SLOTS; ++j)
                if (j < nslots && j != wslot && rx.w[j] != 0.f) hs[j] = ld_state4(a.hist + j * a.hist_stride + i);
            float4 m, h, acc;
            m.x = mean_cfg<CFG>(r, xe.x, f.x, fn.x, gw), m.y = mean_cfg<CFG>(r, xe.y, f.y, fn.y, gw);
            m.z = mean_cfg<CFG>(r, xe.z, f.z, fn.z, gw), m.w = mean_cfg<CFG>(r, xe.w, f.w, fn.w, gw);
            h.x = fmaf(rx.p, xe.x, rx.q * m.x), h.y = fmaf(rx.p, xe.y, rx.q * m.y);
            h.z = fmaf(rx.p, xe.z, rx.q * m.z), h.w = fmaf(rx.p, xe.w, rx.q * m.w);
            acc = make_float4(rx.r * xb.x, rx.r * xb.y, rx.r * xb.z, rx.r * xb.w);
#pragma unroll
            for (int j = 0; j < AZB_STEP_MAX_SLOTS; ++j) {
                if (j < nslots) {
                    if (j == wslot) {
                        AZB_AXPY4(acc, rx.w[j], h);
                    } else if (rx.w[j] != 0.f) {
                        AZB_AXPY4(acc, rx.w[j], hs[j]);
                    }
                }
            }
            if (store_h) stg_stream4(a.hist + wslot * a.hist_stride + i, h);
            store_out4<IT>(a, r, out_p, i, acc);
        }
        return;
    }

    const bool generate = (a.eps == nullptr) && (r.n != 0.0f);
    if (!NOISE || !generate) {
        const float poison = (!NOISE && generate) ? __int_as_float(0x7fc00000) : 0.f;
        const int64_t n4 = a.numel >> 2;
        constexpr int U = 4;
        for (int64_t base = tid; base < n4; base += nthreads * U) {
            float4 x[U], f[U], fn[U], e[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int64_t i = (base + u * nthreads) << 2;
                if (i < a.numel) {
                    x[u] = ld_state4(xe_p + i);
                    const int64_t fi = f_index(a, i);
                    f[u] = Vec4<FT>::load(a.f, fi);
                    fn[u] = CFG ? Vec4<FT>::load(a.fneg, fi) : zero4;
                    e[u] = a.eps ? ldg_stream4(a.eps + i) : make_float4(poison, poison, poison, poison);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int64_t i = (base + u * nthreads) << 2;
                if (i < a.numel) finish4<FT, IT, CFG>(a, r, out_p, gw, i, x[u], f[u], fn[u], e[u]);
            }
        }
        return;
    }

    if constexpr (!NOISE) return;
    // In-register noise.  Work item (g, j): Philox streams 4g..4g+3 at call j produce 16 normals
    // that belong to the four float4 groups at global elements (4j+ii)*T + 4g, ii = 0..3.  The grid is two-dimensional
    // (x over the T/4 stream groups, y over the calls), so no index is ever divided.
    const uint64_t offset = (uint64_t)((a.off_dev ? a.off_dev[0] : 0) + a.off_host + (int64_t)rx.draw * a.off_inc);
    const uint64_t seed = a.off_dev ? (uint64_t)a.off_dev[1] : a.seed;
    const uint64_t ctr0 = offset >> 2;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int64_t T = a.T, T4 = T >> 2;
    const int64_t j_lo = (a.elem_off / T) >> 2;
    const int64_t j_hi = ((a.elem_off + a.numel - 1) / T) >> 2;
    for (int64_t j = j_lo + blockIdx.y; j <= j_hi; j += gridDim.y) {
        for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < T4; g += (int64_t)gridDim.x * blockDim.x) {
            float4 x[4], f[4], fn[4];
            int64_t loc[4];
            bool ok[4];
            const int64_t loc0 = (j << 2) * T + (g << 2) - a.elem_off;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                loc[ii] = loc0 + ii * T;
                ok[ii] = loc[ii] >= 0 && loc[ii] < a.numel;
                fn[ii] = zero4;
                if (ok[ii]) {
                    x[ii] = ld_state4(xe_p + loc[ii]);
                    const int64_t fi = f_index(a, loc[ii]);
                    f[ii] = Vec4<FT>::load(a.f, fi);
                    if (CFG) fn[ii] = Vec4<FT>::load(a.fneg, fi);
                }
            }
            if (!(ok[0] | ok[1] | ok[2] | ok[3])) continue;
            float4 z[4];  // z[e] = normals of stream 4g+e, lanes ii
#pragma unroll
            for (int e = 0; e < 4; ++e) z[e] = normal4(ctr0 + (uint64_t)j, (uint64_t)((g << 2) + e), key);
            if (ok[0]) finish4<FT, IT, CFG>(a, r, out_p, gw, loc[0], x[0], f[0], fn[0], make_float4(z[0].x, z[1].x, z[2].x, z[3].x));
            if (ok[1]) finish4<FT, IT, CFG>(a, r, out_p, gw, loc[1], x[1], f[1], fn[1], make_float4(z[0].y, z[1].y, z[2].y, z[3].y));
            if (ok[2]) finish4<FT, IT, CFG>(a, r, out_p, gw, loc[2], x[2], f[2], fn[2], make_float4(z[0].z, z[1].z, z[2].z, z[3].z));
            if (ok[3]) finish4<FT, IT, CFG>(a, r, out_p, gw, loc[3], x[3], f[3], fn[3], make_float4(z[0].w, z[1].w, z[2].w, z[3].w));
        }
    }
}

// Scalar path for small / unaligned tensors (e.g. the (64,5) MLP case, batch-less shapes).
template <typename FT, typename IT, bool CFG>
__global__ void __launch_bounds__(256) step_scalar_kernel(const StepArgs a) {
    pdl_enter();
    RowX rx;
    const Row r = load_row_n(a, rx);
    const float* xe_p = a.src[rx.flags & 1u];
    const float* xb_p = a.src[(rx.flags >> 1) & 1u];
    float* out_p = a.dst[(rx.flags >> 2) & 1u];
    const float gw = CFG ? __ldg(a.guidance) : 0.f;
    const bool hist = rx.flags & 8u;
    const int wslot = (int)((rx.flags >> 4) & 15u), nslots = (int)((rx.flags >> 12) & 15u);
    const bool store_h = (rx.flags >> 8) & 1u;
    const bool generate = !hist && (a.eps == nullptr) && (r.n != 0.0f);
    const uint64_t offset = (uint64_t)((a.off_dev ? a.off_dev[0] : 0) + a.off_host + (int64_t)rx.draw * a.off_inc);
    const uint64_t seed = a.off_dev ? (uint64_t)a.off_dev[1] : a.seed;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.numel; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = ld_state1(xe_p + i);
        const int64_t fi = f_index(a, i);
        const float f = Vec4<FT>::load1(a.f, fi);
        const float fn = CFG ? Vec4<FT>::load1(a.fneg, fi) : 0.f;
        const float m = mean_cfg<CFG>(r, x, f, fn, gw);
        float o;
        if (hist) {
            const float xb = ld_state1(xb_p + i);
            const float h = fmaf(rx.p, x, rx.q * m);
            o = rx.r * xb;
#pragma unroll
            for (int j = 0; j < AZB_STEP_MAX_SLOTS; ++j) {
                if (j < nslots) {
                    if (j == wslot) o = fmaf(rx.w[j], h, o);
                    else if (rx.w[j] != 0.f) o = fmaf(rx.w[j], ld_state1(a.hist + j * a.hist_stride + i), o);
                }
            }
            if (store_h) a.hist[wslot * a.hist_stride + i] = h;
        } else {
            float e = a.eps ? __ldg(a.eps + i) : (generate ? normal_at(a.elem_off + i, a.T, offset >> 2, key) : 0.0f);
            o = affine_of(r, x, m, e);
        }
        out_p[i] = o;
        const float s = __fmul_rn(r.c_in_next, o);
        Vec4<IT>::store1(a.xin, i, s);
        if (a.xin_copies > 1) Vec4<IT>::store1(a.xin, i + a.numel, s);
    }
}

__global__ void advance_kernel(int32_t* step_idx, int64_t* off, int64_t inc, const unsigned char* table,
                               unsigned char* out, int elem_bytes, int count, int32_t steps) {
    pdl_enter();
    __shared__ int32_t s_next;
    if (threadIdx.x == 0) {
        s_next = *step_idx + 1;
        *step_idx = s_next;
        if (off) *off += inc;
    }
    __syncthreads();
    if (table && out) {
        int32_t row = min(s_next, steps - 1);
        int nbytes = elem_bytes * count;
        for (int b = threadIdx.x; b < nbytes; b += blockDim.x) out[b] = table[(int64_t)row * nbytes + b];
    }
}

__global__ void __launch_bounds__(256) init_noise_kernel(float* x, int64_t numel, float mean, float std, uint64_t seed,
                                                         uint64_t offset, int64_t T, int64_t elem_off, int vec) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (!vec) {
        for (int64_t i = tid; i < numel; i += nthreads)
            x[i] = __fadd_rn(mean, __fmul_rn(std, normal_at(elem_off + i, T, offset >> 2, key)));
        return;
    }
    const int64_t T4 = T >> 2;
    const int64_t j_lo = (elem_off / T) >> 2, j_hi = ((elem_off + numel - 1) / T) >> 2;
    const int64_t nwork = T4 * (j_hi - j_lo + 1);
    for (int64_t w = tid; w < nwork; w += nthreads) {
        const int64_t jj = w / T4, g = w - jj * T4, j = j_lo + jj;
        float4 z[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) z[e] = normal4((offset >> 2) + (uint64_t)j, (uint64_t)((g << 2) + e), key);
        const float* zf = reinterpret_cast<const float*>(z);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            int64_t loc = ((j << 2) + ii) * T + (g << 2) - elem_off;
            if (loc >= 0 && loc < numel) {
                float4 o;
                o.x = __fadd_rn(mean, __fmul_rn(std, zf[0 * 4 + ii]));
                o.y = __fadd_rn(mean, __fmul_rn(std, zf[1 * 4 + ii]));
                o.z = __fadd_rn(mean, __fmul_rn(std, zf[2 * 4 + ii]));
                o.w = __fadd_rn(mean, __fmul_rn(std, zf[3 * 4 + ii]));
                stg_stream4(x + loc, o);
            }
        }
    }
}

int sm_count() { return azb_sm_count(); }

int grid_for(int64_t work_items, int per_sm) {
    int64_t blocks = (work_items + 255) / 256;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename FT, typename IT, bool CFG>
int launch_step(const StepArgs& a, bool vec, cudaStream_t s) {
    if (vec) {
        if (a.eps || a.noise_hint < 0)
            return azb_launch(step_vec4_kernel<FT, IT, CFG, false>, dim3(grid_for(a.numel >> 2, 16)), dim3(256), 0, s, a);
        // the row may ask for in-register noise: x over the T/4 Philox stream groups, y over the calls that cover this
        // tensor, about 16 CTAs per SM in total (the paths without noise treat the grid as a flat one)
        const int64_t T4 = a.T >> 2;
        const int64_t spans = ((a.elem_off + a.numel - 1) / a.T >> 2) - ((a.elem_off / a.T) >> 2) + 1;
        const int64_t cap = (int64_t)sm_count() * 16;
        int64_t gx = (T4 + 255) / 256;
        if (gx > cap) gx = cap;
        int64_t gy = cap / gx < 1 ? 1 : cap / gx;
        if (gy > spans) gy = spans;
        if (gy > 65535) gy = 65535;
        return azb_launch(step_vec4_kernel<FT, IT, CFG, true>, dim3((unsigned)gx, (unsigned)gy), dim3(256), 0, s, a);
    }
    return azb_launch(step_scalar_kernel<FT, IT, CFG>, dim3(grid_for(a.numel, 16)), dim3(256), 0, s, a);
}

template <typename FT, bool CFG>
int dispatch_in(const StepArgs& a, int in_dtype, bool vec, cudaStream_t s) {
    if (a.xin == nullptr) return launch_step<FT, NoOut, CFG>(a, vec, s);
    switch (in_dtype) {
        case AZB_F32: return launch_step<FT, float, CFG>(a, vec, s);
        case AZB_BF16: return launch_step<FT, __nv_bfloat16, CFG>(a, vec, s);
        case AZB_F16: return launch_step<FT, __half, CFG>(a, vec, s);
    }
    return AZB_E_DTYPE;
}

template <typename FT>
int dispatch_cfg(const StepArgs& a, int in_dtype, bool vec, cudaStream_t s) {
    return a.fneg ? dispatch_in<FT, true>(a, in_dtype, vec, s) : dispatch_in<FT, false>(a, in_dtype, vec, s);
}

size_t dtype_size(int d) { return d == AZB_F32 ? 4 : (d == AZB_BF16 || d == AZB_F16) ? 2 : d == AZB_I64 ? 8 : 0; }

}  // namespace

extern "C" int azb_rng_policy(int64_t numel, int64_t* rng_threads, int64_t* offset_inc) {
    if (numel <= 0) return AZB_E_SHAPE;
    int dev = 0, sms = 0, tpsm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    if (e != cudaSuccess) return (int)e;
    int64_t grid = (numel + 255) / 256;
    int64_t cap = (int64_t)sms * (tpsm / 256);
    if (grid > cap) grid = cap;
    int64_t T = grid * 256;
    if (rng_threads) *rng_threads = T;
    if (offset_inc) *offset_inc = ((numel - 1) / (T * 4) + 1) * 4;
    return AZB_OK;
}

extern "C" int azb_step_ex_f32(const AzbStep* d, void* stream) {
    AZB_CHECK_PTR(d);
    AZB_CHECK_PTR(d->src[0]);
    AZB_CHECK_PTR(d->dst[0]);
    AZB_CHECK_PTR(d->f);
    AZB_CHECK_PTR(d->table);
    AZB_CHECK_PTR(d->step_idx);
    if (d->f_neg && !d->guidance) return AZB_E_NULL;
    const int64_t n_per_sample = d->n_per_sample, batch = d->batch;
    if (n_per_sample <= 0 || batch <= 0 || d->f_batch_stride < n_per_sample) return AZB_E_SHAPE;
    if (d->rng_threads <= 0 || (d->rng_threads & 3) || d->rng_elem_offset < 0) return AZB_E_SHAPE;
    if (d->row_floats != AZB_COEF_COLS && d->row_floats < AZB_ROW_COLS) return AZB_E_SHAPE;
    if (d->row_floats % 4 || !azb_aligned(d->table, 16)) return AZB_E_ALIGN;
    if (d->x_in_copies != 1 && d->x_in_copies != 2) return AZB_E_SHAPE;
    if (d->hist && d->hist_stride < n_per_sample * batch) return AZB_E_SHAPE;
    const size_t fsz = dtype_size(d->f_dtype), isz = dtype_size(d->in_dtype);
    if (fsz == 0 || d->f_dtype == AZB_I64) return AZB_E_DTYPE;
    if (d->x_in_next && (isz == 0 || d->in_dtype == AZB_I64)) return AZB_E_DTYPE;
    StepArgs a{};
    a.src[0] = d->src[0], a.src[1] = d->src[1] ? d->src[1] : d->src[0];
    a.dst[0] = d->dst[0], a.dst[1] = d->dst[1] ? d->dst[1] : d->dst[0];
    a.f = d->f, a.fneg = d->f_neg, a.guidance = d->guidance, a.eps = d->eps, a.xin = d->x_in_next, a.hist = d->hist;
    a.n_per_sample = n_per_sample, a.numel = n_per_sample * batch, a.f_bstride = d->f_batch_stride;
    a.hist_stride = d->hist_stride, a.table = d->table, a.step_idx = d->step_idx, a.seed = d->seed;
    a.off_dev = d->philox_state, a.off_host = d->offset_host, a.off_inc = d->offset_inc, a.T = d->rng_threads;
    a.elem_off = d->rng_elem_offset, a.row_floats = d->row_floats, a.xin_copies = d->x_in_copies, a.noise_hint = d->noise_hint;
    a.np_magic = (a.numel < (1ll << 32) && n_per_sample > 1) ? (~0ull / (unsigned long long)n_per_sample) + 1ull : 0ull;
    const bool vec = (n_per_sample % 4 == 0) && (d->f_batch_stride % 4 == 0) && (d->rng_elem_offset % 4 == 0) &&
                     azb_aligned(a.src[0], 16) && azb_aligned(a.src[1], 16) && azb_aligned(a.dst[0], 16) &&
                     azb_aligned(a.dst[1], 16) && azb_aligned(a.f, 4 * fsz) && (!a.fneg || azb_aligned(a.fneg, 4 * fsz)) &&
                     (!a.eps || azb_aligned(a.eps, 16)) && (!a.xin || azb_aligned(a.xin, 4 * isz)) &&
                     (!a.hist || (azb_aligned(a.hist, 16) && d->hist_stride % 4 == 0));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (d->f_dtype) {
        case AZB_F32: return dispatch_cfg<float>(a, d->in_dtype, vec, s);
        case AZB_BF16: return dispatch_cfg<__nv_bfloat16>(a, d->in_dtype, vec, s);
        case AZB_F16: return dispatch_cfg<__half>(a, d->in_dtype, vec, s);
    }
    return AZB_E_DTYPE;
}

extern "C" int azb_step_f32(const float* x_t, const void* f, int f_dtype, int64_t f_batch_stride, const float* eps,
                            float* x_s, void* x_in_next, int in_dtype, int64_t n_per_sample, int64_t batch,
                            const float* coef_table, const int32_t* step_idx, uint64_t seed,
                            const int64_t* philox_state, int64_t offset_host, int64_t rng_threads,
                            int64_t rng_elem_offset, void* stream) {
    AZB_CHECK_PTR(x_t);
    AZB_CHECK_PTR(x_s);
    AzbStep d{};
    d.src[0] = x_t, d.dst[0] = x_s, d.f = f, d.eps = eps, d.x_in_next = x_in_next, d.table = coef_table, d.step_idx = step_idx;
    d.philox_state = philox_state, d.f_batch_stride = f_batch_stride, d.n_per_sample = n_per_sample, d.batch = batch;
    d.offset_host = offset_host, d.rng_threads = rng_threads, d.rng_elem_offset = rng_elem_offset, d.seed = seed;
    d.f_dtype = f_dtype, d.in_dtype = in_dtype, d.row_floats = AZB_COEF_COLS, d.x_in_copies = 1;
    return azb_step_ex_f32(&d, stream);
}

extern "C" int azb_advance(int32_t* step_idx, int64_t* philox_state, int64_t offset_inc, const void* time_table,
                           void* time_out, int time_elem_bytes, int time_count, int32_t steps, void* stream) {
    AZB_CHECK_PTR(step_idx);
    if ((time_table == nullptr) != (time_out == nullptr)) return AZB_E_NULL;
    if (time_table && (time_elem_bytes <= 0 || time_count <= 0 || steps <= 0)) return AZB_E_SHAPE;
    return azb_launch(advance_kernel, dim3(1), dim3(64), 0, reinterpret_cast<cudaStream_t>(stream), step_idx, philox_state,
                      offset_inc, reinterpret_cast<const unsigned char*>(time_table), reinterpret_cast<unsigned char*>(time_out),
                      time_elem_bytes, time_count, steps);
}

extern "C" int azb_init_noise_f32(float* x, int64_t numel, float mean_T, float std_T, uint64_t seed, int64_t offset_host,
                                  int64_t rng_threads, int64_t rng_elem_offset, void* stream) {
    AZB_CHECK_PTR(x);
    if (numel <= 0 || rng_threads <= 0 || (rng_threads & 3) || rng_elem_offset < 0) return AZB_E_SHAPE;
    int vec = (numel % 4 == 0) && (rng_elem_offset % 4 == 0) && azb_aligned(x, 16);
    int64_t work = numel;
    if (vec) {
        int64_t spans = ((rng_elem_offset + numel - 1) / rng_threads >> 2) - ((rng_elem_offset / rng_threads) >> 2) + 1;
        work = (rng_threads >> 2) * spans;
    }
    init_noise_kernel<<<grid_for(work, 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, numel, mean_T, std_T, seed, (uint64_t)offset_host, rng_threads, rng_elem_offset, vec);
    return azb_launch_status();
}
