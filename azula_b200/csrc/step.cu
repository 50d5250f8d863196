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

__device__ __forceinline__ Row load_row(const float* table, const int32_t* step_idx) {
    const float4* r = reinterpret_cast<const float4*>(table + (int64_t)(*step_idx) * AZB_COEF_COLS);
    float4 a = __ldg(r), b = __ldg(r + 1);
    return Row{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
}

// The update of ONE element, every operation rounded separately (no FMA contraction) so that
// it reproduces the eager op sequence bit for bit given the same F and eps:
//   mean = c_skip*x + c_out*F               denoise.py:322 / adm/__init__.py:125-130
//   mean = clip(mean)                       adm/__init__.py:133-134
//   x_s  = alpha_s*mean                     sample.py:212,257
//   x_s += k*(x - alpha_t*mean)             sample.py:213,258
//   x_s += n*eps                            sample.py:214,259
__device__ __forceinline__ float transition(const Row& r, float x, float f, float eps) {
    float m = __fadd_rn(__fmul_rn(r.c_skip, x), __fmul_rn(r.c_out, f));
    if (m == m) m = fminf(fmaxf(m, -r.clip), r.clip);
    float xs = __fmul_rn(r.alpha_s, m);
    xs = __fadd_rn(xs, __fmul_rn(r.k, __fsub_rn(x, __fmul_rn(r.alpha_t, m))));
    xs = __fadd_rn(xs, __fmul_rn(r.n, eps));
    return xs;
}

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

struct StepArgs {
    const float* x;
    const void* f;
    const float* eps;
    float* out;
    void* xin;
    int64_t n_per_sample, numel, f_bstride;
    const float* table;
    const int32_t* step_idx;
    uint64_t seed;
    const int64_t* off_dev;
    int64_t off_host, T, elem_off;
};

__device__ __forceinline__ int64_t f_index(const StepArgs& a, int64_t i) {
    if (a.f_bstride == a.n_per_sample) return i;
    int64_t b = i / a.n_per_sample;
    return b * a.f_bstride + (i - b * a.n_per_sample);
}

template <typename FT, typename IT>
__device__ __forceinline__ void finish4(const StepArgs& a, const Row& r, int64_t i, float4 x, float4 f, float4 e) {
    float4 o;
    o.x = transition(r, x.x, f.x, e.x);
    o.y = transition(r, x.y, f.y, e.y);
    o.z = transition(r, x.z, f.z, e.z);
    o.w = transition(r, x.w, f.w, e.w);
    stg_stream4(a.out + i, o);
    Vec4<IT>::store(a.xin, i,
                    make_float4(__fmul_rn(r.c_in_next, o.x), __fmul_rn(r.c_in_next, o.y), __fmul_rn(r.c_in_next, o.z),
                                __fmul_rn(r.c_in_next, o.w)));
}

// Vector path: numel, n_per_sample, f_bstride, elem_off multiples of 4; 16-byte aligned pointers.
template <typename FT, typename IT>
__global__ void __launch_bounds__(256) step_vec4_kernel(const StepArgs a) {
    pdl_enter();
    const Row r = load_row(a.table, a.step_idx);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const bool generate = (a.eps == nullptr) && (r.n != 0.0f);

    if (!generate) {
        const int64_t n4 = a.numel >> 2;
        constexpr int U = 4;
        for (int64_t base = tid; base < n4; base += nthreads * U) {
            float4 x[U], f[U], e[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int64_t i = (base + u * nthreads) << 2;
                if (i < a.numel) {
                    x[u] = ldg_stream4(a.x + i);
                    f[u] = Vec4<FT>::load(a.f, f_index(a, i));
                    e[u] = a.eps ? ldg_stream4(a.eps + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int64_t i = (base + u * nthreads) << 2;
                if (i < a.numel) finish4<FT, IT>(a, r, i, x[u], f[u], e[u]);
            }
        }
        return;
    }

    // In-register noise.  Work item (g, j): Philox streams 4g..4g+3 at call j produce 16 normals
    // that belong to the four float4 groups at global elements (4j+ii)*T + 4g, ii = 0..3.
    const uint64_t offset = (uint64_t)((a.off_dev ? a.off_dev[0] : 0) + a.off_host);
    const uint64_t seed = a.off_dev ? (uint64_t)a.off_dev[1] : a.seed;
    const uint64_t ctr0 = offset >> 2;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int64_t T = a.T, T4 = T >> 2;
    const int64_t j_lo = (a.elem_off / T) >> 2;
    const int64_t j_hi = ((a.elem_off + a.numel - 1) / T) >> 2;
    const int64_t nwork = T4 * (j_hi - j_lo + 1);
    for (int64_t w = tid; w < nwork; w += nthreads) {
        const int64_t jj = w / T4;
        const int64_t g = w - jj * T4;
        const int64_t j = j_lo + jj;
        float4 x[4], f[4];
        int64_t loc[4];
        bool ok[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            loc[ii] = ((j << 2) + ii) * T + (g << 2) - a.elem_off;
            ok[ii] = loc[ii] >= 0 && loc[ii] < a.numel;
            if (ok[ii]) {
                x[ii] = ldg_stream4(a.x + loc[ii]);
                f[ii] = Vec4<FT>::load(a.f, f_index(a, loc[ii]));
            }
        }
        if (!(ok[0] | ok[1] | ok[2] | ok[3])) continue;
        float4 z[4];  // z[e] = normals of stream 4g+e, lanes ii
#pragma unroll
        for (int e = 0; e < 4; ++e) z[e] = normal4(ctr0 + (uint64_t)j, (uint64_t)((g << 2) + e), key);
        if (ok[0]) finish4<FT, IT>(a, r, loc[0], x[0], f[0], make_float4(z[0].x, z[1].x, z[2].x, z[3].x));
        if (ok[1]) finish4<FT, IT>(a, r, loc[1], x[1], f[1], make_float4(z[0].y, z[1].y, z[2].y, z[3].y));
        if (ok[2]) finish4<FT, IT>(a, r, loc[2], x[2], f[2], make_float4(z[0].z, z[1].z, z[2].z, z[3].z));
        if (ok[3]) finish4<FT, IT>(a, r, loc[3], x[3], f[3], make_float4(z[0].w, z[1].w, z[2].w, z[3].w));
    }
}

// Scalar path for small / unaligned tensors (e.g. the (64,5) MLP case, batch-less shapes).
template <typename FT, typename IT>
__global__ void __launch_bounds__(256) step_scalar_kernel(const StepArgs a) {
    pdl_enter();
    const Row r = load_row(a.table, a.step_idx);
    const bool generate = (a.eps == nullptr) && (r.n != 0.0f);
    const uint64_t offset = (uint64_t)((a.off_dev ? a.off_dev[0] : 0) + a.off_host);
    const uint64_t seed = a.off_dev ? (uint64_t)a.off_dev[1] : a.seed;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.numel; i += (int64_t)gridDim.x * blockDim.x) {
        float x = __ldg(a.x + i);
        float f = Vec4<FT>::load1(a.f, f_index(a, i));
        float e = a.eps ? __ldg(a.eps + i) : (generate ? normal_at(a.elem_off + i, a.T, offset >> 2, key) : 0.0f);
        float o = transition(r, x, f, e);
        a.out[i] = o;
        Vec4<IT>::store1(a.xin, i, __fmul_rn(r.c_in_next, o));
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

int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}

int grid_for(int64_t work_items, int per_sm) {
    int64_t blocks = (work_items + 255) / 256;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename FT, typename IT>
int launch_step(const StepArgs& a, bool vec, cudaStream_t s) {
    if (vec) {
        int64_t T4 = a.T >> 2;
        int64_t spans = ((a.elem_off + a.numel - 1) / a.T >> 2) - ((a.elem_off / a.T) >> 2) + 1;
        int64_t work = a.eps ? (a.numel >> 2) : ((a.numel >> 2) > T4 * spans ? (a.numel >> 2) : T4 * spans);
        // no-noise path consumes 4 float4 per thread and iteration
        return azb_launch(step_vec4_kernel<FT, IT>, dim3(grid_for(work, 16)), dim3(256), 0, s, a);
    }
    return azb_launch(step_scalar_kernel<FT, IT>, dim3(grid_for(a.numel, 16)), dim3(256), 0, s, a);
}

template <typename FT>
int dispatch_in(const StepArgs& a, int in_dtype, bool vec, cudaStream_t s) {
    if (a.xin == nullptr) return launch_step<FT, NoOut>(a, vec, s);
    switch (in_dtype) {
        case AZB_F32: return launch_step<FT, float>(a, vec, s);
        case AZB_BF16: return launch_step<FT, __nv_bfloat16>(a, vec, s);
        case AZB_F16: return launch_step<FT, __half>(a, vec, s);
    }
    return AZB_E_DTYPE;
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

extern "C" int azb_step_f32(const float* x_t, const void* f, int f_dtype, int64_t f_batch_stride, const float* eps,
                            float* x_s, void* x_in_next, int in_dtype, int64_t n_per_sample, int64_t batch,
                            const float* coef_table, const int32_t* step_idx, uint64_t seed,
                            const int64_t* philox_state, int64_t offset_host, int64_t rng_threads,
                            int64_t rng_elem_offset, void* stream) {
    AZB_CHECK_PTR(x_t);
    AZB_CHECK_PTR(f);
    AZB_CHECK_PTR(x_s);
    AZB_CHECK_PTR(coef_table);
    AZB_CHECK_PTR(step_idx);
    if (n_per_sample <= 0 || batch <= 0 || f_batch_stride < n_per_sample) return AZB_E_SHAPE;
    if (rng_threads <= 0 || (rng_threads & 3) || rng_elem_offset < 0) return AZB_E_SHAPE;
    if (!azb_aligned(coef_table, 16)) return AZB_E_ALIGN;
    size_t fsz = dtype_size(f_dtype), isz = dtype_size(in_dtype);
    if (fsz == 0 || f_dtype == AZB_I64) return AZB_E_DTYPE;
    if (x_in_next && (isz == 0 || in_dtype == AZB_I64)) return AZB_E_DTYPE;
    StepArgs a{x_t,        f,        eps,  x_s,           x_in_next,   n_per_sample, n_per_sample * batch, f_batch_stride,
               coef_table, step_idx, seed, philox_state, offset_host, rng_threads,  rng_elem_offset};
    bool vec = (n_per_sample % 4 == 0) && (f_batch_stride % 4 == 0) && (rng_elem_offset % 4 == 0) &&
               azb_aligned(x_t, 16) && azb_aligned(x_s, 16) && azb_aligned(f, 4 * fsz) &&
               (!eps || azb_aligned(eps, 16)) && (!x_in_next || azb_aligned(x_in_next, 4 * isz));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (f_dtype) {
        case AZB_F32: return dispatch_in<float>(a, in_dtype, vec, s);
        case AZB_BF16: return dispatch_in<__nv_bfloat16>(a, in_dtype, vec, s);
        case AZB_F16: return dispatch_in<__half>(a, in_dtype, vec, s);
    }
    return AZB_E_DTYPE;
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
