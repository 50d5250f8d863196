// Small kernels around the tensor-core path of the ADM UNet (sm_100a):
//   - im2col of the 3-channel network input (fp32 NCHW -> bf16 [pixels][64]) so that the first
//     3x3 convolution (azula/plugins/adm/_src/unet.py:471) runs on the same tcgen05 GEMM;
//   - the sinusoidal timestep features (_src/nn.py:90-108);
//   - fp32 linear layers with optional SiLU on the input, for the time-embedding MLP and the
//     per-block emb_layers (_src/unet.py:458-462,198-204), which stay in fp32 because they steer
//     every normalisation of the network.

#include "common.cuh"

namespace {

// out[n][h][w][k], k = (kh*3+kw)*C + c for k < 9C, zero up to K (=64); input zero padded by 1.
// CT > 0: channel count known at compile time (the divisions by c and by 3 become multiplications; the RGB stem is
// CT = 3), 0: generic.
template <int CT>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        int n, int c_rt, int h, int w, int K) {
    pdl_enter();
    const int c = CT > 0 ? CT : c_rt;
    const int vecs = K / 8;
    const int live = (9 * c + 7) / 8;  // vectors that hold at least one tap; the rest of a row is zero padding
    const int64_t total = (int64_t)n * h * w * vecs;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(item % vecs);
        const int64_t pix = item / vecs;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (v < live) {
            const int ow = (int)(pix % w);
            const int oh = (int)((pix / w) % h);
            const int on = (int)(pix / ((int64_t)w * h));
            const float* img = x + (int64_t)on * c * h * w;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = v * 8 + j;
                float val = 0.f;
                if (k < 9 * c) {
                    const int tap = k / c, ch = k - tap * c;
                    const int ih = oh + tap / 3 - 1, iw = ow + tap % 3 - 1;
                    if (ih >= 0 && ih < h && iw >= 0 && iw < w) val = __ldg(img + ((int64_t)ch * h + ih) * w + iw);
                }
                f[j] = val;
            }
            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
            o.x = *reinterpret_cast<uint32_t*>(&t0), o.y = *reinterpret_cast<uint32_t*>(&t1);
            o.z = *reinterpret_cast<uint32_t*>(&t2), o.w = *reinterpret_cast<uint32_t*>(&t3);
        }
        *reinterpret_cast<uint4*>(out + pix * K + v * 8) = o;
    }
}

// emb[r][0:half] = cos(t_r * f_i), emb[r][half:2half] = sin(t_r * f_i), f_i = exp(-ln(max_period) i / half)
__global__ void timestep_features_kernel(const int64_t* __restrict__ t_i64, const float* __restrict__ t_f32, int rows,
                                         int dim, float max_period, float* __restrict__ out) {
    pdl_enter();
    const int half = dim / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * half; i += gridDim.x * blockDim.x) {
        const int r = i / half, j = i - r * half;
        const float t = t_i64 ? (float)t_i64[r] : t_f32[r];
        const float freq = expf(-logf(max_period) * (float)j / (float)half);
        const float arg = t * freq;
        out[(int64_t)r * dim + j] = cosf(arg);
        out[(int64_t)r * dim + half + j] = sinf(arg);
        if ((dim & 1) && j == 0) out[(int64_t)r * dim + dim - 1] = 0.f;
    }
}

// y[m][n] = b[n] + sum_k act(x[m][k]) * W[n][k]; one warp per output column, loops over rows.
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                         const float* __restrict__ b, float* __restrict__ y, int M, int N,
                                                         int K, int silu_in) {
    pdl_enter();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= N) return;
    const float* wrow = W + (int64_t)warp * K;
    for (int m = blockIdx.y; m < M; m += gridDim.y) {
        const float* xrow = x + (int64_t)m * K;
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
        if (lane == 0) y[(int64_t)m * N + warp] = acc + (b ? __ldg(b + warp) : 0.f);
    }
}

__global__ void add_rows_kernel(float* __restrict__ y, const float* __restrict__ table, const int64_t* __restrict__ idx,
                                int rows, int dim) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * dim; i += gridDim.x * blockDim.x) {
        const int r = i / dim, j = i - r * dim;
        y[i] += table[idx[r] * (int64_t)dim + j];
    }
}

}  // namespace

extern "C" int azb_im2col3x3_f32(const float* x, void* out, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k_pad,
                                 void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(out);
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || k_pad % 8 || 9 * c > k_pad) return AZB_E_SHAPE;
    if (!azb_aligned(out, 16)) return AZB_E_ALIGN;
    const int64_t total = n * h * w * (k_pad / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    const dim3 grid((unsigned)blocks), block(256);
    if (c == 3) return azb_launch(im2col3x3_kernel<3>, grid, block, 0, st, x, o, (int)n, 3, (int)h, (int)w, (int)k_pad);
    if (c == 4) return azb_launch(im2col3x3_kernel<4>, grid, block, 0, st, x, o, (int)n, 4, (int)h, (int)w, (int)k_pad);
    return azb_launch(im2col3x3_kernel<0>, grid, block, 0, st, x, o, (int)n, (int)c, (int)h, (int)w, (int)k_pad);
}

extern "C" int azb_timestep_features_f32(const void* t, int t_dtype, int64_t rows, int64_t dim, float max_period,
                                         float* out, void* stream) {
    AZB_CHECK_PTR(t);
    AZB_CHECK_PTR(out);
    if (rows <= 0 || dim < 2) return AZB_E_SHAPE;
    if (t_dtype != AZB_I64 && t_dtype != AZB_F32) return AZB_E_DTYPE;
    const int64_t work = rows * (dim / 2);
    return azb_launch(timestep_features_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
        t_dtype == AZB_I64 ? reinterpret_cast<const int64_t*>(t) : nullptr,
        t_dtype == AZB_F32 ? reinterpret_cast<const float*>(t) : nullptr, (int)rows, (int)dim, max_period, out);
}

extern "C" int azb_linear_f32(const float* x, const float* w, const float* b, float* y, int64_t m, int64_t n, int64_t k,
                              int silu_in, void* stream) {
    AZB_CHECK_PTR(x);
    AZB_CHECK_PTR(w);
    AZB_CHECK_PTR(y);
    if (m <= 0 || n <= 0 || k <= 0 || k % 4) return AZB_E_SHAPE;
    if (!azb_aligned(x, 16) || !azb_aligned(w, 16)) return AZB_E_ALIGN;
    dim3 grid((unsigned)((n + 7) / 8), (unsigned)(m < 64 ? m : 64));
    return azb_launch(linear_f32_kernel, grid, dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, w, b, y, (int)m, (int)n, (int)k,
                                                                                silu_in);
}

extern "C" int azb_zero_bytes(void* ptr, int64_t bytes, void* stream) {
    AZB_CHECK_PTR(ptr);
    if (bytes <= 0) return AZB_E_SHAPE;
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, reinterpret_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? AZB_OK : (int)e;
}

extern "C" int azb_add_rows_f32(float* y, const float* table, const int64_t* idx, int64_t rows, int64_t dim,
                                void* stream) {
    AZB_CHECK_PTR(y);
    AZB_CHECK_PTR(table);
    AZB_CHECK_PTR(idx);
    if (rows <= 0 || dim <= 0) return AZB_E_SHAPE;
    return azb_launch(add_rows_kernel, dim3((unsigned)((rows * dim + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
        y, table, idx, (int)rows, (int)dim);
}
