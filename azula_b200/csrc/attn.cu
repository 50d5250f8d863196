// Self-attention of the ADM UNet's AttentionBlock (azula/plugins/adm/_src/unet.py:328-345,
// 361-381): softmax(q k^T / sqrt(d)) v per (image, head), computed flash-style -- the T x T
// logits never reach HBM (the reference materialises them in fp32, unet.py:343).
//
// Shapes are tiny for a B200 (T <= 1024, d = 64; 0.5 % of the network's FLOPs), so this first
// version uses warp-level mma.sync (m16n8k16, bf16 in / fp32 accumulate) with cp.async double
// buffering; a tcgen05/TMEM version is planned once the convolutions stop dominating.
//
// qkv: (N, T, ld) bf16.  Head `hd` finds q/k/v at channel hd*head_stride + {0, k_delta, v_delta}
// (legacy order: head_stride = 3d, k_delta = d, v_delta = 2d; new order: head_stride = d,
// k_delta = C, v_delta = 2C).  out: (N, T, out_ld) bf16, head hd at channel hd*d.

#include "common.cuh"

namespace {

constexpr int BM = 64;  // queries per CTA (4 warps x 16)
constexpr int BN = 64;  // keys per tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
// fp16 operands (10-bit mantissa, the TF32 mode's attention: qkv arrive as fp16 from the TF32 projection's epilogue)
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
template <bool F16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (F16) mma_f16(d, a, b0, b1);
    else mma_bf16(d, a, b0, b1);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (F16) return pack_f16(lo, hi);
    else return pack_bf16(lo, hi);
}

struct AttnParams {
    const __nv_bfloat16* qkv;  // (fp16 when F16: same 2-byte layout)
    int64_t ld;
    __nv_bfloat16* out;        // (fp32 when F16: the TF32 mode keeps activations in fp32)
    int64_t out_ld;
    int T, heads;
    int head_stride, k_delta, v_delta;
    float scale_log2e;  // (1/sqrt(d)) * log2(e)
};

template <int D>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int row0, int T) {
    constexpr int PITCH = D + 8;
    constexpr int VEC = D / 8;  // 16-byte vectors per row
    for (int i = threadIdx.x; i < BN * VEC; i += 128) {
        const int r = i / VEC, v = i - r * VEC;
        __nv_bfloat16* d = dst + r * PITCH + v * 8;
        if (row0 + r < T) {
            cp_async16(d, src + (int64_t)(row0 + r) * ld + v * 8);
        } else {
            *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
        }
    }
}

template <int D, bool F16 = false>
__global__ void __launch_bounds__(128) attention_kernel(const AttnParams p) {
    pdl_enter();
    constexpr int PITCH = D + 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* sK = sQ + BM * PITCH;      // [2][BN][PITCH]
    __nv_bfloat16* sV = sK + 2 * BN * PITCH;  // [2][BN][PITCH]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int q0 = blockIdx.x * BM;
    const int hd = blockIdx.y, n = blockIdx.z;
    const __nv_bfloat16* base = p.qkv + (int64_t)n * p.T * p.ld + hd * p.head_stride;
    const __nv_bfloat16* gQ = base;
    const __nv_bfloat16* gK = base + p.k_delta;
    const __nv_bfloat16* gV = base + p.v_delta;

    load_tile<D>(sQ, gQ, p.ld, q0, p.T);
    load_tile<D>(sK, gK, p.ld, 0, p.T);
    load_tile<D>(sV, gV, p.ld, 0, p.T);
    cp_async_commit();

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float row_max[2] = {-INFINITY, -INFINITY}, row_sum[2] = {0.f, 0.f};

    const int num_tiles = (p.T + BN - 1) / BN;
    for (int tile = 0; tile < num_tiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < num_tiles) {
            load_tile<D>(sK + (buf ^ 1) * BN * PITCH, gK, p.ld, (tile + 1) * BN, p.T);
            load_tile<D>(sV + (buf ^ 1) * BN * PITCH, gV, p.ld, (tile + 1) * BN, p.T);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __nv_bfloat16* tK = sK + buf * BN * PITCH;
        const __nv_bfloat16* tV = sV + buf * BN * PITCH;

        // S = Q K^T for this warp's 16 rows x 64 keys
        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            uint32_t a[4];
            ldsm_x4(a, sQ + (warp * 16 + (lane & 15)) * PITCH + kk * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int nb = 0; nb < BN / 16; ++nb) {
                uint32_t b[4];
                ldsm_x4(b, tK + (nb * 16 + (lane & 7) + (lane >> 4) * 8) * PITCH + kk * 16 + ((lane >> 3) & 1) * 8);
                mma_16816<F16>(s[2 * nb], a, b[0], b[1]);
                mma_16816<F16>(s[2 * nb + 1], a, b[2], b[3]);
            }
        }
        // mask keys beyond T, scale, online softmax (rows g and g+8 of this warp)
        const int key0 = tile * BN;
        float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = key0 + i * 8 + t4 * 2 + (j & 1);
                float v = s[i][j] * p.scale_log2e;
                if (key >= p.T) v = -INFINITY;
                s[i][j] = v;
                tmax[j >> 1] = fmaxf(tmax[j >> 1], v);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
            const float new_max = fmaxf(row_max[r], tmax[r]);
            corr[r] = exp2f(row_max[r] - new_max);
            row_max[r] = new_max;
            row_sum[r] *= corr[r];
        }
        float psum[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = exp2f(s[i][j] - row_max[j >> 1]);
                s[i][j] = e;
                psum[j >> 1] += e;
            }
        }
        row_sum[0] += psum[0], row_sum[1] += psum[1];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o[i][0] *= corr[0], o[i][1] *= corr[0];
            o[i][2] *= corr[1], o[i][3] *= corr[1];
        }
        // O += P V
#pragma unroll
        for (int kb = 0; kb < BN / 16; ++kb) {
            uint32_t a[4];
            a[0] = pack2<F16>(s[2 * kb][0], s[2 * kb][1]);
            a[1] = pack2<F16>(s[2 * kb][2], s[2 * kb][3]);
            a[2] = pack2<F16>(s[2 * kb + 1][0], s[2 * kb + 1][1]);
            a[3] = pack2<F16>(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
            for (int nb = 0; nb < D / 16; ++nb) {
                uint32_t b[4];
                ldsm_x4_trans(b, tV + (kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + nb * 16 + (lane >> 4) * 8);
                mma_16816<F16>(o[2 * nb], a, b[0], b[1]);
                mma_16816<F16>(o[2 * nb + 1], a, b[2], b[3]);
            }
        }
        __syncthreads();  // everyone done with this buffer before it is refilled
    }

    // finalise: divide by the row sums (summed over the quad), store bf16
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 1);
        row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 2);
    }
    const float inv[2] = {1.f / row_sum[0], 1.f / row_sum[1]};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        if (q >= p.T) continue;
        if constexpr (F16) {
            float* dst = reinterpret_cast<float*>(p.out) + ((int64_t)n * p.T + q) * p.out_ld + hd * D;
#pragma unroll
            for (int i = 0; i < D / 8; ++i)
                *reinterpret_cast<float2*>(dst + i * 8 + t4 * 2) = make_float2(o[i][2 * r] * inv[r], o[i][2 * r + 1] * inv[r]);
        } else {
            __nv_bfloat16* dst = p.out + ((int64_t)n * p.T + q) * p.out_ld + hd * D;
#pragma unroll
            for (int i = 0; i < D / 8; ++i)
                *reinterpret_cast<uint32_t*>(dst + i * 8 + t4 * 2) = pack_bf16(o[i][2 * r] * inv[r], o[i][2 * r + 1] * inv[r]);
        }
    }
}

template <int D, bool F16 = false>
int launch_attn(const AttnParams& p, int n, cudaStream_t s) {
    constexpr int smem = (BM + 4 * BN) * (D + 8) * 2;
    static AzbPerDevice<bool> configured_dev;
    bool& configured = configured_dev.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<D, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    azb_launch(attention_kernel<D, F16>, dim3((p.T + BM - 1) / BM, p.heads, n), dim3(128), smem, s, p);
    return azb_launch_status();
}

}  // namespace

int azb_attention_tc_launch(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t, int64_t heads,
                            int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta, int qk_norm, float qk_eps,
                            void* stream);  // attn_tc.cu

// Attention with the per-head RMS normalisation of q and k folded into the logits (T <= 256, d = 64: DiT tokens):
// replaces azb_segment_rmsnorm_bf16 + azb_attention_bf16 (azula/nn/attention.py:103,110-116).
extern "C" int azb_attention_qknorm_bf16(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t,
                                         int64_t heads, int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta,
                                         float qk_eps, void* stream) {
    AZB_CHECK_PTR(qkv);
    AZB_CHECK_PTR(out);
    if (n <= 0 || t <= 0 || heads <= 0 || n > 65535 || heads > 65535) return AZB_E_SHAPE;
    return azb_attention_tc_launch(qkv, ld, out, out_ld, n, t, heads, d, head_stride, k_delta, v_delta, 1, qk_eps, stream);
}

extern "C" int azb_attention_bf16(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t,
                                  int64_t heads, int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta,
                                  void* stream) {
    AZB_CHECK_PTR(qkv);
    AZB_CHECK_PTR(out);
    if (n <= 0 || t <= 0 || heads <= 0 || n > 65535 || heads > 65535) return AZB_E_SHAPE;
    // head width 64 (every ADM card, DiT-B): tcgen05 kernel; other widths: the mma.sync kernel below
    const int rc = azb_attention_tc_launch(qkv, ld, out, out_ld, n, t, heads, d, head_stride, k_delta, v_delta, 0, 0.f, stream);
    if (rc != AZB_E_UNSUPPORTED) return rc;
    return azb_attention_mma_bf16(qkv, ld, out, out_ld, n, t, heads, d, head_stride, k_delta, v_delta, stream);
}

extern "C" int azb_attention_mma_bf16(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t,
                                      int64_t heads, int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta,
                                      void* stream) {
    AZB_CHECK_PTR(qkv);
    AZB_CHECK_PTR(out);
    if (n <= 0 || t <= 0 || heads <= 0 || n > 65535 || heads > 65535) return AZB_E_SHAPE;
    if (ld % 8 || out_ld % 2 || head_stride % 8 || k_delta % 8 || v_delta % 8) return AZB_E_ALIGN;
    if (!azb_aligned(qkv, 16) || !azb_aligned(out, 4)) return AZB_E_ALIGN;
    AttnParams p{};
    p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv), p.ld = ld;
    p.out = reinterpret_cast<__nv_bfloat16*>(out), p.out_ld = out_ld;
    p.T = (int)t, p.heads = (int)heads;
    p.head_stride = (int)head_stride, p.k_delta = (int)k_delta, p.v_delta = (int)v_delta;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (d) {
        case 16: return launch_attn<16>(p, (int)n, s);
        case 32: return launch_attn<32>(p, (int)n, s);
        case 64: return launch_attn<64>(p, (int)n, s);
        case 128: return launch_attn<128>(p, (int)n, s);
        // num_heads = 4 cards (imagenet_128x128_cond, cards.yaml:19-34): head widths 128 / 192 / 256
        case 192: return launch_attn<192>(p, (int)n, s);
        case 256: return launch_attn<256>(p, (int)n, s);
    }
    return AZB_E_SHAPE;
}

// The TF32 mode's attention: qkv fp16 (N, T, ld), out fp32 (N, T, out_ld); fp32 softmax and accumulation.
extern "C" int azb_attention_f16(const void* qkv, int64_t ld, float* out, int64_t out_ld, int64_t n, int64_t t, int64_t heads,
                                 int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta, void* stream) {
    AZB_CHECK_PTR(qkv);
    AZB_CHECK_PTR(out);
    if (n <= 0 || t <= 0 || heads <= 0 || n > 65535 || heads > 65535) return AZB_E_SHAPE;
    if (ld % 8 || out_ld % 2 || head_stride % 8 || k_delta % 8 || v_delta % 8) return AZB_E_ALIGN;
    if (!azb_aligned(qkv, 16) || !azb_aligned(out, 8)) return AZB_E_ALIGN;
    AttnParams p{};
    p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv), p.ld = ld;
    p.out = reinterpret_cast<__nv_bfloat16*>(out), p.out_ld = out_ld;
    p.T = (int)t, p.heads = (int)heads;
    p.head_stride = (int)head_stride, p.k_delta = (int)k_delta, p.v_delta = (int)v_delta;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (d) {
        case 32: return launch_attn<32, true>(p, (int)n, s);
        case 64: return launch_attn<64, true>(p, (int)n, s);
        case 128: return launch_attn<128, true>(p, (int)n, s);
        case 256: return launch_attn<256, true>(p, (int)n, s);
    }
    return AZB_E_SHAPE;
}
