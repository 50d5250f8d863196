/*
 * azb.h -- C ABI of libazb, the sm_100a engine behind azula_b200.
 *
 * The reference (probabilists/azula @ bec12b8) is pure Python on PyTorch and has no FFI;
 * its "operator API" on the generation path is the duck-typed Python surface
 *   Sampler.step / Sampler.init     azula/sample.py:96-128,163-176,204-216,248-261
 *   Denoiser.forward                azula/denoise.py:293-324, azula/plugins/adm/__init__.py:86-136
 *   backbone(x, t, **kw)            azula/plugins/adm/_src/unet.py:605-634
 * Each entry point below replaces the ATen dispatch sequence issued by the cited lines and
 * is what a reference-side binding (ctypes stub) would call.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless stated otherwise;
 *   - the library allocates nothing and keeps no pointer after a call returns;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *     synchronisation and no host read of device data => legal inside CUDA-graph capture;
 *   - return value: 0 = ok, >0 = cudaError_t of the launch, <0 = AZB_E_* argument error;
 *     nothing throws, prints or exits.  azb_strerror() explains a code.
 */
#ifndef AZB_H_
#define AZB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZB_VERSION 1

enum {
    AZB_OK = 0,
    AZB_E_NULL = -1,       /* required pointer is NULL */
    AZB_E_ALIGN = -2,      /* pointer/stride not aligned as the kernel needs */
    AZB_E_DTYPE = -3,      /* unsupported dtype code */
    AZB_E_SHAPE = -4,      /* unsupported / inconsistent shape */
    AZB_E_DRIVER = -5,     /* CUDA driver entry point (tensor-map encode) unavailable */
    AZB_E_UNSUPPORTED = -6 /* feature not compiled / not available on this device */
};

/* dtype codes */
enum { AZB_F32 = 0, AZB_BF16 = 1, AZB_F16 = 2, AZB_I64 = 3 };

/* Columns of one row of the per-step coefficient table (float32[steps][AZB_COEF_COLS]).
 * The row is built once per sampler from the schedule itself (azula_b200/engine/table.py)
 * with the reference's own scalar operation order, see DESIGN.md "coefficient table". */
enum {
    AZB_C_SKIP = 0,   /* c_skip(t)                     denoise.py:311 / adm/__init__.py:111 */
    AZB_C_OUT = 1,    /* c_out(t)                      denoise.py:310 / adm/__init__.py:110 */
    AZB_ALPHA_S = 2,  /* alpha_s                       sample.py:212,257                    */
    AZB_K = 3,        /* sigma_s*sqrt(1-tau)/sigma_t   sample.py:213,258                    */
    AZB_ALPHA_T = 4,  /* alpha_t                       sample.py:213,258                    */
    AZB_N = 5,        /* sigma_s*sqrt(tau)             sample.py:214,259                    */
    AZB_C_IN_NEXT = 6,/* c_in(s): pre-scale of the next backbone input (denoise.py:309)     */
    AZB_CLIP = 7,     /* mean clipped to [-clip, clip]; +inf = no clip (adm/__init__.py:133) */
    AZB_COEF_COLS = 8
};

int azb_version(void);
const char* azb_strerror(int code);

/* Number of threads T = 256*grid that ATen's randn kernel would use for `numel` elements on
 * the current device (torch/include/ATen/native/cuda/DistributionTemplates.h:50-62), and the
 * Philox offset increment of one such call.  Host-side helpers, no device work. */
int azb_rng_policy(int64_t numel, int64_t* rng_threads, int64_t* offset_inc);

/*
 * Fused transition q(X_s | X_t): replaces the ~86 ATen dispatches of
 * DDPMSampler.step / DDIMSampler.step (azula/sample.py:204-216,248-261) together with the
 * posterior-mean arithmetic of the denoiser (azula/denoise.py:322, adm/__init__.py:125-134):
 *
 *     m   = clamp(c_skip*x_t + c_out*F, -clip, clip)
 *     x_s = alpha_s*m + k*(x_t - alpha_t*m) + n*eps          (each op rounded as in eager)
 *     x_in_next = (c_in_next * x_s) cast to in_dtype           (optional)
 *
 * row = coef_table[*step_idx].  F element (b, j) is read at f[b*f_batch_stride + j] so the
 * kernel can take the first C of 2C channels of a learned-variance output.  eps is read from
 * `eps` when given; otherwise, when n != 0, it is generated in registers with Philox4x32-10
 * laid out exactly like torch.randn_like launched with `rng_threads` threads (azb_rng_policy;
 * element i of x_t is global element rng_elem_offset + i).  Generator state: when
 * `philox_state` (device int64[2] = {offset, seed}) is given the kernel uses seed
 * philox_state[1] and offset philox_state[0] + offset_host, so a captured graph follows the
 * generator without re-capture; otherwise it uses `seed` and `offset_host`.
 */
int azb_step_f32(const float* x_t, const void* f, int f_dtype, int64_t f_batch_stride,
                 const float* eps, float* x_s, void* x_in_next, int in_dtype,
                 int64_t n_per_sample, int64_t batch, const float* coef_table,
                 const int32_t* step_idx, uint64_t seed, const int64_t* philox_state,
                 int64_t offset_host, int64_t rng_threads, int64_t rng_elem_offset, void* stream);

/* Bookkeeping between two steps of a captured loop: ++*step_idx, philox_state[0] += offset_inc,
 * and (optionally) time_out[0..time_count) = time_table[min(*step_idx, steps-1)][0..time_count)
 * (elements of `time_elem_bytes` bytes) -- the backbone's time input of the next step. */
int azb_advance(int32_t* step_idx, int64_t* philox_state, int64_t offset_inc,
                const void* time_table, void* time_out, int time_elem_bytes, int time_count,
                int32_t steps, void* stream);

/* Sampler.init (azula/sample.py:120-128) for scalar mean/var:
 *   x = mean_T + std_T * eps,  eps laid out like torch.randn_like (see azb_step_f32). */
int azb_init_noise_f32(float* x, int64_t numel, float mean_T, float std_T, uint64_t seed,
                       int64_t offset_host, int64_t rng_threads, int64_t rng_elem_offset,
                       void* stream);

/*
 * Convolution (3x3 pad 1 / 1x1) or linear layer as an implicit GEMM on tcgen05 tensor cores:
 * replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Conv1d(k=1) / nn.Linear on the ADM
 * path (azula/plugins/adm/_src/unet.py:182,207,213-215,277,285,471,602).
 *   act    NHWC bf16, pixel stride act_ld elements (n images of h x w pixels, c_in channels)
 *   wpack  bf16 [c_out_rows][taps][k_per_tap], k_per_tap = c_in rounded up to 64, zero padded;
 *          c_out_rows = c_out rounded up to the N tile (16/32/64/128), zero padded
 *   out    out_mode 0: bf16 NHWC with pixel stride out_ld (+ residual bf16 NHWC, stride res_ld)
 *          out_mode 1: fp32 NCHW [n][c_out][h][w]
 *   bias   fp32 [c_out] or NULL.  A linear layer is taps=1 with h=1, w=rows.
 */
int azb_conv_gemm_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                       const void* wpack, int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap,
                       const float* bias, const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                       int out_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AZB_H_ */
