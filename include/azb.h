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

#define AZB_VERSION 4

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

/* activation codes of the GEMM epilogue */
enum { AZB_ACT_NONE = 0, AZB_ACT_SILU = 1, AZB_ACT_RELU = 2, AZB_ACT_RELU2 = 3 };

/* row normalisations of azb_rownorm_mod_bf16 */
enum { AZB_NORM_LAYER = 0, AZB_NORM_RMS = 1 };

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

/*
 * The same kernel behind a descriptor, extended to every sampler of azula/sample.py and to classifier-free guidance
 * (azula/guidance/cfg.py:35-65).  One call = one STAGE of a sampler step (one backbone evaluation followed by one
 * update); row = table[*step_idx] holds row_floats floats: the 8 columns above, then (row_floats >= AZB_ROW_COLS)
 *
 *     AZB_R_P, AZB_R_Q     stored quantity   h = p*x_e + q*m                         sample.py:345,349,525,664,...
 *     AZB_R_R              x_out = r*x_b + sum_j W[j]*H[j]   (H[write_slot] = the fresh h)  sample.py:351,528-535
 *     AZB_R_FLAGS (int32)  bit 0 x_e = src[bit], bit 1 x_b = src[bit], bit 2 out = dst[bit], bit 3 history mode,
 *                          bits 4-7 write slot, bit 8 store h into hist[write_slot], bits 12-15 slots in use
 *     AZB_R_DRAW (int32)   number of noise draws of the loop BEFORE this stage (Philox offset = base + draw*offset_inc)
 *     AZB_R_W .. +8        W[0..7]
 *
 * Affine mode (bit 3 clear) is azb_step_f32's arithmetic on x = src[x_e], bit for bit.  History mode serves Heun
 * (stage A: predictor into the alternate state buffer, slope kept in slot 0; stage B: corrector from both slopes,
 * sample.py:337-352) and the Adams-Bashforth family (ring of `order` slopes, weights solved in float64 at table-build
 * time, sample.py:487-508).  With f_neg the posterior mean is the guided one, each branch clipped on its own as the
 * wrapped denoiser would:  m = m+ + guidance[0] * (m+ - m-),  m+- = clamp(c_skip*x + c_out*F+-).
 * x_in_next receives x_in_copies (1 or 2) consecutive copies of c_in_next * x_out (2: the input of a 2B-batch forward
 * evaluating both guidance branches at once).  Loads of the state buffers are coherent: src and dst may alias.
 */
enum { AZB_R_P = 8, AZB_R_Q = 9, AZB_R_R = 10, AZB_R_FLAGS = 12, AZB_R_DRAW = 13, AZB_R_W = 16, AZB_ROW_COLS = 32 };
enum { AZB_STEP_MAX_SLOTS = 8 };

typedef struct AzbStep {
    const float* src[2];    /* state buffers a row may read (legacy rows: src[0])                       */
    float* dst[2];          /* state buffers a row may write (legacy rows: dst[0])                      */
    const void* f;          /* backbone output F (F+ under guidance)                                    */
    const void* f_neg;      /* F- or NULL                                                               */
    const float* guidance;  /* device scalar omega (required with f_neg)                                */
    const float* eps;       /* explicit noise or NULL (affine mode)                                     */
    void* x_in_next;        /* next backbone input or NULL                                              */
    float* hist;            /* history slots, slot j at hist + j*hist_stride (history mode)             */
    const float* table;     /* float32[stages][row_floats]                                              */
    const int32_t* step_idx;
    const int64_t* philox_state; /* device {offset, seed} or NULL (then `seed`, offset_host)            */
    int64_t f_batch_stride, n_per_sample, batch, hist_stride;
    int64_t offset_host, offset_inc, rng_threads, rng_elem_offset;
    uint64_t seed;
    int32_t f_dtype, in_dtype, row_floats, x_in_copies;
    int32_t noise_hint;     /* < 0: the caller guarantees that no row of the table draws noise (n == 0 everywhere or eps
                             * given): the kernel variant without the in-register generator runs (half the registers,
                             * HBM rate); a row with n != 0 then yields NaN.  0: unknown (generator compiled in).       */
    int32_t reserved_;
} AzbStep;

int azb_step_ex_f32(const AzbStep* desc, void* stream);

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

/*
 * Same convolution, additionally writing what the GroupNorm that CONSUMES `out` needs: for every
 * 32-row slab of every 128-pixel M tile and every output channel, {sum, sum of squares} of the stored
 * (bf16-rounded) values: colsum float2[rows][c_out / stat_gran], rows from azb_conv_colsum_rows();
 * stat_gran = 1 (one entry per channel) or 8 (one entry per block of 8 channels: enough whenever the
 * consuming GroupNorm's groups are multiples of 8 channels, 8x less traffic).  Fusing this into the
 * epilogue removes the separate statistics read pass (native_group_norm's first half,
 * azula/plugins/adm/_src/nn.py:80-87).  bf16 NHWC output only.
 */
int azb_conv_gemm_stats_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                             const void* wpack, int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap,
                             const float* bias, const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                             float* colsum, int stat_gran, void* stream);

/*
 * The general form of the convolution / linear entry, for the in-repo backbones
 * (azula/nn/unet.py:75-81,87-95,168-205 ConvNd + SiLU + gated residual; azula/nn/dit.py:86-91,
 * 104-107 and azula/nn/attention.py:101,116-118 linears):
 *
 *     out = residual + gate[sample] * act_fn(conv(act) + bias)
 *
 *   stride      1 or 2 (3x3 pad 1 only): (h, w) are INPUT extents, output = ceil(h/stride) x ceil(w/stride)
 *   act_fn      AZB_ACT_*
 *   gate        fp32, channel c of sample s at gate[s*gate_ld + c] (gate_ld = 0: one shared row);
 *               sample = output pixel index / gate_rows.  NULL = no gate.  bf16 NHWC output only.
 *   colsum      as in azb_conv_gemm_stats_bf16, or NULL.
 */
int azb_conv2d_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                    const void* wpack, int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap,
                    int stride, const float* bias, int act_fn, const float* gate, int64_t gate_ld,
                    int64_t gate_rows, const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                    int out_mode, float* colsum, int stat_gran, void* stream);

/*
 * The tail of an ADM ResBlock whose skip connection is a 1x1 convolution (azula/plugins/adm/_src/unet.py:
 * 213-215,243-247):  out = conv3x3(act) + conv1x1(act2) + bias  as ONE implicit GEMM whose K dimension is
 * [9 taps x k_per_tap | k2]: wpack is bf16 [c_out_rows][9*k_per_tap + k2] (the 1x1 weights appended, k2 = c_in2
 * rounded up to 64), bias = the sum of both biases.  act2 is NHWC bf16 at the same resolution.  Removes the
 * write + read of the skip tensor and one launch.  colsum as in azb_conv_gemm_stats_bf16 (may be NULL).
 */
int azb_conv_skip_stats_bf16(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld,
                             const void* act2, int64_t c_in2, int64_t act2_ld, const void* wpack, int64_t c_out,
                             int64_t c_out_rows, int64_t k_per_tap, int64_t k2, const float* bias, void* out,
                             int64_t out_ld, float* colsum, int stat_gran, void* stream);

/*
 * Every option of the convolution / linear kernel in one descriptor (plain C struct, zero = "not used"):
 *
 *     out = residual + gate[sample] * act_fn(conv(act) [+ conv1x1(act2)] + bias)
 *
 * Fields as in the flat entry points above, plus the exact GroupNorm accumulators: when gn_acc is given, the
 * epilogue adds {sum, sum of squares} of the stored (bf16-rounded) values of every (image, block of stat_gran
 * channels) of `out` into  int64 gn_acc[n][c_out / stat_gran][4] = {sum hi, sum lo, sumsq hi, sumsq lo}
 * (value * 2^40 = hi * 2^32 + lo) with integer atomics: exact, hence independent of the order in which tiles
 * finish (bit reproducible), and no reduction pass or launch is needed before azb_gn_apply_acc_bf16, which folds
 * the blocks of each group -- of one tensor or of a concatenation of two -- in its prologue.  The caller zeroes
 * gn_acc before the producer runs.  Needs every 32-row slab of an M tile inside one image (azb_conv_colsum_rows).
 *
 * With a workspace the kernel may split the K dimension over 2 or 4 CTAs per output tile when the feature map is
 * small and the reduction long (ADM's 8 x 8 layers, K = 9216 .. 18432): helpers park fp32 partial accumulators in
 * the workspace, the owner folds them in a fixed order before its epilogue (deterministic).
 */
typedef struct AzbConv {
    const void* act;
    int64_t n, h, w, c_in, act_ld;
    const void* wpack;
    int64_t c_out, c_out_rows, k_per_tap;
    int32_t taps, stride, act_fn, out_mode, stat_gran;
    int32_t res_up;           /* 1: `residual` is (n, h / 2, w / 2, c_out): added through a nearest-neighbour 2x upsampling */
    const float* bias;
    const float* gate;
    int64_t gate_ld, gate_rows;
    const void* residual;
    int64_t res_ld;
    void* out;
    int64_t out_ld;
    float* colsum;
    const void* act2;
    int64_t c_in2, act2_ld, k2;
    int64_t* gn_acc;
    void* workspace;          /* optional split-K scratch, 256-byte aligned, ZERO-initialised once by the caller (the */
    int64_t workspace_bytes;  /* kernel leaves its flags zeroed); 16 MiB covers every shape.  NULL = never split K    */
    /* Fused input transform (3 x 3, stride 1, feature maps >= 16 x 8, c_in % 64 == 0; AZB_E_UNSUPPORTED otherwise --
     * ask azb_conv_choice first): the convolution reads  bf16(act(a[n][c] * x + b[n][c]))  instead of x, i.e. the
     * GroupNorm (+ scale / shift) + SiLU that precedes it in the reference (_src/unet.py:177-181,203-207,238-243)
     * without a normalisation pass over HBM.  in_coef: fp32 [n][c_in][2] from azb_gn_coef_f32; in_silu as given there. */
    const float* in_coef;
    int32_t in_silu;
    /* in_up = 1 (needs in_coef): `act` is (n, h / 2, w / 2, c_in) and the convolution reads its nearest-neighbour 2x
     * upsampling, (h, w) being the upsampled extents: conv(up(act(A x + B))) of an upsampling ResBlock
     * (_src/unet.py:101-109,229-233) without the 4x larger intermediate. */
    int32_t in_up;
    /* Fused per-pixel normalisation of the input (same conditions as in_coef, instead of it): the convolution reads
     *     bf16((1 + a[n][c]) * norm_C(x)[pixel] + b[n][c])
     * i.e. the LayerNorm / RMSNorm over the channels of every pixel and the Ada-Norm-Zero modulation that open
     * UNetBlock._forward (azula/nn/unet.py:97-104, azula/nn/layers.py LayerNorm / RMSNorm) without a pass over HBM.
     * in_norm: 0 none, 1 LayerNorm, 2 RMSNorm (eps = in_eps); in_rowstat: fp32 [pixels][c_in / 64][2] = {sum, sum of
     * squares} of every 64-channel block of x, as written by the producer of x through `rowstat`; in_mod: fp32
     * [a(c_in) | b(c_in)] per sample, in_mod_ld floats between samples (0 = one shared row). */
    int32_t in_norm;
    float in_eps;
    const float* in_rowstat;
    const float* in_mod;
    int64_t in_mod_ld;
    /* rowstat (c_out % 64 == 0; needs the row-domain epilogue, AzbConvChoice.epi == 2, AZB_E_UNSUPPORTED otherwise): the
     * epilogue also writes {sum, sum of squares} of the stored (bf16-rounded) values of every (pixel, 64-channel block) of
     * `out` to fp32 rowstat[pixels][c_out / 64][2]. */
    float* rowstat;
    /* out_up = 1 (row-domain epilogue, N tile >= 128, c_out % 64 == 0; AZB_E_UNSUPPORTED otherwise): `out` is
     * (n, 2 h, 2 w, c_out) with pixel stride out_ld and receives the nearest-neighbour 2x upsampling of the result -- the
     * nn.Upsample that follows the last block of an ascent level (azula/nn/unet.py:186-190) without its own pass. */
    int32_t out_up;
    int32_t reserved_;
} AzbConv;

int azb_conv_bf16(const AzbConv* desc, void* stream);

/* What the launcher would do for `desc` (nothing is launched, no pointer is dereferenced): halo = 1 when the 3 x 3
 * operand is staged as halo tiles (the condition for in_coef), pair / lean / block_n / splits as described above. */
typedef struct AzbConvChoice {
    int32_t halo, pair, lean, block_n, splits, tiles;
    int32_t epi;  /* epilogue: 0 transposing (generic), 1 lean, 2 row domain with TMA stores */
} AzbConvChoice;
int azb_conv_choice(const AzbConv* desc, AzbConvChoice* choice);

/* Tuning hooks of the launchers (process-wide, not thread-safe; meant for A/B measurements).
 * value -1 restores the automatic choice.
 *   AZB_CONV_KNOB_PAIR      0: never use CTA pairs, 1: whenever the shape allows (even number of 128-pixel tiles,
 *                           N tile >= 128, no split-K), -1: when the reduction is long enough to pay
 *   AZB_CONV_KNOB_PREFETCH  k-blocks of weight prefetch into L2 (0 = off)
 *   AZB_CONV_KNOB_SPLITK    0: never split K even when a workspace is given
 *   AZB_GN_KNOB_WAVE        GroupNorm apply: resident CTAs per SM of the single-wave grid (0 = short CTAs of 16
 *                           vectors per thread, the pre-wave policy)
 *   AZB_CONV_KNOB_BLOCKN    force the N tile (16 .. 256; ignored unless it divides the padded C_out)
 *   AZB_CONV_KNOB_LEAN      0: always the generic epilogue (all switches at run time)
 *   AZB_CONV_KNOB_HALO      0: never stage 3 x 3 operands as halo tiles (tap-wise TMA loads instead), 1: wherever the shape
 *                           allows, -1: where an input transform needs them or the feature map has >= 32 patches of 8 x 16 */
#define AZB_CONV_KNOB_PAIR 0
#define AZB_CONV_KNOB_PREFETCH 1
#define AZB_CONV_KNOB_SPLITK 2
#define AZB_GN_KNOB_WAVE 3
#define AZB_CONV_KNOB_BLOCKN 4
#define AZB_CONV_KNOB_LEAN 5
#define AZB_CONV_KNOB_HALO 6
#define AZB_CONV_KNOB_HALO_SA 7    /* halo kernels: A slots (2 .. 4; the rest of the ring holds weight stages) */
#define AZB_CONV_KNOB_HALO_SB 8    /* halo kernels: cap on the weight stages */
#define AZB_CONV_KNOB_HALO_SPREAD 9 /* halo kernels: 0 = the fused 1 x 1 blocks follow the last halo item (default: spread) */
#define AZB_KNOB_PDL 10 /* 0: plain stream-ordered launches instead of programmatic dependent launches */
#define AZB_CONV_KNOB_ROWEPI 11 /* row-domain epilogue with TMA stores: 0 never, 1 / -1 (default) wherever possible */
#define AZB_CONV_KNOBS 12
int azb_conv_tuning(int knob, int value);
/* Diagnostics (scripts/graph_trace.py): `buf` = device array of uint64 {launch counter, 7 unused, then 4 words per
 * convolution launch in stream order: earliest CTA entry, earliest CTA start after the programmatic-launch wait, latest
 * CTA end (%globaltimer ns; initialise mins to ~0, max to 0), unused} or NULL to switch tracing off.  Launch descriptors
 * built (or graphs captured) while a buffer is set carry it; no reference counterpart. */
int azb_debug_trace(void* buf);

/*
 * The REFERENCE-NUMERICS mode: the same contraction with fp32 operands in HBM and tcgen05.mma.kind::tf32 (10-bit operand
 * mantissa, fp32 accumulation) -- what cuDNN runs for the reference's fp32 modules under PyTorch's default flags
 * (torch.backends.cudnn.allow_tf32; azula/plugins/adm/_src/unet.py:182,207,213-215,277,285,471,602).
 *   act    fp32 NHWC, pixel stride act_ld floats; c_in % 4 == 0 (the 3-channel network input is padded to 4)
 *   wpack  fp32 [c_out_rows][taps][k_per_tap], k_per_tap = c_in rounded up to 32, zero padded
 *   out    out_mode 0: fp32 NHWC (pixel stride out_ld), or fp16 NHWC when out_f16 (operands of azb_attention_f16);
 *          out_mode 1: fp32 NCHW.  residual: fp32 NHWC (stride res_ld) or NULL; bias fp32 or NULL; act_fn AZB_ACT_*.
 * (h, w) are the INPUT extents; stride 1 or 2.
 */
int azb_conv_tf32(const void* act, int64_t n, int64_t h, int64_t w, int64_t c_in, int64_t act_ld, const void* wpack,
                  int64_t c_out, int64_t c_out_rows, int taps, int64_t k_per_tap, int stride, const float* bias, int act_fn,
                  const void* residual, int64_t res_ld, void* out, int64_t out_ld, int out_mode, int out_f16, void* stream);

/* GroupNorm32 of the reference-numerics mode, fp32 NHWC in and out (azula/plugins/adm/_src/nn.py:80-87 and the uses in
 * _src/unet.py:177-181,203-207,229-243,276,599-601): statistics {mean, rstd} per (image, group), then
 *   y = act(((x - mean) rstd gamma + beta) (1 + scale) + shift)   followed by nearest 2x upsampling (mode 1) or 2 x 2
 * average pooling (mode 2) of the result; stats NULL = resampling only.  scale_shift: [scale(c) | shift(c)] per sample. */
int azb_gn_stats_f32(const float* x, int64_t ld, int64_t n, int64_t hw, int64_t c, int64_t groups, float eps, float* stats,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* (workspace: optional scratch, 256-byte aligned, ZERO-initialised once by the caller and left zeroed by the kernel --
 * 256-byte-rounded n * groups * 4 bytes of arrival counters + n * groups * 16 * 16 bytes of fp64 partial sums let large
 * maps be reduced by up to 16 CTAs per (image, group), folded in a fixed order by the last one to arrive.) */
int azb_gn_apply_f32(const float* x, int64_t x_ld, float* y, int64_t y_ld, int64_t n, int64_t h, int64_t w, int64_t c,
                     int64_t groups, const float* stats, const float* gamma, const float* beta, const float* scale_shift,
                     int64_t ss_stride, int silu, int mode, void* stream);
/* fp32 (n, c, h, w) -> fp32 (n, h, w, c_pad) with zero padded channels (the network input of the TF32 mode). */
int azb_nchw_to_nhwc_f32(const float* x, float* y, int64_t n, int64_t c, int64_t h, int64_t w, int64_t c_pad, void* stream);
/* azb_attention_bf16 with fp16 operands (N, T, ld) and an fp32 result (N, T, out_ld): the reference-numerics mode's
 * QKVAttention (_src/unet.py:328-345,361-381); d in {32, 64, 128, 256}. */
int azb_attention_f16(const void* qkv, int64_t ld, float* out, int64_t out_ld, int64_t n, int64_t t, int64_t heads, int64_t d,
                      int64_t head_stride, int64_t k_delta, int64_t v_delta, void* stream);

/* Rows of the colsum buffer for an (n, h, w) activation; *slab_in_image = 1 when every 32-row slab
 * lies inside one image (the condition for azb_gn_finalize_f32), else 0.  Host-side helper. */
int azb_conv_colsum_rows(int64_t n, int64_t h, int64_t w, int64_t* rows, int64_t* slab_in_image);

/* stats[n][g] = {mean, rstd} from the column sums of one or two convolutions whose outputs form the
 * channel ranges [0, c_a) and [c_a, c_a + c_b) of the normalised tensor (a decoder concatenation,
 * azula/plugins/adm/_src/unet.py:631); colsum_b may be NULL with c_b = 0.  Deterministic. */
int azb_gn_finalize_f32(const float* colsum_a, int64_t c_a, int gran_a, const float* colsum_b, int64_t c_b, int gran_b,
                        int64_t n, int64_t h, int64_t w, int64_t groups, float eps, float* stats, void* stream);

/*
 * GroupNorm statistics over NHWC bf16 (N, HW, C), pixel stride ld: stats[n][g] = {mean, rstd}.
 * With azb_gn_apply_bf16 this replaces native_group_norm + SiLU + scale/shift + resampling
 * (azula/plugins/adm/_src/nn.py:80-87, _src/unet.py:177-181,203-207,229-243,276,599-601).
 * `partial` is a float workspace of azb_gn_stats_workspace() elements; `counters` int32[n],
 * zero-initialised once (the kernel leaves it zeroed).  Deterministic (no float atomics).
 */
int azb_gn_stats_workspace(int64_t n, int64_t hw, int64_t c, int64_t groups, int64_t* partial_floats);
int azb_gn_stats_bf16(const void* x, int64_t ld, int64_t n, int64_t hw, int64_t c, int64_t groups, float eps,
                      float* partial, float* stats, int32_t* counters, void* stream);

/*
 * y = act((x - mean) * rstd * gamma + beta) * (1 + scale) + shift ... folded to act(A*x + B):
 *   stats NULL          -> identity transform (used to resample the residual branch)
 *   scale_shift         -> fp32 rows [scale(C) | shift(C)]; image n uses row n*ss_stride (0 = shared);
 *                          when ss_step is given, *ss_step * ss_step_stride is added (per-step table)
 *   silu                -> SiLU after the affine map
 *   mode                -> 0 same size, 1 nearest x2 upsample, 2 2x2 average pool (of the activated values)
 * x (N,H,W,C) stride x_ld -> y (N,Ho,Wo,C) stride y_ld, both bf16 NHWC.
 */
int azb_gn_apply_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t n, int64_t h, int64_t w,
                      int64_t c, int64_t groups, const float* stats, const float* gamma, const float* beta,
                      const float* scale_shift, int64_t ss_stride, const int32_t* ss_step, int64_t ss_step_stride,
                      int silu, int mode, void* stream);

/* azb_gn_apply_bf16 with the statistics taken from the exact accumulators written by azb_conv_bf16 (gn_acc, see
 * AzbConv): x is the concatenation of channel ranges [0, c_a) and [c_a, c_a + c_b) whose producers accumulated
 * into acc_a and acc_b (acc_b NULL, c_b = 0 for a single tensor), `gran` channels per accumulator entry.  Mean and
 * rstd of every group are derived in the kernel's prologue in double precision: GroupNorm without a statistics
 * pass and without a reduction launch (azula/plugins/adm/_src/nn.py:80-87). */
int azb_gn_apply_acc_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t n, int64_t h, int64_t w,
                          int64_t c, int64_t groups, const int64_t* acc_a, int64_t c_a, const int64_t* acc_b,
                          int64_t c_b, int64_t gran, float eps, const float* gamma, const float* beta,
                          const float* scale_shift, int64_t ss_stride, int silu, int mode, void* stream);

/* azb_gn_apply_acc_bf16 in pooling mode (2) that reads x ONCE and writes both branches of a downsampling ResBlock
 * (_src/unet.py:229-233): y = avgpool2x2(act(A x + B)) and y_raw = avgpool2x2(x), both (n, h / 2, w / 2, c). */
int azb_gn_pool_acc_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, void* y_raw, int64_t y_raw_ld, int64_t n,
                         int64_t h, int64_t w, int64_t c, int64_t groups, const int64_t* acc_a, int64_t c_a,
                         const int64_t* acc_b, int64_t c_b, int64_t gran, float eps, const float* gamma, const float* beta,
                         const float* scale_shift, int64_t ss_stride, int silu, void* stream);

/* The per-(image, channel) coefficients {A, B} of azb_gn_apply_acc_bf16's transform y = act(A x + B), written to
 * coef fp32 [n][c][2] for a convolution that applies it to its input on the fly (AzbConv::in_coef): GroupNorm +
 * scale / shift + SiLU cost one launch over n * c values instead of a pass over the activation.  With silu the
 * pair is halved (SiLU(2h) = h + h tanh(h)); pass the same flag as AzbConv::in_silu. */
int azb_gn_coef_f32(int64_t n, int64_t h, int64_t w, int64_t c, int64_t groups, const int64_t* acc_a, int64_t c_a,
                    const int64_t* acc_b, int64_t c_b, int64_t gran, float eps, const float* gamma, const float* beta,
                    const float* scale_shift, int64_t ss_stride, int silu, float* coef, void* stream);

/*
 * softmax(q k^T / sqrt(d)) v per (image, head) without materialising the T x T logits
 * (QKVAttentionLegacy / QKVAttention, azula/plugins/adm/_src/unet.py:328-345,361-381).
 * qkv (N, T, ld) bf16; head hd reads q/k/v at channel hd*head_stride + {0, k_delta, v_delta};
 * out (N, T, out_ld) bf16, head hd at channel hd*d.  d in {16, 32, 64, 128}.
 */
int azb_attention_bf16(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t,
                       int64_t heads, int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta,
                       void* stream);
/* The same with the per-head RMS normalisation of q and k (torch.nn.RMSNorm(elementwise_affine = False), eps = qk_eps;
 * azula/nn/attention.py:103) folded into the logits: s'_ij = s_ij rq_i rk_j.  Replaces the in-place pass
 * azb_segment_rmsnorm_bf16 for sequences of T <= 256 tokens and head width 64; AZB_E_UNSUPPORTED otherwise. */
int azb_attention_qknorm_bf16(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t, int64_t heads,
                              int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta, float qk_eps, void* stream);

/* The same contract served by the warp-level mma.sync kernel only (azb_attention_bf16 uses the tcgen05 / TMEM
 * kernel for d = 64 and this one for the other widths); exported so that tests can compare the two. */
int azb_attention_mma_bf16(const void* qkv, int64_t ld, void* out, int64_t out_ld, int64_t n, int64_t t,
                           int64_t heads, int64_t d, int64_t head_stride, int64_t k_delta, int64_t v_delta,
                           void* stream);

/* im2col of the fp32 NCHW network input for the first 3x3 conv (_src/unet.py:471):
 * out[n][h][w][k] bf16, k = (kh*3+kw)*c + ch for k < 9c, zero up to k_pad. */
int azb_im2col3x3_f32(const float* x, void* out, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k_pad,
                      void* stream);

/* Sinusoidal timestep features [cos | sin] (_src/nn.py:90-108); t is int64 or fp32 [rows]. */
int azb_timestep_features_f32(const void* t, int t_dtype, int64_t rows, int64_t dim, float max_period,
                              float* out, void* stream);

/* y[m][n] = b[n] + sum_k act(x[m][k]) w[n][k] in fp32 (act = SiLU when silu_in): the time-embedding
 * MLP and the per-block emb_layers (_src/unet.py:458-462,198-204). */
int azb_linear_f32(const float* x, const float* w, const float* b, float* y, int64_t m, int64_t n, int64_t k,
                   int silu_in, void* stream);

/* Zeroes `bytes` bytes on the stream (a memset node when captured): clears the GroupNorm accumulators of a
 * forward pass before its first producer runs. */
int azb_zero_bytes(void* ptr, int64_t bytes, void* stream);

/* y[r][:] += table[idx[r]][:] (class-label embedding, _src/unet.py:621-623). */
int azb_add_rows_f32(float* y, const float* table, const int64_t* idx, int64_t rows, int64_t dim, void* stream);

/*
 * y[r][c] = (1 + a[s][c]) * norm(x[r])[c] + b[s][c] over rows of C contiguous bf16 channels (row stride
 * x_ld / y_ld elements), fp32 arithmetic: the Ada-Norm-Zero prologue of UNetBlock (azula/nn/unet.py:
 * 99-104 with azula/nn/layers.py:152-155: standardisation over the channel dimension with the UNBIASED
 * variance of torch.var_mean, no affine) and of DiTBlock (azula/nn/dit.py:102-103 with
 * torch.nn.RMSNorm(elementwise_affine=False)).  mod is fp32 with [a(C) | b(C) | ...] per sample,
 * sample s = r / rows_per_sample at mod + s*mod_ld (mod_ld = 0: shared); mod NULL = plain normalisation.
 */
int azb_rownorm_mod_bf16(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t rows, int64_t c, int kind,
                         float eps, const float* mod, int64_t mod_ld, int64_t rows_per_sample, void* stream);

/* In-place RMS normalisation (no affine) of `segs` contiguous segments of d channels per row, starting at
 * channel 0 of x (rows x ld, bf16): the query-key normalisation of MultiheadSelfAttention
 * (azula/nn/attention.py:103 on the q and k thirds of the qkv projection: segs = 2 * heads). */
int azb_segment_rmsnorm_bf16(void* x, int64_t ld, int64_t rows, int64_t segs, int64_t d, float eps, void* stream);

/* The same in-place pass over the q and k thirds of a qkv projection (segments of d channels, 2 * heads of them), with the
 * rotary positional embedding of MultiheadSelfAttention (azula/nn/attention.py:105-108,124-156) applied after the
 * normalisation (norm = 0: rotation only): channel pair (2 i, 2 i + 1) of head h at token l turns by the angle whose
 * {cos, sin} (fp32) is rot[(l mod rows_per_sample)][h * d / 2 + i]; rot NULL = azb_segment_rmsnorm_bf16. */
int azb_qk_norm_rope_bf16(void* x, int64_t ld, int64_t rows, int64_t heads, int64_t d, int norm, float eps,
                          const float* rot, int64_t rows_per_sample, void* stream);

/* Patchify(channel_last) of azula/nn/vit.py:97 (azula/nn/layers.py:198-222): fp32 NCHW (n, c, hp*p, wp*q)
 * -> bf16 tokens (n*hp*wp, k_pad), token (i, j) channel (z*p + a)*q + b = x[n][z][i*p + a][j*q + b], zero
 * padded to k_pad channels. */
int azb_patchify_f32(const float* x, void* tokens, int64_t n, int64_t c, int64_t hp, int64_t wp, int64_t p,
                     int64_t q, int64_t k_pad, void* stream);

/* Unpatchify(channel_last) of azula/nn/vit.py:105 from the channel-major fp32 GEMM output
 * yt[(z*p + a)*q + b][token] (out_mode 1 of the convolution entry) -> fp32 NCHW (n, c, hp*p, wp*q). */
int azb_unpatchify_f32(const float* yt, float* out, int64_t n, int64_t c, int64_t hp, int64_t wp, int64_t p,
                       int64_t q, void* stream);

/* y[m][j] = b[j] + sum_k act(x[m][xoff[j] + k]) w[j][k]: azb_linear_f32 whose output column j reads its K
 * inputs at column offset xoff[j] of a row of x (row stride x_ld): every block's second Ada-Norm-Zero
 * linear in ONE launch (azula/nn/unet.py:65-70, azula/nn/dit.py:57-63).  xoff NULL = 0. */
int azb_linear_gather_f32(const float* x, int64_t x_ld, const int32_t* xoff, const float* w, const float* b,
                          float* y, int64_t m, int64_t n, int64_t k, int silu_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AZB_H_ */
