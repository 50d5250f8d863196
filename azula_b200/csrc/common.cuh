// Shared device/host helpers for libazb (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/azb.h"

#define AZB_CHECK_PTR(p) \
    do {                 \
        if ((p) == nullptr) return AZB_E_NULL; \
    } while (0)

extern int azb_knob[AZB_CONV_KNOBS];  // api.cu
extern void* azb_trace_buf;           // api.cu (azb_debug_trace)
__device__ __forceinline__ unsigned long long azb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

static inline int azb_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? AZB_OK : (int)e;
}

// Programmatic dependent launch: every kernel of the sampling step is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs are scheduled (and run their prologue) while the
// previous kernel of the stream drains, instead of after a full launch gap -- ~250 dependent launches per step.  The
// kernel side of the contract: pdl_trigger() first thing (lets the NEXT kernel begin launching once all CTAs of this
// one have started), pdl_wait() before the first access to global memory (waits until the previous kernel has completed
// and its writes are visible).  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_trigger();
    pdl_wait();
}

template <typename... KArgs, typename... Args>
static inline int azb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = azb_knob[AZB_KNOB_PDL] != 0 ? 1 : 0;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    return e == cudaSuccess ? AZB_OK : (int)e;
}

// Per-device host-side caches.  cudaFuncSetAttribute, SM counts and cluster occupancy belong to ONE device: a process
// that drives several GPUs (or moves from one to another) must not reuse the first device's values.  Entries are
// idempotent (every thread computes the same value), so unsynchronised concurrent first use is benign.
constexpr int AZB_MAX_DEVICES = 64;
static inline int azb_current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= AZB_MAX_DEVICES) return 0;
    return d;
}
template <typename T>
struct AzbPerDevice {
    T v[AZB_MAX_DEVICES] = {};
    T& get() { return v[azb_current_device()]; }
};
static inline int azb_sm_count() {
    static AzbPerDevice<int> sms;
    int& n = sms.get();
    if (!n) {
        int dev = azb_current_device();
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

static inline bool azb_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Streaming 128-bit accesses: data touched once per step, keep it out of L1.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream2u(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Coherent (not .nc) streaming load: for kernels that may run in place (output aliases the input).
__device__ __forceinline__ float4 ld_stream4_coherent(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    // plain (coherent) streaming load: the apply pass may run in place (y == x)
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t lo16) { return __uint_as_float(lo16 << 16); }
__device__ __forceinline__ float tanh_approx(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
