// Library-level entry points of the C ABI (include/azb.h).
#include "common.cuh"

extern "C" int azb_version(void) { return AZB_VERSION; }

// Tuning knobs (azb_conv_tuning): -1 = automatic.
int azb_knob[AZB_CONV_KNOBS] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};

extern "C" int azb_conv_tuning(int knob, int value) {
    if (knob < 0 || knob >= AZB_CONV_KNOBS) return AZB_E_SHAPE;
    azb_knob[knob] = value;
    return AZB_OK;
}

// Diagnostics: device buffer of 64-bit words {launch counter, then per convolution launch: earliest CTA entry, earliest
// CTA start (after the programmatic-launch wait), latest CTA end, CTAs} in %globaltimer nanoseconds; see azb.h.
void* azb_trace_buf = nullptr;
extern "C" int azb_debug_trace(void* buf) {
    azb_trace_buf = buf;
    return AZB_OK;
}

extern "C" const char* azb_strerror(int code) {
    switch (code) {
        case AZB_OK: return "ok";
        case AZB_E_NULL: return "azb: required pointer is NULL";
        case AZB_E_ALIGN: return "azb: pointer or stride is not aligned as required";
        case AZB_E_DTYPE: return "azb: unsupported dtype code";
        case AZB_E_SHAPE: return "azb: unsupported or inconsistent shape";
        case AZB_E_DRIVER: return "azb: CUDA driver entry point unavailable";
        case AZB_E_UNSUPPORTED: return "azb: not supported on this device/build";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "azb: unknown error";
}
