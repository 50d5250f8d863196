// Library-level entry points of the C ABI (include/azb.h).
#include "common.cuh"

extern "C" int azb_version(void) { return AZB_VERSION; }

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
