// Version / error-string / device-check entry points of the C ABI.
#include "common.cuh"

extern "C" int eml_version(void) { return 15; }

extern "C" const char *eml_error_string(int code) {
    switch (code) {
        case EML_OK: return "ok";
        case EML_E_NULL: return "emlight_b200: required pointer is NULL";
        case EML_E_SHAPE: return "emlight_b200: unsupported size or shape";
        case EML_E_ALIGN: return "emlight_b200: pointer or pitch not aligned as documented";
        case EML_E_ARG: return "emlight_b200: invalid scalar argument";
        case EML_E_WORKSPACE: return "emlight_b200: workspace too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "emlight_b200: unknown error code";
}

extern "C" int eml_device_ok(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    return major == 10 ? EML_OK : EML_E_ARG;
}
