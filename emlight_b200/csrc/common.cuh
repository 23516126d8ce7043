// Shared helpers for the emlight_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/emlight_b200.h"

#define EML_CHECK_PTR(p) do { if ((p) == nullptr) return EML_E_NULL; } while (0)
#define EML_CHECK_ALIGN16(p) do { if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return EML_E_ALIGN; } while (0)

static inline int eml_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EML_OK : static_cast<int>(e);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Debug / A-B switches read from the environment (e.g. EML_NO_PERSIST=1 selects the non-persistent kernels).
static inline bool eml_env_flag(const char *name) {
    const char *v = getenv(name);
    return v != nullptr && v[0] != '\0' && v[0] != '0';
}
