// Version / error-string / device-check entry points of the C ABI.
#include "common.cuh"

extern "C" int eml_version(void) { return 23; }

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

// ------------------------------------------------------------------------------------------------ fused optimiser step
// torch.optim.Adam (amsgrad=False, weight_decay=0) over ONE flat fp32 buffer that holds every parameter of a network
// (RegressionNetwork/train.py:55-57: Adam(lr 1e-4, betas (0.9, 0.999)); GenProjector/models/pix2pix_model.py:56-70: betas (0, 0.9)),
// with the 1/world_size of the gradient all-reduce folded in (grad_scale): one pass over p, g, m, v instead of torch's ~10
// multi-tensor launches plus a separate divide.  Same operation order as torch's single-tensor Adam:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g g;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
namespace {
__global__ void __launch_bounds__(256) adam_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                        float *__restrict__ v, long n4, long n, float b1, float b2, float eps,
                                                        float step_size, float inv_bc2_sqrt, float gscale) {
    const long stride = static_cast<long>(gridDim.x) * blockDim.x;
    for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4 *>(p)[i], M = reinterpret_cast<float4 *>(m)[i], V = reinterpret_cast<float4 *>(v)[i];
        const float4 G = reinterpret_cast<const float4 *>(g)[i];
        float *pp = &P.x, *mm = &M.x, *vv = &V.x;
        const float *gg = &G.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gr = gg[e] * gscale;
            mm[e] = fmaf(mm[e], b1, (1.f - b1) * gr);                     // lerp(m, g, 1 - b1) as torch computes it: m + (g - m)(1 - b1) differs in
            vv[e] = fmaf(vv[e], b2, (1.f - b2) * gr * gr);                // the last bit only; the test bounds the trajectory difference
            pp[e] -= step_size * (mm[e] / (sqrtf(vv[e]) * inv_bc2_sqrt + eps));
        }
        reinterpret_cast<float4 *>(p)[i] = P; reinterpret_cast<float4 *>(m)[i] = M; reinterpret_cast<float4 *>(v)[i] = V;
    }
    if (blockIdx.x == 0 && threadIdx.x < static_cast<unsigned>(n - 4 * n4)) {            // tail (n % 4 elements)
        const long i = 4 * n4 + threadIdx.x;
        const float gr = g[i] * gscale;
        m[i] = fmaf(m[i], b1, (1.f - b1) * gr);
        v[i] = fmaf(v[i], b2, (1.f - b2) * gr * gr);
        p[i] -= step_size * (m[i] / (sqrtf(v[i]) * inv_bc2_sqrt + eps));
    }
}
}  // namespace

extern "C" int eml_adam_step(float *p, const float *g, float *m, float *v, long n, float lr, float beta1, float beta2, float eps,
                             int step, float grad_scale, void *stream) {
    EML_CHECK_PTR(p); EML_CHECK_PTR(g); EML_CHECK_PTR(m); EML_CHECK_PTR(v);
    EML_CHECK_ALIGN16(p); EML_CHECK_ALIGN16(g); EML_CHECK_ALIGN16(m); EML_CHECK_ALIGN16(v);
    if (n <= 0 || step <= 0 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f)) return EML_E_ARG;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    const long n4 = n / 4;
    long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adam_flat_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, n4, n, beta1, beta2, eps, static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)), grad_scale);
    return eml_launch_status();
}
