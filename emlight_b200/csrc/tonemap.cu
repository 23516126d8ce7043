// HDR tonemapping with a per-image percentile, sm_100a.
// Replaces RegressionNetwork/util.py:36-66 `TonemapHDR.__call__`: p = x^(1/gamma); r = percentile_q of the strictly positive p
// (numpy's linear interpolation between the two neighbouring order statistics); alpha = max_mapping / (r + 1e-10);
// out = clip(alpha * p, 0, 1).  The percentile is an exact selection, not a sort: one CTA per image runs an MSB-first radix select
// (4 passes of 8 bits over the image's own values, which stay L2-resident) on the bit patterns of the positive floats, twice when
// the percentile falls between two order statistics.
#include "common.cuh"

namespace {

constexpr int T_THREADS = 1024;

__global__ void __launch_bounds__(256) tonemap_power_kernel(const float *__restrict__ x, float *__restrict__ p, long n, float inv_gamma,
                                                            int use_gamma) {
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
        p[i] = use_gamma ? powf(x[i], inv_gamma) : x[i];
}

// k-th smallest (0-based) of the positive values of v[0..n): radix select over the uint32 patterns (monotonic for positive floats).
__device__ float select_kth(const float *__restrict__ v, long n, long k, unsigned *s_hist, unsigned *s_state) {
    unsigned prefix = 0, mask = 0;
    long rank = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += T_THREADS) s_hist[i] = 0;
        __syncthreads();
        for (long i = threadIdx.x; i < n; i += T_THREADS) {
            const float f = v[i];
            if (f > 0.f) {
                const unsigned u = __float_as_uint(f);
                if ((u & mask) == prefix) atomicAdd(&s_hist[(u >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long acc = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (acc + s_hist[b] > rank) break;
                acc += s_hist[b];
            }
            s_state[0] = static_cast<unsigned>(b);
            s_state[1] = static_cast<unsigned>(acc);
        }
        __syncthreads();
        prefix |= s_state[0] << shift;
        mask |= 255u << shift;
        rank -= s_state[1];
        __syncthreads();
    }
    return __uint_as_float(prefix);
}

__global__ void __launch_bounds__(T_THREADS) tonemap_alpha_kernel(const float *__restrict__ p, long per_image, float q, float max_mapping,
                                                                  float *__restrict__ alpha) {
    __shared__ unsigned s_hist[256], s_state[2], s_cnt[32];
    const float *v = p + blockIdx.x * per_image;
    unsigned c = 0;
    for (long i = threadIdx.x; i < per_image; i += T_THREADS) c += v[i] > 0.f ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    long n = 0;
    for (int w = 0; w < T_THREADS / 32; ++w) n += s_cnt[w];
    __syncthreads();
    double r = 0.0;                                                 // no positive value: every p is 0 (or invalid) -> percentile 0
    if (n > 0) {
        const double pos = (n - 1) * static_cast<double>(q) / 100.0;    // numpy 'linear': virtual index into the sorted positives
        const long lo = static_cast<long>(floor(pos)), hi = static_cast<long>(ceil(pos));
        const double vlo = select_kth(v, per_image, lo, s_hist, s_state);
        const double vhi = hi == lo ? vlo : static_cast<double>(select_kth(v, per_image, hi, s_hist, s_state));
        r = vlo + (vhi - vlo) * (pos - lo);
    }
    if (threadIdx.x == 0) alpha[blockIdx.x] = static_cast<float>(max_mapping / (r + 1e-10));
}

__global__ void __launch_bounds__(256) tonemap_scale_kernel(float *__restrict__ p, long per_image, long n, const float *__restrict__ alpha,
                                                            int clip) {
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
        float v = alpha[i / per_image] * p[i];
        if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
        p[i] = v;
    }
}

}  // namespace

extern "C" int eml_tonemap_hdr(const float *x, float *out, float *alpha, int B, long per_image, float gamma, float percentile,
                               float max_mapping, int use_gamma, int clip, int alpha_given, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out); EML_CHECK_PTR(alpha);
    if (B <= 0 || per_image <= 0) return EML_E_SHAPE;
    if (!(gamma > 0.f) || !(percentile >= 0.f && percentile <= 100.f)) return EML_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long n = B * per_image;
    const long blocks = (n + 255) / 256;
    const unsigned grid = static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16);
    tonemap_power_kernel<<<grid, 256, 0, st>>>(x, out, n, 1.f / gamma, use_gamma);
    if (!alpha_given) tonemap_alpha_kernel<<<static_cast<unsigned>(B), T_THREADS, 0, st>>>(out, per_image, percentile, max_mapping, alpha);
    tonemap_scale_kernel<<<grid, 256, 0, st>>>(out, per_image, n, alpha, clip);
    return eml_launch_status();
}
