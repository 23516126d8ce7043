// Ground-truth light parameters from an HDR panorama (the inverse of the SG render), sm_100a.
// Replaces RegressionNetwork/representation/distribution_representation.py:89-119 `extract_mesh.compute`: steradian-weight the
// panorama, threshold at 5 % of the brightest weighted intensity, sum the lit pixels into their nearest anchor (LUT built by the host
// exactly like :77-86), everything else into the ambient term, then distribution / intensity / rgb_ratio.
// One CTA per panorama: pass 1 = block max, pass 2 = shared-memory double accumulators (ln x 3 + 3), pass 3 = the few divisions.
#include "common.cuh"

namespace {

constexpr int X_THREADS = 1024;
constexpr int X_MAX_LN = 512;

__global__ void __launch_bounds__(X_THREADS) extract_params_kernel(const float *__restrict__ hdr, const int *__restrict__ idx,
                                                                   const double *__restrict__ ster, int H, int W, int ln,
                                                                   float *__restrict__ dist, float *__restrict__ intensity,
                                                                   float *__restrict__ rgb_ratio, float *__restrict__ ambient,
                                                                   unsigned char *__restrict__ map) {
    __shared__ double s_acc[X_MAX_LN * 3 + 3];
    __shared__ double s_red[32];
    __shared__ double s_max, s_esum;
    const int b = blockIdx.x, tid = threadIdx.x, P = H * W;
    const float *img = hdr + static_cast<long>(b) * P * 3;
    for (int i = tid; i < ln * 3 + 3; i += X_THREADS) s_acc[i] = 0.0;
    // ---- pass 1: brightest steradian-weighted intensity (0.3 R + 0.59 G + 0.11 B, :91-94)
    double mx = -1e300;
    for (int p = tid; p < P; p += X_THREADS) {
        const double s = ster[p / W];
        const double it = 0.3 * (s * img[3 * p]) + 0.59 * (s * img[3 * p + 1]) + 0.11 * (s * img[3 * p + 2]);
        mx = fmax(mx, it);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        double m = s_red[0];
        for (int w = 1; w < X_THREADS / 32; ++w) m = fmax(m, s_red[w]);
        s_max = m;
    }
    __syncthreads();
    const double thr = s_max * 0.05;
    // ---- pass 2: lit pixels -> nearest anchor, the rest -> ambient (:95-107)
    for (int p = tid; p < P; p += X_THREADS) {
        const double s = ster[p / W];
        const double r = s * img[3 * p], g = s * img[3 * p + 1], bl = s * img[3 * p + 2];
        const bool lit = 0.3 * r + 0.59 * g + 0.11 * bl > thr;
        if (map != nullptr) map[static_cast<long>(b) * P + p] = lit ? 1 : 0;
        double *dst = lit ? &s_acc[idx[p] * 3] : &s_acc[ln * 3];
        atomicAdd(dst, r); atomicAdd(dst + 1, g); atomicAdd(dst + 2, bl);
    }
    __syncthreads();
    // ---- pass 3: distribution = anchor energy / total energy, intensity = |sum of anchors|, rgb_ratio (:109-113)
    double e = 0.0, cr = 0.0, cg = 0.0, cb = 0.0;
    for (int k = tid; k < ln; k += X_THREADS) {
        e += 0.3 * s_acc[3 * k] + 0.59 * s_acc[3 * k + 1] + 0.11 * s_acc[3 * k + 2];
        cr += s_acc[3 * k]; cg += s_acc[3 * k + 1]; cb += s_acc[3 * k + 2];
    }
    __shared__ double s_part[4][32];
    e = warp_sum_d(e); cr = warp_sum_d(cr); cg = warp_sum_d(cg); cb = warp_sum_d(cb);
    if ((tid & 31) == 0) { s_part[0][tid >> 5] = e; s_part[1][tid >> 5] = cr; s_part[2][tid >> 5] = cg; s_part[3][tid >> 5] = cb; }
    __syncthreads();
    if (tid == 0) {
        double t[4] = {0, 0, 0, 0};
        for (int q = 0; q < 4; ++q)
            for (int w = 0; w < X_THREADS / 32; ++w) t[q] += s_part[q][w];
        s_esum = t[0];
        const double nrm = sqrt(t[1] * t[1] + t[2] * t[2] + t[3] * t[3]);
        intensity[b] = static_cast<float>(nrm);
        rgb_ratio[3 * b] = static_cast<float>(t[1] / nrm); rgb_ratio[3 * b + 1] = static_cast<float>(t[2] / nrm);
        rgb_ratio[3 * b + 2] = static_cast<float>(t[3] / nrm);
        ambient[3 * b] = static_cast<float>(s_acc[ln * 3]); ambient[3 * b + 1] = static_cast<float>(s_acc[ln * 3 + 1]);
        ambient[3 * b + 2] = static_cast<float>(s_acc[ln * 3 + 2]);
    }
    __syncthreads();
    for (int k = tid; k < ln; k += X_THREADS)
        dist[static_cast<long>(b) * ln + k] =
            static_cast<float>((0.3 * s_acc[3 * k] + 0.59 * s_acc[3 * k + 1] + 0.11 * s_acc[3 * k + 2]) / s_esum);
}

}  // namespace

extern "C" int eml_extract_params(const float *hdr, const int *idx, const double *ster, int B, int H, int W, int ln, float *dist,
                                  float *intensity, float *rgb_ratio, float *ambient, unsigned char *map, void *stream) {
    EML_CHECK_PTR(hdr); EML_CHECK_PTR(idx); EML_CHECK_PTR(ster); EML_CHECK_PTR(dist); EML_CHECK_PTR(intensity);
    EML_CHECK_PTR(rgb_ratio); EML_CHECK_PTR(ambient);
    if (B <= 0 || H <= 0 || W <= 0 || ln <= 0 || ln > X_MAX_LN) return EML_E_SHAPE;
    extract_params_kernel<<<static_cast<unsigned>(B), X_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(hdr, idx, ster, H, W, ln, dist,
                                                                                                        intensity, rgb_ratio, ambient, map);
    return eml_launch_status();
}
