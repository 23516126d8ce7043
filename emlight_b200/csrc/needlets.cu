// Spherical-needlet kernels (Needlets/sphere_needlets.py, mat_gen2.py, gt_gen_j3.py), sm_100a.
//
//  * needlet_basis_kernel: SN_matrix = [Y_00 | psi_jk] (sphere_needlets.py:196-238) by the addition theorem
//        psi_jk(x) = sqrt(lambda_j) sum_l b(l/B^j) (2l+1)/(4 pi) P_l(x . xi_jk)
//    (the reference sums conj(Y_lm(x)) Y_lm(xi_jk) over m ring by ring, :34-104; the two agree to 1e-15, oracle/needlets_oracle.py)
//    in float64: one thread per (grid point, cubature point), three-term Legendre recurrence, coefficients c[j][l] from the host.
//    The reference spends hours of Python per 128x256 grid here (mat_gen2.py:27 is commented out for that reason).
//  * split_bf16_kernel: fp32 (rows, cols) -> bf16 hi / lo (rows, Kp) operands of the TMA-fed tcgen05 GEMM (gemm_tma.cu), which does
//    the projection  coef = (SN * omega)^T pano  (gt_gen_j3.py:39-43) and the reconstruction  rec = SN coef  (mat_gen2.py:55).
//  * needlet_sparsify_kernel: per image and level block, zero the coefficients below frac * max|.| (mat_gen2.py:43-51).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int NB_MAX_L = 64;

__global__ void __launch_bounds__(256) needlet_basis_kernel(const double *__restrict__ xyz, long P, const double *__restrict__ centres,
                                                            const int *__restrict__ level, int K, const double *__restrict__ coef,
                                                            int nlev, int lmax, double *__restrict__ out, long out_pitch) {
    __shared__ double s_c[8 * (NB_MAX_L + 1)];
    for (int i = threadIdx.x; i < nlev * (lmax + 1); i += blockDim.x) s_c[i] = coef[i];
    __syncthreads();
    const long total = P * (K + 1);
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const long p = idx / (K + 1);
        const int k = static_cast<int>(idx - p * (K + 1)) - 1;
        double v;
        if (k < 0) {
            v = 0.28209479177387814;                           // Y_00 = 1 / sqrt(4 pi)
        } else {
            double t = xyz[3 * p] * centres[3 * k] + xyz[3 * p + 1] * centres[3 * k + 1] + xyz[3 * p + 2] * centres[3 * k + 2];
            t = fmin(1.0, fmax(-1.0, t));
            const double *c = s_c + level[k] * (lmax + 1);
            double p0 = 1.0, p1 = t;
            v = c[1] * p1;
            for (int l = 2; l <= lmax; ++l) {
                const double p2 = ((2 * l - 1) * t * p1 - (l - 1) * p0) / l;
                p0 = p1; p1 = p2;
                v += c[l] * p2;                                 // c[l] == 0 outside the level's band [l_st, l_en]
            }
        }
        out[p * out_pitch + (k + 1)] = v;
    }
}

// x (rows, cols) fp32 with row stride `ld` -> hi/lo (rows, Kp) bf16, columns >= cols zero.  lo may be NULL.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float *__restrict__ x, long rows, int cols, long ld,
                                                         __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int Kp) {
    const long total = rows * Kp;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const long r = idx / Kp;
        const int c = static_cast<int>(idx - r * Kp);
        const float v = c < cols ? x[r * ld + c] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[idx] = h;
        if (lo != nullptr) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// coef (B, n, ch) contiguous; block (b, blk) handles rows [lo, hi) x ch values: keep |v| > frac * max|v|, zero the rest.
__global__ void __launch_bounds__(256) needlet_sparsify_kernel(float *__restrict__ coef, int n, int ch, const int *__restrict__ ranges,
                                                               float frac) {
    __shared__ float s_max[8];
    const int b = blockIdx.x, blk = blockIdx.y;
    const int lo = ranges[2 * blk], hi = ranges[2 * blk + 1];
    float *base = coef + (static_cast<long>(b) * n + lo) * ch;
    const int cnt = (hi - lo) * ch;
    float m = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) m = fmaxf(m, fabsf(base[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int w = 0; w < 8; ++w) m = fmaxf(m, s_max[w]);
    const float thr = m * frac;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x)
        if (!(fabsf(base[i]) > thr)) base[i] = 0.f;
}

}  // namespace

extern "C" int eml_needlet_basis(const double *xyz, long P, const double *centres, const int *level, int K, const double *coef,
                                 int nlev, int lmax, double *out, long out_pitch, void *stream) {
    EML_CHECK_PTR(xyz); EML_CHECK_PTR(centres); EML_CHECK_PTR(level); EML_CHECK_PTR(coef); EML_CHECK_PTR(out);
    if (P <= 0 || K <= 0 || nlev <= 0 || nlev > 8 || lmax < 1 || lmax > NB_MAX_L || out_pitch < K + 1) return EML_E_SHAPE;
    const long total = P * (K + 1);
    const long blocks = (total + 255) / 256;
    needlet_basis_kernel<<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        xyz, P, centres, level, K, coef, nlev, lmax, out, out_pitch);
    return eml_launch_status();
}

extern "C" int eml_split_bf16(const float *x, long rows, int cols, long ld, void *hi, void *lo, int Kp, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(hi);
    if (rows <= 0 || cols <= 0 || Kp < cols || ld < cols) return EML_E_SHAPE;
    const long total = rows * Kp;
    const long blocks = (total + 255) / 256;
    split_bf16_kernel<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, rows, cols, ld, static_cast<__nv_bfloat16 *>(hi), static_cast<__nv_bfloat16 *>(lo), Kp);
    return eml_launch_status();
}

extern "C" int eml_needlet_sparsify(float *coef, int B, int n, int ch, const int *ranges, int nranges, float frac, void *stream) {
    EML_CHECK_PTR(coef); EML_CHECK_PTR(ranges);
    if (B <= 0 || n <= 0 || ch <= 0 || nranges <= 0 || nranges > 65535) return EML_E_SHAPE;
    if (!(frac >= 0.f)) return EML_E_ARG;
    needlet_sparsify_kernel<<<dim3(static_cast<unsigned>(B), static_cast<unsigned>(nranges)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        coef, n, ch, ranges, frac);
    return eml_launch_status();
}
