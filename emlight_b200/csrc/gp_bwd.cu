// Adjoint kernels of the GenProjector training path (emlight_b200/gp_train.py): the gather's adjoint (col2im through the sampling
// table of SphereConv2D, sphere_cnn.py:111-124), activation / bias, SPADE modulation + parameter-free BatchNorm
// (normalization.py:101-115 under .train()), InstanceNorm (normalization.py:45).  All of them are single-pass HBM-bound kernels over
// NHWC fp32 tensors; the contractions of the backward run on the tcgen05 GEMM (gemm_tma.cu).
//
// Written in a restricted subset on purpose -- 1-D launches, no shared memory, no warp intrinsics, per-thread register accumulation
// followed by one atomicAdd per (worker, channel): with EML_EMULATE defined the SAME source compiles with g++ into a host library
// whose launches are plain loops, which is how tests/test_gp_bwd_emulated.py checks the index arithmetic on a box without a GPU.
// Reduction pattern ("column workers"): worker w handles channel c = w % C of rows strip, strip + nstrips, ... so that adjacent
// threads touch adjacent channels of the same row (coalesced) and the number of atomics is the number of workers, not of elements.
#ifdef EML_EMULATE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/emlight_b200.h"
#define __global__
#define __restrict__
struct EmuIdx { unsigned x; };
static EmuIdx blockIdx, threadIdx, blockDim, gridDim;
static inline float atomicAdd(float *p, float v) { float o = *p; *p += v; return o; }
static inline double atomicAdd(double *p, double v) { double o = *p; *p += v; return o; }
static inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
#define EML_CHECK_PTR(p) do { if ((p) == nullptr) return EML_E_NULL; } while (0)
static inline int eml_launch_status() { return EML_OK; }
#define EML_LAUNCH(kern, grid, block, stream, ...)                                                                     \
    do {                                                                                                               \
        (void)(stream);                                                                                                \
        gridDim.x = static_cast<unsigned>(grid); blockDim.x = static_cast<unsigned>(block);                            \
        for (unsigned b_ = 0; b_ < gridDim.x; ++b_)                                                                    \
            for (unsigned t_ = 0; t_ < blockDim.x; ++t_) { blockIdx.x = b_; threadIdx.x = t_; kern(__VA_ARGS__); }     \
    } while (0)
#define EML_API(name) name##_emu
#else
#include "common.cuh"
#define EML_LAUNCH(kern, grid, block, stream, ...) \
    kern<<<static_cast<unsigned>(grid), (block), 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__)
#define EML_API(name) name
#endif

namespace {

constexpr int THREADS = 256;
constexpr long MAX_WORKERS = 148L * 1024;          // one resident wave of 256-thread blocks at 4 blocks per SM

struct Strips { long nstrips; long workers; long blocks; };

// column workers for an (M rows) x (C channels) reduction
inline Strips make_strips(long M, int C) {
    Strips s;
    s.nstrips = MAX_WORKERS / C;
    if (s.nstrips < 1) s.nstrips = 1;
    if (s.nstrips > M) s.nstrips = M;
    s.workers = s.nstrips * C;
    s.blocks = (s.workers + THREADS - 1) / THREADS;
    return s;
}

inline long blocks_for(long n) { return (n + THREADS - 1) / THREADS; }

#define GLOBAL_TID (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x)

// derivative of act at pre-activation u (0 none, 1 ReLU, 2 LeakyReLU(0.2)); the same test works on the activation's OUTPUT,
// which has the sign of u
#define ACT_SLOPE(u, act) ((act) == 0 ? 1.0f : ((u) > 0.0f ? 1.0f : ((act) == 1 ? 0.0f : 0.2f)))

// ------------------------------------------------------------------------------------------------ col2im through the sampling table
// one thread per (m, tap, channel quad): dx[b, idx[p,tap,t], c..c+3] += w[p,tap,t] * dA[m, tap*Cp + c..c+3]
__global__ void col2im_lut_kernel(const float *__restrict__ dA, int Cp, const int *__restrict__ lut_idx,
                                  const float *__restrict__ lut_w, float *dx, int dx_pitch, long total, long out_pixels,
                                  long in_pixels) {
    const long tid = GLOBAL_TID;
    if (tid >= total) return;
    const int q4 = Cp >> 2;
    const int cq = static_cast<int>(tid % q4);
    const long rest = tid / q4;
    const int tap = static_cast<int>(rest % 9);
    const long m = rest / 9;
    const long p = m % out_pixels;
    const long b = m / out_pixels;
    const float *src = dA + m * 9L * Cp + static_cast<long>(tap) * Cp + 4 * cq;
    const float v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
    const long l = (p * 9 + tap) * 4;
    for (int t = 0; t < 4; ++t) {
        const int q = lut_idx[l + t];
        const float w = lut_w[l + t];
        if (q < 0 || w == 0.0f) continue;
        float *dst = dx + (b * in_pixels + q) * dx_pitch + 4 * cq;
        atomicAdd(dst + 0, w * v0);
        atomicAdd(dst + 1, w * v1);
        atomicAdd(dst + 2, w * v2);
        atomicAdd(dst + 3, w * v3);
    }
}

// ------------------------------------------------------------------------------------------------ dx *= act'(x + bias), sum per channel
__global__ void act_bwd_kernel(float *dx, int dx_pitch, const float *__restrict__ x, int x_pitch,
                               const float *__restrict__ bias, int act, long M, int C, long nstrips, double *sums) {
    const long w = GLOBAL_TID;
    if (w >= nstrips * C) return;
    const int c = static_cast<int>(w % C);
    const float bc = bias ? bias[c] : 0.0f;
    double acc = 0.0;
    for (long r = w / C; r < M; r += nstrips) {
        float g = dx[r * dx_pitch + c];
        if (act) {
            const float u = x[r * x_pitch + c] + bc;
            g *= ACT_SLOPE(u, act);
            dx[r * dx_pitch + c] = g;
        }
        acc += g;
    }
    if (sums) atomicAdd(sums + c, acc);
}

// ------------------------------------------------------------------------------------------------ out = act(raw + bias) backward
__global__ void bias_act_bwd_kernel(const float *__restrict__ g, int g_pitch, const float *__restrict__ out, int out_pitch,
                                    int act, float *dx, int dx_pitch, long M, int C, long nstrips, double *sums) {
    const long w = GLOBAL_TID;
    if (w >= nstrips * C) return;
    const int c = static_cast<int>(w % C);
    double acc = 0.0;
    for (long r = w / C; r < M; r += nstrips) {
        const float o = out[r * out_pitch + c];
        const float v = g[r * g_pitch + c] * ACT_SLOPE(o, act);
        dx[r * dx_pitch + c] = v;
        acc += v;
    }
    if (sums) atomicAdd(sums + c, acc);
}

// ------------------------------------------------------------------------------------------------ SPADE modulation backward, pass 1
// out = lrelu?(xhat * (1 + gamma + bias_gamma) + beta + bias_beta), xhat = (x - mean) * inv_std:
//   g' = g * lrelu'(out); d_gamma = g' * xhat; d_beta = g'; d_xhat = g' * (1 + gamma + bias_gamma)
//   sums[0] += sum g' xhat (-> d bias_gamma), sums[1] += sum g' (-> d bias_beta), sums[2] += sum d_xhat, sums[3] += sum d_xhat xhat
__global__ void spade_bwd_kernel(const float *__restrict__ g, int g_pitch, const float *__restrict__ out, int out_pitch,
                                 const float *__restrict__ x, int x_pitch, const float *__restrict__ mean,
                                 const float *__restrict__ inv_std, const float *__restrict__ gb, int gb_pitch,
                                 const float *__restrict__ bias_gamma, float *d_gb, float *d_xhat, int dxh_pitch, long M,
                                 int C, int leaky, long nstrips, double *sums) {
    const long w = GLOBAL_TID;
    if (w >= nstrips * C) return;
    const int c = static_cast<int>(w % C);
    const float mu = mean[c], is = inv_std[c], bg = bias_gamma ? bias_gamma[c] : 0.0f;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (long r = w / C; r < M; r += nstrips) {
        float gv = g[r * g_pitch + c];
        if (leaky) gv *= ACT_SLOPE(out[r * out_pitch + c], 2);
        const float xh = (x[r * x_pitch + c] - mu) * is;
        const float dg = gv * xh;
        const float dxh = gv * (1.0f + gb[r * gb_pitch + c] + bg);
        d_gb[r * gb_pitch + c] = dg;
        d_gb[r * gb_pitch + C + c] = gv;
        d_xhat[r * dxh_pitch + c] = dxh;
        s0 += dg; s1 += gv; s2 += dxh; s3 += static_cast<double>(dxh) * xh;
    }
    atomicAdd(sums + c, s0);
    atomicAdd(sums + C + c, s1);
    atomicAdd(sums + 2 * C + c, s2);
    atomicAdd(sums + 3 * C + c, s3);
}

// ------------------------------------------------------------------------------------------------ parameter-free BatchNorm backward
// dx = inv_std * (d_xhat - s0/n - xhat * s1/n)   (batch statistics; sums == NULL: running statistics, dx = inv_std * d_xhat)
__global__ void bn_free_bwd_kernel(const float *__restrict__ d_xhat, int dxh_pitch, const float *__restrict__ x, int x_pitch,
                                   const float *__restrict__ mean, const float *__restrict__ inv_std,
                                   const double *__restrict__ sums, double count, float *dx, int dx_pitch, long M, int C) {
    const long tid = GLOBAL_TID;
    if (tid >= M * C) return;
    const int c = static_cast<int>(tid % C);
    const long r = tid / C;
    const float is = inv_std[c];
    float v = d_xhat[r * dxh_pitch + c];
    if (sums) {
        const float xh = (x[r * x_pitch + c] - mean[c]) * is;
        v -= static_cast<float>(sums[c] / count) + xh * static_cast<float>(sums[C + c] / count);
    }
    dx[r * dx_pitch + c] = is * v;
}

// ------------------------------------------------------------------------------------------------ InstanceNorm (+LeakyReLU) backward
// reduce: per (image, channel) sums[b][0..3][c] += sum g', sum g' xhat, sum raw, sum raw^2   (g' = g * lrelu'(out), xhat from out)
__global__ void inorm_bwd_reduce_kernel(const float *__restrict__ g, int g_pitch, const float *__restrict__ out, int out_pitch,
                                        const float *__restrict__ raw, int raw_pitch, long B, long HW, int C, int leaky,
                                        long nstrips, double *sums) {
    const long w = GLOBAL_TID;                        // workers = B * nstrips * C
    const long per_image = nstrips * C;
    if (w >= B * per_image) return;
    const long b = w / per_image;
    const long wi = w % per_image;
    const int c = static_cast<int>(wi % C);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (long r = wi / C; r < HW; r += nstrips) {
        const long row = b * HW + r;
        const float o = out[row * out_pitch + c];
        float gv = g[row * g_pitch + c];
        float xh = o;
        if (leaky) { gv *= ACT_SLOPE(o, 2); if (o < 0.0f) xh = o * 5.0f; }
        const float rv = raw[row * raw_pitch + c];
        s0 += gv; s1 += static_cast<double>(gv) * xh; s2 += rv; s3 += static_cast<double>(rv) * rv;
    }
    double *dst = sums + b * 4 * C;
    atomicAdd(dst + c, s0);
    atomicAdd(dst + C + c, s1);
    atomicAdd(dst + 2 * C + c, s2);
    atomicAdd(dst + 3 * C + c, s3);
}

__global__ void inorm_bwd_apply_kernel(const float *__restrict__ g, int g_pitch, const float *__restrict__ out, int out_pitch,
                                       const double *__restrict__ sums, long B, long HW, int C, float eps, int leaky, float *dx,
                                       int dx_pitch) {
    const long tid = GLOBAL_TID;
    if (tid >= B * HW * C) return;
    const int c = static_cast<int>(tid % C);
    const long row = tid / C;
    const long b = row / HW;
    const double *s = sums + b * 4 * C;
    const double n = static_cast<double>(HW);
    const double m = s[2 * C + c] / n;
    double var = s[3 * C + c] / n - m * m;
    if (var < 0.0) var = 0.0;
    const float is = rsqrtf(static_cast<float>(var) + eps);
    const float o = out[row * out_pitch + c];
    float gv = g[row * g_pitch + c];
    float xh = o;
    if (leaky) { gv *= ACT_SLOPE(o, 2); if (o < 0.0f) xh = o * 5.0f; }
    dx[row * dx_pitch + c] = is * (gv - static_cast<float>(s[c] / n) - xh * static_cast<float>(s[C + c] / n));
}

}  // namespace

// =================================================================================================== C ABI
extern "C" int EML_API(eml_col2im_lut)(const float *dA, int Cp, const int *lut_idx, const float *lut_w, float *dx, int dx_pitch,
                                       int B, long out_pixels, long in_pixels, void *stream) {
    EML_CHECK_PTR(dA); EML_CHECK_PTR(lut_idx); EML_CHECK_PTR(lut_w); EML_CHECK_PTR(dx);
    if (B <= 0 || out_pixels <= 0 || in_pixels <= 0 || Cp <= 0 || (Cp & 3) || dx_pitch < Cp) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * out_pixels * 9 * (Cp >> 2);
    if (blocks_for(total) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(col2im_lut_kernel, blocks_for(total), THREADS, stream, dA, Cp, lut_idx, lut_w, dx, dx_pitch, total, out_pixels, in_pixels);
    return eml_launch_status();
}

extern "C" int EML_API(eml_act_bwd)(float *dx, int dx_pitch, const float *x, int x_pitch, const float *bias, int act, long M, int C,
                                    double *bias_sums, void *stream) {
    EML_CHECK_PTR(dx);
    if (act < 0 || act > 2) return EML_E_ARG;
    if (act) EML_CHECK_PTR(x);
    if (M <= 0 || C <= 0 || dx_pitch < C || (act && x_pitch < C)) return EML_E_SHAPE;
    if (!act && !bias_sums) return EML_OK;
    const Strips s = make_strips(M, C);
    EML_LAUNCH(act_bwd_kernel, s.blocks, THREADS, stream, dx, dx_pitch, x, x_pitch, bias, act, M, C, s.nstrips, bias_sums);
    return eml_launch_status();
}

extern "C" int EML_API(eml_bias_act_bwd)(const float *g, int g_pitch, const float *out, int out_pitch, int act, float *dx,
                                         int dx_pitch, long M, int C, double *bias_sums, void *stream) {
    EML_CHECK_PTR(g); EML_CHECK_PTR(out); EML_CHECK_PTR(dx);
    if (act < 0 || act > 2) return EML_E_ARG;
    if (M <= 0 || C <= 0 || g_pitch < C || out_pitch < C || dx_pitch < C) return EML_E_SHAPE;
    const Strips s = make_strips(M, C);
    EML_LAUNCH(bias_act_bwd_kernel, s.blocks, THREADS, stream, g, g_pitch, out, out_pitch, act, dx, dx_pitch, M, C, s.nstrips, bias_sums);
    return eml_launch_status();
}

extern "C" int EML_API(eml_spade_bwd)(const float *g, int g_pitch, const float *out, int out_pitch, const float *x, int x_pitch,
                                      const float *mean, const float *inv_std, const float *gb, int gb_pitch,
                                      const float *bias_gamma, float *d_gb, float *d_xhat, int dxh_pitch, long M, int C,
                                      int leaky_relu, double *sums, void *stream) {
    EML_CHECK_PTR(g); EML_CHECK_PTR(out); EML_CHECK_PTR(x); EML_CHECK_PTR(mean); EML_CHECK_PTR(inv_std); EML_CHECK_PTR(gb);
    EML_CHECK_PTR(d_gb); EML_CHECK_PTR(d_xhat); EML_CHECK_PTR(sums);
    if (M <= 0 || C <= 0 || g_pitch < C || out_pitch < C || x_pitch < C || gb_pitch < 2 * C || dxh_pitch < C) return EML_E_SHAPE;
    const Strips s = make_strips(M, C);
    EML_LAUNCH(spade_bwd_kernel, s.blocks, THREADS, stream, g, g_pitch, out, out_pitch, x, x_pitch, mean, inv_std, gb, gb_pitch,
               bias_gamma, d_gb, d_xhat, dxh_pitch, M, C, leaky_relu, s.nstrips, sums);
    return eml_launch_status();
}

extern "C" int EML_API(eml_bn_free_bwd)(const float *d_xhat, int dxh_pitch, const float *x, int x_pitch, const float *mean,
                                        const float *inv_std, const double *sums, double count, float *dx, int dx_pitch, long M,
                                        int C, void *stream) {
    EML_CHECK_PTR(d_xhat); EML_CHECK_PTR(inv_std); EML_CHECK_PTR(dx);
    if (sums) { EML_CHECK_PTR(x); EML_CHECK_PTR(mean); if (!(count > 0.0)) return EML_E_ARG; }
    if (M <= 0 || C <= 0 || dxh_pitch < C || dx_pitch < C || (sums && x_pitch < C)) return EML_E_SHAPE;
    if (blocks_for(M * C) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(bn_free_bwd_kernel, blocks_for(M * C), THREADS, stream, d_xhat, dxh_pitch, x, x_pitch, mean, inv_std, sums, count, dx,
               dx_pitch, M, C);
    return eml_launch_status();
}

extern "C" int EML_API(eml_instance_norm_bwd)(const float *g, int g_pitch, const float *out, int out_pitch, const float *raw,
                                              int raw_pitch, int B, long HW, int C, float eps, int leaky_relu, double *sums,
                                              float *dx, int dx_pitch, void *stream) {
    EML_CHECK_PTR(g); EML_CHECK_PTR(out); EML_CHECK_PTR(raw); EML_CHECK_PTR(sums); EML_CHECK_PTR(dx);
    if (B <= 0 || HW <= 0 || C <= 0 || g_pitch < C || out_pitch < C || raw_pitch < C || dx_pitch < C) return EML_E_SHAPE;
    long nstrips = MAX_WORKERS / (static_cast<long>(B) * C);
    if (nstrips < 1) nstrips = 1;
    if (nstrips > HW) nstrips = HW;
    const long workers = static_cast<long>(B) * nstrips * C;
    EML_LAUNCH(inorm_bwd_reduce_kernel, blocks_for(workers), THREADS, stream, g, g_pitch, out, out_pitch, raw, raw_pitch,
               static_cast<long>(B), HW, C, leaky_relu, nstrips, sums);
    int rc = eml_launch_status();
    if (rc != EML_OK) return rc;
    const long total = static_cast<long>(B) * HW * C;
    if (blocks_for(total) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(inorm_bwd_apply_kernel, blocks_for(total), THREADS, stream, g, g_pitch, out, out_pitch, sums, static_cast<long>(B), HW,
               C, eps, leaky_relu, dx, dx_pitch);
    return eml_launch_status();
}
