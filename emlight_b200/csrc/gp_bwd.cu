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
#include <string.h>
static inline unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float u2f(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#else
#include "common.cuh"
#define EML_LAUNCH(kern, grid, block, stream, ...) \
    kern<<<static_cast<unsigned>(grid), (block), 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__)
#define EML_API(name) name
__device__ __forceinline__ unsigned f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(unsigned u) { return __uint_as_float(u); }
// spade_ops.cu: the shared-memory-transposed form of eml_im2col_lut_bf16_t (CUDA build only)
bool eml_im2col_t_tiled_ok(int x_pitch, int C, int Cp, const void *bias, int act, const void *hi, const void *lo, long Mp);
int eml_im2col_t_tiled(const float *x, int x_pitch, int Cp, const int *lut_idx, const float *lut_w, void *hi, void *lo, long Mp, long M,
                       long out_pixels, long in_pixels, cudaStream_t st);
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

// ------------------------------------------------------------------------------------------------ nearest x2 upsample backward
// dx[b,h,w,c] = sum_{i,j<2} g[b,2h+i,2w+j,c]   (generator.py:42,70 self.up)
__global__ void upsample2_bwd_kernel(const float *__restrict__ g, int g_pitch, float *dx, int dx_pitch, long total, int H, int W, int C) {
    const long tid = GLOBAL_TID;
    if (tid >= total) return;
    const int c = static_cast<int>(tid % C);
    long t = tid / C;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H);
    const long b = t / H;
    const long W2 = 2L * W;
    const float *r0 = g + ((b * 2 * H + 2 * h) * W2 + 2 * w) * g_pitch + c;
    const float *r1 = r0 + W2 * g_pitch;
    dx[tid / C * dx_pitch + c] = (r0[0] + r0[g_pitch]) + (r1[0] + r1[g_pitch]);
}

// ------------------------------------------------------------------------------------------------ out = (tanh(raw + bias) + 1) * scale backward
// NCHW gradient / output -> NHWC gradient of raw; bias_sums[c] += sum (column workers over the HW*B rows)
__global__ void tanh_nchw_bwd_kernel(const float *__restrict__ g, const float *__restrict__ out, float scale, float *d_raw, int pitch,
                                     long B, long HW, int C, long nstrips, double *sums) {
    const long w = GLOBAL_TID;
    if (w >= nstrips * C) return;
    const int c = static_cast<int>(w % C);
    double acc = 0.0;
    for (long r = w / C; r < B * HW; r += nstrips) {
        const long b = r / HW, p = r % HW;
        const long src = (b * C + c) * HW + p;
        const float t = out[src] / scale - 1.0f;
        const float v = g[src] * scale * (1.0f - t * t);
        d_raw[r * pitch + c] = v;
        acc += v;
    }
    if (sums) atomicAdd(sums + c, acc);
}

// ------------------------------------------------------------------------------------------------ pooling backward (gather form)
// mode 0: avg_pool2d(3, stride 2, pad 1, count_include_pad=False): dx[y,x] = sum over windows (yo,xo) containing (y,x) of g / n(yo,xo)
// mode 1: max_pool2d(2,2): dx[y,x] = g[y/2,x/2] if (y,x) is the FIRST maximum of its window in row-major order (ATen's choice)
__global__ void pool2d_bwd_kernel(const float *__restrict__ g, int g_pitch, const float *__restrict__ x, int x_pitch, float *dx,
                                  int dx_pitch, long total, int Hi, int Wi, int Ho, int Wo, int C, int mode) {
    const long tid = GLOBAL_TID;
    if (tid >= total) return;
    const int c = static_cast<int>(tid % C);
    long t = tid / C;
    const int xx = static_cast<int>(t % Wi); t /= Wi;
    const int y = static_cast<int>(t % Hi);
    const long b = t / Hi;
    float v = 0.0f;
    if (mode == 0) {
        for (int yo = y / 2; yo <= (y + 1) / 2; ++yo) {
            if (yo >= Ho) continue;
            const int ny = (2 * yo - 1 >= 0 ? 1 : 0) + 1 + (2 * yo + 1 < Hi ? 1 : 0);
            for (int xo = xx / 2; xo <= (xx + 1) / 2; ++xo) {
                if (xo >= Wo) continue;
                const int nx = (2 * xo - 1 >= 0 ? 1 : 0) + 1 + (2 * xo + 1 < Wi ? 1 : 0);
                v += g[((b * Ho + yo) * static_cast<long>(Wo) + xo) * g_pitch + c] / static_cast<float>(ny * nx);
            }
        }
    } else {
        const int yo = y / 2, xo = xx / 2;
        const float *win = x + ((b * Hi + 2 * yo) * static_cast<long>(Wi) + 2 * xo) * x_pitch + c;
        const float v00 = win[0], v01 = win[x_pitch], v10 = win[static_cast<long>(Wi) * x_pitch], v11 = win[(static_cast<long>(Wi) + 1) * x_pitch];
        int arg = 0; float m = v00;
        if (v01 > m) { m = v01; arg = 1; }
        if (v10 > m) { m = v10; arg = 2; }
        if (v11 > m) { m = v11; arg = 3; }
        if (arg == (y & 1) * 2 + (xx & 1)) v = g[((b * Ho + yo) * static_cast<long>(Wo) + xo) * g_pitch + c];
    }
    dx[tid / C * dx_pitch + c] = v;
}

// ------------------------------------------------------------------------------------------------ loss seeds
// d a = coef * d(eml_loss_reduce(mode))/da  for the element-wise modes 0..4 (mode 5, the cosine distance, couples the channels)
__global__ void loss_seed_kernel(const float *__restrict__ a, int a_pitch, const float *__restrict__ b, int b_pitch,
                                 const float *__restrict__ mask, long M, int C, int mode, float coef,
                                 const float *__restrict__ coef_dev, float *da, int da_pitch) {
    const long tid = GLOBAL_TID;
    if (tid >= M * C) return;
    if (coef_dev) coef *= coef_dev[0];
    const int c = static_cast<int>(tid % C);
    const long r = tid / C;
    const float av = a[r * a_pitch + c];
    float d;
    if (mode == 0) d = 1.0f;
    else if (mode == 1) d = av < 1.0f ? 1.0f : 0.0f;
    else if (mode == 2) d = av > -1.0f ? -1.0f : 0.0f;
    else {
        const float df = av - b[r * b_pitch + c];
        d = df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f);
        if (mode == 4) { const float m = mask[r]; d *= m + (1.0f - m) * 50.0f; }
    }
    da[r * da_pitch + c] = coef * d;
}

// cosine distance over the channels as ATen computes it, 1 - ahat . bhat with ahat = a / max(|a|, eps), bhat = b / max(|b|, eps):
//   |a| > eps: d/da = -(bhat - (ahat . bhat) ahat) / |a|;   |a| <= eps: d/da = -bhat / eps       (normalise first: no |a|^3 underflow)
__global__ void cos_seed_kernel(const float *__restrict__ a, int a_pitch, const float *__restrict__ b, int b_pitch, long M, int C,
                                float eps, float coef, const float *__restrict__ coef_dev, float *da, int da_pitch) {
    const long r = GLOBAL_TID;
    if (r >= M) return;
    if (coef_dev) coef *= coef_dev[0];
    float na = 0.0f, nb = 0.0f;
    for (int c = 0; c < C; ++c) {
        const float av = a[r * a_pitch + c], bv = b[r * b_pitch + c];
        na += av * av; nb += bv * bv;
    }
    const float an = sqrtf(na);
    const float ia = 1.0f / fmaxf(an, eps), ib = 1.0f / fmaxf(sqrtf(nb), eps);
    float cosv = 0.0f;
    for (int c = 0; c < C; ++c) cosv += (a[r * a_pitch + c] * ia) * (b[r * b_pitch + c] * ib);
    for (int c = 0; c < C; ++c) {
        const float ah = a[r * a_pitch + c] * ia, bh = b[r * b_pitch + c] * ib;
        const float d = an > eps ? (bh - cosv * ah) * ia : bh * ia;
        da[r * da_pitch + c] = -coef * d;
    }
}

// ------------------------------------------------------------------------------------------------ transposed bf16 im2col (weight gradient)
// The operand of the weight-gradient GEMM dWk^T = A^T dY has K = pixels, so it wants A TRANSPOSED: rows = (tap, channel), columns =
// output pixels m.  This is eml_im2col_lut_bf16 writing the transpose directly, split into bf16 hi / lo (round to nearest even, the
// same split as eml_split_bf16): thread = (m fastest, tap, channel quad) -> the four row writes of a warp are 64 contiguous bytes.
//   hi/lo[(tap*Cp + c) * Mp + m] = bf16 split of sum_t w[p,tap,t] * act(x[b, idx[p,tap,t], c] + bias[c])     (columns >= M stay 0)
#define BF16_RNE(f) static_cast<unsigned short>((f2u(f) + 0x7FFFu + ((f2u(f) >> 16) & 1u)) >> 16)
__global__ void im2col_lut_bf16_t_kernel(const float *__restrict__ x, int x_pitch, int C, int Cp, const int *__restrict__ lut_idx,
                                         const float *__restrict__ lut_w, const float *__restrict__ bias, int act,
                                         unsigned short *hi, unsigned short *lo, long Mp, long M, long out_pixels, long in_pixels,
                                         long total) {
    const long tid = GLOBAL_TID;
    if (tid >= total) return;
    const long m = tid % M;
    const long rest = tid / M;
    const int tap = static_cast<int>(rest % 9);
    const int cq = static_cast<int>(rest / 9);
    const long p = m % out_pixels;
    const long b = m / out_pixels;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const long l = (p * 9 + tap) * 4;
    for (int t = 0; t < 4; ++t) {
        const int q = lut_idx[l + t];
        const float w = lut_w[l + t];
        if (q < 0 || w == 0.0f) continue;
        const float *src = x + (b * in_pixels + q) * x_pitch + 4 * cq;
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * cq + j;
            if (c >= C) break;
            float u = src[j] + (bias ? bias[c] : 0.0f);
            if (act == 1) u = u > 0.0f ? u : 0.0f;
            else if (act == 2) u = u > 0.0f ? u : 0.2f * u;
            acc[j] += w * u;
        }
    }
    for (int j = 0; j < 4; ++j) {
        const long o = (static_cast<long>(tap) * Cp + 4 * cq + j) * Mp + m;
        const unsigned short h = BF16_RNE(acc[j]);
        hi[o] = h;
        if (lo) { const float r = acc[j] - u2f(static_cast<unsigned>(h) << 16); lo[o] = BF16_RNE(r); }
    }
}

// ------------------------------------------------------------------------------------------------ col2im, gather form (no atomics)
// The sampling table inverted on the host into CSR over INPUT pixels: for input pixel q the entries e in [offs[q], offs[q+1]) name the
// im2col cell src[e] = p*9 + tap that read q with weight w[e].  One thread per (b, q, channel quad) sums its entries in a fixed order:
// deterministic, one write per element, reads of dA rows coalesced over the channel quads.
//   dx[b, q, c] = sum_e w[e] * dA[b*out_pixels + src[e]/9, (src[e]%9)*Cp + c]
__global__ void col2im_csr_kernel(const float *__restrict__ dA, int Cp, const int *__restrict__ offs, const int *__restrict__ src,
                                  const float *__restrict__ w, float *dx, int dx_pitch, long total, long out_pixels, long in_pixels) {
    const long tid = GLOBAL_TID;
    if (tid >= total) return;
    const int q4 = Cp >> 2;
    const int cq = static_cast<int>(tid % q4);
    const long bq = tid / q4;
    const long q = bq % in_pixels;
    const long b = bq / in_pixels;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    const float *base = dA + b * out_pixels * 9L * Cp + 4 * cq;
    for (int e = offs[q]; e < offs[q + 1]; ++e) {
        const float *row = base + static_cast<long>(src[e]) * Cp;          // (p*9 + tap) * Cp
        const float we = w[e];
        a0 += we * row[0]; a1 += we * row[1]; a2 += we * row[2]; a3 += we * row[3];
    }
    float *dst = dx + bq * dx_pitch + 4 * cq;
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
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

extern "C" int EML_API(eml_upsample2_bwd)(const float *g, int g_pitch, float *dx, int dx_pitch, int B, int H, int W, int C, void *stream) {
    EML_CHECK_PTR(g); EML_CHECK_PTR(dx);
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || g_pitch < C || dx_pitch < C) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * H * W * C;
    if (blocks_for(total) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(upsample2_bwd_kernel, blocks_for(total), THREADS, stream, g, g_pitch, dx, dx_pitch, total, H, W, C);
    return eml_launch_status();
}

extern "C" int EML_API(eml_tanh_nchw_bwd)(const float *g_nchw, const float *out_nchw, float scale, float *d_raw, int pitch, int B,
                                          long HW, int C, double *bias_sums, void *stream) {
    EML_CHECK_PTR(g_nchw); EML_CHECK_PTR(out_nchw); EML_CHECK_PTR(d_raw);
    if (B <= 0 || HW <= 0 || C <= 0 || pitch < C) return EML_E_SHAPE;
    if (!(scale > 0.0f)) return EML_E_ARG;
    const Strips s = make_strips(static_cast<long>(B) * HW, C);
    EML_LAUNCH(tanh_nchw_bwd_kernel, s.blocks, THREADS, stream, g_nchw, out_nchw, scale, d_raw, pitch, static_cast<long>(B), HW, C,
               s.nstrips, bias_sums);
    return eml_launch_status();
}

extern "C" int EML_API(eml_pool2d_bwd)(const float *g, int g_pitch, const float *x, int x_pitch, float *dx, int dx_pitch, int Hi, int Wi,
                                       int C, int B, int mode, void *stream) {
    EML_CHECK_PTR(g); EML_CHECK_PTR(dx);
    if (mode != 0 && mode != 1) return EML_E_ARG;
    if (mode == 1) EML_CHECK_PTR(x);
    if (B <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || g_pitch < C || dx_pitch < C || (mode == 1 && x_pitch < C)) return EML_E_SHAPE;
    if (mode == 1 && ((Hi | Wi) & 1)) return EML_E_SHAPE;
    const int Ho = mode == 0 ? (Hi + 1) / 2 : Hi / 2, Wo = mode == 0 ? (Wi + 1) / 2 : Wi / 2;
    const long total = static_cast<long>(B) * Hi * Wi * C;
    if (blocks_for(total) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(pool2d_bwd_kernel, blocks_for(total), THREADS, stream, g, g_pitch, x, x_pitch, dx, dx_pitch, total, Hi, Wi, Ho, Wo, C, mode);
    return eml_launch_status();
}

extern "C" int EML_API(eml_loss_seed)(const float *a, int a_pitch, const float *b, int b_pitch, const float *mask, long M, int C, int mode,
                                      float coef, const float *coef_dev, float *da, int da_pitch, void *stream) {
    EML_CHECK_PTR(a); EML_CHECK_PTR(da);
    if (mode < 0 || mode > 5) return EML_E_ARG;
    if (mode >= 3) EML_CHECK_PTR(b);
    if (mode == 4) EML_CHECK_PTR(mask);
    if (M <= 0 || C <= 0 || a_pitch < C || da_pitch < C || (mode >= 3 && b_pitch < C)) return EML_E_SHAPE;
    if (mode == 5) {
        EML_LAUNCH(cos_seed_kernel, blocks_for(M), THREADS, stream, a, a_pitch, b, b_pitch, M, C, 1e-20f, coef, coef_dev, da, da_pitch);
    } else {
        if (blocks_for(M * C) > 0x7fffffffL) return EML_E_SHAPE;
        EML_LAUNCH(loss_seed_kernel, blocks_for(M * C), THREADS, stream, a, a_pitch, b, b_pitch, mask, M, C, mode, coef, coef_dev, da, da_pitch);
    }
    return eml_launch_status();
}

extern "C" int EML_API(eml_im2col_lut_bf16_t)(const float *x, int x_pitch, int C, int Cp, const int *lut_idx, const float *lut_w,
                                              const float *bias, int act, void *At_hi, void *At_lo, long Mp, int B, long out_pixels,
                                              long in_pixels, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(lut_idx); EML_CHECK_PTR(lut_w); EML_CHECK_PTR(At_hi);
    if (act < 0 || act > 2) return EML_E_ARG;
    const long M = static_cast<long>(B) * out_pixels;
    if (B <= 0 || C <= 0 || Cp < C || (Cp & 3) || x_pitch < C || out_pixels <= 0 || in_pixels <= 0 || Mp < M) return EML_E_SHAPE;
#ifndef EML_EMULATE
    if (eml_im2col_t_tiled_ok(x_pitch, C, Cp, bias, act, At_hi, At_lo, Mp))
        return eml_im2col_t_tiled(x, x_pitch, Cp, lut_idx, lut_w, At_hi, At_lo, Mp, M, out_pixels, in_pixels, static_cast<cudaStream_t>(stream));
#endif
    const long total = M * 9 * (Cp >> 2);
    if (blocks_for(total) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(im2col_lut_bf16_t_kernel, blocks_for(total), THREADS, stream, x, x_pitch, C, Cp, lut_idx, lut_w, bias, act,
               static_cast<unsigned short *>(At_hi), static_cast<unsigned short *>(At_lo), Mp, M, out_pixels, in_pixels, total);
    return eml_launch_status();
}

extern "C" int EML_API(eml_col2im_csr)(const float *dA, int Cp, const int *offs, const int *src, const float *w, float *dx, int dx_pitch,
                                       int B, long out_pixels, long in_pixels, void *stream) {
    EML_CHECK_PTR(dA); EML_CHECK_PTR(offs); EML_CHECK_PTR(src); EML_CHECK_PTR(w); EML_CHECK_PTR(dx);
    if (B <= 0 || out_pixels <= 0 || in_pixels <= 0 || Cp <= 0 || (Cp & 3) || dx_pitch < Cp) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * in_pixels * (Cp >> 2);
    if (blocks_for(total) > 0x7fffffffL) return EML_E_SHAPE;
    EML_LAUNCH(col2im_csr_kernel, blocks_for(total), THREADS, stream, dA, Cp, offs, src, w, dx, dx_pitch, total, out_pixels, in_pixels);
    return eml_launch_status();
}
