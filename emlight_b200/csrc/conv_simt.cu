// fp32 FFMA companions of the tensor-core convolution (sm_100a):
//   * eml_conv_forward_simt  -- the EML_PREC_FP32 mode of eml_conv_forward: same gather / affine / ReLU /
//     zero-padding / pooling / statistics semantics as conv_gemm.cu, evaluated with plain fp32 FMAs.
//     It is the bit-for-bit-stable "exact" mode (anchor argmax parity) and the on-device cross-check of the
//     tcgen05 path at sizes the CPU oracle cannot reach.  Not tuned: one thread per output pixel, 16 output
//     channels per block.y, weights staged in shared memory.
//   * stem (conv0 3x3 on the NCHW image, DenseNet.py:89-92), BatchNorm folding, head pooling, linear layers.
#include "common.cuh"

namespace {

constexpr int SIMT_NB = 16;       // output channels per block
constexpr int SIMT_THREADS = 128;

struct SimtArgs {
    const float *in, *scale, *shift, *w;
    float *out;
    double *stats;
    long stats_stride;
    long M;
    int H, W, C_in, in_pitch, C_out, out_pitch, out_choff, mode, relu, taps;
};

__global__ void __launch_bounds__(SIMT_THREADS) conv_simt_kernel(const SimtArgs a) {
    extern __shared__ float s_w[];                       // [SIMT_NB][taps][C_in]
    __shared__ double s_red[2][SIMT_NB][SIMT_THREADS / 32];
    const int n0 = blockIdx.y * SIMT_NB;
    const int K = a.taps * a.C_in;
    for (int i = threadIdx.x; i < SIMT_NB * K; i += blockDim.x) {
        const int n = i / K, r = i - n * K;
        const int tap = r / a.C_in, c = r - tap * a.C_in;
        s_w[i] = (n0 + n < a.C_out) ? a.w[(static_cast<long>(n0 + n) * a.C_in + c) * a.taps + tap] : 0.f;
    }
    __syncthreads();
    const long m = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const bool row_ok = m < a.M;
    float acc[SIMT_NB];
#pragma unroll
    for (int n = 0; n < SIMT_NB; ++n) acc[n] = 0.f;
    if (row_ok) {
        int x = 0, y = 0;
        long base = m;
        if (a.mode == EML_CONV_3x3) { x = static_cast<int>(m % a.W); y = static_cast<int>((m / a.W) % a.H); }
        if (a.mode == EML_CONV_POOL2) {
            const int Wo = a.W >> 1, Ho = a.H >> 1;
            const int xo = static_cast<int>(m % Wo);
            const long t = m / Wo;
            const int yo = static_cast<int>(t % Ho);
            base = ((t / Ho) * a.H + 2 * yo) * a.W + 2 * xo;
        }
        for (int tap = 0; tap < a.taps; ++tap) {
            long src = base;
            if (a.mode == EML_CONV_3x3) {
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                if (x + dx < 0 || x + dx >= a.W || y + dy < 0 || y + dy >= a.H) continue;   // zero padding after the affine
                src = base + dy * a.W + dx;
            }
            const float *p = a.in + src * a.in_pitch;
            const float *wt = s_w + tap * a.C_in;
            for (int c = 0; c < a.C_in; ++c) {
                const float sc = a.scale ? a.scale[c] : 1.f, sh = a.shift ? a.shift[c] : 0.f;
                float v;
                if (a.mode == EML_CONV_POOL2) {
                    v = 0.f;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        float u = fmaf(p[static_cast<long>((t >> 1) * a.W + (t & 1)) * a.in_pitch + c], sc, sh);
                        v += a.relu ? fmaxf(u, 0.f) : u;
                    }
                    v *= 0.25f;
                } else {
                    v = fmaf(p[c], sc, sh);
                    if (a.relu) v = fmaxf(v, 0.f);
                }
#pragma unroll
                for (int n = 0; n < SIMT_NB; ++n) acc[n] = fmaf(v, wt[n * K + c], acc[n]);
            }
        }
        float *o = a.out + m * a.out_pitch + a.out_choff + n0;
#pragma unroll
        for (int n = 0; n < SIMT_NB; ++n)
            if (n0 + n < a.C_out) o[n] = acc[n];
    }
    if (a.stats != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int n = 0; n < SIMT_NB; ++n) {
            double v = row_ok ? static_cast<double>(acc[n]) : 0.0;
            double s1 = warp_sum_d(v), s2 = warp_sum_d(v * v);
            if (lane == 0) { s_red[0][n][warp] = s1; s_red[1][n][warp] = s2; }
        }
        __syncthreads();
        if (threadIdx.x < SIMT_NB && n0 + threadIdx.x < a.C_out) {
            double s1 = 0.0, s2 = 0.0;
            for (int w = 0; w < SIMT_THREADS / 32; ++w) { s1 += s_red[0][threadIdx.x][w]; s2 += s_red[1][threadIdx.x][w]; }
            atomicAdd(a.stats + n0 + threadIdx.x, s1);
            atomicAdd(a.stats + a.stats_stride + n0 + threadIdx.x, s2);
        }
    }
}

// ---------------------------------------------------------------------------------------------- stem
// One thread per pixel, all C_out (<= 32) channels; NCHW fp32 in (coalesced along x), NHWC out.
__global__ void __launch_bounds__(128) stem_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                   const float *__restrict__ scale, const float *__restrict__ shift,
                                                   float *__restrict__ out, int out_pitch, double *stats_raw,
                                                   double *stats_out, long so_stride, int B, int H, int W, int C_out, int write_out,
                                                   int relu) {
    __shared__ float s_w[32 * 27];
    __shared__ float s_sc[32], s_sh[32];
    __shared__ double s_red[4][32][4];
    for (int i = threadIdx.x; i < C_out * 27; i += blockDim.x) s_w[i] = w[i];      // [o][c][ky][kx]
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
        s_sc[i] = (scale && i < C_out) ? scale[i] : 1.f;
        s_sh[i] = (shift && i < C_out) ? shift[i] : 0.f;
    }
    __syncthreads();
    const long P = static_cast<long>(B) * H * W;
    const long m = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const bool ok = m < P;
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.f;
    if (ok) {
        const int xx = static_cast<int>(m % W), yy = static_cast<int>((m / W) % H);
        const long b = m / (static_cast<long>(W) * H);
        const float *xb = x + b * 3 * H * W;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int sy = yy + ky - 1, sx = xx + kx - 1;
                    float v = 0.f;
                    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = __ldg(xb + (static_cast<long>(c) * H + sy) * W + sx);
#pragma unroll
                    for (int o = 0; o < 32; ++o)
                        if (o < C_out) acc[o] = fmaf(v, s_w[o * 27 + c * 9 + ky * 3 + kx], acc[o]);
                }
    }
    float res[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        res[o] = fmaf(acc[o], s_sc[o], s_sh[o]);
        if (relu) res[o] = fmaxf(res[o], 0.f);
    }
    if (ok && write_out) {
        float *op = out + m * out_pitch;
        if ((out_pitch & 3) == 0 && (C_out & 3) == 0) {
#pragma unroll
            for (int o = 0; o < 32; o += 4)
                if (o < C_out) *reinterpret_cast<float4 *>(op + o) = make_float4(res[o], res[o + 1], res[o + 2], res[o + 3]);
        } else {
#pragma unroll
            for (int o = 0; o < 32; ++o)
                if (o < C_out) op[o] = res[o];
        }
    }
    if (stats_raw != nullptr || stats_out != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            if (o >= C_out) break;
            double a0 = ok ? static_cast<double>(acc[o]) : 0.0, r0 = ok ? static_cast<double>(res[o]) : 0.0;
            double t0 = warp_sum_d(a0), t1 = warp_sum_d(a0 * a0), t2 = warp_sum_d(r0), t3 = warp_sum_d(r0 * r0);
            if (lane == 0) { s_red[0][o][warp] = t0; s_red[1][o][warp] = t1; s_red[2][o][warp] = t2; s_red[3][o][warp] = t3; }
        }
        __syncthreads();
        if (threadIdx.x < C_out) {
            double t[4] = {0, 0, 0, 0};
            for (int q = 0; q < 4; ++q)
                for (int wv = 0; wv < 4; ++wv) t[q] += s_red[q][threadIdx.x][wv];
            if (stats_raw) { atomicAdd(stats_raw + threadIdx.x, t[0]); atomicAdd(stats_raw + C_out + threadIdx.x, t[1]); }
            if (stats_out) { atomicAdd(stats_out + threadIdx.x, t[2]); atomicAdd(stats_out + so_stride + threadIdx.x, t[3]); }
        }
    }
}

// Variant 2 of the stem: the default for C_out = 24 since round 2 (timed on B200: 55.66 -> 55.26 ms per B=256 step, outputs bit-identical to
// variant 1 -- same fmaf order -- tests/test_experiments_gpu.py; EML_STEM_V1=1 selects the old kernel).
// SASS of stem_kernel shows 786 LDS for 896 FFMA -- the weights are read from shared memory one scalar per FMA ([o][tap] layout,
// stride 27), so the kernel is bound by shared-memory issue (1 LDS / clk / SM against 4 FFMA warps / clk / SM).  Here the weights
// are stored [tap][o]: one broadcast LDS.128 feeds four FMAs, and the output count is a template parameter so that the 8 dead
// accumulators of C_out = 24 are gone.
template <int CO>
__global__ void __launch_bounds__(128) stem_kernel_v2(const float *__restrict__ x, const float *__restrict__ w,
                                                      const float *__restrict__ scale, const float *__restrict__ shift,
                                                      float *__restrict__ out, int out_pitch, double *stats_raw,
                                                      double *stats_out, long so_stride, int B, int H, int W, int write_out,
                                                      int relu) {
    static_assert(CO % 4 == 0 && CO <= 32, "stem_kernel_v2: C_out must be a multiple of 4, at most 32");
    __shared__ __align__(16) float s_w[27 * CO];
    __shared__ float s_sc[CO], s_sh[CO];
    __shared__ double s_red[4][CO][4];
    for (int i = threadIdx.x; i < CO * 27; i += blockDim.x) s_w[(i % 27) * CO + i / 27] = w[i];      // w is [o][c][ky][kx]
    for (int i = threadIdx.x; i < CO; i += blockDim.x) {
        s_sc[i] = scale ? scale[i] : 1.f;
        s_sh[i] = shift ? shift[i] : 0.f;
    }
    __syncthreads();
    const long P = static_cast<long>(B) * H * W;
    const long m = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const bool ok = m < P;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    if (ok) {
        const int xx = static_cast<int>(m % W), yy = static_cast<int>((m / W) % H);
        const long b = m / (static_cast<long>(W) * H);
        const float *xb = x + b * 3 * H * W;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int sy = yy + ky - 1, sx = xx + kx - 1;
                    float v = 0.f;
                    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = __ldg(xb + (static_cast<long>(c) * H + sy) * W + sx);
                    const float4 *wt = reinterpret_cast<const float4 *>(s_w + (c * 9 + ky * 3 + kx) * CO);
#pragma unroll
                    for (int o4 = 0; o4 < CO / 4; ++o4) {
                        const float4 wv = wt[o4];
                        acc[4 * o4 + 0] = fmaf(v, wv.x, acc[4 * o4 + 0]);
                        acc[4 * o4 + 1] = fmaf(v, wv.y, acc[4 * o4 + 1]);
                        acc[4 * o4 + 2] = fmaf(v, wv.z, acc[4 * o4 + 2]);
                        acc[4 * o4 + 3] = fmaf(v, wv.w, acc[4 * o4 + 3]);
                    }
                }
    }
    float res[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        res[o] = fmaf(acc[o], s_sc[o], s_sh[o]);
        if (relu) res[o] = fmaxf(res[o], 0.f);
    }
    if (ok && write_out) {
        float *op = out + m * out_pitch;
        if ((out_pitch & 3) == 0) {
#pragma unroll
            for (int o = 0; o < CO; o += 4) *reinterpret_cast<float4 *>(op + o) = make_float4(res[o], res[o + 1], res[o + 2], res[o + 3]);
        } else {
#pragma unroll
            for (int o = 0; o < CO; ++o) op[o] = res[o];
        }
    }
    if (stats_raw != nullptr || stats_out != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            double a0 = ok ? static_cast<double>(acc[o]) : 0.0, r0 = ok ? static_cast<double>(res[o]) : 0.0;
            double t0 = warp_sum_d(a0), t1 = warp_sum_d(a0 * a0), t2 = warp_sum_d(r0), t3 = warp_sum_d(r0 * r0);
            if (lane == 0) { s_red[0][o][warp] = t0; s_red[1][o][warp] = t1; s_red[2][o][warp] = t2; s_red[3][o][warp] = t3; }
        }
        __syncthreads();
        if (threadIdx.x < CO) {
            double t[4] = {0, 0, 0, 0};
            for (int q = 0; q < 4; ++q)
                for (int wv = 0; wv < 4; ++wv) t[q] += s_red[q][threadIdx.x][wv];
            if (stats_raw) { atomicAdd(stats_raw + threadIdx.x, t[0]); atomicAdd(stats_raw + CO + threadIdx.x, t[1]); }
            if (stats_out) { atomicAdd(stats_out + threadIdx.x, t[2]); atomicAdd(stats_out + so_stride + threadIdx.x, t[3]); }
        }
    }
}

// ---------------------------------------------------------------------------------------------- BN fold
__global__ void bn_fold_kernel(const double *stats, long sstride, double count, const float *rmean, const float *rvar,
                               const float *gamma, const float *beta, const float *pre_scale,
                               const float *pre_shift, float *scale, float *shift, float *bmean, float *bvar,
                               int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double a = pre_scale ? static_cast<double>(pre_scale[c]) : 1.0;
    const double b = pre_shift ? static_cast<double>(pre_shift[c]) : 0.0;
    double mean_u, var_u;                      // statistics of u = a*t + b
    if (stats != nullptr) {
        const double mt = stats[c] / count;
        double vt = stats[sstride + c] / count - mt * mt;
        if (vt < 0.0) vt = 0.0;
        mean_u = a * mt + b;
        var_u = a * a * vt;                    // biased variance, as F.batch_norm(training=True) normalises with
    } else {
        mean_u = static_cast<double>(rmean[c]);
        var_u = static_cast<double>(rvar[c]);
    }
    const double g = gamma ? static_cast<double>(gamma[c]) : 1.0;
    const double be = beta ? static_cast<double>(beta[c]) : 0.0;
    const double inv = g / sqrt(var_u + static_cast<double>(eps));
    // BN(u) = inv*(u - mean_u) + beta = inv*a*t + inv*(b - mean_u) + beta
    scale[c] = static_cast<float>(inv * a);
    shift[c] = static_cast<float>(inv * (b - mean_u) + be);
    if (bmean) bmean[c] = static_cast<float>(mean_u);
    if (bvar) bvar[c] = static_cast<float>(var_u);
}

// ---------------------------------------------------------------------------------------------- head pool
// out[b, (yo*Wp + xo)*C + c] = mean_{pool x pool} relu(scale[c]*in + shift[c])
__global__ void head_pool_kernel(const float *__restrict__ in, int in_pitch, const float *__restrict__ scale,
                                 const float *__restrict__ shift, float *__restrict__ out, int B, int H, int W,
                                 int C, int pool) {
    const int Hp = H / pool, Wp = W / pool;
    const long total = static_cast<long>(B) * Hp * Wp * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long t = i / C;
        const int xo = static_cast<int>(t % Wp); t /= Wp;
        const int yo = static_cast<int>(t % Hp);
        const long b = t / Hp;
        const float sc = scale ? scale[c] : 1.f, sh = shift ? shift[c] : 0.f;
        float s = 0.f;
        for (int dy = 0; dy < pool; ++dy)
            for (int dx = 0; dx < pool; ++dx) {
                const long p = (b * H + yo * pool + dy) * W + xo * pool + dx;
                s += fmaxf(fmaf(in[p * in_pitch + c], sc, sh), 0.f);
            }
        out[i] = s / static_cast<float>(pool * pool);
    }
}

// ---------------------------------------------------------------------------------------------- linear
// out (M,N) = a (M,K) @ w (N,K)^T + bias.  64x64 tile, 16-wide k-slab, 4x4 micro-tile per thread.
__global__ void __launch_bounds__(256) linear_kernel(const float *__restrict__ a, const float *__restrict__ w,
                                                     const float *__restrict__ bias, float *__restrict__ out,
                                                     int M, int N, int K) {
    __shared__ float sa[16][64 + 4], sw[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, k = i & 15;
            sa[k][r] = (m0 + r < M && k0 + k < K) ? a[static_cast<long>(m0 + r) * K + k0 + k] : 0.f;
            sw[k][r] = (n0 + r < N && k0 + k < K) ? w[static_cast<long>(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; wv[i] = sw[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) out[static_cast<long>(m) * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
        }
}

// Few rows (M <= 16: the ConvEncoder's fc at batch <= 16, generator.py:124): the 64x64 tiling leaves N/64 CTAs walking all of K
// serially (1.4 ms for 16 x 8192 x 256).  Here one warp owns one output feature: it streams that weight row once (float4 per lane,
// coalesced), keeps M partial sums per lane against the L1/L2-resident activations and reduces them with shuffles.
__global__ void __launch_bounds__(256) linear_rows_kernel(const float *__restrict__ a, const float *__restrict__ w,
                                                          const float *__restrict__ bias, float *__restrict__ out,
                                                          int M, int N, int K) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = 0.f;
    const float *wr = w + static_cast<long>(n) * K;
    if ((K & 3) == 0) {
        for (int k = lane * 4; k < K; k += 128) {
            const float4 wv = __ldg(reinterpret_cast<const float4 *>(wr + k));
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                if (m < M) {
                    const float4 av = __ldg(reinterpret_cast<const float4 *>(a + static_cast<long>(m) * K + k));
                    acc[m] = fmaf(av.x, wv.x, fmaf(av.y, wv.y, fmaf(av.z, wv.z, fmaf(av.w, wv.w, acc[m]))));
                }
            }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            const float wv = __ldg(wr + k);
#pragma unroll
            for (int m = 0; m < 16; ++m)
                if (m < M) acc[m] = fmaf(__ldg(a + static_cast<long>(m) * K + k), wv, acc[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        float v = acc[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && m < M) out[static_cast<long>(m) * N + n] = v + (bias ? bias[n] : 0.f);
    }
}

}  // namespace

int eml_conv_forward_simt(const eml_conv_params *p, cudaStream_t st) {
    EML_CHECK_PTR(p->w_oihw);
    SimtArgs a{};
    a.in = p->in; a.scale = p->scale; a.shift = p->shift; a.w = p->w_oihw; a.out = p->out; a.stats = p->stats;
    a.stats_stride = p->stats_stride > 0 ? p->stats_stride : p->C_out;
    a.H = p->H; a.W = p->W; a.C_in = p->C_in; a.in_pitch = p->in_pitch; a.C_out = p->C_out;
    a.out_pitch = p->out_pitch; a.out_choff = p->out_choff; a.mode = p->mode; a.relu = p->relu;
    a.taps = p->mode == EML_CONV_3x3 ? 9 : 1;
    a.M = (p->mode == EML_CONV_POOL2) ? static_cast<long>(p->B) * (p->H / 2) * (p->W / 2)
                                      : static_cast<long>(p->B) * p->H * p->W;
    const size_t smem = sizeof(float) * SIMT_NB * a.taps * a.C_in;
    if (smem > 200 * 1024) return EML_E_SHAPE;
    cudaError_t e = cudaFuncSetAttribute(conv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    dim3 grid(static_cast<unsigned>((a.M + SIMT_THREADS - 1) / SIMT_THREADS), (a.C_out + SIMT_NB - 1) / SIMT_NB);
    conv_simt_kernel<<<grid, SIMT_THREADS, smem, st>>>(a);
    return eml_launch_status();
}

extern "C" int eml_stem_forward(const float *x_nchw, const float *w_oihw, const float *scale, const float *shift,
                                float *out, int out_pitch, double *stats_raw, double *stats_out,
                                long stats_out_stride, int B, int H, int W, int C_out, int write_out, int relu,
                                void *stream) {
    EML_CHECK_PTR(x_nchw); EML_CHECK_PTR(w_oihw);
    if (write_out) { EML_CHECK_PTR(out); EML_CHECK_ALIGN16(out); }
    if (B <= 0 || H <= 0 || W <= 0 || C_out <= 0 || C_out > 32 || out_pitch < C_out) return EML_E_SHAPE;
    const long P = static_cast<long>(B) * H * W;
    static const bool v2 = !eml_env_flag("EML_STEM_V1");
    if (v2 && C_out == 24) {
        stem_kernel_v2<24><<<static_cast<unsigned>((P + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
            x_nchw, w_oihw, scale, shift, out, out_pitch, stats_raw, stats_out, stats_out_stride > 0 ? stats_out_stride : C_out, B, H, W,
            write_out, relu);
        return eml_launch_status();
    }
    stem_kernel<<<static_cast<unsigned>((P + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x_nchw, w_oihw, scale, shift, out, out_pitch, stats_raw, stats_out,
        stats_out_stride > 0 ? stats_out_stride : C_out, B, H, W, C_out, write_out, relu);
    return eml_launch_status();
}

extern "C" int eml_bn_fold(const double *stats, long stats_stride, double count, const float *running_mean, const float *running_var,
                           const float *gamma, const float *beta, const float *pre_scale, const float *pre_shift,
                           float *scale, float *shift, float *batch_mean, float *batch_var, int C, float eps,
                           void *stream) {
    EML_CHECK_PTR(scale); EML_CHECK_PTR(shift);
    if (C <= 0) return EML_E_SHAPE;
    if (stats == nullptr && (running_mean == nullptr || running_var == nullptr)) return EML_E_NULL;
    if (stats != nullptr && !(count > 0)) return EML_E_ARG;
    bn_fold_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, stats_stride > 0 ? stats_stride : C, count, running_mean, running_var, gamma, beta, pre_scale, pre_shift, scale, shift, batch_mean,
        batch_var, C, eps);
    return eml_launch_status();
}

extern "C" int eml_head_pool(const float *in, int in_pitch, const float *scale, const float *shift, float *out,
                             int B, int H, int W, int C, int pool, void *stream) {
    EML_CHECK_PTR(in); EML_CHECK_PTR(out);
    if (B <= 0 || C <= 0 || pool <= 0 || H % pool || W % pool || in_pitch < C) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * (H / pool) * (W / pool) * C;
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    head_pool_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, in_pitch, scale, shift, out, B, H, W, C, pool);
    return eml_launch_status();
}

extern "C" int eml_linear_fp32(const float *a, const float *w, const float *bias, float *out, int M, int N, int K,
                               void *stream) {
    EML_CHECK_PTR(a); EML_CHECK_PTR(w); EML_CHECK_PTR(out);
    if (M <= 0 || N <= 0 || K <= 0) return EML_E_SHAPE;
    if (M <= 16 && (((K & 3) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) || (K & 3))) {
        linear_rows_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, w, bias, out, M, N, K);
        return eml_launch_status();
    }
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    linear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, w, bias, out, M, N, K);
    return eml_launch_status();
}
