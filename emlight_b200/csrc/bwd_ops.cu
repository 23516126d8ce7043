// DenseNet-BC backward support kernels (fp32 FFMA), sm_100a.  Activations / gradients NHWC fp32.
//
// The two data-gradient convolutions reuse the tcgen05 implicit-GEMM kernels (eml_conv_forward with transposed / flipped
// weights).  What remains is BatchNorm's training-mode backward (two per-channel reductions, then an elementwise apply that
// also applies the ReLU mask and accumulates into the dense block's gradient slab) and the weight gradients, which are
// reductions over all pixels (K = B*H*W) producing tiny outputs (48 x C_in, 12 x 48 x 9): here plain register-tiled FFMA
// kernels with one atomicAdd per output per block.  Everything the forward did not store is recomputed from the stored
// slab: u = pre_a*x + pre_b (folded last_norm), xhat = (u - mean)*inv, z = gamma*xhat + beta, relu mask = z > 0.
#include "common.cuh"

namespace {

struct BnArgs {
    const float *grad; int g_pitch;      // gradient w.r.t. the BN(+ReLU) output; pool=1: indexed by the 2x2-pooled pixel, scaled by 1/4
    const float *x; int x_pitch;         // stored input of the BN
    const float *pre_a, *pre_b;          // NULL => identity
    const float *mean, *inv;             // batch statistics of u
    const float *gamma, *beta;
    int relu, pool, H, W;
    long M;                              // rows of x
    int C;
};

__device__ __forceinline__ float bn_masked_grad(const BnArgs &a, long m, int c, float &xhat) {
    const float xv = a.x[m * a.x_pitch + c];
    const float u = (a.pre_a ? a.pre_a[c] : 1.f) * xv + (a.pre_b ? a.pre_b[c] : 0.f);
    xhat = (u - a.mean[c]) * a.inv[c];
    long gm = m;
    float gs = 1.f;
    if (a.pool) {
        const int xx = static_cast<int>(m % a.W);
        const long t = m / a.W;
        const int yy = static_cast<int>(t % a.H);
        const long b = t / a.H;
        gm = (b * (a.H >> 1) + (yy >> 1)) * (a.W >> 1) + (xx >> 1);
        gs = 0.25f;
    }
    float g = a.grad[gm * a.g_pitch + c] * gs;
    if (a.relu && !(fmaf(a.gamma[c], xhat, a.beta[c]) > 0.f)) g = 0.f;
    return g;
}

// sums[c] += sum_m g ; sums[stride + c] += sum_m g * xhat.  Block = 32 channels x 8 row lanes, rows strided by gridDim.y.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnArgs a, double *sums, long stride) {
    __shared__ double s1[8][32], s2[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    double a1 = 0.0, a2 = 0.0;
    if (c < a.C) {
        for (long m = blockIdx.y * 8L + rl; m < a.M; m += gridDim.y * 8L) {
            float xhat;
            const float g = bn_masked_grad(a, m, c, xhat);
            a1 += g; a2 += static_cast<double>(g) * xhat;
        }
    }
    s1[rl][threadIdx.x & 31] = a1; s2[rl][threadIdx.x & 31] = a2;
    __syncthreads();
    if (threadIdx.x < 32 && c < a.C) {
        double t1 = 0.0, t2 = 0.0;
        for (int r = 0; r < 8; ++r) { t1 += s1[r][threadIdx.x]; t2 += s2[r][threadIdx.x]; }
        atomicAdd(sums + c, t1);
        atomicAdd(sums + stride + c, t2);
    }
}

// du = gamma*inv*(g - S1/M - xhat*S2/M); out (+)= du * (to_stored ? pre_a : 1).  Block = 32 channels x 8 row lanes (coalesced
// 128-byte channel segments, no per-element division).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnArgs a, const double *sums, long stride, float *out, int o_pitch,
                                                           int accumulate, int to_stored) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    if (c >= a.C) return;
    const double invM = 1.0 / static_cast<double>(a.M);
    const float m1 = static_cast<float>(sums[c] * invM), m2 = static_cast<float>(sums[stride + c] * invM);
    float k = a.gamma[c] * a.inv[c];
    if (to_stored && a.pre_a) k *= a.pre_a[c];
    for (long m = blockIdx.y * 8L + rl; m < a.M; m += gridDim.y * 8L) {
        float xhat;
        const float g = bn_masked_grad(a, m, c, xhat);
        const float du = k * (g - m1 - xhat * m2);
        float *o = out + m * o_pitch + c;
        *o = accumulate ? *o + du : du;
    }
}


// ---- float4 variants (all pitches multiples of 4, 16-byte aligned bases): thread = 4 channels, 4 rows in flight ---------------
__device__ __forceinline__ long bn_grad_row(const BnArgs &a, long m, float &gs) {
    gs = 1.f;
    if (!a.pool) return m;
    const int xx = static_cast<int>(m % a.W);
    const long t = m / a.W;
    const int yy = static_cast<int>(t % a.H);
    const long b = t / a.H;
    gs = 0.25f;
    return (b * (a.H >> 1) + (yy >> 1)) * (a.W >> 1) + (xx >> 1);
}
struct Bn4 { float pa[4], pb[4], mean[4], inv[4], gamma[4], beta[4]; };
__device__ __forceinline__ void bn_load4(const BnArgs &a, int c, Bn4 &p) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool ok = c + e < a.C;
        p.pa[e] = (ok && a.pre_a) ? a.pre_a[c + e] : 1.f;
        p.pb[e] = (ok && a.pre_b) ? a.pre_b[c + e] : 0.f;
        p.mean[e] = ok ? a.mean[c + e] : 0.f;
        p.inv[e] = ok ? a.inv[c + e] : 0.f;
        p.gamma[e] = ok ? a.gamma[c + e] : 0.f;
        p.beta[e] = ok ? a.beta[c + e] : 0.f;
    }
}
__device__ __forceinline__ void bn_eval4(const BnArgs &a, const Bn4 &p, const float4 xv, const float4 gv, float gs, float (&g)[4], float (&xh)[4]) {
    const float x4[4] = {xv.x, xv.y, xv.z, xv.w}, g4[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        xh[e] = (fmaf(p.pa[e], x4[e], p.pb[e]) - p.mean[e]) * p.inv[e];
        g[e] = g4[e] * gs;
        if (a.relu && !(fmaf(p.gamma[e], xh[e], p.beta[e]) > 0.f)) g[e] = 0.f;
    }
}

// thread -> (channel quad q, row lane rl).  Wide tensors (more than 32 quads): a warp = 32 consecutive quads of one row, 8 row lanes,
// block column blockIdx.x.  Narrow ones (the 48-channel bottleneck of every dense layer: 12 quads) would leave 20 of 32 lanes idle that
// way -- there the block's 256 threads are dealt as (256 / nq) row lanes x nq quads, one block column.
struct BnMap { int c, rl, rows, nq, q; bool active; };
__device__ __forceinline__ BnMap bn_map(const BnArgs &a) {
    BnMap m;
    const int cq = (a.C + 3) >> 2;
    if (cq <= 32) {
        m.nq = cq; m.rows = 256 / cq; m.q = threadIdx.x % cq; m.rl = threadIdx.x / cq;
        m.c = m.q * 4; m.active = m.rl < m.rows;
    } else {
        m.nq = 32; m.rows = 8; m.q = threadIdx.x & 31; m.rl = threadIdx.x >> 5;
        m.c = (blockIdx.x * 32 + m.q) * 4; m.active = m.c < a.C;
    }
    return m;
}

__global__ void __launch_bounds__(256) bn_bwd_reduce4_kernel(const BnArgs a, double *sums, long stride) {
    __shared__ double s1[256][4], s2[256][4];
    const BnMap mp = bn_map(a);
    const int c = mp.c, rl = mp.rl;
    double a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
    if (mp.active) {
        Bn4 p; bn_load4(a, c, p);
        for (long m0 = (static_cast<long>(blockIdx.y) * mp.rows + rl) * 4; m0 < a.M; m0 += static_cast<long>(gridDim.y) * mp.rows * 4) {
            float4 xv[4], gv[4]; float gs[4]; bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long m = m0 + u;
                ok[u] = m < a.M;
                if (!ok[u]) continue;
                const long gm = bn_grad_row(a, m, gs[u]);
                xv[u] = __ldg(reinterpret_cast<const float4 *>(a.x + m * a.x_pitch + c));
                gv[u] = __ldg(reinterpret_cast<const float4 *>(a.grad + gm * a.g_pitch + c));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ok[u]) continue;
                float g[4], xh[4];
                bn_eval4(a, p, xv[u], gv[u], gs[u], g, xh);
#pragma unroll
                for (int e = 0; e < 4; ++e) { a1[e] += g[e]; a2[e] += static_cast<double>(g[e]) * xh[e]; }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { s1[threadIdx.x][e] = a1[e]; s2[threadIdx.x][e] = a2[e]; }      // idle threads hold zeros
    __syncthreads();
    if (threadIdx.x < mp.nq * 4) {
        const int q = threadIdx.x >> 2, e = threadIdx.x & 3;
        const int cc = (mp.nq == 32 && ((a.C + 3) >> 2) > 32 ? blockIdx.x * 128 : 0) + q * 4 + e;
        if (cc < a.C) {
            double t1 = 0.0, t2 = 0.0;
            for (int r = 0; r < mp.rows; ++r) { t1 += s1[r * mp.nq + q][e]; t2 += s2[r * mp.nq + q][e]; }
            atomicAdd(sums + cc, t1);
            atomicAdd(sums + stride + cc, t2);
        }
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply4_kernel(const BnArgs a, const double *sums, long stride, float *out, int o_pitch,
                                                            int accumulate, int to_stored) {
    const BnMap mp = bn_map(a);
    const int c = mp.c, rl = mp.rl;
    if (!mp.active) return;
    Bn4 p; bn_load4(a, c, p);
    const double invM = 1.0 / static_cast<double>(a.M);
    float k[4], m1[4], m2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool okc = c + e < a.C;
        m1[e] = okc ? static_cast<float>(sums[c + e] * invM) : 0.f;
        m2[e] = okc ? static_cast<float>(sums[stride + c + e] * invM) : 0.f;
        k[e] = p.gamma[e] * p.inv[e] * (to_stored ? p.pa[e] : 1.f);
    }
    const bool full = c + 3 < a.C;
    for (long m0 = (static_cast<long>(blockIdx.y) * mp.rows + rl) * 4; m0 < a.M; m0 += static_cast<long>(gridDim.y) * mp.rows * 4) {
        float4 xv[4], gv[4], ov[4]; float gs[4]; bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long m = m0 + u;
            ok[u] = m < a.M;
            if (!ok[u]) continue;
            const long gm = bn_grad_row(a, m, gs[u]);
            xv[u] = __ldg(reinterpret_cast<const float4 *>(a.x + m * a.x_pitch + c));
            gv[u] = *reinterpret_cast<const float4 *>(a.grad + gm * a.g_pitch + c);       // may alias `out` (in-place BN2 backward)
            if (accumulate) ov[u] = *reinterpret_cast<const float4 *>(out + m * o_pitch + c);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            float g[4], xh[4], du[4];
            bn_eval4(a, p, xv[u], gv[u], gs[u], g, xh);
#pragma unroll
            for (int e = 0; e < 4; ++e) du[e] = k[e] * (g[e] - m1[e] - xh[e] * m2[e]);
            float *o = out + (m0 + u) * o_pitch + c;
            if (full) {
                float4 r = make_float4(du[0], du[1], du[2], du[3]);
                if (accumulate) { r.x += ov[u].x; r.y += ov[u].y; r.z += ov[u].z; r.w += ov[u].w; }
                *reinterpret_cast<float4 *>(o) = r;
            } else {                                           // tail quad: the neighbouring channels belong to other layers
                const float o4[4] = {ov[u].x, ov[u].y, ov[u].z, ov[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c + e < a.C) o[e] = accumulate ? o4[e] + du[e] : du[e];
            }
        }
    }
}

inline bool bn_vec_ok(const BnArgs &a, const float *out, int o_pitch) {
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return (a.g_pitch & 3) == 0 && (a.x_pitch & 3) == 0 && al(a.grad) && al(a.x) && (out == nullptr || ((o_pitch & 3) == 0 && al(out)));
}

// ---------------------------------------------------------------------------------------------- weight gradients
// dW[n, c] += sum_m G[m, n] * A(m, c);  A = act(scale*x + shift) (optionally the average over the 2x2 pool window; then G / m index
// pooled pixels).  Block tile 48 (n) x 64 (c); thread tile 3 x 4; rows in slabs of 32 staged in shared memory.
constexpr int WG_TN = 48, WG_TC = 64, WG_R = 32;
__global__ void __launch_bounds__(256) wgrad_1x1_kernel(const float *__restrict__ G, int g_pitch, int N, const float *__restrict__ x,
                                                        int x_pitch, int C, const float *__restrict__ scale,
                                                        const float *__restrict__ shift, int relu, int pool, int H, int W,
                                                        float *__restrict__ dW, long M) {
    __shared__ float sG[WG_R][WG_TN + 1], sA[WG_R][WG_TC + 1];
    const int c0 = blockIdx.y * WG_TC, n0 = blockIdx.z * WG_TN;
    const int tn = (threadIdx.x >> 4) * 3, tc = (threadIdx.x & 15) * 4;       // 16 x 16 threads -> 48 x 64 outputs
    float acc[3][4] = {};
    const long rows_per = (M + gridDim.x - 1) / gridDim.x;
    const long r_beg = blockIdx.x * rows_per, r_end = min(M, r_beg + rows_per);
    for (long r0 = r_beg; r0 < r_end; r0 += WG_R) {
        for (int i = threadIdx.x; i < WG_R * WG_TN; i += 256) {
            const int r = i / WG_TN, n = i - r * WG_TN;
            const long m = r0 + r;
            sG[r][n] = (m < r_end && n0 + n < N) ? G[m * g_pitch + n0 + n] : 0.f;
        }
        for (int i = threadIdx.x; i < WG_R * WG_TC; i += 256) {
            const int r = i / WG_TC, c = i - r * WG_TC;
            const long m = r0 + r;
            float v = 0.f;
            if (m < r_end && c0 + c < C) {
                const float sc = scale ? scale[c0 + c] : 1.f, sh = shift ? shift[c0 + c] : 0.f;
                if (pool) {
                    const int Wo = W >> 1, Ho = H >> 1;
                    const int xo = static_cast<int>(m % Wo);
                    const long t = m / Wo;
                    const int yo = static_cast<int>(t % Ho);
                    const long base = ((t / Ho) * H + 2 * yo) * W + 2 * xo;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float u = fmaf(x[(base + (q >> 1) * W + (q & 1)) * x_pitch + c0 + c], sc, sh);
                        v += relu ? fmaxf(u, 0.f) : u;
                    }
                    v *= 0.25f;
                } else {
                    v = fmaf(x[m * x_pitch + c0 + c], sc, sh);
                    if (relu) v = fmaxf(v, 0.f);
                }
            }
            sA[r][c] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < WG_R; ++r) {
            float g[3], av[4];
#pragma unroll
            for (int i = 0; i < 3; ++i) g[i] = sG[r][tn + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) av[j] = sA[r][tc + j];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], av[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n0 + tn + i < N && c0 + tc + j < C) atomicAdd(dW + static_cast<long>(n0 + tn + i) * C + c0 + tc + j, acc[i][j]);
}

// dW[n, c, tap] += sum_m dY[m, n] * (scale[c]*b[m + tap, c] + shift[c])  (0 outside the image); N <= 16, C <= 64.
// Block tile: all taps, thread = (n-group of 4, c) -> 9 x 4 accumulators.  Rows in slabs of 32 pixels of one image row.
__global__ void __launch_bounds__(256) wgrad_3x3_kernel(const float *__restrict__ dY, int dy_pitch, int N, const float *__restrict__ bt,
                                                        int b_pitch, int C, const float *__restrict__ scale,
                                                        const float *__restrict__ shift, float *__restrict__ dW, int B, int H, int W) {
    constexpr int SEG = 32;
    __shared__ float sY[SEG][16 + 1];
    __shared__ float sB[3][SEG + 2][64 + 1];
    const int c = threadIdx.x & 63, ng = threadIdx.x >> 6;                  // 64 channels x 4 n-groups (4 n each)
    float acc[9][4] = {};
    const int segs_x = (W + SEG - 1) / SEG;
    const long nseg = static_cast<long>(B) * H * segs_x;
    for (long s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int sx = static_cast<int>(s % segs_x);
        const long t = s / segs_x;
        const int y = static_cast<int>(t % H);
        const long b = t / H;
        const int x0 = sx * SEG;
        for (int i = threadIdx.x; i < SEG * 16; i += 256) {
            const int p = i >> 4, n = i & 15;
            sY[p][n] = (x0 + p < W && n < N) ? dY[((b * H + y) * static_cast<long>(W) + x0 + p) * dy_pitch + n] : 0.f;
        }
        for (int i = threadIdx.x; i < 3 * (SEG + 2) * 64; i += 256) {
            const int cc = i & 63;
            const int p = (i >> 6) % (SEG + 2), ry = (i >> 6) / (SEG + 2);
            const int iy = y + ry - 1, ix = x0 + p - 1;
            float v = 0.f;
            if (cc < C && iy >= 0 && iy < H && ix >= 0 && ix < W)
                v = fmaf(bt[((b * H + iy) * static_cast<long>(W) + ix) * b_pitch + cc], scale ? scale[cc] : 1.f, shift ? shift[cc] : 0.f);
            sB[ry][p][cc] = v;
        }
        __syncthreads();
        for (int p = 0; p < SEG; ++p) {
            float yv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) yv[i] = sY[p][ng * 4 + i];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float bv = sB[tap / 3][p + tap % 3][c];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[tap][i] = fmaf(yv[i], bv, acc[tap][i]);
            }
        }
        __syncthreads();
    }
    if (c < C)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (ng * 4 + i < N) atomicAdd(dW + (static_cast<long>(ng * 4 + i) * C + c) * 9 + tap, acc[tap][i]);
}

// stem: dW0[o, ci, ky, kx] += sum_m dZ[m, o] * x_nchw[b, ci, y+ky-1, x+kx-1];  thread = (o < 32, k < 27) looped, pixels strided.
__global__ void __launch_bounds__(256) wgrad_stem_kernel(const float *__restrict__ dZ, int dz_pitch, int O, const float *__restrict__ x,
                                                         float *__restrict__ dW, int B, int H, int W) {
    constexpr int SEG = 64;
    __shared__ float sZ[SEG][32 + 1];
    __shared__ float sX[3][3][SEG + 2];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};                 // outputs idx = threadIdx.x + 256*j over O*27 (<= 864)
    const int segs_x = (W + SEG - 1) / SEG;
    const long nseg = static_cast<long>(B) * H * segs_x;
    for (long s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int sx = static_cast<int>(s % segs_x);
        const long t = s / segs_x;
        const int y = static_cast<int>(t % H);
        const long b = t / H;
        const int x0 = sx * SEG;
        for (int i = threadIdx.x; i < SEG * 32; i += 256) {
            const int p = i >> 5, o = i & 31;
            sZ[p][o] = (x0 + p < W && o < O) ? dZ[((b * H + y) * static_cast<long>(W) + x0 + p) * dz_pitch + o] : 0.f;
        }
        for (int i = threadIdx.x; i < 9 * (SEG + 2); i += 256) {
            const int p = i % (SEG + 2), r = i / (SEG + 2);
            const int ci = r / 3, ry = r % 3;
            const int iy = y + ry - 1, ix = x0 + p - 1;
            sX[ci][ry][p] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[((b * 3 + ci) * H + iy) * static_cast<long>(W) + ix] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = threadIdx.x + 256 * j;
            if (idx < O * 27) {
                const int o = idx / 27, k = idx - o * 27;
                const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
                float a = 0.f;
                for (int p = 0; p < SEG; ++p) a = fmaf(sZ[p][o], sX[ci][ky][p + kx], a);
                acc[j] += a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int idx = threadIdx.x + 256 * j;
        if (idx < O * 27) atomicAdd(dW + idx, acc[j]);
    }
}

BnArgs make_bn(const float *grad, int g_pitch, const float *x, int x_pitch, const float *pre_a, const float *pre_b, const float *mean,
               const float *inv, const float *gamma, const float *beta, int relu, int pool, int H, int W, long M, int C) {
    BnArgs a{};
    a.grad = grad; a.g_pitch = g_pitch; a.x = x; a.x_pitch = x_pitch; a.pre_a = pre_a; a.pre_b = pre_b; a.mean = mean; a.inv = inv;
    a.gamma = gamma; a.beta = beta; a.relu = relu; a.pool = pool; a.H = H; a.W = W; a.M = M; a.C = C;
    return a;
}

}  // namespace

extern "C" int eml_bn_bwd_reduce(const float *grad, int g_pitch, const float *x, int x_pitch, const float *pre_a, const float *pre_b,
                                 const float *mean, const float *inv_std, const float *gamma, const float *beta, int relu, int pool,
                                 int H, int W, long M, int C, double *sums, long sums_stride, void *stream) {
    EML_CHECK_PTR(grad); EML_CHECK_PTR(x); EML_CHECK_PTR(mean); EML_CHECK_PTR(inv_std); EML_CHECK_PTR(gamma); EML_CHECK_PTR(beta);
    EML_CHECK_PTR(sums);
    if (M <= 0 || C <= 0 || x_pitch < C || g_pitch < C || (pool && (H <= 0 || W <= 0 || ((H | W) & 1)))) return EML_E_SHAPE;
    const BnArgs a = make_bn(grad, g_pitch, x, x_pitch, pre_a, pre_b, mean, inv_std, gamma, beta, relu, pool, H, W, M, C);
    if (bn_vec_ok(a, nullptr, 0)) {
        const int cq = (C + 3) / 4;
        const int gx = cq <= 32 ? 1 : (C + 127) / 128;                   // bn_map: narrow tensors use one block column
        const int rows4 = (cq <= 32 ? 256 / cq : 8) * 4;                 // rows per block and pass
        long gy = (M + rows4 * 16 - 1) / (rows4 * 16);
        const long cap = (148L * 8 + gx - 1) / gx;
        if (gy > cap) gy = cap;
        dim3 grid(gx, static_cast<unsigned>(gy < 1 ? 1 : gy));
        bn_bwd_reduce4_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, sums, sums_stride > 0 ? sums_stride : C);
        return eml_launch_status();
    }
    long gy = (M + 8 * 64 - 1) / (8 * 64);
    if (gy > 148 * 4) gy = 148 * 4;
    dim3 grid((C + 31) / 32, static_cast<unsigned>(gy < 1 ? 1 : gy));
    bn_bwd_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, sums, sums_stride > 0 ? sums_stride : C);
    return eml_launch_status();
}

extern "C" int eml_bn_bwd_apply(const float *grad, int g_pitch, const float *x, int x_pitch, const float *pre_a, const float *pre_b,
                                const float *mean, const float *inv_std, const float *gamma, const float *beta, int relu, int pool,
                                int H, int W, long M, int C, const double *sums, long sums_stride, float *out, int out_pitch,
                                int accumulate, int to_stored, void *stream) {
    EML_CHECK_PTR(grad); EML_CHECK_PTR(x); EML_CHECK_PTR(mean); EML_CHECK_PTR(inv_std); EML_CHECK_PTR(gamma); EML_CHECK_PTR(beta);
    EML_CHECK_PTR(sums); EML_CHECK_PTR(out);
    if (M <= 0 || C <= 0 || x_pitch < C || g_pitch < C || out_pitch < C) return EML_E_SHAPE;
    const BnArgs a = make_bn(grad, g_pitch, x, x_pitch, pre_a, pre_b, mean, inv_std, gamma, beta, relu, pool, H, W, M, C);
    if (bn_vec_ok(a, out, out_pitch)) {
        const int cq = (C + 3) / 4;
        const int gx = cq <= 32 ? 1 : (C + 127) / 128;
        const int rows4 = (cq <= 32 ? 256 / cq : 8) * 4;
        long gy4 = (M + rows4 * 4 - 1) / (rows4 * 4);
        const long cap = (148L * 16 + gx - 1) / gx;
        if (gy4 > cap) gy4 = cap;
        dim3 grid4(gx, static_cast<unsigned>(gy4 < 1 ? 1 : gy4));
        bn_bwd_apply4_kernel<<<grid4, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, sums, sums_stride > 0 ? sums_stride : C, out, out_pitch,
                                                                                 accumulate, to_stored);
        return eml_launch_status();
    }
    long gy = (M + 8 * 32 - 1) / (8 * 32);
    if (gy > 148 * 8) gy = 148 * 8;
    dim3 grid((C + 31) / 32, static_cast<unsigned>(gy < 1 ? 1 : gy));
    bn_bwd_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        a, sums, sums_stride > 0 ? sums_stride : C, out, out_pitch, accumulate, to_stored);
    return eml_launch_status();
}

bool eml_wgrad1x1_tc_supported(int N, int C, int pool, long M);                                         // wgrad1x1_tc.cu
int eml_wgrad1x1_tc(const float *G, int g_pitch, int N, const float *x, int x_pitch, int C, const float *scale, const float *shift,
                    int relu, float *dW, long M, int precision, cudaStream_t st);

extern "C" int eml_wgrad_1x1(const float *G, int g_pitch, int N, const float *x, int x_pitch, int C, const float *scale,
                             const float *shift, int relu, int pool, int H, int W, float *dW, long M, int precision, void *stream) {
    EML_CHECK_PTR(G); EML_CHECK_PTR(x); EML_CHECK_PTR(dW);
    if (M <= 0 || N <= 0 || C <= 0 || g_pitch < N || x_pitch < C) return EML_E_SHAPE;
    if (precision != EML_PREC_FP32 && eml_wgrad1x1_tc_supported(N, C, pool, M))
        return eml_wgrad1x1_tc(G, g_pitch, N, x, x_pitch, C, scale, shift, relu, dW, M, precision, static_cast<cudaStream_t>(stream));
    long gx = (M + 4095) / 4096;
    if (gx > 148 * 2) gx = 148 * 2;
    dim3 grid(static_cast<unsigned>(gx < 1 ? 1 : gx), (C + WG_TC - 1) / WG_TC, (N + WG_TN - 1) / WG_TN);
    wgrad_1x1_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(G, g_pitch, N, x, x_pitch, C, scale, shift, relu, pool, H, W, dW, M);
    return eml_launch_status();
}

// out[b, y, x, c] = mean over the 2x2 window of relu(scale[c] * in[b, 2y+dy, 2x+dx, c] + shift[c])  -- the transition's pooled activation
// (RegressionNetwork/DenseNet.py:14-21: norm -> relu -> conv -> avg_pool2d, the pooling commuted in front of the 1x1 conv), materialised
// once for the transition's weight gradient so that it can run on the tensor-core wgrad kernel instead of the SIMT one.
__global__ void __launch_bounds__(256) pool_act_kernel(const float *__restrict__ in, int in_pitch, const float *__restrict__ scale,
                                                       const float *__restrict__ shift, int B, int H, int W, int C, float *__restrict__ out,
                                                       int out_pitch) {
    const int Hp = H >> 1, Wp = W >> 1, Q = (C + 3) >> 2;
    const long total = static_cast<long>(B) * Hp * Wp * Q;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(idx % Q);
        const long mp = idx / Q;
        const int xp = static_cast<int>(mp % Wp);
        const long t = mp / Wp;
        const int yp = static_cast<int>(t % Hp);
        const long b = t / Hp;
        const int c = q * 4;
        float sc[4], sh[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) { sc[e] = c + e < C ? scale[c + e] : 0.f; sh[e] = c + e < C ? shift[c + e] : 0.f; }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float *p = in + ((b * H + 2 * yp + dy) * W + 2 * xp + dx) * in_pitch + c;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (c + 3 < C) { const float4 f = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
                else for (int e = 0; e < 4; ++e) if (c + e < C) v[e] = p[e];
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] += fmaxf(fmaf(v[e], sc[e], sh[e]), 0.f);
            }
        *reinterpret_cast<float4 *>(out + mp * out_pitch + c) = make_float4(0.25f * acc[0], 0.25f * acc[1], 0.25f * acc[2], 0.25f * acc[3]);
    }
}

extern "C" int eml_pool_act(const float *in, int in_pitch, const float *scale, const float *shift, int B, int H, int W, int C, float *out,
                            int out_pitch, void *stream) {
    EML_CHECK_PTR(in); EML_CHECK_PTR(scale); EML_CHECK_PTR(shift); EML_CHECK_PTR(out);
    EML_CHECK_ALIGN16(in); EML_CHECK_ALIGN16(out);
    if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || (in_pitch & 3) || (out_pitch & 3) || in_pitch < C ||
        out_pitch < ((C + 3) & ~3))
        return EML_E_SHAPE;
    const long total = static_cast<long>(B) * (H / 2) * (W / 2) * ((C + 3) / 4);
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pool_act_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, in_pitch, scale, shift, B, H, W, C, out, out_pitch);
    return eml_launch_status();
}

bool eml_wgrad3x3_tc_supported(int N, int C);                                                            // wgrad3x3_tc.cu
int eml_wgrad3x3_tc(const float *dY, int dy_pitch, int N, const float *b, int b_pitch, const float *scale, const float *shift,
                    float *dW, int B, int H, int W, int precision, cudaStream_t st);

extern "C" int eml_wgrad_3x3(const float *dY, int dy_pitch, int N, const float *b, int b_pitch, int C, const float *scale,
                             const float *shift, float *dW, int B, int H, int W, int precision, void *stream) {
    EML_CHECK_PTR(dY); EML_CHECK_PTR(b); EML_CHECK_PTR(dW);
    if (B <= 0 || H <= 0 || W <= 0 || N <= 0 || N > 16 || C <= 0 || C > 64 || dy_pitch < N || b_pitch < C) return EML_E_SHAPE;
    if (precision != EML_PREC_FP32 && eml_wgrad3x3_tc_supported(N, C) && (b_pitch & 3) == 0)
        return eml_wgrad3x3_tc(dY, dy_pitch, N, b, b_pitch, scale, shift, dW, B, H, W, precision, static_cast<cudaStream_t>(stream));
    wgrad_3x3_kernel<<<148 * 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(dY, dy_pitch, N, b, b_pitch, C, scale, shift, dW, B, H, W);
    return eml_launch_status();
}

extern "C" int eml_wgrad_stem(const float *dZ, int dz_pitch, int O, const float *x_nchw, float *dW, int B, int H, int W, void *stream) {
    EML_CHECK_PTR(dZ); EML_CHECK_PTR(x_nchw); EML_CHECK_PTR(dW);
    if (B <= 0 || H <= 0 || W <= 0 || O <= 0 || O > 32 || dz_pitch < O) return EML_E_SHAPE;
    wgrad_stem_kernel<<<148 * 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(dZ, dz_pitch, O, x_nchw, dW, B, H, W);
    return eml_launch_status();
}
