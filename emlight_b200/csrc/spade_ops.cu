// GenProjector (SPADE / SphereNet generator) support kernels, sm_100a.  All activations NHWC fp32.
//
// SphereConv2D (GenProjector/models/networks/spherenet/sphere_cnn.py:87-124) = bilinear resampling on a per-pixel
// tangent-plane 3x3 pattern + a stride-3 3x3 convolution, i.e.  y[m, o] = b_o + sum_{tap,c} W[o,c,tap] * S[m,tap,c]
// with S[m,tap,c] = sum_{t<4} w_t(m,tap) * x[src_t(m,tap), c].  The reference materialises S as a 9x larger image with
// grid_sample; here `eml_im2col_lut` builds the (M, 9*Cp) operand of the tcgen05 GEMM (eml_conv_forward, 1x1 mode)
// straight from a per-(H,W,stride) lookup table of 4 taps + weights per (pixel, tap) -- the same kernel serves the
// regular stride-2 3x3 convolutions of the ConvEncoder with a one-tap table.  The producer's per-channel bias and the
// consumer's activation (ReLU of mlp_shared, LeakyReLU(0.2) of the SPADE blocks) are applied to the source values
// before blending, so those intermediate tensors are never written.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : 0.2f * v;
    return v;
}

// A[m, tap*Cp + c] = sum_t w[m,tap,t] * act(x[b, idx[m,tap,t], c] + bias[c]);  one thread = one float4 of channels.
__global__ void __launch_bounds__(256) im2col_lut_kernel(const float *__restrict__ x, int x_pitch, int C, int Cp,
                                                         const int *__restrict__ idx, const float *__restrict__ wgt,
                                                         const float *__restrict__ bias, int act, float *__restrict__ A,
                                                         long Mo_img, long in_img_pixels, long total_quads) {
    const int cq = Cp >> 2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total_quads;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(i % cq);
        const long mt = i / cq;                       // (m, tap)
        const int tap = static_cast<int>(mt % 9);
        const long m = mt / 9;
        const long b = m / Mo_img;
        const long mp = m - b * Mo_img;               // pixel within the output image
        const int c = q * 4;
        const int4 id = *reinterpret_cast<const int4 *>(idx + (mp * 9 + tap) * 4);
        const float4 w = *reinterpret_cast<const float4 *>(wgt + (mp * 9 + tap) * 4);
        float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) {
            bs.x = bias[c];
            if (c + 1 < C) bs.y = bias[c + 1];
            if (c + 2 < C) bs.z = bias[c + 2];
            if (c + 3 < C) bs.w = bias[c + 3];
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *xb = x + b * in_img_pixels * x_pitch + c;
        const int ids[4] = {id.x, id.y, id.z, id.w};
        const float ws[4] = {w.x, w.y, w.z, w.w};
        const bool full = c + 3 < C && (x_pitch & 3) == 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (ids[t] < 0) continue;                 // zero padding (grid_sample padding_mode='zeros' / conv padding)
            const float *p = xb + static_cast<long>(ids[t]) * x_pitch;
            float4 v;
            if (full) {
                v = __ldg(reinterpret_cast<const float4 *>(p));
            } else {
                v.x = p[0];
                v.y = c + 1 < C ? p[1] : 0.f;
                v.z = c + 2 < C ? p[2] : 0.f;
                v.w = c + 3 < C ? p[3] : 0.f;
            }
            acc.x = fmaf(ws[t], apply_act(v.x + bs.x, act), acc.x);
            acc.y = fmaf(ws[t], apply_act(v.y + bs.y, act), acc.y);
            acc.z = fmaf(ws[t], apply_act(v.z + bs.z, act), acc.z);
            acc.w = fmaf(ws[t], apply_act(v.w + bs.w, act), acc.w);
        }
        if (c + 1 >= C) acc.y = 0.f;
        if (c + 2 >= C) acc.z = 0.f;
        if (c + 3 >= C) acc.w = 0.f;
        *reinterpret_cast<float4 *>(A + (m * 9 + tap) * Cp + c) = acc;
    }
}


// Same gather, emitting the GEMM operand already split into bf16 hi / lo matrices of row length Kp (a multiple of 64;
// columns >= 9*Cp are zero) -- the layout gemm_tma.cu's tensor maps describe.  A_lo may be NULL (single-pass bf16).
// A block takes I2_TP consecutive output pixels: their 9 LUT entries go to shared memory once, then the threads walk the
// tile's (pixel, k-quad) items FLATTENED -- consecutive threads = consecutive 8-byte pieces of an A row, whatever C is
// (the 3-channel label map has ONE quad per tap: a warp-per-pixel loop left 31 lanes idle there) -- four items per
// thread in flight (16 independent 16-byte gathers) before any is converted and stored.
constexpr int I2_TP = 32;
__global__ void __launch_bounds__(256) im2col_lut_bf16_kernel(const float *__restrict__ x, int x_pitch, int C, int Cp,
                                                              const int *__restrict__ idx, const float *__restrict__ wgt,
                                                              const float *__restrict__ bias, int act,
                                                              __nv_bfloat16 *__restrict__ A_hi, __nv_bfloat16 *__restrict__ A_lo, int Kp,
                                                              int Mo_img, long in_img_pixels, long M) {
    __shared__ int4 s_idx[I2_TP * 9];
    __shared__ float4 s_w[I2_TP * 9];
    __shared__ long s_xoff[I2_TP];
    const int cq = Cp >> 2, kq = Kp >> 2, kdata = 9 * cq;
    const int cq_shift = (cq & (cq - 1)) == 0 ? __ffs(cq) - 1 : -1;
    const bool vec = (x_pitch & 3) == 0;
    const long ntiles = (M + I2_TP - 1) / I2_TP;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long m0 = tile * I2_TP;
        const int npx = static_cast<int>(min(static_cast<long>(I2_TP), M - m0));
        __syncthreads();                                             // previous tile's readers are done
        for (int e = threadIdx.x; e < npx * 9; e += blockDim.x) {
            const int px = e / 9, tap = e - px * 9;
            const long m = m0 + px;
            const long b = m / Mo_img;
            const long mp = m - b * Mo_img;
            s_idx[e] = __ldg(reinterpret_cast<const int4 *>(idx + (mp * 9 + tap) * 4));
            s_w[e] = __ldg(reinterpret_cast<const float4 *>(wgt + (mp * 9 + tap) * 4));
            if (tap == 0) s_xoff[px] = b * in_img_pixels * x_pitch;
        }
        __syncthreads();
        const int items = npx * kq;
        for (int it0 = threadIdx.x; it0 < items; it0 += 4 * blockDim.x) {
            float4 v[4][4], bs[4], w[4];
            int qk[4], px[4], c[4];
            unsigned live = 0;                                       // bit u*4+t: gather t of item u is in flight
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u * blockDim.x;
                px[u] = -1;
                if (it >= items) continue;
                px[u] = it / kq;
                qk[u] = it - px[u] * kq;
                if (qk[u] >= kdata) continue;                        // K padding: zeros
                const int tap = cq_shift >= 0 ? (qk[u] >> cq_shift) : qk[u] / cq;
                c[u] = (qk[u] - tap * cq) * 4;
                const int4 id = s_idx[px[u] * 9 + tap];
                w[u] = s_w[px[u] * 9 + tap];
                const int ids[4] = {id.x, id.y, id.z, id.w};
                const float *xb = x + s_xoff[px[u]] + c[u];
                const bool full = vec && c[u] + 3 < C;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (ids[t] < 0) continue;                        // zero padding (grid_sample padding_mode='zeros' / conv padding)
                    const float *p = xb + static_cast<long>(ids[t]) * x_pitch;
                    if (full) {
                        v[u][t] = __ldg(reinterpret_cast<const float4 *>(p));
                    } else {
                        v[u][t].x = p[0];
                        v[u][t].y = c[u] + 1 < C ? p[1] : 0.f;
                        v[u][t].z = c[u] + 2 < C ? p[2] : 0.f;
                        v[u][t].w = c[u] + 3 < C ? p[3] : 0.f;
                    }
                    live |= 1u << (u * 4 + t);
                }
                bs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias != nullptr) {
                    bs[u].x = bias[c[u]];
                    if (c[u] + 1 < C) bs[u].y = bias[c[u] + 1];
                    if (c[u] + 2 < C) bs[u].z = bias[c[u] + 2];
                    if (c[u] + 3 < C) bs[u].w = bias[c[u] + 3];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (px[u] < 0) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qk[u] < kdata) {
                    const float ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (!((live >> (u * 4 + t)) & 1u)) continue;
                        acc.x = fmaf(ws[t], apply_act(v[u][t].x + bs[u].x, act), acc.x);
                        acc.y = fmaf(ws[t], apply_act(v[u][t].y + bs[u].y, act), acc.y);
                        acc.z = fmaf(ws[t], apply_act(v[u][t].z + bs[u].z, act), acc.z);
                        acc.w = fmaf(ws[t], apply_act(v[u][t].w + bs[u].w, act), acc.w);
                    }
                    if (c[u] + 1 >= C) acc.y = 0.f;
                    if (c[u] + 2 >= C) acc.z = 0.f;
                    if (c[u] + 3 >= C) acc.w = 0.f;
                }
                __nv_bfloat162 h01 = __floats2bfloat162_rn(acc.x, acc.y), h23 = __floats2bfloat162_rn(acc.z, acc.w);
                uint2 hv;
                hv.x = *reinterpret_cast<uint32_t *>(&h01); hv.y = *reinterpret_cast<uint32_t *>(&h23);
                const long ko = (m0 + px[u]) * Kp + qk[u] * 4;
                *reinterpret_cast<uint2 *>(A_hi + ko) = hv;
                if (A_lo != nullptr) {
                    __nv_bfloat162 l01 = __floats2bfloat162_rn(acc.x - __low2float(h01), acc.y - __high2float(h01));
                    __nv_bfloat162 l23 = __floats2bfloat162_rn(acc.z - __low2float(h23), acc.w - __high2float(h23));
                    uint2 lv;
                    lv.x = *reinterpret_cast<uint32_t *>(&l01); lv.y = *reinterpret_cast<uint32_t *>(&l23);
                    *reinterpret_cast<uint2 *>(A_lo + ko) = lv;
                }
            }
        }
    }
}

// out = ((x - mean_c) * inv_c) * (1 + g + bg_c) + (b + bb_c), optional LeakyReLU(0.2)   (normalization.py:101-115)
// gb holds gamma in channels [0,C) and beta in [C,2C) of each pixel (one GEMM with concatenated weights).
__global__ void __launch_bounds__(256) spade_modulate_kernel(const float *__restrict__ x, int x_pitch,
                                                             const float *__restrict__ mean, const float *__restrict__ inv,
                                                             const float *__restrict__ gb, int gb_pitch,
                                                             const float *__restrict__ bias_g, const float *__restrict__ bias_b,
                                                             float *__restrict__ out, int out_pitch, long M, int C, int lrelu) {
    const long total = M * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long m = i / C;
        const float n = (x[m * x_pitch + c] - mean[c]) * inv[c];
        const float g = gb[m * gb_pitch + c] + bias_g[c];
        const float b = gb[m * gb_pitch + C + c] + bias_b[c];
        float v = fmaf(n, 1.f + g, b);
        if (lrelu) v = v > 0.f ? v : 0.2f * v;
        out[m * out_pitch + c] = v;
    }
}

// Fast path of the same operand: no bias, no activation (eml_bias_act applies them ONCE per source value beforehand instead of once
// per bilinear tap and filter tap, 36x), whole channel quads.  The general kernel above is issue-bound (ncu: 67 % issue slots, L2 at
// 13 %): here a warp owns one output pixel, lanes stride the row's k-quads (no division by the row length), the LUT is staged as
// ready-made float offsets, and an item is 2 LDS + 4 LDG + 16 FMA + the bf16 split.
__global__ void __launch_bounds__(256) im2col_lut_bf16_plain_kernel(const float *__restrict__ x, int x_pitch, int Cp,
                                                                    const int *__restrict__ idx, const float *__restrict__ wgt,
                                                                    __nv_bfloat16 *__restrict__ A_hi, __nv_bfloat16 *__restrict__ A_lo, int Kp,
                                                                    int Mo_img, long in_img_pixels, long M) {
    __shared__ int4 s_off[I2_TP * 9];
    __shared__ float4 s_w[I2_TP * 9];
    __shared__ long s_xoff[I2_TP];
    const int cq = Cp >> 2, kq = Kp >> 2, kdata = 9 * cq;
    const int cq_shift = (cq & (cq - 1)) == 0 ? __ffs(cq) - 1 : -1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long ntiles = (M + I2_TP - 1) / I2_TP;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long m0 = tile * I2_TP;
        const int npx = static_cast<int>(min(static_cast<long>(I2_TP), M - m0));
        __syncthreads();
        for (int e = threadIdx.x; e < npx * 9; e += blockDim.x) {
            const int px = e / 9, tap = e - px * 9;
            const long m = m0 + px;
            const long b = m / Mo_img;
            const long mp = m - b * Mo_img;
            int4 id = __ldg(reinterpret_cast<const int4 *>(idx + (mp * 9 + tap) * 4));
            id.x = id.x < 0 ? -1 : id.x * x_pitch; id.y = id.y < 0 ? -1 : id.y * x_pitch;
            id.z = id.z < 0 ? -1 : id.z * x_pitch; id.w = id.w < 0 ? -1 : id.w * x_pitch;
            s_off[e] = id;
            s_w[e] = __ldg(reinterpret_cast<const float4 *>(wgt + (mp * 9 + tap) * 4));
            if (tap == 0) s_xoff[px] = b * in_img_pixels * x_pitch;
        }
        __syncthreads();
        for (int px = warp; px < npx; px += 8) {
            const float *xb = x + s_xoff[px];
            __nv_bfloat16 *row_hi = A_hi + (m0 + px) * Kp;
            __nv_bfloat16 *row_lo = A_lo ? A_lo + (m0 + px) * Kp : nullptr;
            for (int q0 = lane; q0 < kq; q0 += 128) {
                float4 v[4][4], w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int qk = q0 + 32 * u;
#pragma unroll
                    for (int t = 0; t < 4; ++t) v[u][t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (qk < kdata) {
                        const int tap = cq_shift >= 0 ? (qk >> cq_shift) : qk / cq;
                        const int c = (qk - tap * cq) * 4;
                        const int4 o = s_off[px * 9 + tap];
                        w[u] = s_w[px * 9 + tap];
                        if (o.x >= 0) v[u][0] = __ldg(reinterpret_cast<const float4 *>(xb + o.x + c));
                        if (o.y >= 0) v[u][1] = __ldg(reinterpret_cast<const float4 *>(xb + o.y + c));
                        if (o.z >= 0) v[u][2] = __ldg(reinterpret_cast<const float4 *>(xb + o.z + c));
                        if (o.w >= 0) v[u][3] = __ldg(reinterpret_cast<const float4 *>(xb + o.w + c));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int qk = q0 + 32 * u;
                    if (qk >= kq) continue;
                    const float ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        acc.x = fmaf(ws[t], v[u][t].x, acc.x); acc.y = fmaf(ws[t], v[u][t].y, acc.y);
                        acc.z = fmaf(ws[t], v[u][t].z, acc.z); acc.w = fmaf(ws[t], v[u][t].w, acc.w);
                    }
                    __nv_bfloat162 h01 = __floats2bfloat162_rn(acc.x, acc.y), h23 = __floats2bfloat162_rn(acc.z, acc.w);
                    uint2 hv;
                    hv.x = *reinterpret_cast<uint32_t *>(&h01); hv.y = *reinterpret_cast<uint32_t *>(&h23);
                    *reinterpret_cast<uint2 *>(row_hi + qk * 4) = hv;
                    if (row_lo != nullptr) {
                        __nv_bfloat162 l01 = __floats2bfloat162_rn(acc.x - __low2float(h01), acc.y - __high2float(h01));
                        __nv_bfloat162 l23 = __floats2bfloat162_rn(acc.z - __low2float(h23), acc.w - __high2float(h23));
                        uint2 lv;
                        lv.x = *reinterpret_cast<uint32_t *>(&l01); lv.y = *reinterpret_cast<uint32_t *>(&l23);
                        *reinterpret_cast<uint2 *>(row_lo + qk * 4) = lv;
                    }
                }
            }
        }
    }
}

// The TRANSPOSED operand of the weight-gradient GEMM (rows = (tap, channel), columns = output pixels; eml_im2col_lut_bf16_t in
// gp_bwd.cu holds the reference form, one thread per (pixel, tap, quad) with 2-byte stores).  Tiled: a block takes 64 pixels x one
// filter tap x 64 channels, gathers like the fast path above (16 lanes = 16 consecutive channel quads of one source pixel: 256-byte
// reads), splits into bf16 hi / lo, transposes through shared memory and writes 64 rows of 128 contiguous bytes.  No bias, no
// activation (eml_bias_act first), whole quads, Mp % 8 == 0.
constexpr int IT_PX = 64, IT_CH = 64, IT_LD = IT_PX + 2;
__global__ void __launch_bounds__(256) im2col_lut_bf16_t_tiled_kernel(const float *__restrict__ x, int x_pitch, int Cp,
                                                                      const int *__restrict__ idx, const float *__restrict__ wgt,
                                                                      unsigned short *__restrict__ hi, unsigned short *__restrict__ lo,
                                                                      long Mp, long M, long out_pixels, long in_pixels, int cgroups) {
    __shared__ __align__(16) unsigned short s_hi[IT_CH][IT_LD];
    __shared__ __align__(16) unsigned short s_lo[IT_CH][IT_LD];
    long blk = blockIdx.x;
    const int cg = static_cast<int>(blk % cgroups); blk /= cgroups;
    const int tap = static_cast<int>(blk % 9);
    const long m0 = (blk / 9) * IT_PX;
    const int c0 = cg * IT_CH;
    const int nch = min(IT_CH, Cp - c0);
    const int cql = threadIdx.x & 15, pxl0 = threadIdx.x >> 4;
    float4 v[4][4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long m = m0 + pxl0 + 16 * i;
#pragma unroll
        for (int t = 0; t < 4; ++t) v[i][t] = make_float4(0.f, 0.f, 0.f, 0.f);
        w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && cql * 4 < nch) {
            const long b = m / out_pixels;
            const long p = m - b * out_pixels;
            const int4 id = __ldg(reinterpret_cast<const int4 *>(idx + (p * 9 + tap) * 4));
            w[i] = __ldg(reinterpret_cast<const float4 *>(wgt + (p * 9 + tap) * 4));
            const float *xb = x + b * in_pixels * x_pitch + c0 + cql * 4;
            if (id.x >= 0) v[i][0] = __ldg(reinterpret_cast<const float4 *>(xb + static_cast<long>(id.x) * x_pitch));
            if (id.y >= 0) v[i][1] = __ldg(reinterpret_cast<const float4 *>(xb + static_cast<long>(id.y) * x_pitch));
            if (id.z >= 0) v[i][2] = __ldg(reinterpret_cast<const float4 *>(xb + static_cast<long>(id.z) * x_pitch));
            if (id.w >= 0) v[i][3] = __ldg(reinterpret_cast<const float4 *>(xb + static_cast<long>(id.w) * x_pitch));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int px = pxl0 + 16 * i;
        const float ws[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            acc[0] = fmaf(ws[t], v[i][t].x, acc[0]); acc[1] = fmaf(ws[t], v[i][t].y, acc[1]);
            acc[2] = fmaf(ws[t], v[i][t].z, acc[2]); acc[3] = fmaf(ws[t], v[i][t].w, acc[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16 h = __float2bfloat16_rn(acc[j]);
            s_hi[cql * 4 + j][px] = __bfloat16_as_ushort(h);
            s_lo[cql * 4 + j][px] = __bfloat16_as_ushort(__float2bfloat16_rn(acc[j] - __bfloat162float(h)));
        }
    }
    __syncthreads();
    const int r = threadIdx.x >> 2, seg = threadIdx.x & 3;
    if (r >= nch) return;
    const long mcol = m0 + seg * 16;
    if (mcol >= M) return;
    const long o = (static_cast<long>(tap) * Cp + c0 + r) * Mp + mcol;
    const uint32_t *rh = reinterpret_cast<const uint32_t *>(&s_hi[r][seg * 16]);
    const uint32_t *rl = reinterpret_cast<const uint32_t *>(&s_lo[r][seg * 16]);
    if (mcol + 16 <= M) {
        *reinterpret_cast<uint4 *>(hi + o) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
        *reinterpret_cast<uint4 *>(hi + o + 8) = make_uint4(rh[4], rh[5], rh[6], rh[7]);
        if (lo != nullptr) {
            *reinterpret_cast<uint4 *>(lo + o) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
            *reinterpret_cast<uint4 *>(lo + o + 8) = make_uint4(rl[4], rl[5], rl[6], rl[7]);
        }
    } else {
        for (int k = 0; mcol + k < M; ++k) {
            hi[o + k] = s_hi[r][seg * 16 + k];
            if (lo != nullptr) lo[o + k] = s_lo[r][seg * 16 + k];
        }
    }
}

// out = act(x + bias[c])  (act: 0 none, 1 ReLU, 2 LeakyReLU(0.2)): discriminator model0 (SphereConv + LeakyReLU, discriminator.py:91-92),
// VGG conv + ReLU, and the input transform of a SphereConv applied once per value (see the fast path above).  `out` may be `x`.
__global__ void __launch_bounds__(256) bias_act_kernel(const float *x, int x_pitch, const float *__restrict__ bias, int act,
                                                       float *out, int out_pitch, long M, int C) {
    if (((x_pitch | out_pitch | C) & 3) == 0) {
        const int cq = C >> 2;
        const long total = M * cq;
        for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
            const int c = static_cast<int>(i % cq) * 4;
            const long m = i / cq;
            float4 v = *reinterpret_cast<const float4 *>(x + m * x_pitch + c);
            if (bias != nullptr) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
            *reinterpret_cast<float4 *>(out + m * out_pitch + c) = v;
        }
        return;
    }
    const long total = M * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long m = i / C;
        out[m * out_pitch + c] = apply_act(x[m * x_pitch + c] + (bias ? bias[c] : 0.f), act);
    }
}

// out = a + bias_a (+ r + bias_r)     (SPADEResnetBlock output x_s + dx, architecture.py:51-58)
__global__ void __launch_bounds__(256) bias_residual_kernel(const float *__restrict__ a, int a_pitch, const float *__restrict__ bias_a,
                                                            const float *__restrict__ r, int r_pitch, const float *__restrict__ bias_r,
                                                            float *__restrict__ out, int out_pitch, long M, int C) {
    const long total = M * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long m = i / C;
        float v = a[m * a_pitch + c] + (bias_a ? bias_a[c] : 0.f);
        if (r != nullptr) v += r[m * r_pitch + c] + (bias_r ? bias_r[c] : 0.f);
        out[m * out_pitch + c] = v;
    }
}

// nearest-neighbour resize NHWC -> NHWC (F.interpolate(mode='nearest'): src = floor(dst * in / out)); also the x2 upsample
__global__ void __launch_bounds__(256) resize_nearest_kernel(const float *__restrict__ x, int x_pitch, int Hi, int Wi,
                                                             float *__restrict__ out, int out_pitch, int Ho, int Wo, int C, long B,
                                                             int src_nchw) {
    const long total = B * Ho * Wo * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long t = i / C;
        const int xo = static_cast<int>(t % Wo); t /= Wo;
        const int yo = static_cast<int>(t % Ho);
        const long b = t / Ho;
        // PyTorch: src = min(floor(dst * scale), in - 1) with scale = in / out evaluated in float
        const int ys = min(static_cast<int>(floorf(yo * (static_cast<float>(Hi) / Ho))), Hi - 1);
        const int xs = min(static_cast<int>(floorf(xo * (static_cast<float>(Wi) / Wo))), Wi - 1);
        const float v = src_nchw ? x[((b * C + c) * Hi + ys) * static_cast<long>(Wi) + xs]
                                 : x[((b * Hi + ys) * static_cast<long>(Wi) + xs) * x_pitch + c];
        out[((b * Ho + yo) * static_cast<long>(Wo) + xo) * out_pitch + c] = v;
    }
}

// bilinear resize NCHW -> NHWC, align_corners=False (F.interpolate(mode='bilinear'), generator.py:116)
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float *__restrict__ x, int Hi, int Wi, float *__restrict__ out,
                                                              int out_pitch, int Ho, int Wo, int C, long B) {
    const long total = B * Ho * Wo * C;
    const float sy = static_cast<float>(Hi) / Ho, sx = static_cast<float>(Wi) / Wo;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long t = i / C;
        const int xo = static_cast<int>(t % Wo); t /= Wo;
        const int yo = static_cast<int>(t % Ho);
        const long b = t / Ho;
        const float fy = fmaxf((yo + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((xo + 0.5f) * sx - 0.5f, 0.f);
        const int y0 = min(static_cast<int>(fy), Hi - 1), x0 = min(static_cast<int>(fx), Wi - 1);
        const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
        const float ly = fy - y0, lx = fx - x0;
        const float *p = x + (b * C + c) * static_cast<long>(Hi) * Wi;
        const float v = (1.f - ly) * ((1.f - lx) * p[y0 * static_cast<long>(Wi) + x0] + lx * p[y0 * static_cast<long>(Wi) + x1]) +
                        ly * ((1.f - lx) * p[y1 * static_cast<long>(Wi) + x0] + lx * p[y1 * static_cast<long>(Wi) + x1]);
        out[((b * Ho + yo) * static_cast<long>(Wo) + xo) * out_pitch + c] = v;
    }
}

// InstanceNorm2d(affine=False, eps) + optional LeakyReLU(0.2): one block per (b, 32-channel group); two passes over H*W.
__global__ void __launch_bounds__(256) instance_norm_kernel(const float *__restrict__ x, int x_pitch, float *__restrict__ out,
                                                            int out_pitch, int HW, int C, float eps, int lrelu) {
    __shared__ double s1[8][32], s2[8][32];
    __shared__ float s_mean[32], s_inv[32];
    const long b = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int row = threadIdx.x >> 5;                              // 8 pixel lanes
    const float *xb = x + b * HW * static_cast<long>(x_pitch);
    double a1 = 0.0, a2 = 0.0;
    if (c < C)
        for (int p = row; p < HW; p += 8) {
            const double v = xb[static_cast<long>(p) * x_pitch + c];
            a1 += v; a2 += v * v;
        }
    s1[row][threadIdx.x & 31] = a1; s2[row][threadIdx.x & 31] = a2;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t1 = 0.0, t2 = 0.0;
        for (int r = 0; r < 8; ++r) { t1 += s1[r][threadIdx.x]; t2 += s2[r][threadIdx.x]; }
        const double mean = t1 / HW;
        double var = t2 / HW - mean * mean;                       // biased variance, like F.instance_norm
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = static_cast<float>(mean);
        s_inv[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
    if (c < C) {
        const float mean = s_mean[threadIdx.x & 31], inv = s_inv[threadIdx.x & 31];
        float *ob = out + b * HW * static_cast<long>(out_pitch);
        for (int p = row; p < HW; p += 8) {
            float v = (xb[static_cast<long>(p) * x_pitch + c] - mean) * inv;
            if (lrelu) v = v > 0.f ? v : 0.2f * v;
            ob[static_cast<long>(p) * out_pitch + c] = v;
        }
    }
}

// out_nchw[b,c,p] = (tanh(x[b,p,c] + bias[c]) + 1) * scale      (generator.py:85-86)
__global__ void __launch_bounds__(256) tanh_out_kernel(const float *__restrict__ x, int x_pitch, const float *__restrict__ bias,
                                                       float *__restrict__ out, long B, int HW, int C, float scale) {
    const long total = B * C * HW;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % HW);
        long t = i / HW;
        const int c = static_cast<int>(t % C);
        const long b = t / C;
        out[i] = (tanhf(x[(b * HW + p) * x_pitch + c] + (bias ? bias[c] : 0.f)) + 1.f) * scale;
    }
}


// mode 0: avg_pool2d(k=3, s=2, p=1, count_include_pad=False) (discriminator.py:48-51); mode 1: max_pool2d(k=2, s=2) (VGG19)
__global__ void __launch_bounds__(256) pool_kernel(const float *__restrict__ x, int x_pitch, int Hi, int Wi, float *__restrict__ out,
                                                   int out_pitch, int Ho, int Wo, int C, long B, int mode) {
    const long total = B * Ho * Wo * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long t = i / C;
        const int xo = static_cast<int>(t % Wo); t /= Wo;
        const int yo = static_cast<int>(t % Ho);
        const long b = t / Ho;
        float r;
        if (mode == 0) {
            float s = 0.f; int n = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int y = 2 * yo + dy, xx = 2 * xo + dx;
                    if (y >= 0 && y < Hi && xx >= 0 && xx < Wi) { s += x[((b * Hi + y) * static_cast<long>(Wi) + xx) * x_pitch + c]; ++n; }
                }
            r = s / static_cast<float>(n);
        } else {
            r = -INFINITY;
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) r = fmaxf(r, x[((b * Hi + 2 * yo + dy) * static_cast<long>(Wi) + 2 * xo + dx) * x_pitch + c]);
        }
        out[((b * Ho + yo) * static_cast<long>(Wo) + xo) * out_pitch + c] = r;
    }
}

// Scalar loss reductions over NHWC tensors a (and b), accumulated in double (GenProjector/models/networks/loss.py:57-82,109-114;
// pix2pix_model.py:101-122).  acc += sum over the M*C elements (mode 5: over the M pixels) of
//   0: a                       (generator hinge: -mean(D(fake)))        1: min(a - 1, 0)   (D hinge, real)
//   2: min(-a - 1, 0)          (D hinge, fake)                          3: |a - b|         (L1: VGG / feature matching)
//   4: |a - b| * (m + (1-m)*50), m = mask[pixel]   (mask-weighted feature matching; the weight is >= 0 so it factors out of |.|)
//   5: 1 - <a,b> / max(|a||b|, eps)                (cosine distance over the channel dimension)
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float *__restrict__ a, int a_pitch, const float *__restrict__ b, int b_pitch,
                                                          const float *__restrict__ mask, long M, int C, int mode, double *acc) {
    __shared__ double s_part[8];
    double part = 0.0;
    if (mode == 5) {
        for (long m = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; m < M; m += static_cast<long>(gridDim.x) * blockDim.x) {
            float ab = 0.f, aa = 0.f, bb = 0.f;
            for (int c = 0; c < C; ++c) { const float u = a[m * a_pitch + c], v = b[m * b_pitch + c]; ab = fmaf(u, v, ab); aa = fmaf(u, u, aa); bb = fmaf(v, v, bb); }
            part += 1.0 - static_cast<double>(ab / fmaxf(sqrtf(aa) * sqrtf(bb), 1e-20f));
        }
    } else {
        const long total = M * C;
        for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
            const int c = static_cast<int>(i % C);
            const long m = i / C;
            const float u = a[m * a_pitch + c];
            float v;
            if (mode == 0) v = u;
            else if (mode == 1) v = fminf(u - 1.f, 0.f);
            else if (mode == 2) v = fminf(-u - 1.f, 0.f);
            else {
                v = fabsf(u - b[m * b_pitch + c]);
                if (mode == 4) { const float mk = mask[m]; v *= mk + (1.f - mk) * 50.f; }
            }
            part += v;
        }
    }
    part = warp_sum_d(part);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        atomicAdd(acc, t);
    }
}

inline unsigned grid_for(long total, int per_block = 256) {
    long b = (total + per_block - 1) / per_block;
    const long cap = 148L * 16;
    return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------ spectral norm
// torch.nn.utils.spectral_norm on a (O, K) weight matrix (architecture.py:37-40, normalization.py:29): one power iteration
//     v <- normalize(W^T u),  u <- normalize(W v),  sigma = u . (W v)
// as four small deterministic kernels behind ONE C-ABI call (the torch formulation is ~12 tensor ops per wrapped convolution and
// forward: ~140 convolutions per G+D iteration).  (1) t = W^T u: one thread per column, rows walked 8 at a time (coalesced over the
// columns); (2) one block: v = t / max(|t|, eps); (3) s = W v: one warp per row; (4) one block: u = s / max(|s|, eps), sigma = u . s.
__global__ void __launch_bounds__(256) sn_wtu_kernel(const float *__restrict__ w, const float *__restrict__ u, float *__restrict__ t, int O, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int o = 0;
    for (; o + 8 <= O; o += 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(__ldg(w + static_cast<long>(o + i) * K + k), __ldg(u + o + i), acc[i]);
    }
    for (; o < O; ++o) acc[0] = fmaf(__ldg(w + static_cast<long>(o) * K + k), __ldg(u + o), acc[0]);
    t[k] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}
__device__ __forceinline__ float sn_block_sum(float v, float *red) {           // 1024 threads, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = threadIdx.x < 32 ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (threadIdx.x == 0) red[0] = r;
    }
    __syncthreads();
    r = red[0];
    __syncthreads();
    return r;
}
// out = in / max(|in|, eps); if sigma != NULL also *sigma = out . in
__global__ void __launch_bounds__(1024) sn_normalize_kernel(const float *__restrict__ in, float *__restrict__ out, int n, float eps, float *sigma) {
    __shared__ float red[32];
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) ss = fmaf(in[i], in[i], ss);
    const float nrm = sqrtf(sn_block_sum(ss, red));
    const float inv = 1.f / fmaxf(nrm, eps);
    float dot = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) { const float o = in[i] * inv; out[i] = o; dot = fmaf(o, in[i], dot); }
    if (sigma != nullptr) {
        dot = sn_block_sum(dot, red);
        if (threadIdx.x == 0) *sigma = dot;
    }
}
__global__ void __launch_bounds__(256) sn_wv_kernel(const float *__restrict__ w, const float *__restrict__ v, float *__restrict__ s, int O, int K) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= O) return;
    const float *wr = w + static_cast<long>(o) * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(wr + k), __ldg(v + k), acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) s[o] = acc;
}
// eval mode: sigma = u . s with the stored u (no update)
__global__ void __launch_bounds__(1024) sn_dot_kernel(const float *__restrict__ a, const float *__restrict__ b, int n, float *out) {
    __shared__ float red[32];
    float d = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) d = fmaf(a[i], b[i], d);
    d = sn_block_sum(d, red);
    if (threadIdx.x == 0) *out = d;
}

}  // namespace

extern "C" int eml_im2col_lut(const float *x, int x_pitch, int C, int Cp, const int *lut_idx, const float *lut_w,
                              const float *bias, int act, float *A, int B, long out_pixels, long in_pixels, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(lut_idx); EML_CHECK_PTR(lut_w); EML_CHECK_PTR(A);
    EML_CHECK_ALIGN16(lut_idx); EML_CHECK_ALIGN16(lut_w); EML_CHECK_ALIGN16(A);
    if (B <= 0 || C <= 0 || Cp < C || (Cp & 3) || x_pitch < C || out_pixels <= 0 || in_pixels <= 0) return EML_E_SHAPE;
    if (act < 0 || act > 2) return EML_E_ARG;
    if ((x_pitch & 3) == 0) EML_CHECK_ALIGN16(x);
    const long total = static_cast<long>(B) * out_pixels * 9 * (Cp / 4);
    im2col_lut_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, C, Cp, lut_idx, lut_w, bias, act, A,
                                                                                   out_pixels, in_pixels, total);
    return eml_launch_status();
}


extern "C" int eml_im2col_lut_bf16(const float *x, int x_pitch, int C, int Cp, const int *lut_idx, const float *lut_w,
                                   const float *bias, int act, void *A_hi, void *A_lo, int Kp, int B, long out_pixels,
                                   long in_pixels, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(lut_idx); EML_CHECK_PTR(lut_w); EML_CHECK_PTR(A_hi);
    EML_CHECK_ALIGN16(lut_idx); EML_CHECK_ALIGN16(lut_w); EML_CHECK_ALIGN16(A_hi);
    if (A_lo) EML_CHECK_ALIGN16(A_lo);
    if (B <= 0 || C <= 0 || Cp < C || (Cp & 3) || x_pitch < C || out_pixels <= 0 || in_pixels <= 0) return EML_E_SHAPE;
    if (Kp < 9 * Cp || (Kp & 63)) return EML_E_SHAPE;
    if (act < 0 || act > 2) return EML_E_ARG;
    if ((x_pitch & 3) == 0) EML_CHECK_ALIGN16(x);
    if (out_pixels >= (1L << 31)) return EML_E_SHAPE;
    const long M = static_cast<long>(B) * out_pixels;
    if (bias == nullptr && act == 0 && C == Cp && (x_pitch & 3) == 0 && in_pixels * x_pitch < (1L << 31) && !eml_env_flag("EML_IM2COL_GENERAL")) {
        im2col_lut_bf16_plain_kernel<<<grid_for(M, I2_TP), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            x, x_pitch, Cp, lut_idx, lut_w, static_cast<__nv_bfloat16 *>(A_hi), static_cast<__nv_bfloat16 *>(A_lo), Kp,
            static_cast<int>(out_pixels), in_pixels, M);
        return eml_launch_status();
    }
    im2col_lut_bf16_kernel<<<grid_for(M, I2_TP), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, x_pitch, C, Cp, lut_idx, lut_w, bias, act, static_cast<__nv_bfloat16 *>(A_hi), static_cast<__nv_bfloat16 *>(A_lo), Kp,
        static_cast<int>(out_pixels), in_pixels, M);
    return eml_launch_status();
}

extern "C" int eml_spade_modulate(const float *x, int x_pitch, const float *mean, const float *inv_std, const float *gamma_beta,
                                  int gb_pitch, const float *bias_gamma, const float *bias_beta, float *out, int out_pitch,
                                  long M, int C, int leaky_relu, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(mean); EML_CHECK_PTR(inv_std); EML_CHECK_PTR(gamma_beta); EML_CHECK_PTR(bias_gamma);
    EML_CHECK_PTR(bias_beta); EML_CHECK_PTR(out);
    if (M <= 0 || C <= 0 || x_pitch < C || gb_pitch < 2 * C || out_pitch < C) return EML_E_SHAPE;
    spade_modulate_kernel<<<grid_for(M * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, mean, inv_std, gamma_beta, gb_pitch,
                                                                                       bias_gamma, bias_beta, out, out_pitch, M, C, leaky_relu);
    return eml_launch_status();
}

extern "C" int eml_spectral_norm(const float *w, int O, int K, float *u, float *v, int training, float eps, float *scratch, float *sigma,
                                 void *stream) {
    EML_CHECK_PTR(w); EML_CHECK_PTR(u); EML_CHECK_PTR(v); EML_CHECK_PTR(scratch); EML_CHECK_PTR(sigma);
    if (O <= 0 || K <= 0) return EML_E_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *t = scratch, *sv = scratch + K;                                     // K + O floats
    if (training) {
        sn_wtu_kernel<<<(K + 255) / 256, 256, 0, st>>>(w, u, t, O, K);
        sn_normalize_kernel<<<1, 1024, 0, st>>>(t, v, K, eps, nullptr);
        sn_wv_kernel<<<(O + 7) / 8, 256, 0, st>>>(w, v, sv, O, K);
        sn_normalize_kernel<<<1, 1024, 0, st>>>(sv, u, O, eps, sigma);
    } else {
        sn_wv_kernel<<<(O + 7) / 8, 256, 0, st>>>(w, v, sv, O, K);
        sn_dot_kernel<<<1, 1024, 0, st>>>(u, sv, O, sigma);
    }
    return eml_launch_status();
}

// called by eml_im2col_lut_bf16_t (gp_bwd.cu) when the operand qualifies for the tiled kernel
bool eml_im2col_t_tiled_ok(int x_pitch, int C, int Cp, const void *bias, int act, const void *hi, const void *lo, long Mp) {
    return bias == nullptr && act == 0 && C == Cp && (x_pitch & 3) == 0 && (Mp & 7) == 0 && (reinterpret_cast<uintptr_t>(hi) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(lo) & 15) == 0 && !eml_env_flag("EML_IM2COL_GENERAL");
}
int eml_im2col_t_tiled(const float *x, int x_pitch, int Cp, const int *lut_idx, const float *lut_w, void *hi, void *lo, long Mp, long M,
                       long out_pixels, long in_pixels, cudaStream_t st) {
    const int cgroups = (Cp + IT_CH - 1) / IT_CH;
    const long blocks = ((M + IT_PX - 1) / IT_PX) * 9 * cgroups;
    if (blocks > 0x7fffffffL) return EML_E_SHAPE;
    im2col_lut_bf16_t_tiled_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, x_pitch, Cp, lut_idx, lut_w, static_cast<unsigned short *>(hi),
                                                                               static_cast<unsigned short *>(lo), Mp, M, out_pixels, in_pixels, cgroups);
    return eml_launch_status();
}

extern "C" int eml_bias_act(const float *x, int x_pitch, const float *bias, int act, float *out, int out_pitch, long M, int C, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out);
    if (M <= 0 || C <= 0 || x_pitch < C || out_pitch < C) return EML_E_SHAPE;
    if (act < 0 || act > 2) return EML_E_ARG;
    if (((x_pitch | out_pitch | C) & 3) == 0) { EML_CHECK_ALIGN16(x); EML_CHECK_ALIGN16(out); if (bias) EML_CHECK_ALIGN16(bias); }
    bias_act_kernel<<<grid_for(M * ((C + 3) / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, bias, act, out, out_pitch, M, C);
    return eml_launch_status();
}

extern "C" int eml_bias_residual(const float *a, int a_pitch, const float *bias_a, const float *r, int r_pitch, const float *bias_r,
                                 float *out, int out_pitch, long M, int C, void *stream) {
    EML_CHECK_PTR(a); EML_CHECK_PTR(out);
    if (M <= 0 || C <= 0 || a_pitch < C || out_pitch < C || (r != nullptr && r_pitch < C)) return EML_E_SHAPE;
    bias_residual_kernel<<<grid_for(M * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, a_pitch, bias_a, r, r_pitch, bias_r, out,
                                                                                      out_pitch, M, C);
    return eml_launch_status();
}

extern "C" int eml_resize_nearest(const float *x, int x_pitch, int Hi, int Wi, float *out, int out_pitch, int Ho, int Wo, int C, int B,
                                  int src_is_nchw, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out);
    if (B <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || out_pitch < C || (!src_is_nchw && x_pitch < C)) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * Ho * Wo * C;
    resize_nearest_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, Hi, Wi, out, out_pitch, Ho, Wo, C, B,
                                                                                       src_is_nchw);
    return eml_launch_status();
}

extern "C" int eml_resize_bilinear_nchw(const float *x, int Hi, int Wi, float *out, int out_pitch, int Ho, int Wo, int C, int B,
                                        void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out);
    if (B <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || out_pitch < C) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * Ho * Wo * C;
    resize_bilinear_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, Hi, Wi, out, out_pitch, Ho, Wo, C, B);
    return eml_launch_status();
}

extern "C" int eml_instance_norm(const float *x, int x_pitch, float *out, int out_pitch, int B, int HW, int C, float eps,
                                 int leaky_relu, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out);
    if (B <= 0 || C <= 0 || HW <= 0 || x_pitch < C || out_pitch < C) return EML_E_SHAPE;
    dim3 grid((C + 31) / 32, B);
    instance_norm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, out, out_pitch, HW, C, eps, leaky_relu);
    return eml_launch_status();
}

extern "C" int eml_tanh_to_nchw(const float *x, int x_pitch, const float *bias, float *out, int B, int HW, int C, float scale,
                                void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out);
    if (B <= 0 || C <= 0 || HW <= 0 || x_pitch < C) return EML_E_SHAPE;
    const long total = static_cast<long>(B) * C * HW;
    tanh_out_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, bias, out, B, HW, C, scale);
    return eml_launch_status();
}

extern "C" int eml_pool2d(const float *x, int x_pitch, int Hi, int Wi, float *out, int out_pitch, int C, int B, int mode, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(out);
    if (B <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || x_pitch < C || out_pitch < C) return EML_E_SHAPE;
    if (mode != 0 && mode != 1) return EML_E_ARG;
    if (mode == 1 && ((Hi | Wi) & 1)) return EML_E_SHAPE;
    const int Ho = mode == 0 ? (Hi + 1) / 2 : Hi / 2, Wo = mode == 0 ? (Wi + 1) / 2 : Wi / 2;
    const long total = static_cast<long>(B) * Ho * Wo * C;
    pool_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, Hi, Wi, out, out_pitch, Ho, Wo, C, B, mode);
    return eml_launch_status();
}

extern "C" int eml_loss_reduce(const float *a, int a_pitch, const float *b, int b_pitch, const float *mask, long M, int C, int mode,
                               double *acc, void *stream) {
    EML_CHECK_PTR(a); EML_CHECK_PTR(acc);
    if (M <= 0 || C <= 0 || a_pitch < C) return EML_E_SHAPE;
    if (mode < 0 || mode > 5) return EML_E_ARG;
    if (mode >= 3) { EML_CHECK_PTR(b); if (b_pitch < C) return EML_E_SHAPE; }
    if (mode == 4) EML_CHECK_PTR(mask);
    loss_reduce_kernel<<<grid_for(mode == 5 ? M : M * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, a_pitch, b, b_pitch, mask, M, C, mode, acc);
    return eml_launch_status();
}

// Per-channel sum / sum of squares of an NHWC tensor (SPADE's parameter-free BatchNorm in training mode, normalization.py:80:
// batch statistics over (B, H, W)).  blockDim = (channel lanes, row lanes); float partials over <= 64 rows, double beyond.
__global__ void __launch_bounds__(256) channel_stats_kernel(const float *__restrict__ x, int x_pitch, long M, int C, double *__restrict__ sums) {
    __shared__ double s_acc[2][256];
    const int cl = threadIdx.x, rl = threadIdx.y, nrl = blockDim.y;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + cl;
        double a1 = 0.0, a2 = 0.0;
        if (c < C) {
            for (long m0 = static_cast<long>(blockIdx.x) * nrl * 64; m0 < M; m0 += static_cast<long>(gridDim.x) * nrl * 64) {
                float p1 = 0.f, p2 = 0.f;
                for (int i = 0; i < 64; ++i) {
                    const long m = m0 + static_cast<long>(i) * nrl + rl;
                    if (m < M) { const float v = x[m * x_pitch + c]; p1 += v; p2 = fmaf(v, v, p2); }
                }
                a1 += p1; a2 += p2;
            }
        }
        const int t = rl * blockDim.x + cl;
        s_acc[0][t] = a1; s_acc[1][t] = a2;
        __syncthreads();
        if (rl == 0 && c < C) {
            for (int r = 1; r < nrl; ++r) { a1 += s_acc[0][r * blockDim.x + cl]; a2 += s_acc[1][r * blockDim.x + cl]; }
            atomicAdd(sums + c, a1);
            atomicAdd(sums + C + c, a2);
        }
        __syncthreads();
    }
}

extern "C" int eml_channel_stats(const float *x, int x_pitch, long M, int C, double *sums, void *stream) {
    EML_CHECK_PTR(x); EML_CHECK_PTR(sums);
    if (M <= 0 || C <= 0 || x_pitch < C) return EML_E_SHAPE;
    int bx = 32;
    while (bx < C && bx < 256) bx <<= 1;
    const dim3 block(bx, 256 / bx);
    const long chunks = (M + block.y * 64 - 1) / (block.y * 64);
    const unsigned grid = static_cast<unsigned>(chunks < 148 * 4 ? chunks : 148 * 4);
    channel_stats_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, x_pitch, M, C, sums);
    return eml_launch_status();
}
