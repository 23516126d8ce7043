// One-kernel dense layer for inference-mode BatchNorm (RegressionNetwork/DenseNet.py:26-55), sm_100a.
//
// The reference layer is  x -> norm1 -> relu1 -> conv1 (1x1, C_in -> 48) -> norm2 -> conv2 (3x3, 48 -> 12): there is NO
// nonlinearity between conv1 and conv2 (DenseNet.py:30-43), so with running-statistics BatchNorm everything after relu1 is
// one linear map of a = relu(scale1 * x + shift1):
//     y[r,x,o] = sum_{dy,dx,c} Weff[(dy,dx,o), c] * a[r+dy-1, x+dx-1, c]  +  sum_{(dy,dx) inside the image} beta[(dy,dx), o]
//     Weff[(dy,dx,o), c] = sum_b W2[o,b,dy,dx] * scale2[b] * W1[b,c]          beta[(dy,dx),o] = sum_b W2[o,b,dy,dx] * shift2[b]
// (zero padding applies to the norm2 output, hence the position-dependent bias).  The 48-channel bottleneck -- 384 B per
// pixel of HBM traffic per layer, 37 % of the whole network's bytes -- never exists.
//
// Mapping.  Taps go to the N dimension, not K:  Z[p, (dy,dx,o)] = sum_c a[p,c] * Weff[(dy,dx,o), c]  is a plain 1x1-style GEMM
// (M = 128 pixels of one image row, N = 108 -> 112, K = C_in) whose A operand is read from shared memory ONCE per k-step
// (a 9-tap K loop re-reads it 9x; the shared-memory operand pipe is what bounded the old 3x3 kernel), and
//     y[r,x,o] = sum_{dy,dx} Z[(r+dy-1, x+dx-1), (dy,dx,o)]
// is a stencil over accumulator rows, done by the epilogue:
//   * a CTA walks DOWN a band of R image rows (full width, R+2 rows of Z); per row tile the epilogue thread that owns pixel x
//     adds the tile's dy=2 columns to the partial sum of output row r-1 (-> complete, emitted), its dy=1 columns to row r and
//     stores its dy=0 columns as the start of row r+1.  The two in-flight partial rows live in TENSOR MEMORY next to the
//     accumulators (tcgen05.st / tcgen05.ld, lane-private, no shared-memory traffic);
//   * the completed row U[x, (dx,o)] goes through one shared-memory row buffer, where y[x,o] = U[x-1,0,o] + U[x,1,o] + U[x+1,2,o]
//     + bias is formed and written with coalesced float4 stores into the slab at channel offset C_in.
// Roles (800 threads, 1 CTA/SM, persistent over bands): warps 0-15 producers (NHWC gather, norm1 affine, ReLU folded into
// cvt.rz.relu.bf16x2, bf16 hi/lo split, SWIZZLE_128B K-major ring; register double-buffered loads), warp 16 MMA issuer
// (resident composite weights, 3 tcgen05.mma per k-step in bf16x3), warps 17-24 epilogue (one warpgroup per half row).
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int F_TILE_M = 128;
constexpr int F_CHUNK_K = 64;
constexpr int F_MAX_STAGES = 4;
constexpr int F_PRODUCERS = 512;
constexpr int F_PWARPS = F_PRODUCERS / 32;
constexpr int F_EPI = 256;                              // 2 warpgroups
constexpr int F_THREADS = F_PRODUCERS + 32 + F_EPI;     // 800
constexpr int F_A_TILE = F_TILE_M * F_CHUNK_K * 2;      // 16 KB (one bf16 image)
constexpr int F_G = 12;                                 // growth rate (output channels)
constexpr int F_GRP = 3 * F_G;                          // 36 columns per dy group, ordered (dx, o)
constexpr int F_NPAD = 112;
constexpr int F_ZSTRIDE = 128;                          // TMEM columns between the two Z buffers
constexpr int F_UBASE = 256;                            // TMEM column of the partial-row slots
constexpr int F_USTRIDE = 48;                           // [slot(2)][half(2)] x 48 columns (36 used)
constexpr int F_MAX_C = 320;                            // 5 K-chunks
constexpr int F_WCHUNK = 2 * F_NPAD * 128;              // bytes of one packed weight chunk [hi | lo]
constexpr int F_SROW = F_GRP;                           // floats per pixel in the row buffer

struct FArgs {
    const float *in;
    const float *scale;
    const float *shift;
    const unsigned char *wpack;
    const float *bias9;          // (3 row classes, 3 column classes, 12)
    float *out;
    int B, H, W, R;              // R = output rows per band (H % R == 0)
    int C_in, in_pitch, out_pitch, out_choff;
    int nchunks, stages;
    long nbands;
};

__device__ __forceinline__ void f_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void f_tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void f_tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void f_tmem_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void f_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// relu + truncation to bf16 in one instruction: d = {bf16(max(hi_in,0)), bf16(max(lo_in,0))}, round toward zero, so that for
// v >= 0 the residual v - hi is >= 0 and exactly representable, and for v < 0 both parts are 0 after a second .relu convert.
__device__ __forceinline__ uint32_t cvt_rz_relu_bf16x2(float lo_in, float hi_in) {
    uint32_t d;
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_in), "f"(lo_in));
    return d;
}
__device__ __forceinline__ uint32_t cvt_rn_relu_bf16x2(float lo_in, float hi_in) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_in), "f"(lo_in));
    return d;
}

// norm1 affine + ReLU + bf16 hi/lo split of 4 consecutive channels -> two 8-byte shared-memory stores.
template <bool SPLIT>
__device__ __forceinline__ void f_store_quad(unsigned char *a_hi, unsigned char *a_lo, uint32_t off, float4 v, float4 sc, float4 sh) {
    const float o0 = fmaf(v.x, sc.x, sh.x), o1 = fmaf(v.y, sc.y, sh.y), o2 = fmaf(v.z, sc.z, sh.z), o3 = fmaf(v.w, sc.w, sh.w);
    uint2 hv;
    if (SPLIT) {
        hv.x = cvt_rz_relu_bf16x2(o0, o1);
        hv.y = cvt_rz_relu_bf16x2(o2, o3);
        uint2 lv;
        lv.x = cvt_rn_relu_bf16x2(o0 - __uint_as_float(hv.x << 16), o1 - __uint_as_float(hv.x & 0xffff0000u));
        lv.y = cvt_rn_relu_bf16x2(o2 - __uint_as_float(hv.y << 16), o3 - __uint_as_float(hv.y & 0xffff0000u));
        *reinterpret_cast<uint2 *>(a_lo + off) = lv;
    } else {
        hv.x = cvt_rn_relu_bf16x2(o0, o1);
        hv.y = cvt_rn_relu_bf16x2(o2, o3);
    }
    *reinterpret_cast<uint2 *>(a_hi + off) = hv;
}

// Walks the CTA's bands -> row tiles -> K chunks in the order every role agrees on.
struct BandIter {
    long band;
    long m0;         // first pixel (b*H*W + r*W + x) of the current tile
    int nt, t, c;    // tiles in the band, current tile, current chunk
    bool valid;
};
__device__ __forceinline__ void band_init(BandIter &it, const FArgs &a, long band) {
    it.band = band;
    it.valid = band < a.nbands;
    it.t = 0; it.c = 0; it.nt = 0; it.m0 = 0;
    if (!it.valid) return;
    const int bpi = a.H / a.R;
    const long img = band / bpi;
    const int r0 = static_cast<int>(band % bpi) * a.R;
    const int lo = max(r0 - 1, 0), hi = min(r0 + a.R, a.H - 1);
    it.nt = (hi - lo + 1) * (a.W / F_TILE_M);
    it.m0 = (img * a.H + lo) * a.W;
}
__device__ __forceinline__ void band_next_chunk(BandIter &it, const FArgs &a, int grid) {
    if (++it.c < a.nchunks) return;
    it.c = 0;
    it.m0 += F_TILE_M;
    if (++it.t < it.nt) return;
    band_init(it, a, it.band + grid);
}

template <bool SPLIT>
__global__ void __launch_bounds__(F_THREADS, 1) dense_layer_kernel(const FArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2 * F_MAX_STAGES + 1 + 4];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_scale[F_MAX_C], s_shift[F_MAX_C], s_bias[9 * F_G + 4];

    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * F_A_TILE;
    const int NST = a.stages;
    unsigned char *w_sm = smem + NST * STAGE_BYTES;
    float *s_row = reinterpret_cast<float *>(w_sm + static_cast<size_t>(a.nchunks) * F_WCHUNK);   // [(W + 2) pixels][36]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TPR = a.W / F_TILE_M;                        // tiles per image row (1 or 2)

    const uint32_t bar_full = smem_u32(&s_bar[0]);
    const uint32_t bar_empty = smem_u32(&s_bar[F_MAX_STAGES]);
    const uint32_t bar_w = smem_u32(&s_bar[2 * F_MAX_STAGES]);
    const uint32_t bar_zfull = smem_u32(&s_bar[2 * F_MAX_STAGES + 1]);     // [2]
    const uint32_t bar_zempty = smem_u32(&s_bar[2 * F_MAX_STAGES + 3]);    // [2]

    if (tid == 0) {
        for (int s = 0; s < F_MAX_STAGES; ++s) { mbar_init(bar_full + 8 * s, F_PWARPS); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_w, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_zfull + 8 * i, 1); mbar_init(bar_zempty + 8 * i, 4); }
        fence_mbar_init();
    }
    for (int i = tid; i < F_MAX_C; i += F_THREADS) {
        s_scale[i] = i < a.C_in ? a.scale[i] : 0.f;        // channels past C_in convert to exact zeros (their weights are 0 too)
        s_shift[i] = i < a.C_in ? a.shift[i] : 0.f;
    }
    for (int i = tid; i < 9 * F_G; i += F_THREADS) s_bias[i] = a.bias9[i];
    if (tid < F_SROW) { s_row[tid] = 0.f; s_row[(a.W + 1) * F_SROW + tid] = 0.f; }   // zero pixels left and right of the row
    if (warp == F_PWARPS) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int grid = static_cast<int>(gridDim.x);

    if (warp < F_PWARPS) {
        // =========================================================== PRODUCERS: thread = (pixel row, 4 channels of each k-step)
        const int row = tid >> 2, sub = tid & 3, r7 = row & 7;
        const uint32_t st_base = static_cast<uint32_t>((row >> 3) * 1024 + r7 * 128 + (sub & 1) * 8);
        const int jsub = sub >> 1;
        auto issue = [&](const BandIter &it, float4 (&v)[4]) {
            const int c0 = it.c * F_CHUNK_K + sub * 4;
            const float *p = a.in + (it.m0 + row) * a.in_pitch + c0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + k * 16 < a.C_in) v[k] = __ldg(reinterpret_cast<const float4 *>(p + k * 16));
            }
        };
        auto process = [&](int c, const float4 (&v)[4], int s) {
            unsigned char *a_hi = smem + static_cast<size_t>(s) * STAGE_BYTES;
            unsigned char *a_lo = a_hi + F_A_TILE;
            const int ks = min(4, (a.C_in - c * F_CHUNK_K + 15) >> 4);
            const int q0 = (c * F_CHUNK_K >> 2) + sub;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < ks) {
                    const float4 sc = *reinterpret_cast<const float4 *>(&s_scale[(q0 + k * 4) << 2]);
                    const float4 sh = *reinterpret_cast<const float4 *>(&s_shift[(q0 + k * 4) << 2]);
                    f_store_quad<SPLIT>(a_hi, a_lo, st_base + ((((k << 1) + jsub) ^ r7) << 4), v[k], sc, sh);
                }
            }
        };
        float4 cur[4], nxt[4];
        BandIter it, nx;
        band_init(it, a, blockIdx.x);
        if (it.valid) issue(it, cur);
        uint32_t g = 0;
        while (it.valid) {
            nx = it;
            band_next_chunk(nx, a, grid);
            if (nx.valid) issue(nx, nxt);
            const int s = g % NST;
            const uint32_t ph = (g / NST) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            process(it.c, cur, s);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) f_mbar_arrive(bar_full + 8 * s);
#pragma unroll
            for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
            it = nx; ++g;
        }
    } else if (warp == F_PWARPS) {
        // =========================================================== MMA ISSUER
        const bool leader = elect_one();
        if (leader) {
            const uint32_t bytes = static_cast<uint32_t>(a.nchunks) * F_WCHUNK;
            mbar_expect_tx(bar_w, bytes);
            bulk_g2s(smem_u32(w_sm), a.wpack, bytes, bar_w);
        }
        mbar_wait(bar_w, 0);
        const uint32_t idesc = make_idesc_bf16(F_TILE_M, F_NPAD);
        const uint64_t dA0 = make_sw128_desc(smem_u32(smem));
        const uint64_t dB0 = make_sw128_desc(smem_u32(w_sm));
        const uint32_t stage16 = STAGE_BYTES >> 4, alo16 = F_A_TILE >> 4, wchunk16 = F_WCHUNK >> 4, blo16 = (F_NPAD * 128) >> 4;
        uint32_t g = 0, j = 0;
        for (long band = blockIdx.x; band < a.nbands; band += grid) {
            BandIter it;
            band_init(it, a, band);
            for (int t = 0; t < it.nt; ++t, ++j) {
                const uint32_t zb = j & 1, zph = (j >> 1) & 1;
                mbar_wait(bar_zempty + 8 * zb, zph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + zb * F_ZSTRIDE;
                for (int c = 0; c < a.nchunks; ++c, ++g) {
                    const uint32_t s = g % NST;
                    const uint32_t ph = (g / NST) & 1;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    if (leader) {
                        const int ks = min(4, (a.C_in - c * F_CHUNK_K + 15) >> 4);
                        const uint64_t da_hi = dA0 + static_cast<uint64_t>(s * stage16), da_lo = da_hi + alo16;
                        const uint64_t db_hi = dB0 + static_cast<uint64_t>(static_cast<uint32_t>(c) * wchunk16), db_lo = db_hi + blo16;
                        for (int k = 0; k < ks; ++k) {
                            const uint64_t adv = static_cast<uint64_t>(k * 2);
                            umma_bf16(d_tmem, da_hi + adv, db_hi + adv, idesc, (c | k) != 0 ? 1u : 0u);
                            if (SPLIT) {
                                umma_bf16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
                                umma_bf16(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                            }
                        }
                        umma_commit(bar_empty + 8 * s);
                        if (c == a.nchunks - 1) umma_commit(bar_zfull + 8 * zb);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =========================================================== EPILOGUE: warpgroup wg owns half-row wg; thread = pixel
        const int ew = warp - (F_PWARPS + 1);                 // 0..7
        const int wg = ew >> 2, q = warp & 3;                 // TMEM lane quarter = warp % 4
        const int et = tid - (F_PRODUCERS + 32);              // 0..255
        const int nE = F_TILE_M * TPR;                        // epilogue threads that take part (128 or 256)
        const bool active = wg < TPR;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int px = q * 32 + lane;                         // pixel inside the tile
        if (active) {
            uint32_t j = 0;                                   // tile counter (all tiles of this CTA, both halves)
            for (long band = blockIdx.x; band < a.nbands; band += grid) {
                const int bpi = a.H / a.R;
                const long img = band / bpi;
                const int r0 = static_cast<int>(band % bpi) * a.R;
                const int rlo = max(r0 - 1, 0), rhi = min(r0 + a.R, a.H - 1), rend = r0 + a.R - 1;    // rend = last output row
                for (int rho = rlo; rho <= rhi; ++rho) {
                    // ---- this warpgroup's tile of Z row rho
                    const uint32_t jt = j + static_cast<uint32_t>(wg);
                    j += static_cast<uint32_t>(TPR);
                    const uint32_t zb = jt & 1, zph = (jt >> 1) & 1;
                    const bool emit = rho - 1 >= r0;                          // output row rho-1 completes now
                    const bool upd = rho >= r0 && rho <= rend;                // output row rho receives its dy=1 part
                    const bool upd_add = rho > rlo;                           // ... on top of the dy=0 part stored by row rho-1
                    const bool init_next = rho + 1 <= rend;                   // output row rho+1 starts with this row's dy=0 part
                    const bool flush = rho == a.H - 1 && upd;                 // bottom image row: nothing below completes it
                    const uint32_t zc = lane_addr + zb * F_ZSTRIDE;
                    const uint32_t us0 = lane_addr + F_UBASE + ((((rho + 1) & 1) * 2 + wg) * F_USTRIDE);   // U[rho-1] in, U[rho+1] out
                    const uint32_t us1 = lane_addr + F_UBASE + (((rho & 1) * 2 + wg) * F_USTRIDE);         // U[rho]
                    float *srow = s_row + (wg * F_TILE_M + px + 1) * F_SROW;
                    mbar_wait(bar_zfull + 8 * zb, zph);
                    __syncwarp();
                    tc_fence_after();
#pragma unroll
                    for (int piece = 0; piece < 2; ++piece) {
                        const uint32_t off = piece * 16;
                        float z[16], u[16];
                        if (emit) {
                            tmem_ld16(zc + 2 * F_GRP + off, z);
                            tmem_ld16(us0 + off, u);
#pragma unroll
                            for (int e = 0; e < 16; e += 4)
                                *reinterpret_cast<float4 *>(srow + off + e) = make_float4(z[e] + u[e], z[e + 1] + u[e + 1], z[e + 2] + u[e + 2], z[e + 3] + u[e + 3]);
                        }
                        if (init_next) {
                            tmem_ld16(zc + off, z);
                            f_tmem_st16(us0 + off, z);
                        }
                        if (upd) {
                            tmem_ld16(zc + F_GRP + off, z);
                            if (upd_add) {
                                tmem_ld16(us1 + off, u);
#pragma unroll
                                for (int e = 0; e < 16; ++e) z[e] += u[e];
                            }
                            f_tmem_st16(us1 + off, z);
                        }
                    }
                    {
                        const uint32_t off = 32;
                        float z[4], u[4];
                        if (emit) {
                            f_tmem_ld4(zc + 2 * F_GRP + off, z);
                            f_tmem_ld4(us0 + off, u);
                            *reinterpret_cast<float4 *>(srow + off) = make_float4(z[0] + u[0], z[1] + u[1], z[2] + u[2], z[3] + u[3]);
                        }
                        if (init_next) {
                            f_tmem_ld4(zc + off, z);
                            f_tmem_st4(us0 + off, z);
                        }
                        if (upd) {
                            f_tmem_ld4(zc + F_GRP + off, z);
                            if (upd_add) {
                                f_tmem_ld4(us1 + off, u);
#pragma unroll
                                for (int e = 0; e < 4; ++e) z[e] += u[e];
                            }
                            f_tmem_st4(us1 + off, z);
                        }
                    }
                    f_tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) f_mbar_arrive(bar_zempty + 8 * zb);        // Z buffer drained: the MMA warp may refill it
                    // ---- completed rows: U row -> shared row buffer -> 3-tap horizontal sum + bias -> slab
                    for (int pass = 0; pass < 2; ++pass) {
                        int orow;
                        if (pass == 0) { if (!emit) continue; orow = rho - 1; }
                        else {
                            if (!flush) continue;
                            orow = rho;
                            float z[16];
#pragma unroll
                            for (int piece = 0; piece < 2; ++piece) {
                                tmem_ld16(us1 + piece * 16, z);
#pragma unroll
                                for (int e = 0; e < 16; e += 4)
                                    *reinterpret_cast<float4 *>(srow + piece * 16 + e) = make_float4(z[e], z[e + 1], z[e + 2], z[e + 3]);
                            }
                            float z4[4];
                            f_tmem_ld4(us1 + 32, z4);
                            *reinterpret_cast<float4 *>(srow + 32) = make_float4(z4[0], z4[1], z4[2], z4[3]);
                        }
                        asm volatile("bar.sync 1, %0;" ::"r"(nE) : "memory");     // the whole row is staged
                        const int rc = orow == 0 ? 0 : (orow == a.H - 1 ? 2 : 1);
                        float *obase = a.out + ((img * a.H + orow) * a.W) * a.out_pitch + a.out_choff;
                        for (int f = et; f < a.W * 3; f += nE) {
                            const int x = f / 3, qd = f - x * 3;
                            const float *s0 = s_row + x * F_SROW + qd * 4;         // pixel x-1 (buffer index x), dx = 0
                            const float4 v0 = *reinterpret_cast<const float4 *>(s0);
                            const float4 v1 = *reinterpret_cast<const float4 *>(s0 + F_SROW + F_G);
                            const float4 v2 = *reinterpret_cast<const float4 *>(s0 + 2 * F_SROW + 2 * F_G);
                            const int cc = x == 0 ? 0 : (x == a.W - 1 ? 2 : 1);
                            const float4 bb = *reinterpret_cast<const float4 *>(&s_bias[(rc * 3 + cc) * F_G + qd * 4]);
                            float4 o;
                            o.x = v0.x + v1.x + v2.x + bb.x; o.y = v0.y + v1.y + v2.y + bb.y;
                            o.z = v0.z + v1.z + v2.z + bb.z; o.w = v0.w + v1.w + v2.w + bb.w;
                            *reinterpret_cast<float4 *>(obase + static_cast<long>(x) * a.out_pitch + qd * 4) = o;
                        }
                        asm volatile("bar.sync 1, %0;" ::"r"(nE) : "memory");     // row buffer free again
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == F_PWARPS) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

size_t fused_smem(int nchunks, int W, bool split, int stages) {
    return static_cast<size_t>(stages) * (split ? 2 : 1) * F_A_TILE + static_cast<size_t>(nchunks) * F_WCHUNK +
           static_cast<size_t>(W + 2) * F_SROW * 4 + 1024;
}
int fused_stages(int nchunks, int W, bool split) {
    for (int st = F_MAX_STAGES; st >= 2; --st)
        if (fused_smem(nchunks, W, split, st) <= 227 * 1024) return st;
    return 0;
}

}  // namespace

extern "C" int eml_dense_layer_supported(int H, int W, int C_in, int growth, int precision) {
    if (growth != F_G || (W != 128 && W != 256) || H < 2 || C_in <= 0 || C_in > F_MAX_C || (C_in & 3)) return 0;
    if (precision != EML_PREC_BF16 && precision != EML_PREC_BF16X3) return 0;
    return fused_stages((C_in + F_CHUNK_K - 1) / F_CHUNK_K, W, precision == EML_PREC_BF16X3) >= 2 ? 1 : 0;
}

extern "C" int eml_dense_layer_forward(const eml_dense_layer_params *p, void *stream) {
    EML_CHECK_PTR(p); EML_CHECK_PTR(p->in); EML_CHECK_PTR(p->out); EML_CHECK_PTR(p->scale); EML_CHECK_PTR(p->shift);
    EML_CHECK_PTR(p->wpack); EML_CHECK_PTR(p->bias9);
    EML_CHECK_ALIGN16(p->in); EML_CHECK_ALIGN16(p->out); EML_CHECK_ALIGN16(p->wpack);
    if (p->B <= 0 || !eml_dense_layer_supported(p->H, p->W, p->C_in, p->growth, p->precision)) return EML_E_SHAPE;
    if ((p->in_pitch & 3) || p->in_pitch < p->C_in || (p->out_pitch & 3) || (p->out_choff & 3) || p->out_choff < 0 ||
        p->out_pitch < p->out_choff + F_G)
        return EML_E_ALIGN;
    if (static_cast<long>(p->B) * p->H * p->W >= (1L << 31)) return EML_E_SHAPE;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // rows per band: minimise (waves of bands over the SMs) x (rows of Z per band, R + 2 halo rows)
    int R = p->H;
    long best = -1;
    for (int r = 1; r <= p->H; ++r) {
        if (p->H % r) continue;
        const long nb = static_cast<long>(p->B) * (p->H / r);
        const long cost = ((nb + sms - 1) / sms) * (r + 2);
        if (best < 0 || cost < best || (cost == best && r > R)) { best = cost; R = r; }
    }
    if (const char *env = getenv("EML_DENSE_ROWS")) {     // debug / test switch: force the band height
        const int r = atoi(env);
        if (r > 0 && p->H % r == 0) R = r;
    }
    FArgs a{};
    a.in = p->in; a.scale = p->scale; a.shift = p->shift; a.wpack = static_cast<const unsigned char *>(p->wpack);
    a.bias9 = p->bias9; a.out = p->out;
    a.B = p->B; a.H = p->H; a.W = p->W; a.R = R;
    a.C_in = p->C_in; a.in_pitch = p->in_pitch; a.out_pitch = p->out_pitch; a.out_choff = p->out_choff;
    a.nchunks = (p->C_in + F_CHUNK_K - 1) / F_CHUNK_K;
    const bool split = p->precision == EML_PREC_BF16X3;
    a.stages = fused_stages(a.nchunks, p->W, split);
    a.nbands = static_cast<long>(p->B) * (p->H / R);
    const size_t smem = fused_smem(a.nchunks, p->W, split, a.stages);
    const unsigned grid = static_cast<unsigned>(a.nbands < sms ? a.nbands : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(dense_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        dense_layer_kernel<true><<<grid, F_THREADS, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(dense_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        dense_layer_kernel<false><<<grid, F_THREADS, smem, st>>>(a);
    }
    return eml_launch_status();
}
