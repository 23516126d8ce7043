// 3x3 (pad 1) convolution of the 48-channel bottleneck as a tcgen05 implicit GEMM WITHOUT im2col (sm_100a).
//
// Replaces conv2 of RegressionNetwork/DenseNet.py:41-43 (norm2 -> conv2, no ReLU between them).
// One CTA = 128 consecutive output pixels of one image row.  The three input rows it needs (130 pixels each,
// zero border included) are read ONCE with linear float4 loads, pushed through the fused BatchNorm affine
// (+ optional ReLU), split into bf16 hi/lo and stored in a planar K-major NO-SWIZZLE layout
//
//      A[row ry][16-byte k-chunk kc][pixel px] : 8 bf16 channels            (address = ((ry*KC + kc)*PP + px) * 16)
//
// In that layout the UMMA "M" rows (pixels) are uniformly 16 bytes apart (SBO = 128 = 8 rows x 16 B) and k-chunks
// are LBO = PP*16 bytes apart, so the operand of filter tap (dy,dx) is the SAME buffer viewed through a descriptor
// whose start address is shifted by (dy*KC*PP + dx)*16 bytes: 9 taps x 3 k-steps (x3 bf16x3 passes) = 81 MMAs
// (M=128, N=16, K=16) read shifted windows of one staging buffer -- no per-tap copies, each input element is
// transformed exactly once (the generic gather kernel redid the affine + split for every tap: 10x the instructions).
// The packed weights of all 9 taps (27 KB with hi/lo) arrive with one cp.async.bulk and stay resident.
//
// Algorithmic bytes per output pixel: 48*4 read + 12*4 written = 240 B (HBM-bound; the 3x halo re-read is L2 traffic).
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int ROWS_TILE = 128;
constexpr int ROWS_PP = 131;            // pixels per staged row incl. halo (130) + 1 pad: spreads planes over banks
constexpr int ROWS_THREADS = 256;

struct RowsArgs {
    const float *in;
    const float *scale;
    const float *shift;
    const unsigned char *wplanar;       // [hi | lo] each 9 * KC * N_pad * 16 bytes
    float *out;
    double *stats;
    long stats_stride;
    int B, H, W;
    int in_pitch;
    int C_out, N_pad, out_pitch, out_choff;
    int relu;
};

template <int C_IN, bool SPLIT>
__global__ void __launch_bounds__(ROWS_THREADS) conv3x3_rows_kernel(const RowsArgs a) {
    constexpr int KC = C_IN / 8;                     // 16-byte k-chunks per pixel
    constexpr int CQ = C_IN / 4;                     // float4 quads per pixel
    constexpr int PLANE = ROWS_PP * 16;              // bytes between k-chunks (LBO)
    constexpr int A_BYTES = 3 * KC * PLANE;          // one staging image (hi or lo)
    constexpr int ROW_F4 = 130 * CQ;                 // float4 per staged row
    constexpr int TOTAL_F4 = 3 * ROW_F4;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_scale[C_IN];
    __shared__ __align__(16) float s_shift[C_IN];

    unsigned char *a_hi = smem;
    unsigned char *a_lo = smem + A_BYTES;
    unsigned char *w_sm = smem + (SPLIT ? 2 : 1) * A_BYTES;
    const int w_img = 9 * KC * a.N_pad * 16;         // bytes of one weight image (hi or lo)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_w = smem_u32(&s_bar[0]), bar_acc = smem_u32(&s_bar[1]);

    // tile -> (b, y, x0)
    const int tiles_x = a.W / ROWS_TILE;
    const int tx = blockIdx.x % tiles_x;
    const long t2 = blockIdx.x / tiles_x;
    const int y = static_cast<int>(t2 % a.H);
    const long b = t2 / a.H;
    const int x0 = tx * ROWS_TILE;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_acc, 1);
        fence_mbar_init();
        const uint32_t bytes = static_cast<uint32_t>((SPLIT ? 2 : 1) * w_img);
        mbar_expect_tx(bar_w, bytes);
        bulk_g2s(smem_u32(w_sm), a.wplanar, bytes, bar_w);
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 32);
    }
    for (int i = tid; i < C_IN; i += ROWS_THREADS) {
        s_scale[i] = a.scale ? a.scale[i] : 1.f;
        s_shift[i] = a.shift ? a.shift[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    // ---- stage the 3 x 130-pixel halo: linear float4 reads, affine, bf16 split, planar stores
    const float *img = a.in + b * a.H * static_cast<long>(a.W) * a.in_pitch;
    constexpr int ITERS = (TOTAL_F4 + ROWS_THREADS - 1) / ROWS_THREADS;
    constexpr int BATCH = 5;
#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += BATCH) {
        float4 v[BATCH];
        int off[BATCH], q4[BATCH];
        bool ok[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int f = tid + (it0 + j) * ROWS_THREADS;
            const int ry = f / ROW_F4;
            const int rem = f - ry * ROW_F4;
            const int px = rem / CQ;
            const int q = rem - px * CQ;
            const int iy = y + ry - 1, ix = x0 + px - 1;
            off[j] = f < TOTAL_F4 ? ((ry * KC + (q >> 1)) * ROWS_PP + px) * 16 + (q & 1) * 8 : -1;
            q4[j] = q * 4;
            ok[j] = f < TOTAL_F4 && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[j]) v[j] = __ldg(reinterpret_cast<const float4 *>(img + (static_cast<long>(iy) * a.W + ix) * a.in_pitch + q4[j]));
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            if (off[j] < 0) continue;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);          // zero padding is applied AFTER the affine
            if (ok[j]) {
                const float4 sc = *reinterpret_cast<const float4 *>(s_scale + q4[j]);
                const float4 sh = *reinterpret_cast<const float4 *>(s_shift + q4[j]);
                o.x = fmaf(v[j].x, sc.x, sh.x); o.y = fmaf(v[j].y, sc.y, sh.y);
                o.z = fmaf(v[j].z, sc.z, sh.z); o.w = fmaf(v[j].w, sc.w, sh.w);
                if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            }
            store_quad<SPLIT>(a_hi, a_lo, static_cast<uint32_t>(off[j]), o);
        }
    }
    fence_proxy_async();
    __syncthreads();

    // ---- 9 taps x 3 k-steps: shifted descriptor windows over the staging buffer (warp-uniform, one elected lane issues)
    if (warp == 0) {
        mbar_wait(bar_w, 0);
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(ROWS_TILE, a.N_pad);
            const uint64_t dA_hi = make_nosw_desc(smem_u32(a_hi), PLANE, 128);
            const uint64_t dA_lo = make_nosw_desc(smem_u32(a_lo), PLANE, 128);
            const uint64_t dW_hi = make_nosw_desc(smem_u32(w_sm), static_cast<uint32_t>(a.N_pad * 16), 128);
            const uint64_t dW_lo = dW_hi + static_cast<uint64_t>(w_img >> 4);
            uint32_t acc = 0;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
                for (int ks = 0; ks < KC / 2; ++ks) {
                    const uint64_t aoff = static_cast<uint32_t>((dy * KC + 2 * ks) * ROWS_PP + dx);
                    const uint64_t woff = static_cast<uint32_t>((tap * KC + 2 * ks) * a.N_pad);
                    umma_bf16(tmem_base, dA_hi + aoff, dW_hi + woff, idesc, acc);
                    acc = 1;
                    if (SPLIT) {
                        umma_bf16(tmem_base, dA_lo + aoff, dW_hi + woff, idesc, 1u);
                        umma_bf16(tmem_base, dA_hi + aoff, dW_lo + woff, idesc, 1u);
                    }
                }
            }
            umma_commit(bar_acc);
        }
        __syncwarp();
    }

    // ---- epilogue: warps 0-3 own TMEM lanes [32w, 32w+32) = pixels x0 + 32w + lane
    mbar_wait(bar_acc, 0);
    __syncwarp();
    tc_fence_after();
    float *tile = reinterpret_cast<float *>(smem);           // [128][N_pad+1] statistics staging (MMAs have retired)
    const int tp = a.N_pad + 1;
    if (warp < 4) {
        const int row = warp * 32 + lane;
        float *orow = a.out + ((b * a.H + y) * static_cast<long>(a.W) + x0 + row) * a.out_pitch + a.out_choff;
        const bool vec_ok = ((a.out_pitch | a.out_choff) & 3) == 0;
        for (int g = 0; g < a.N_pad; g += 16) {
            float v[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(g), v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = g + q * 4;
                if (vec_ok && n + 3 < a.C_out) {
                    *reinterpret_cast<float4 *>(orow + n) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (n + e < a.C_out) orow[n + e] = v[q * 4 + e];
                }
            }
            if (a.stats != nullptr) {
#pragma unroll
                for (int e = 0; e < 16; ++e) tile[row * tp + g + e] = v[e];
            }
        }
    }
    if (a.stats != nullptr) {
        __syncthreads();
        for (int n = tid; n < a.C_out; n += ROWS_THREADS) {
            double s1 = 0.0, s2 = 0.0;
            for (int r = 0; r < ROWS_TILE; ++r) {
                const double x = static_cast<double>(tile[r * tp + n]);
                s1 += x; s2 += x * x;
            }
            atomicAdd(a.stats + n, s1);
            atomicAdd(a.stats + a.stats_stride + n, s2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, 32);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent ROLLING-ROW variant (statistics epilogue for C_out <= 16: per-thread double sums over the CTA's pixels, round 2).
// A CTA owns a vertical band (image b, 128-pixel column block, RB consecutive output rows) and walks down it: output row
// y needs input rows y-1, y, y+1, of which y-1 and y are already staged from the previous step -- so every new output row
// stages exactly ONE new halo row (130 px) into a 4-slot ring instead of three (3x fewer loads, conversions and
// shared-memory stores, and no 3x L2 re-read of the bottleneck tensor).  Tap (dy,dx) of output row r reads ring slot
// (r+dy) mod 4 through a shifted no-swizzle descriptor, exactly as in the kernel above.
// Tensor-side shared-memory traffic is cut as well: the weights are staged as one N=32 operand [B_hi ; B_lo], so
//     D[:, 0:16]  += A_hi * B_hi      D[:, 16:32] += A_hi * B_lo      (ONE M128 N32 K16 MMA, A_hi read once)
//     D[:, 0:16]  += A_lo * B_hi                                       (one M128 N16 K16 MMA)
// and the epilogue adds the two 16-column halves.  2 MMAs and 76 operand wavefronts per k-step instead of 3 and 108.
// Roles: warps 0-7 producers, warp 8 MMA issuer, warps 9-12 epilogue; TMEM 2 x 32 columns.
constexpr int RP_THREADS = 256 + 32 + 128;
constexpr int RING = 4;

__device__ __forceinline__ void rows_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct RollArgs {
    RowsArgs r;
    const unsigned char *wcat;     // SPLIT: [tap][kc][32 rows (hi 0-15, lo 16-31)][16 B]; else planar hi image
    int rb;                        // output rows per band
    int bands_y;                   // bands per image column block
    long nunits;                   // B * tiles_x * bands_y
};

template <int C_IN, bool SPLIT>
__global__ void __launch_bounds__(RP_THREADS, 1) conv3x3_roll_kernel(const RollArgs ra) {
    const RowsArgs &a = ra.r;
    constexpr int KC = C_IN / 8;
    constexpr int CQ = C_IN / 4;
    constexpr int PLANE = ROWS_PP * 16;
    constexpr int SLOT_BYTES = KC * PLANE;                 // one staged row, one image (hi or lo)
    constexpr int IMG_BYTES = RING * SLOT_BYTES;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[1 + 2 * RING + 4];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_scale[C_IN];
    __shared__ __align__(16) float s_shift[C_IN];

    unsigned char *ring_hi = smem;
    unsigned char *ring_lo = smem + IMG_BYTES;             // only when SPLIT
    unsigned char *w_sm = smem + (SPLIT ? 2 : 1) * IMG_BYTES;
    const int wn = 2 * a.N_pad;                            // rows of the staged weight operand: always the [hi ; lo] image
    const uint32_t acc_cols = (SPLIT ? 2 * a.N_pad : a.N_pad) <= 32 ? 32u : ((SPLIT ? 2 * a.N_pad : a.N_pad) <= 64 ? 64u : 128u);
    const int w_bytes = 9 * KC * wn * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_w = smem_u32(&s_bar[0]);
    const uint32_t bar_rfull = smem_u32(&s_bar[1]);                 // [RING] row staged
    const uint32_t bar_rempty = smem_u32(&s_bar[1 + RING]);         // [RING] last MMA reading the row retired
    const uint32_t bar_accfull = smem_u32(&s_bar[1 + 2 * RING]);    // [2]
    const uint32_t bar_accempty = smem_u32(&s_bar[3 + 2 * RING]);   // [2]

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < RING; ++i) { mbar_init(bar_rfull + 8 * i, 8); mbar_init(bar_rempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accempty + 8 * i, 4); }
        fence_mbar_init();
    }
    if (warp == 8) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 2 * acc_cols);
    }
    for (int i = tid; i < C_IN; i += RP_THREADS) {
        s_scale[i] = a.scale ? a.scale[i] : 1.f;
        s_shift[i] = a.shift ? a.shift[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int tiles_x = (a.W + ROWS_TILE - 1) / ROWS_TILE;

    // unit -> (b, tx, band); rows of the band: y0 .. y0+nrows-1
    auto decode = [&](long unit, long &b, int &x0, int &y0, int &nrows) {
        const int band = static_cast<int>(unit % ra.bands_y);
        const long t = unit / ra.bands_y;
        x0 = static_cast<int>(t % tiles_x) * ROWS_TILE;
        b = t / tiles_x;
        y0 = band * ra.rb;
        nrows = min(ra.rb, a.H - y0);
    };

    if (warp < 8) {
        // ======================================================= PRODUCERS: one halo row per step
        constexpr int SLOTS = 256 / CQ;
        constexpr int NPX = (130 + SLOTS - 1) / SLOTS;
        const int q = tid % CQ, ps = tid / CQ;
        const bool active = ps < SLOTS;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + q * 4);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + q * 4);
        const uint32_t soff0 = static_cast<uint32_t>(((q >> 1) * ROWS_PP + ps) * 16 + (q & 1) * 8);
        const long gstride = static_cast<long>(SLOTS) * a.in_pitch;
        uint32_t R = 0;                                         // rows staged so far by this CTA
        for (long unit = blockIdx.x; unit < ra.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            float4 v[2][NPX];
            unsigned okm[2];
            auto load_row = [&](int iy, float4 (&dst)[NPX], unsigned &m) {
                const bool row_ok = active && iy >= 0 && iy < a.H;
                const float *rp = a.in + ((b * a.H + iy) * static_cast<long>(a.W) + (x0 - 1 + ps)) * a.in_pitch + q * 4;
                m = 0;
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    const int px = ps + i * SLOTS;
                    const int ix = x0 - 1 + px;
                    const bool ok = row_ok && px < 130 && ix >= 0 && ix < a.W;
                    dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) { dst[i] = __ldg(reinterpret_cast<const float4 *>(rp + i * gstride)); m |= 1u << i; }
                }
            };
            auto store_row = [&](uint32_t Rg, const float4 (&src)[NPX], unsigned m) {
                const uint32_t slot = Rg % RING, ph = (Rg / RING) & 1;
                mbar_wait(bar_rempty + 8 * slot, ph ^ 1);
                unsigned char *hi = ring_hi + slot * SLOT_BYTES;
                unsigned char *lo = ring_lo + slot * SLOT_BYTES;
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    if (!active || ps + i * SLOTS >= 130) continue;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);           // zero padding is applied AFTER the affine
                    if ((m >> i) & 1u) {
                        o.x = fmaf(src[i].x, sc.x, sh.x); o.y = fmaf(src[i].y, sc.y, sh.y);
                        o.z = fmaf(src[i].z, sc.z, sh.z); o.w = fmaf(src[i].w, sc.w, sh.w);
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    }
                    store_quad<SPLIT>(hi, lo, soff0 + static_cast<uint32_t>(i * SLOTS * 16), o);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) rows_mbar_arrive(bar_rfull + 8 * slot);
            };
            // rows y0-1 .. y0+nrows (nrows+2 rows), software-pipelined: row k+1's loads fly while row k is converted
            const int total = nrows + 2;
            load_row(y0 - 1, v[0], okm[0]);
            for (int k = 0; k < total; ++k) {
                const int cur = k & 1;
                if (k + 1 < total) {
                    if (cur == 0) load_row(y0 + k, v[1], okm[1]); else load_row(y0 + k, v[0], okm[0]);
                }
                if (cur == 0) store_row(R, v[0], okm[0]); else store_row(R, v[1], okm[1]);
                ++R;
            }
        }
    } else if (warp == 8) {
        // ======================================================= MMA ISSUER (warp-uniform loops, one elected lane issues)
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(bar_w, static_cast<uint32_t>(w_bytes));
            bulk_g2s(smem_u32(w_sm), ra.wcat, static_cast<uint32_t>(w_bytes), bar_w);
        }
        mbar_wait(bar_w, 0);
        const uint32_t idesc_cat = make_idesc_bf16(ROWS_TILE, SPLIT ? wn : a.N_pad);    // non-split reads only the hi rows
        const uint32_t idesc_16 = make_idesc_bf16(ROWS_TILE, a.N_pad);
        // descriptor bases; per MMA only the 14-bit start-address field (16-byte units) moves
        const uint64_t dA_hi = make_nosw_desc(smem_u32(ring_hi), PLANE, 128);
        const uint64_t dA_lo = make_nosw_desc(smem_u32(ring_lo), PLANE, 128);
        const uint64_t dW = make_nosw_desc(smem_u32(w_sm), static_cast<uint32_t>(wn * 16), 128);
        uint32_t R0 = 0, j = 0;
        for (long unit = blockIdx.x; unit < ra.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            // rows R0 (y0-1) and R0+1 (y0) must be present before the first output row
            for (int r = 0; r < 2; ++r) mbar_wait(bar_rfull + 8 * ((R0 + r) % RING), ((R0 + r) / RING) & 1);
            for (int i = 0; i < nrows; ++i, ++j) {
                const uint32_t buf = j & 1, aph = (j >> 1) & 1;
                const uint32_t Rn = R0 + i + 2;                                      // newest row needed (y+1)
                mbar_wait(bar_rfull + 8 * (Rn % RING), (Rn / RING) & 1);
                mbar_wait(bar_accempty + 8 * buf, aph ^ 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t d_tmem = tmem_base + buf * acc_cols;
                    uint32_t acc = 0;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t slot16 = ((R0 + i + dy) % RING) * (SLOT_BYTES / 16);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                            for (int ks = 0; ks < KC / 2; ++ks) {
                                const uint64_t aoff = slot16 + static_cast<uint32_t>(2 * ks * ROWS_PP + dx);
                                const uint64_t woff = static_cast<uint32_t>(((dy * 3 + dx) * KC + 2 * ks) * wn);
                                umma_bf16(d_tmem, dA_hi + aoff, dW + woff, idesc_cat, acc);       // hi*[hi|lo] -> cols [0, 2*N_pad)
                                if (SPLIT) umma_bf16(d_tmem, dA_lo + aoff, dW + woff, idesc_16, 1u);  // lo*hi -> cols [0, N_pad)
                                acc = 1;
                            }
                        }
                    }
                    umma_commit(bar_rempty + 8 * ((R0 + i) % RING));                     // row y-1 is dead after this output row
                    if (i == nrows - 1) {                                                // band ends: rows y and y+1 die as well
                        umma_commit(bar_rempty + 8 * ((R0 + i + 1) % RING));
                        umma_commit(bar_rempty + 8 * ((R0 + i + 2) % RING));
                    }
                    umma_commit(bar_accfull + 8 * buf);
                }
                __syncwarp();
            }
            R0 += static_cast<uint32_t>(nrows + 2);
        }
    } else {
        // ======================================================= EPILOGUE
        const int q4 = warp & 3;
        const bool vec_ok = ((a.out_pitch | a.out_choff) & 3) == 0;
        // batch-statistic BatchNorm downstream (training forward): per-channel sum / sum of squares of the outputs, accumulated per thread
        // (= per pixel column) in double over the CTA's whole pixel range, reduced once at the end (N_pad <= 16 only: conv2's 12 channels)
        const bool want_stats = a.stats != nullptr;
        double st1[16], st2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { st1[e] = 0.0; st2[e] = 0.0; }
        uint32_t j = 0;
        for (long unit = blockIdx.x; unit < ra.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            for (int i = 0; i < nrows; ++i, ++j) {
                const uint32_t buf = j & 1, aph = (j >> 1) & 1;
                mbar_wait(bar_accfull + 8 * buf, aph);
                __syncwarp();
                tc_fence_after();
                const int x = x0 + q4 * 32 + lane;
                float *orow = a.out + ((b * a.H + (y0 + i)) * static_cast<long>(a.W) + (x < a.W ? x : 0)) * a.out_pitch + a.out_choff;
                for (int g16 = 0; g16 < a.N_pad; g16 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + buf * acc_cols + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(g16), v);
                    if (SPLIT) {
                        float w2[16];
                        tmem_ld16(tmem_base + buf * acc_cols + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(a.N_pad + g16), w2);
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] += w2[e];
                    }
                    if (g16 + 16 >= a.N_pad) {                                       // last column group read: release the accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) rows_mbar_arrive(bar_accempty + 8 * buf);
                    }
                    if (x < a.W) {
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const int n = g16 + qq * 4;
                            if (vec_ok && n + 3 < a.C_out) {
                                *reinterpret_cast<float4 *>(orow + n) = make_float4(v[qq * 4], v[qq * 4 + 1], v[qq * 4 + 2], v[qq * 4 + 3]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (n + e < a.C_out) orow[n + e] = v[qq * 4 + e];
                            }
                        }
                        if (want_stats && g16 == 0) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) { const double d = static_cast<double>(v[e]); st1[e] += d; st2[e] += d * d; }
                        }
                    }
                }
            }
        }
        if (want_stats) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                double s1 = st1[e], s2 = st2[e];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                }
                if (lane == 0 && e < a.C_out) {
                    atomicAdd(a.stats + e, s1);
                    atomicAdd(a.stats + a.stats_stride + e, s2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(tmem_base, 2 * acc_cols);
    }
}

// OIHW fp32 -> [tap][kc][2*N_pad rows: n<N_pad hi(n), n>=N_pad lo(n-N_pad)][8 bf16]  (the concatenated operand of the rolling kernel)
__global__ void pack_cat_kernel(const float *__restrict__ w, unsigned char *__restrict__ out, int C_out, int C_in, int C_in_pad, int N_pad) {
    const int KC = C_in_pad / 8;
    const int rows = 2 * N_pad;
    const int total = 9 * KC * rows * 8;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int e = idx & 7;
        const int n = (idx >> 3) % rows;
        const int kc = (idx / (8 * rows)) % KC;
        const int tap = idx / (8 * rows * KC);
        const int c = kc * 8 + e, nn = n % N_pad;
        float v = 0.f;
        if (nn < C_out && c < C_in) v = w[(static_cast<long>(nn) * C_in + c) * 9 + tap];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        *reinterpret_cast<__nv_bfloat16 *>(out + static_cast<size_t>(idx) * 2) = n < N_pad ? hi : lo;
    }
}

// OIHW fp32 -> planar [tap][kc][n][8 bf16], hi image then lo image.
__global__ void pack_planar_kernel(const float *__restrict__ w, unsigned char *__restrict__ out, int C_out, int C_in,
                                   int N_pad) {
    const int KC = C_in / 8;
    const int total = 9 * KC * N_pad * 8;
    const size_t img = static_cast<size_t>(9) * KC * N_pad * 16;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int e = idx & 7;
        const int n = (idx >> 3) % N_pad;
        const int kc = (idx / (8 * N_pad)) % KC;
        const int tap = idx / (8 * N_pad * KC);
        const int c = kc * 8 + e;
        float v = 0.f;
        if (n < C_out) v = w[(static_cast<long>(n) * C_in + c) * 9 + tap];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const size_t off = (static_cast<size_t>((tap * KC + kc) * N_pad + n) * 8 + e) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(out + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(out + img + off) = lo;
    }
}

}  // namespace

// ---- entry points used by conv_gemm.cu's dispatcher
static int rows_cin_pad(int C_in) { return C_in == 48 ? 48 : (C_in <= 16 ? 16 : 0); }      // channel counts the kernels are built for

static size_t rows_planar_bytes(int C_out, int C_in) {
    const int N_pad = (C_out + 15) & ~15;
    return static_cast<size_t>(2) * 9 * (rows_cin_pad(C_in) / 8) * N_pad * 16;
}

size_t eml_rows_wpack_bytes(int C_out, int C_in) {
    if (rows_cin_pad(C_in) == 0 || C_out > 64) return 0;
    // planar hi|lo image (non-persistent kernel) followed by the [hi;lo] concatenated image (rolling kernel)
    const int N_pad = (C_out + 15) & ~15;
    return rows_planar_bytes(C_out, C_in) + static_cast<size_t>(9) * (rows_cin_pad(C_in) / 8) * 2 * N_pad * 16;
}

int eml_rows_pack(const float *w_oihw, unsigned char *dst, int C_out, int C_in, cudaStream_t st) {
    const int N_pad = (C_out + 15) & ~15;
    if (C_in == 48) pack_planar_kernel<<<32, 256, 0, st>>>(w_oihw, dst, C_out, C_in, N_pad);     // image of the statistics-epilogue kernel
    pack_cat_kernel<<<32, 256, 0, st>>>(w_oihw, dst + rows_planar_bytes(C_out, C_in), C_out, C_in, rows_cin_pad(C_in), N_pad);
    return eml_launch_status();
}

bool eml_rows_supported(const eml_conv_params *p) {
    if (p->mode != EML_CONV_3x3) return false;
    const bool persist_ok = (p->stats == nullptr || (p->C_out <= 16 && !eml_env_flag("EML_NO_PERSIST_STATS"))) && !eml_env_flag("EML_NO_PERSIST");
    // 12 -> 48 data gradient (conv2's dgrad): input pitch must give 16 readable, zero-padded channels
    if (p->C_in <= 16 && p->in_pitch >= 16 && p->C_out <= 64 && persist_ok && p->scale == nullptr && p->shift == nullptr &&
        (p->precision == EML_PREC_BF16 || p->precision == EML_PREC_BF16X3))
        return true;
    return p->C_in == 48 && p->C_out <= 16 && ((p->W % ROWS_TILE) == 0 || persist_ok) &&
           (p->precision == EML_PREC_BF16 || p->precision == EML_PREC_BF16X3);
}

int eml_rows_forward(const eml_conv_params *p, const unsigned char *wplanar, cudaStream_t st) {
    RowsArgs a{};
    a.in = p->in; a.scale = p->scale; a.shift = p->shift; a.wplanar = wplanar; a.out = p->out;
    a.stats = p->stats; a.stats_stride = p->stats_stride > 0 ? p->stats_stride : p->C_out;
    a.B = p->B; a.H = p->H; a.W = p->W; a.in_pitch = p->in_pitch;
    a.C_out = p->C_out; a.N_pad = (p->C_out + 15) & ~15; a.out_pitch = p->out_pitch; a.out_choff = p->out_choff;
    a.relu = p->relu;
    const bool split = p->precision == EML_PREC_BF16X3;
    constexpr int KC = 48 / 8;
    const size_t a_bytes = static_cast<size_t>(3) * KC * ROWS_PP * 16;
    size_t smem = (split ? 2 : 1) * (a_bytes + static_cast<size_t>(9) * KC * a.N_pad * 16);
    const size_t stats_bytes = a.stats ? static_cast<size_t>(ROWS_TILE) * (a.N_pad + 1) * 4 : 0;
    if (stats_bytes > smem) smem = stats_bytes;
    const long tiles = static_cast<long>(p->B) * p->H * (p->W / ROWS_TILE);
    if (tiles >= (1L << 31)) return EML_E_SHAPE;
    cudaError_t e;
    if ((a.stats == nullptr || (p->C_out <= 16 && !eml_env_flag("EML_NO_PERSIST_STATS"))) && !eml_env_flag("EML_NO_PERSIST")) {
        RollArgs ra{};
        ra.r = a;
        ra.wcat = wplanar + rows_planar_bytes(p->C_out, p->C_in);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int tiles_x = (p->W + ROWS_TILE - 1) / ROWS_TILE;
        // band height: 32 rows (6 % halo overhead) unless that leaves fewer than ~8 units per SM, then 16
        int rb = 32;
        if (static_cast<long>(p->B) * tiles_x * ((p->H + rb - 1) / rb) < 8L * sms) rb = 16;
        if (rb > p->H) rb = p->H;
        ra.rb = rb;
        ra.bands_y = (p->H + rb - 1) / rb;
        ra.nunits = static_cast<long>(p->B) * tiles_x * ra.bands_y;
        const unsigned grid = static_cast<unsigned>(ra.nunits < sms ? ra.nunits : sms);
        const int kc = rows_cin_pad(p->C_in) / 8;
        const size_t ring = static_cast<size_t>(RING) * kc * ROWS_PP * 16;
        const size_t psm = (split ? 2 : 1) * ring + static_cast<size_t>(9) * kc * 2 * a.N_pad * 16;
        auto go = [&](auto kern) -> int {
            cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(psm));
            if (ee != cudaSuccess) return static_cast<int>(ee);
            kern<<<grid, RP_THREADS, psm, st>>>(ra);
            return eml_launch_status();
        };
        if (p->C_in == 48) return split ? go(conv3x3_roll_kernel<48, true>) : go(conv3x3_roll_kernel<48, false>);
        return split ? go(conv3x3_roll_kernel<16, true>) : go(conv3x3_roll_kernel<16, false>);
    }
    if (p->C_in != 48) return EML_E_SHAPE;
    if (split) {
        e = cudaFuncSetAttribute(conv3x3_rows_kernel<48, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        conv3x3_rows_kernel<48, true><<<static_cast<unsigned>(tiles), ROWS_THREADS, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(conv3x3_rows_kernel<48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        conv3x3_rows_kernel<48, false><<<static_cast<unsigned>(tiles), ROWS_THREADS, smem, st>>>(a);
    }
    return eml_launch_status();
}
