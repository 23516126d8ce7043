// Weight gradient of conv2 (3x3, 48 -> 12) on tensor cores (sm_100a):
//     dW[n, c, tap] = sum_p dY[p, n] * N2[p + off(tap), c],      N2 = scale*b + shift inside the image, 0 outside
// Nine GEMMs (one per filter tap) whose reduction dimension is the pixel index.  The staging is the forward rolling-row ring
// (conv3x3_rows.cu): halo rows of N2 as planar tiles [16-byte channel octet][pixel], one NEW row per output row.  Read as a
// tcgen05 MN-major NO-SWIZZLE operand -- octets SBO = one plane apart, pixels 16 B apart (LBO = 128 B per 8 pixels) -- tap (dy,dx)
// is again just a start-address shift (ring slot (r+dy)%4, +dx pixels).  dY rows are staged the same way (2 octets).
//     D_tap[c, n]  (TMEM: lane = channel c, 16 columns per tap)  +=  N2_tap^T[c, k] * dY[k, n]      k = 16 pixels per MMA
// All nine 128x16 accumulators stay in TMEM for the CTA's whole pixel range; one epilogue at the end does the atomicAdds.
// M is 128 (UMMA minimum for the lane-per-row accumulator layout used here), so channel octets 6..15 of the A operand are
// whatever follows the slot in shared memory: those accumulator rows (c >= 48) are simply never written out.
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int T_TILE = 128;
constexpr int T_PP = 131;
constexpr int T_THREADS = 256 + 32 + 128;
constexpr int T_RING = 4;
constexpr int T_C = 48, T_KC = 6, T_CQ = 12;
constexpr int T_PLANE = T_PP * 16;
constexpr int T_SLOT = T_KC * T_PLANE;                 // one staged N2 row, one image
constexpr int T_IMG = T_RING * T_SLOT;
constexpr int T_PAD = 10 * T_PLANE;                    // readable slack behind the lo ring (garbage octets of the M=128 operand)
constexpr int T_YPLANE = T_TILE * 16;                  // dY: one octet plane of a 128-pixel row
constexpr int T_YBUF = 2 * T_YPLANE;                   // 2 octets (16 channels), one image

struct TArgs {
    const float *dY; int dy_pitch, N;                  // (B,H,W,dy_pitch), N <= 16 gradient channels
    const float *b; int b_pitch;                       // bottleneck (B,H,W,b_pitch), 48 channels
    const float *scale, *shift;                        // BN2 folded affine
    float *dW;                                         // (N, 48, 3, 3)
    int B, H, W, rb, bands_y;
    long nunits;
};

__device__ __forceinline__ void t_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// MN-major no-swizzle: SBO = bytes between 8-element MN blocks (channel octets), LBO = bytes between 8-row K groups (8 pixels)
__device__ __forceinline__ uint64_t make_mn_nosw_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
__host__ __device__ constexpr uint32_t t_idesc(int M, int N) {          // bf16 x bf16 -> f32, A and B MN-major
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

template <bool SPLIT>
__global__ void __launch_bounds__(T_THREADS, 1) wgrad3x3_tc_kernel(const TArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2 * T_RING + 4 + 1];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_scale[T_C];
    __shared__ __align__(16) float s_shift[T_C];
    unsigned char *ring_hi = smem;
    unsigned char *ring_lo = smem + T_IMG;
    unsigned char *ybuf = smem + (SPLIT ? 2 : 1) * T_IMG + T_PAD;        // [2 slots][hi | lo][2 planes][128 px][16 B]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_rfull = smem_u32(&s_bar[0]), bar_rempty = smem_u32(&s_bar[T_RING]);
    const uint32_t bar_yfull = smem_u32(&s_bar[2 * T_RING]), bar_yempty = smem_u32(&s_bar[2 * T_RING + 2]);
    const uint32_t bar_done = smem_u32(&s_bar[2 * T_RING + 4]);

    if (tid == 0) {
        for (int i = 0; i < T_RING; ++i) { mbar_init(bar_rfull + 8 * i, 8); mbar_init(bar_rempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_yfull + 8 * i, 8); mbar_init(bar_yempty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 8) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 256);                              // 9 taps x 16 columns
    }
    for (int i = tid; i < T_C; i += T_THREADS) {
        s_scale[i] = a.scale ? a.scale[i] : 1.f;
        s_shift[i] = a.shift ? a.shift[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int tiles_x = (a.W + T_TILE - 1) / T_TILE;
    auto decode = [&](long unit, long &b, int &x0, int &y0, int &nrows) {
        const int band = static_cast<int>(unit % a.bands_y);
        const long t = unit / a.bands_y;
        x0 = static_cast<int>(t % tiles_x) * T_TILE;
        b = t / tiles_x;
        y0 = band * a.rb;
        nrows = min(a.rb, a.H - y0);
    };

    if (warp < 8) {
        // ======================================================= PRODUCERS
        constexpr int SLOTS = 256 / T_CQ;                               // 21 pixel slots x 12 quads
        constexpr int NPX = (130 + SLOTS - 1) / SLOTS;
        const int q = tid % T_CQ, ps = tid / T_CQ;
        const bool active = ps < SLOTS;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + q * 4);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + q * 4);
        const uint32_t soff0 = static_cast<uint32_t>(((q >> 1) * T_PP + ps) * 16 + (q & 1) * 8);
        const long gstride = static_cast<long>(SLOTS) * a.b_pitch;
        uint32_t R = 0, j = 0;
        for (long unit = blockIdx.x; unit < a.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            auto stage_n2 = [&](int iy, uint32_t Rg) {
                const bool row_ok = active && iy >= 0 && iy < a.H;
                const float *rp = a.b + ((b * a.H + iy) * static_cast<long>(a.W) + (x0 - 1 + ps)) * a.b_pitch + q * 4;
                float4 v[NPX]; unsigned m = 0;
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    const int px = ps + i * SLOTS, ix = x0 - 1 + px;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row_ok && px < 130 && ix >= 0 && ix < a.W) { v[i] = __ldg(reinterpret_cast<const float4 *>(rp + i * gstride)); m |= 1u << i; }
                }
                const uint32_t slot = Rg % T_RING, ph = (Rg / T_RING) & 1;
                mbar_wait(bar_rempty + 8 * slot, ph ^ 1);
                unsigned char *hi = ring_hi + slot * T_SLOT, *lo = ring_lo + slot * T_SLOT;
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    if (!active || ps + i * SLOTS >= 130) continue;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((m >> i) & 1u) {
                        o.x = fmaf(v[i].x, sc.x, sh.x); o.y = fmaf(v[i].y, sc.y, sh.y);
                        o.z = fmaf(v[i].z, sc.z, sh.z); o.w = fmaf(v[i].w, sc.w, sh.w);
                    }
                    store_quad<SPLIT>(hi, lo, soff0 + static_cast<uint32_t>(i * SLOTS * 16), o);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) t_mbar_arrive(bar_rfull + 8 * slot);
            };
            auto stage_dy = [&](int y, uint32_t jj) {
                // 128 pixels x 4 quads (16 channels) = 512 quads: thread -> (pixel p, quad qq), 2 per thread
                const uint32_t slot = jj & 1, ph = (jj >> 1) & 1;
                float4 v[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = tid + i * 256;
                    const int p = e >> 2, qq = e & 3;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (x0 + p < a.W && qq * 4 < a.N) {
                        const float *src = a.dY + ((b * a.H + y) * static_cast<long>(a.W) + x0 + p) * a.dy_pitch + qq * 4;
                        if (qq * 4 + 3 < a.N && (a.dy_pitch & 3) == 0) v[i] = __ldg(reinterpret_cast<const float4 *>(src));
                        else { v[i].x = src[0]; if (qq * 4 + 1 < a.N) v[i].y = src[1]; if (qq * 4 + 2 < a.N) v[i].z = src[2]; }
                    }
                }
                mbar_wait(bar_yempty + 8 * slot, ph ^ 1);
                unsigned char *hi = ybuf + slot * (SPLIT ? 2 : 1) * T_YBUF, *lo = hi + T_YBUF;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = tid + i * 256;
                    const int p = e >> 2, qq = e & 3;
                    store_quad<SPLIT>(hi, lo, static_cast<uint32_t>((qq >> 1) * T_YPLANE + p * 16 + (qq & 1) * 8), v[i]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) t_mbar_arrive(bar_yfull + 8 * slot);
            };
            // rows y0-1 .. y0+nrows of N2 (nrows+2), dY rows y0 .. y0+nrows-1
            stage_n2(y0 - 1, R++);
            stage_n2(y0, R++);
            for (int i = 0; i < nrows; ++i, ++j) {
                stage_n2(y0 + i + 1, R++);
                stage_dy(y0 + i, j);
            }
        }
    } else if (warp == 8) {
        // ======================================================= MMA ISSUER
        const bool leader = elect_one();
        const uint32_t idesc = t_idesc(128, 16);
        const uint64_t dA_hi = make_mn_nosw_desc(smem_u32(ring_hi), 128, T_PLANE);
        const uint64_t dA_lo = make_mn_nosw_desc(smem_u32(ring_lo), 128, T_PLANE);
        const uint64_t dY0 = make_mn_nosw_desc(smem_u32(ybuf), 128, T_YPLANE);
        uint32_t R0 = 0, j = 0;
        for (long unit = blockIdx.x; unit < a.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            for (int r = 0; r < 2; ++r) mbar_wait(bar_rfull + 8 * ((R0 + r) % T_RING), ((R0 + r) / T_RING) & 1);
            for (int i = 0; i < nrows; ++i, ++j) {
                const uint32_t Rn = R0 + i + 2;
                mbar_wait(bar_rfull + 8 * (Rn % T_RING), (Rn / T_RING) & 1);
                mbar_wait(bar_yfull + 8 * (j & 1), (j >> 1) & 1);
                tc_fence_after();
                if (leader) {
                    const uint64_t dy_hi = dY0 + static_cast<uint64_t>((j & 1) * ((SPLIT ? 2 : 1) * T_YBUF >> 4));
                    const uint64_t dy_lo = dy_hi + (T_YBUF >> 4);
                    const uint32_t first = (j == 0) ? 0u : 1u;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t slot16 = ((R0 + i + dy) % T_RING) * (T_SLOT / 16);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint32_t d_tmem = tmem_base + (dy * 3 + dx) * 16;
#pragma unroll
                            for (int k = 0; k < T_TILE / 16; ++k) {
                                const uint64_t aoff = slot16 + static_cast<uint32_t>(dx + k * 16);      // +dx pixels, +16 pixels per k-step
                                const uint64_t yoff = static_cast<uint32_t>(k * 16);
                                umma_bf16(d_tmem, dA_hi + aoff, dy_hi + yoff, idesc, (k == 0) ? first : 1u);
                                if (SPLIT) {
                                    umma_bf16(d_tmem, dA_lo + aoff, dy_hi + yoff, idesc, 1u);
                                    umma_bf16(d_tmem, dA_hi + aoff, dy_lo + yoff, idesc, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(bar_rempty + 8 * ((R0 + i) % T_RING));
                    if (i == nrows - 1) {
                        umma_commit(bar_rempty + 8 * ((R0 + i + 1) % T_RING));
                        umma_commit(bar_rempty + 8 * ((R0 + i + 2) % T_RING));
                    }
                    umma_commit(bar_yempty + 8 * (j & 1));
                }
                __syncwarp();
            }
            R0 += static_cast<uint32_t>(nrows + 2);
        }
        if (leader) umma_commit(bar_done);
        __syncwarp();
    } else {
        // ======================================================= EPILOGUE (once)
        const int q4 = warp & 3;
        mbar_wait(bar_done, 0);
        __syncwarp();
        tc_fence_after();
        const int c = q4 * 32 + lane;
        for (int tap = 0; tap < 9; ++tap) {
            float v[16];
            tmem_ld16(tmem_base + tap * 16 + (static_cast<uint32_t>(q4 * 32) << 16), v);
            if (c < T_C) {
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    if (n < a.N) atomicAdd(a.dW + (static_cast<long>(n) * T_C + c) * 9 + tap, v[n]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}


// =====================================================================================================================
// v2 (default): filter taps on BOTH GEMM dimensions.
//   * dx on N: the dY row is staged three times, shifted by dx pixels, as six octet planes (dx, octet) -> B is [k = ring pixel]
//     [n' = dx*16 + n] with N = 48, and the A operand is the UNSHIFTED ring row:
//         D[c, dx*16+n] += sum_k N2row[k, c] * dY[k - dx, n]          (k = ring pixel = image x - x0 + 1, 144 = 9 k-steps)
//   * dy on M: the ring slots are contiguous (slot = 6 octet planes), so the 16 octets of an M = 128 operand that starts at
//     slot s ARE rows s, s+1 and two thirds of s+2.  Two MMAs per k-step cover all three dy; where the 4-slot ring wraps the
//     pair lands in a different accumulator tile (6 tiles x 48 TMEM columns), summed once in the epilogue.
//   54 MMAs (N = 48) per 128-pixel row instead of 216 (N = 16); same staging traffic, dY stores x3 (dY is 16 of 64 channels).
constexpr int U_PP = 144;                               // 130 staged pixels, padded to 9 k-steps of 16 (130..143 stay zero)
constexpr int U_PLANE = U_PP * 16;
constexpr int U_SLOT = T_KC * U_PLANE;
constexpr int U_IMG = T_RING * U_SLOT;
constexpr int U_YBUF = 6 * U_PLANE;                     // (dx, octet) planes of one dY row, one image
constexpr int U_PAD = 4 * U_PLANE;                      // garbage octets of the last M = 128 operand stay inside the allocation
constexpr int U_TILES = 6, U_N = 48;

// which accumulator tiles a ring row at slot s feeds: {first plane, tile} for the two MMAs
__device__ __forceinline__ void u_pair(uint32_t s, int &pA, int &tA, int &pB, int &tB) {
    if (s <= 1) { pA = static_cast<int>(s) * T_KC; tA = 0; pB = (static_cast<int>(s) + 2) * T_KC + 4; tB = 1; }
    else if (s == 2) { pA = 2 * T_KC; tA = 2; pB = 0; tB = 3; }
    else { pA = 3 * T_KC; tA = 4; pB = 0; tB = 5; }
}
// accumulator row m of tile t -> (dy, c), or dy = -1 for a garbage row
__device__ __forceinline__ void u_row(int t, int m, int &dy, int &c) {
    dy = -1; c = 0;
    switch (t) {
    case 0: dy = m / T_C; c = m % T_C; break;                                   // 0..47 dy0 | 48..95 dy1 | 96..127 dy2 (c < 32)
    case 1: if (m < 16) { dy = 2; c = 32 + m; } break;
    case 2: if (m < 2 * T_C) { dy = m / T_C; c = m % T_C; } break;
    case 3: if (m < T_C) { dy = 2; c = m; } break;
    case 4: if (m < T_C) { dy = 0; c = m; } break;
    default: if (m < 2 * T_C) { dy = 1 + m / T_C; c = m % T_C; } break;
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(T_THREADS, 1) wgrad3x3_tc2_kernel(const TArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2 * T_RING + 4 + 1];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_scale[T_C];
    __shared__ __align__(16) float s_shift[T_C];
    constexpr int NIMG = SPLIT ? 2 : 1;
    constexpr int SMEM_TOTAL = NIMG * U_IMG + 2 * NIMG * U_YBUF + U_PAD;
    unsigned char *ring_hi = smem;
    unsigned char *ring_lo = smem + U_IMG;
    unsigned char *ybuf = smem + NIMG * U_IMG;                           // [2 slots][hi | lo][6 planes][144 px][16 B]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_rfull = smem_u32(&s_bar[0]), bar_rempty = smem_u32(&s_bar[T_RING]);
    const uint32_t bar_yfull = smem_u32(&s_bar[2 * T_RING]), bar_yempty = smem_u32(&s_bar[2 * T_RING + 2]);
    const uint32_t bar_done = smem_u32(&s_bar[2 * T_RING + 4]);

    if (tid == 0) {
        for (int i = 0; i < T_RING; ++i) { mbar_init(bar_rfull + 8 * i, 8); mbar_init(bar_rempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_yfull + 8 * i, 8); mbar_init(bar_yempty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 8) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 512);                              // 6 tiles x 48 columns
    }
    for (int i = tid; i < T_C; i += T_THREADS) {
        s_scale[i] = a.scale ? a.scale[i] : 1.f;
        s_shift[i] = a.shift ? a.shift[i] : 0.f;
    }
    // everything the MMAs may read without a producer having written it (padding pixels, shifted-dY borders, the garbage octets
    // of the M = 128 operands) must be finite: zero the whole allocation once
    for (int i = tid; i < SMEM_TOTAL / 16; i += T_THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int tiles_x = (a.W + T_TILE - 1) / T_TILE;
    auto decode = [&](long unit, long &b, int &x0, int &y0, int &nrows) {
        const int band = static_cast<int>(unit % a.bands_y);
        const long t = unit / a.bands_y;
        x0 = static_cast<int>(t % tiles_x) * T_TILE;
        b = t / tiles_x;
        y0 = band * a.rb;
        nrows = min(a.rb, a.H - y0);
    };

    if (warp < 8) {
        // ======================================================= PRODUCERS
        constexpr int SLOTS = 256 / T_CQ;                               // 21 pixel slots x 12 quads
        constexpr int NPX = (130 + SLOTS - 1) / SLOTS;
        const int q = tid % T_CQ, ps = tid / T_CQ;
        const bool active = ps < SLOTS;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + q * 4);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + q * 4);
        const uint32_t soff0 = static_cast<uint32_t>(((q >> 1) * U_PP + ps) * 16 + (q & 1) * 8);
        const long gstride = static_cast<long>(SLOTS) * a.b_pitch;
        uint32_t R = 0, j = 0;
        for (long unit = blockIdx.x; unit < a.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            auto stage_n2 = [&](int iy, uint32_t Rg) {
                const bool row_ok = active && iy >= 0 && iy < a.H;
                const float *rp = a.b + ((b * a.H + iy) * static_cast<long>(a.W) + (x0 - 1 + ps)) * a.b_pitch + q * 4;
                float4 v[NPX]; unsigned m = 0;
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    const int px = ps + i * SLOTS, ix = x0 - 1 + px;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row_ok && px < 130 && ix >= 0 && ix < a.W) { v[i] = __ldg(reinterpret_cast<const float4 *>(rp + i * gstride)); m |= 1u << i; }
                }
                const uint32_t slot = Rg % T_RING, ph = (Rg / T_RING) & 1;
                mbar_wait(bar_rempty + 8 * slot, ph ^ 1);
                unsigned char *hi = ring_hi + slot * U_SLOT, *lo = ring_lo + slot * U_SLOT;
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    if (!active || ps + i * SLOTS >= 130) continue;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((m >> i) & 1u) {
                        o.x = fmaf(v[i].x, sc.x, sh.x); o.y = fmaf(v[i].y, sc.y, sh.y);
                        o.z = fmaf(v[i].z, sc.z, sh.z); o.w = fmaf(v[i].w, sc.w, sh.w);
                    }
                    store_quad<SPLIT>(hi, lo, soff0 + static_cast<uint32_t>(i * SLOTS * 16), o);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) t_mbar_arrive(bar_rfull + 8 * slot);
            };
            auto stage_dy = [&](int y, uint32_t jj) {
                // 128 pixels x 4 quads (16 channels): thread -> (pixel p, quad qq), 2 per thread; each goes to ring pixel p + dx of
                // plane (dx, octet) for dx = 0, 1, 2.  Ring pixels < dx and >= 128 + dx of plane dx are never written: zero.
                const uint32_t slot = jj & 1, ph = (jj >> 1) & 1;
                float4 v[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = tid + i * 256;
                    const int p = e >> 2, qq = e & 3;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (x0 + p < a.W && qq * 4 < a.N) {
                        const float *src = a.dY + ((b * a.H + y) * static_cast<long>(a.W) + x0 + p) * a.dy_pitch + qq * 4;
                        if (qq * 4 + 3 < a.N && (a.dy_pitch & 3) == 0) v[i] = __ldg(reinterpret_cast<const float4 *>(src));
                        else { v[i].x = src[0]; if (qq * 4 + 1 < a.N) v[i].y = src[1]; if (qq * 4 + 2 < a.N) v[i].z = src[2]; }
                    }
                }
                mbar_wait(bar_yempty + 8 * slot, ph ^ 1);
                unsigned char *hi = ybuf + slot * NIMG * U_YBUF, *lo = hi + U_YBUF;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = tid + i * 256;
                    const int p = e >> 2, qq = e & 3;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
                        store_quad<SPLIT>(hi, lo, static_cast<uint32_t>((dx * 2 + (qq >> 1)) * U_PLANE + (p + dx) * 16 + (qq & 1) * 8), v[i]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) t_mbar_arrive(bar_yfull + 8 * slot);
            };
            stage_n2(y0 - 1, R++);
            stage_n2(y0, R++);
            for (int i = 0; i < nrows; ++i, ++j) {
                stage_n2(y0 + i + 1, R++);
                stage_dy(y0 + i, j);
            }
        }
    } else if (warp == 8) {
        // ======================================================= MMA ISSUER
        const bool leader = elect_one();
        const uint32_t idesc = t_idesc(128, U_N);
        const uint64_t dA_hi = make_mn_nosw_desc(smem_u32(ring_hi), 128, U_PLANE);
        const uint64_t dA_lo = make_mn_nosw_desc(smem_u32(ring_lo), 128, U_PLANE);
        const uint64_t dY0 = make_mn_nosw_desc(smem_u32(ybuf), 128, U_PLANE);
        uint32_t R0 = 0, j = 0, used = 0;
        for (long unit = blockIdx.x; unit < a.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            const int ksteps = (min(T_TILE, a.W - x0) + 2 + 15) / 16;          // ring pixels 0 .. width+1 carry data
            for (int r = 0; r < 2; ++r) mbar_wait(bar_rfull + 8 * ((R0 + r) % T_RING), ((R0 + r) / T_RING) & 1);
            for (int i = 0; i < nrows; ++i, ++j) {
                const uint32_t Rn = R0 + i + 2;
                mbar_wait(bar_rfull + 8 * (Rn % T_RING), (Rn / T_RING) & 1);
                mbar_wait(bar_yfull + 8 * (j & 1), (j >> 1) & 1);
                tc_fence_after();
                int pA, tA, pB, tB;
                u_pair((R0 + i) % T_RING, pA, tA, pB, tB);
                if (leader) {
                    const uint64_t dy_hi = dY0 + static_cast<uint64_t>((j & 1) * (NIMG * U_YBUF >> 4));
                    const uint64_t dy_lo = dy_hi + (U_YBUF >> 4);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int pl = h ? pB : pA, t = h ? tB : tA;
                        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(t * U_N);
                        const uint32_t first = (used >> t) & 1u;
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t aoff = static_cast<uint32_t>(pl * (U_PLANE / 16) + k * 16);   // +16 ring pixels per k-step
                            const uint64_t yoff = static_cast<uint32_t>(k * 16);
                            umma_bf16(d_tmem, dA_hi + aoff, dy_hi + yoff, idesc, (k == 0) ? first : 1u);
                            if (SPLIT) {
                                umma_bf16(d_tmem, dA_lo + aoff, dy_hi + yoff, idesc, 1u);
                                umma_bf16(d_tmem, dA_hi + aoff, dy_lo + yoff, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(bar_rempty + 8 * ((R0 + i) % T_RING));
                    if (i == nrows - 1) {
                        umma_commit(bar_rempty + 8 * ((R0 + i + 1) % T_RING));
                        umma_commit(bar_rempty + 8 * ((R0 + i + 2) % T_RING));
                    }
                    umma_commit(bar_yempty + 8 * (j & 1));
                }
                used |= (1u << tA) | (1u << tB);
                __syncwarp();
            }
            R0 += static_cast<uint32_t>(nrows + 2);
        }
        if (leader) umma_commit(bar_done);
        __syncwarp();
    } else {
        // ======================================================= EPILOGUE (once)
        // which tiles were written is a function of the unit list alone: replay it
        uint32_t used = 0, R0 = 0;
        for (long unit = blockIdx.x; unit < a.nunits; unit += gridDim.x) {
            long b; int x0, y0, nrows;
            decode(unit, b, x0, y0, nrows);
            for (int i = 0; i < nrows && used != (1u << U_TILES) - 1u; ++i) {
                int pA, tA, pB, tB;
                u_pair((R0 + i) % T_RING, pA, tA, pB, tB);
                used |= (1u << tA) | (1u << tB);
            }
            R0 += static_cast<uint32_t>(nrows + 2);
        }
        const int q4 = warp & 3, et = tid - (T_THREADS - 128);
        mbar_wait(bar_done, 0);
        __syncwarp();
        tc_fence_after();
        // every MMA has completed and the producers are done: the ring becomes the (dy, c, dx*16+n) sum of the six tiles
        float *acc = reinterpret_cast<float *>(smem);
        for (int i = et; i < 3 * T_C * U_N; i += 128) acc[i] = 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int m = q4 * 32 + lane;
        for (int t = 0; t < U_TILES; ++t) {
            if (!((used >> t) & 1u)) continue;                                  // uniform over the CTA
            int dy, c;
            u_row(t, m, dy, c);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float v[16];
                tmem_ld16(tmem_base + static_cast<uint32_t>(t * U_N + g * 16) + (static_cast<uint32_t>(q4 * 32) << 16), v);
                if (dy >= 0) {
                    float *dst = acc + (dy * T_C + c) * U_N + g * 16;
#pragma unroll
                    for (int n = 0; n < 16; ++n) dst[n] += v[n];
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        for (int i = et; i < 3 * T_C * 3 * a.N; i += 128) {
            const int n = i % a.N, dx = (i / a.N) % 3, c = (i / (3 * a.N)) % T_C, dy = i / (3 * a.N * T_C);
            atomicAdd(a.dW + (static_cast<long>(n) * T_C + c) * 9 + dy * 3 + dx, acc[(dy * T_C + c) * U_N + dx * 16 + n]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

bool eml_wgrad3x3_tc_supported(int N, int C) { return N <= 16 && C == T_C && !eml_env_flag("EML_WGRAD_SIMT"); }

int eml_wgrad3x3_tc(const float *dY, int dy_pitch, int N, const float *b, int b_pitch, const float *scale, const float *shift,
                    float *dW, int B, int H, int W, int precision, cudaStream_t st) {
    TArgs a{};
    a.dY = dY; a.dy_pitch = dy_pitch; a.N = N; a.b = b; a.b_pitch = b_pitch; a.scale = scale; a.shift = shift; a.dW = dW;
    a.B = B; a.H = H; a.W = W;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles_x = (W + T_TILE - 1) / T_TILE;
    int rb = 32;
    if (static_cast<long>(B) * tiles_x * ((H + rb - 1) / rb) < 8L * sms) rb = 16;
    if (rb > H) rb = H;
    a.rb = rb;
    a.bands_y = (H + rb - 1) / rb;
    a.nunits = static_cast<long>(B) * tiles_x * a.bands_y;
    const bool split = precision != EML_PREC_BF16;
    const unsigned grid = static_cast<unsigned>(a.nunits < sms ? a.nunits : sms);
    cudaError_t e;
    if (!eml_env_flag("EML_WGRAD3X3_V1")) {
        const size_t smem2 = static_cast<size_t>(split ? 2 : 1) * (U_IMG + 2 * U_YBUF) + U_PAD;
        if (split) {
            e = cudaFuncSetAttribute(wgrad3x3_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
            if (e != cudaSuccess) return static_cast<int>(e);
            wgrad3x3_tc2_kernel<true><<<grid, T_THREADS, smem2, st>>>(a);
        } else {
            e = cudaFuncSetAttribute(wgrad3x3_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
            if (e != cudaSuccess) return static_cast<int>(e);
            wgrad3x3_tc2_kernel<false><<<grid, T_THREADS, smem2, st>>>(a);
        }
        return eml_launch_status();
    }
    const size_t smem = static_cast<size_t>(split ? 2 : 1) * T_IMG + T_PAD + static_cast<size_t>(2) * (split ? 2 : 1) * T_YBUF;
    if (split) {
        e = cudaFuncSetAttribute(wgrad3x3_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        wgrad3x3_tc_kernel<true><<<grid, T_THREADS, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(wgrad3x3_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        wgrad3x3_tc_kernel<false><<<grid, T_THREADS, smem, st>>>(a);
    }
    return eml_launch_status();
}
