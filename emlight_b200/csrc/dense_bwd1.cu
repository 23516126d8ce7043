// Fused backward of a dense layer's first half -- norm1 -> relu1 -> conv1 (1x1, C_in -> 48) of RegressionNetwork/DenseNet.py:30-37 --
// with respect to the dense block's concatenation slab, training-mode BatchNorm, sm_100a.
//
// What it replaces (per layer, round 1): dA = conv1x1(dN, W1^T) written to HBM (C_in floats per pixel), eml_bn_bwd_reduce (reads dA and
// the slab), eml_bn_bwd_apply (reads dA, the slab and the gradient slab dS, writes dS): 7 C_in + 48 floats of HBM traffic per pixel.
// Here dA never exists in memory and the slab / dS are touched once: 3 C_in + 48 floats per pixel.
//
// Algebra.  With u = pre_a x + pre_b (the folded last_norm in front of blocks 2, 3), xhat = (u - mean) inv = e x + f, z = gamma xhat +
// beta = sc x + sh, a = relu(z), and g = dA * [z > 0], BatchNorm's backward is
//     du = gamma inv (g - S1/n - xhat S2/n),      S1 = sum_p g,  S2 = sum_p g xhat      (= d beta, d gamma).
// The first term depends on the pixel's own mask and is accumulated into dS right here:  dS[p, c] += k1[c] g[p, c],  k1 = gamma inv.
// The other two are an AFFINE function of the stored x with per-channel coefficients,  -(k1/n)(S1 + f S2) - (k1/n) e S2 x[p, c],  and a
// channel's coefficients can be summed over all the layers that consume it: they are applied ONCE per channel, just before that
// channel's gradient is read (eml_dense_bwd1_gather for the 12 channels a layer produced, eml_dense_bwd1_correct for the block input).
//
// Mapping (one persistent CTA per SM, tile = 128 consecutive pixels, stage = 32 channels of the tile):
//   warp 0      TMA      per tile one bulk copy of the dN tile (128 x 48 fp32, contiguous); per stage two tensor-map boxes (32 ch x 128 px
//                        of the slab and of dS, SWIZZLE_128B) into a ring slot
//   warp 1      MMA      D[128 px, 32 ch] = dN[128, 48] W1[32 ch, 48]^T: A = dN as bf16 hi/lo in TENSOR MEMORY (written once per tile by the
//                        converter warpgroup), B = the layer's W1 resident in shared memory; 3 k-steps x 3 products (bf16x3)
//   warps 4-11  EPILOGUE two warpgroups alternate stages; thread = pixel: tcgen05.ld D, read its slab / dS rows (conflict-free through
//                        the swizzle), mask, dS += k1 g written back into the ring slot, per-channel sums by a transpose-reduce over the
//                        warp (31 shuffles per 32 channels) into shared-memory accumulators; one thread stores the slot with a TMA store
//   warps 12-15 CONVERT  dN tile fp32 (shared) -> bf16 hi/lo (TMEM A slot, double buffered)
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int B1_TILE = 128;
constexpr int B1_SC = 32;                       // channels per stage
constexpr int B1_XBYTES = B1_TILE * B1_SC * 4;  // 16 KB (one box)
constexpr int B1_STAGE = 2 * B1_XBYTES;         // slab box + dS box
constexpr int B1_NB = 48;                       // bottleneck channels (K of the GEMM)
constexpr int B1_DNBYTES = B1_TILE * B1_NB * 4; // 24 KB
constexpr int B1_THREADS = 16 * 32;
constexpr int B1_MAXC = 352;
constexpr int B1_MAXST = 4;
constexpr int B1_ND = 4;                        // D accumulator slots (32 columns each)
constexpr int B1_ACOL = 0, B1_DCOL = 96;        // TMEM: 2 A slots x 48 columns, then 4 D slots x 32 columns

struct B1Args {
    const float *dN;
    const unsigned char *wpack;                 // [hi plane | lo plane], each Cpad rows x 128 B (64 bf16, K-major SWIZZLE_128B; k < 48 used)
    const float *vec;                           // 5 x vstride: sc, sh, e, f, k1
    double *sums; long sums_stride;
    int C_in, Cpad, nstg, stages, ndn, vstride;
    long ntiles;
};

__device__ __forceinline__ void b1_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void b1_tma_load(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void b1_tma_store(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void b1_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void b1_tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ uint32_t b1_pack_bf16(float lo_elem, float hi_elem) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t *>(&h);
}
// Transpose-reduce over the warp: on return lane l holds sum over the 32 lanes of v[l].  31 shuffles.
__device__ __forceinline__ float b1_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <bool SPLIT>
__global__ void __launch_bounds__(B1_THREADS, 1) dense_bwd1_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                  const __grid_constant__ CUtensorMap tm_ds, const B1Args a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2 * B1_MAXST + 2 + 2 + 2 + 2 + 2 * B1_ND + 1];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_vec[5][B1_MAXC];
    __shared__ float s_sum[2][B1_MAXC];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int NST = a.stages, NDN = a.ndn;
    unsigned char *dn_sm = smem + static_cast<size_t>(NST) * B1_STAGE;
    unsigned char *w_sm = dn_sm + static_cast<size_t>(NDN) * B1_DNBYTES;          // 1024-aligned: every piece before it is a multiple of 8 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grid = static_cast<int>(gridDim.x);

    const uint32_t bar_full = smem_u32(&s_bar[0]);                       // [NST]  slab + dS boxes landed
    const uint32_t bar_empty = smem_u32(&s_bar[B1_MAXST]);               // [NST]  slot stored back and free
    const uint32_t bar_dnfull = smem_u32(&s_bar[2 * B1_MAXST]);          // [2]    dN tile landed in shared memory
    const uint32_t bar_dnempty = smem_u32(&s_bar[2 * B1_MAXST + 2]);     // [2]    ... and has been converted
    const uint32_t bar_aready = smem_u32(&s_bar[2 * B1_MAXST + 4]);      // [2]    TMEM A slot written
    const uint32_t bar_afree = smem_u32(&s_bar[2 * B1_MAXST + 6]);       // [2]    the tile's MMAs are done with it
    const uint32_t bar_dfull = smem_u32(&s_bar[2 * B1_MAXST + 8]);       // [ND]   D slot computed
    const uint32_t bar_dempty = smem_u32(&s_bar[2 * B1_MAXST + 8 + B1_ND]);   // [ND] D slot read by the epilogue
    const uint32_t bar_w = smem_u32(&s_bar[2 * B1_MAXST + 8 + 2 * B1_ND]);

    if (tid == 0) {
        for (int s = 0; s < B1_MAXST; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_dnfull + 8 * i, 1); mbar_init(bar_dnempty + 8 * i, 4);
            mbar_init(bar_aready + 8 * i, 4); mbar_init(bar_afree + 8 * i, 1);
        }
        for (int i = 0; i < B1_ND; ++i) { mbar_init(bar_dfull + 8 * i, 1); mbar_init(bar_dempty + 8 * i, 4); }
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 5 * B1_MAXC; i += B1_THREADS) {
        const int r = i / B1_MAXC, c = i - r * B1_MAXC;
        s_vec[r][c] = c < a.C_in ? a.vec[static_cast<long>(r) * a.vstride + c] : 0.f;    // channels past C_in: sc = sh = 0 -> mask false, g = 0
    }
    for (int i = tid; i < 2 * B1_MAXC; i += B1_THREADS) s_sum[i / B1_MAXC][i % B1_MAXC] = 0.f;
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 256);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // =========================================================== TMA PRODUCER
        if (lane == 0) {
            uint32_t g = 0, jt = 0;
            const uint32_t nst_u = static_cast<uint32_t>(NST), ndn_u = static_cast<uint32_t>(NDN);
            for (long t = blockIdx.x; t < a.ntiles; t += grid, ++jt) {
                const uint32_t db = jt % ndn_u, dph = (jt / ndn_u) & 1;
                mbar_wait(bar_dnempty + 8 * db, dph ^ 1);
                mbar_expect_tx(bar_dnfull + 8 * db, B1_DNBYTES);
                bulk_g2s(smem_u32(dn_sm + static_cast<size_t>(db) * B1_DNBYTES), a.dN + t * (B1_TILE * B1_NB), B1_DNBYTES, bar_dnfull + 8 * db);
                const int m0 = static_cast<int>(t * B1_TILE);
                for (int j = 0; j < a.nstg; ++j, ++g) {
                    const uint32_t s = g % nst_u, ph = (g / nst_u) & 1;
                    const uint32_t dst = smem_u32(smem + static_cast<size_t>(s) * B1_STAGE);
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_expect_tx(bar_full + 8 * s, B1_STAGE);
                    b1_tma_load(dst, &tm_x, j * B1_SC, m0, bar_full + 8 * s);
                    b1_tma_load(dst + B1_XBYTES, &tm_ds, j * B1_SC, m0, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA ISSUER
        const bool leader = elect_one();
        if (leader) {
            const uint32_t bytes = static_cast<uint32_t>(a.Cpad) * 128u * (SPLIT ? 2u : 1u);
            mbar_expect_tx(bar_w, bytes);
            bulk_g2s(smem_u32(w_sm), a.wpack, bytes, bar_w);
        }
        mbar_wait(bar_w, 0);
        const uint32_t idesc = make_idesc_bf16(B1_TILE, B1_SC);
        const uint64_t dB0 = make_sw128_desc(smem_u32(w_sm));
        const uint32_t blo16 = (static_cast<uint32_t>(a.Cpad) * 128u) >> 4;
        uint32_t g = 0, jt = 0;
        for (long t = blockIdx.x; t < a.ntiles; t += grid, ++jt) {
            const uint32_t da = jt & 1, aph = (jt >> 1) & 1;
            mbar_wait(bar_aready + 8 * da, aph);
            tc_fence_after();
            const uint32_t ta = tmem_base + B1_ACOL + da * 48u;
            for (int j = 0; j < a.nstg; ++j, ++g) {
                const uint32_t ds = g % B1_ND, dph = (g / B1_ND) & 1;
                mbar_wait(bar_dempty + 8 * ds, dph ^ 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t td = tmem_base + B1_DCOL + ds * 32u;
                    const uint64_t db_hi = dB0 + static_cast<uint64_t>(static_cast<uint32_t>(j) * ((B1_SC * 128u) >> 4));
                    const uint64_t db_lo = db_hi + blo16;
#pragma unroll
                    for (int k = 0; k < B1_NB / 16; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        b1_umma_ts(td, ta + 8 * k, db_hi + adv, idesc, k != 0 ? 1u : 0u);
                        if (SPLIT) {
                            b1_umma_ts(td, ta + 24 + 8 * k, db_hi + adv, idesc, 1u);
                            b1_umma_ts(td, ta + 8 * k, db_lo + adv, idesc, 1u);
                        }
                    }
                    umma_commit(bar_dfull + 8 * ds);
                    if (j == a.nstg - 1) umma_commit(bar_afree + 8 * da);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 12) {
        // =========================================================== dN CONVERTER: fp32 tile (shared) -> bf16 hi/lo in the TMEM A slot
        const int w4 = warp & 3;
        const uint32_t row = static_cast<uint32_t>(w4 * 32 + lane);
        const uint32_t ndn_u = static_cast<uint32_t>(NDN);
        uint32_t jt = 0;
        for (long t = blockIdx.x; t < a.ntiles; t += grid, ++jt) {
            const uint32_t db = jt % ndn_u, dph = (jt / ndn_u) & 1, da = jt & 1, aph = (jt >> 1) & 1;
            const uint32_t src = smem_u32(dn_sm + static_cast<size_t>(db) * B1_DNBYTES) + row * (B1_NB * 4);
            mbar_wait(bar_dnfull + 8 * db, dph);
            uint32_t hi[3][8], lo[3][8];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // (rows are 192 B apart: a quarter-warp's 16-byte loads hit two bank groups, 4-way conflicts -- 48 of ~400 wavefronts per tile)
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"(src + static_cast<uint32_t>((4 * k + q) * 16)) : "memory");
                    const uint32_t h0 = b1_pack_bf16(v.x, v.y), h1 = b1_pack_bf16(v.z, v.w);
                    hi[k][2 * q] = h0; hi[k][2 * q + 1] = h1;
                    if (SPLIT) {
                        lo[k][2 * q] = b1_pack_bf16(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u));
                        lo[k][2 * q + 1] = b1_pack_bf16(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) b1_arrive(bar_dnempty + 8 * db);                  // the shared-memory tile has been read
            mbar_wait(bar_afree + 8 * da, aph ^ 1);                          // MMAs of tile jt - 2 are done with this A slot
            tc_fence_after();
            const uint32_t ta = tmem_base + B1_ACOL + da * 48u + (static_cast<uint32_t>(w4 * 32) << 16);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                b1_tmem_st8(ta + 8 * k, hi[k]);
                if (SPLIT) b1_tmem_st8(ta + 24 + 8 * k, lo[k]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) b1_arrive(bar_aready + 8 * da);
        }
    } else if (warp >= 4) {
        // =========================================================== EPILOGUE: warpgroup wg takes stages g = wg (mod 2); thread = pixel
        const int wg = (warp - 4) >> 2, w4 = warp & 3;
        const uint32_t row = static_cast<uint32_t>(w4 * 32 + lane);
        const uint32_t r7 = row & 7u;
        const uint32_t lane_addr = tmem_base + B1_DCOL + (static_cast<uint32_t>(w4 * 32) << 16);
        const uint32_t nst_u = static_cast<uint32_t>(NST);
        uint32_t total = 0;
        for (long t = blockIdx.x; t < a.ntiles; t += grid) total += static_cast<uint32_t>(a.nstg);
        const uint32_t nstg_u = static_cast<uint32_t>(a.nstg);
        for (uint32_t g = static_cast<uint32_t>(wg); g < total; g += 2) {
            const uint32_t s = g % nst_u, ph = (g / nst_u) & 1, ds = g % B1_ND, dph = (g / B1_ND) & 1;
            const int j = static_cast<int>(g % nstg_u);
            const long t = blockIdx.x + static_cast<long>(g / nstg_u) * grid;
            const uint32_t xs = smem_u32(smem + static_cast<size_t>(s) * B1_STAGE) + row * 128u;
            const uint32_t dss = xs + B1_XBYTES;
            mbar_wait(bar_dfull + 8 * ds, dph);
            tc_fence_after();
            float g1[32], g2[32];
            {
                float d0[16], d1[16];
                tmem_ld16(lane_addr + ds * 32u, d0);
                tmem_ld16(lane_addr + ds * 32u + 16u, d1);
#pragma unroll
                for (int e = 0; e < 16; ++e) { g1[e] = d0[e]; g1[16 + e] = d1[e]; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) b1_arrive(bar_dempty + 8 * ds);                   // D slot drained: the MMA warp may refill it
            mbar_wait(bar_full + 8 * s, ph);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t off = (static_cast<uint32_t>(q) ^ r7) << 4;
                float4 xv, dv;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xv.x), "=f"(xv.y), "=f"(xv.z), "=f"(xv.w) : "r"(xs + off) : "memory");
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dv.x), "=f"(dv.y), "=f"(dv.z), "=f"(dv.w) : "r"(dss + off) : "memory");
                const int c = j * B1_SC + 4 * q;
                const float4 sc = *reinterpret_cast<const float4 *>(&s_vec[0][c]), sh = *reinterpret_cast<const float4 *>(&s_vec[1][c]);
                const float4 ee = *reinterpret_cast<const float4 *>(&s_vec[2][c]), ff = *reinterpret_cast<const float4 *>(&s_vec[3][c]);
                const float4 kk = *reinterpret_cast<const float4 *>(&s_vec[4][c]);
                const float x4[4] = {xv.x, xv.y, xv.z, xv.w}, s4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
                const float e4[4] = {ee.x, ee.y, ee.z, ee.w}, f4[4] = {ff.x, ff.y, ff.z, ff.w}, k4[4] = {kk.x, kk.y, kk.z, kk.w};
                float o4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float gg = fmaf(x4[e], s4[e], h4[e]) > 0.f ? g1[4 * q + e] : 0.f;      // same fmaf as the forward's relu(sc x + sh)
                    g1[4 * q + e] = gg;
                    g2[4 * q + e] = gg * fmaf(x4[e], e4[e], f4[e]);
                    o4[e] = fmaf(k4[e], gg, o4[e]);
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dss + off), "f"(o4[0]), "f"(o4[1]), "f"(o4[2]), "f"(o4[3]) : "memory");
            }
            fence_proxy_async();                                             // the updated dS rows are about to be read by the TMA store
            const float t1 = b1_reduce32(g1, lane);
            const float t2 = b1_reduce32(g2, lane);
            atomicAdd(&s_sum[0][j * B1_SC + lane], t1);
            atomicAdd(&s_sum[1][j * B1_SC + lane], t2);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");      // all 128 rows of the slot are written
            if ((tid & 127) == 0) {
                b1_tma_store(&tm_ds, smem_u32(smem + static_cast<size_t>(s) * B1_STAGE) + B1_XBYTES, j * B1_SC, static_cast<int>(t * B1_TILE));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the store has read the slot: it may be refilled
                b1_arrive(bar_empty + 8 * s);
            }
        }
        if ((tid & 127) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the kernel ends
    }
    tc_fence_before();
    __syncthreads();
    for (int c = tid; c < a.C_in; c += B1_THREADS) {
        atomicAdd(a.sums + c, static_cast<double>(s_sum[0][c]));
        atomicAdd(a.sums + a.sums_stride + c, static_cast<double>(s_sum[1][c]));
    }
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn b1_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
int b1_make_map(CUtensorMap *map, const float *base, int C, long M, int pitch) {
    EncodeTiledFn enc = b1_get_encode();
    if (enc == nullptr) return EML_E_ARG;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(M)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch) * 4};
    const cuuint32_t box[2] = {B1_SC, B1_TILE};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? EML_OK : EML_E_ARG;
}

// W1 (48, C_in) fp32 -> [hi plane | lo plane] of Cpad rows (row = input channel c) x 64 bf16 (k = bottleneck channel, zero past 48)
__global__ void dense_bwd1_pack_kernel(const float *__restrict__ w1, unsigned char *__restrict__ out, int C_in, int Cpad) {
    const long total = static_cast<long>(Cpad) * 64;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx & 63), c = static_cast<int>(idx >> 6);
        const float v = (k < B1_NB && c < C_in) ? w1[static_cast<long>(k) * C_in + c] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const uint32_t off = sw128_offset(c, k);
        *reinterpret_cast<__nv_bfloat16 *>(out + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(out + static_cast<size_t>(Cpad) * 128 + off) = lo;
    }
}

// vec (5, vstride): sc, sh (the forward's folded affine, copied), e = pre_a inv, f = (pre_b - mean) inv, k1 = gamma inv
__global__ void dense_bwd1_prep_kernel(const float *sc, const float *sh, const float *pre_a, const float *pre_b, const float *mean,
                                       const float *inv, const float *gamma, int C, int vstride, float *vec) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float pa = pre_a ? pre_a[c] : 1.f, pb = pre_b ? pre_b[c] : 0.f;
    vec[c] = sc[c];
    vec[vstride + c] = sh[c];
    vec[2 * vstride + c] = pa * inv[c];
    vec[3 * vstride + c] = (pb - mean[c]) * inv[c];
    vec[4 * vstride + c] = gamma[c] * inv[c];
}

// After a layer's fused pass: d gamma = S2, d beta = S1; the deferred affine terms of every channel the layer read
//   coefA[c] -= k1/n (S1 + f S2),  coefB[c] -= k1/n e S2
__global__ void dense_bwd1_accum_kernel(const double *sums, long stride, const float *vec, int vstride, double n, int C, float *coefA,
                                        float *coefB, float *dgamma, float *dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double s1 = sums[c], s2 = sums[stride + c];
    const double e = vec[2 * vstride + c], f = vec[3 * vstride + c], k1 = vec[4 * vstride + c];
    dgamma[c] = static_cast<float>(s2);
    dbeta[c] = static_cast<float>(s1);
    coefA[c] -= static_cast<float>(k1 / n * (s1 + f * s2));
    coefB[c] -= static_cast<float>(k1 / n * e * s2);
}

// out[m, 0..n) = dS[m, c0 + i] + coefA[c0 + i] + coefB[c0 + i] * x[m, c0 + i]; out[m, n..out_pitch) = 0   (out == dS + c0 allowed when
// out_pitch == ds_pitch and in_place: the correction of the block-input channels)
__global__ void __launch_bounds__(256) dense_bwd1_gather_kernel(const float *__restrict__ dS, int ds_pitch, const float *__restrict__ x, int x_pitch,
                                                                const float *__restrict__ coefA, const float *__restrict__ coefB, int c0, int n,
                                                                float *out, int out_pitch, int out_n, long M) {
    const long total = M * out_n;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const long m = idx / out_n;
        const int i = static_cast<int>(idx - m * out_n);
        float v = 0.f;
        if (i < n) {
            const int c = c0 + i;
            v = dS[m * ds_pitch + c] + coefA[c] + coefB[c] * x[m * x_pitch + c];
        }
        out[m * out_pitch + i] = v;
    }
}

// float4 variant: c0, n, out_n, every pitch multiples of 4 and 16-byte aligned bases -- thread = (pixel, channel quad)
__global__ void __launch_bounds__(256) dense_bwd1_gather4_kernel(const float *__restrict__ dS, int ds_pitch, const float *__restrict__ x, int x_pitch,
                                                                 const float *__restrict__ coefA, const float *__restrict__ coefB, int c0, int n,
                                                                 float *out, int out_pitch, int out_n, long M) {
    const int Q = out_n >> 2;
    const long total = M * Q;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const long m = idx / Q;
        const int i = static_cast<int>(idx - m * Q) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            const int c = c0 + i;
            const float4 d = *reinterpret_cast<const float4 *>(dS + m * ds_pitch + c);
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + m * x_pitch + c));
            const float4 a = *reinterpret_cast<const float4 *>(coefA + c), b = *reinterpret_cast<const float4 *>(coefB + c);
            v = make_float4(fmaf(b.x, xv.x, d.x + a.x), fmaf(b.y, xv.y, d.y + a.y), fmaf(b.z, xv.z, d.z + a.z), fmaf(b.w, xv.w, d.w + a.w));
        }
        *reinterpret_cast<float4 *>(out + m * out_pitch + i) = v;
    }
}

}  // namespace

extern "C" size_t eml_dense_bwd1_wpack_bytes(int C_in) { return static_cast<size_t>((C_in + 31) & ~31) * 256; }

extern "C" int eml_dense_bwd1_supported(int C_in, long M, int precision) {
    if (C_in <= 0 || C_in > B1_MAXC || M <= 0 || (M % B1_TILE) != 0) return 0;
    if (precision != EML_PREC_BF16 && precision != EML_PREC_BF16X3) return 0;
    return eml_env_flag("EML_NO_FUSED_BWD1") ? 0 : 1;
}

extern "C" int eml_dense_bwd1_pack(const float *w1, void *wpack, int C_in, void *stream) {
    EML_CHECK_PTR(w1); EML_CHECK_PTR(wpack); EML_CHECK_ALIGN16(wpack);
    if (C_in <= 0 || C_in > B1_MAXC) return EML_E_SHAPE;
    dense_bwd1_pack_kernel<<<32, 256, 0, static_cast<cudaStream_t>(stream)>>>(w1, static_cast<unsigned char *>(wpack), C_in, (C_in + 31) & ~31);
    return eml_launch_status();
}

extern "C" int eml_dense_bwd1_prep(const float *sc, const float *sh, const float *pre_a, const float *pre_b, const float *mean,
                                   const float *inv_std, const float *gamma, int C, int vstride, float *vec, void *stream) {
    EML_CHECK_PTR(sc); EML_CHECK_PTR(sh); EML_CHECK_PTR(mean); EML_CHECK_PTR(inv_std); EML_CHECK_PTR(gamma); EML_CHECK_PTR(vec);
    if (C <= 0 || vstride < C) return EML_E_SHAPE;
    dense_bwd1_prep_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sc, sh, pre_a, pre_b, mean, inv_std, gamma, C, vstride, vec);
    return eml_launch_status();
}

extern "C" int eml_dense_bwd1_accum(const double *sums, long sums_stride, const float *vec, int vstride, double n, int C, float *coefA,
                                    float *coefB, float *dgamma, float *dbeta, void *stream) {
    EML_CHECK_PTR(sums); EML_CHECK_PTR(vec); EML_CHECK_PTR(coefA); EML_CHECK_PTR(coefB); EML_CHECK_PTR(dgamma); EML_CHECK_PTR(dbeta);
    if (C <= 0 || n <= 0) return EML_E_SHAPE;
    dense_bwd1_accum_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, sums_stride, vec, vstride, n, C, coefA, coefB, dgamma, dbeta);
    return eml_launch_status();
}

extern "C" int eml_dense_bwd1_gather(const float *dS, int ds_pitch, const float *x, int x_pitch, const float *coefA, const float *coefB,
                                     int c0, int n, float *out, int out_pitch, int out_n, long M, void *stream) {
    EML_CHECK_PTR(dS); EML_CHECK_PTR(x); EML_CHECK_PTR(coefA); EML_CHECK_PTR(coefB); EML_CHECK_PTR(out);
    if (n <= 0 || out_n < n || out_pitch < out_n || M <= 0 || c0 < 0) return EML_E_SHAPE;
    const bool vec = ((c0 | n | out_n | ds_pitch | x_pitch | out_pitch) & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(dS) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(coefA) | reinterpret_cast<uintptr_t>(coefB)) & 15) == 0;
    if (vec) {
        long blocks4 = (M * (out_n / 4) + 255) / 256;
        if (blocks4 > 148 * 16) blocks4 = 148 * 16;
        dense_bwd1_gather4_kernel<<<static_cast<unsigned>(blocks4), 256, 0, static_cast<cudaStream_t>(stream)>>>(dS, ds_pitch, x, x_pitch, coefA, coefB, c0, n,
                                                                                                              out, out_pitch, out_n, M);
        return eml_launch_status();
    }
    long blocks = (M * out_n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dense_bwd1_gather_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(dS, ds_pitch, x, x_pitch, coefA, coefB, c0, n,
                                                                                                         out, out_pitch, out_n, M);
    return eml_launch_status();
}

extern "C" int eml_dense_bwd1(const float *dN, const float *x, int x_pitch, float *dS, int ds_pitch, const void *wpack, const float *vec,
                              int vstride, int C_in, long M, double *sums, long sums_stride, int precision, void *stream) {
    EML_CHECK_PTR(dN); EML_CHECK_PTR(x); EML_CHECK_PTR(dS); EML_CHECK_PTR(wpack); EML_CHECK_PTR(vec); EML_CHECK_PTR(sums);
    EML_CHECK_ALIGN16(dN); EML_CHECK_ALIGN16(x); EML_CHECK_ALIGN16(dS); EML_CHECK_ALIGN16(wpack);
    if (!eml_dense_bwd1_supported(C_in, M, precision) || (x_pitch & 3) || (ds_pitch & 3) || x_pitch < C_in || ds_pitch < ((C_in + 3) & ~3) || vstride < C_in)
        return EML_E_SHAPE;
    CUtensorMap tm_x, tm_ds;
    int rc = b1_make_map(&tm_x, x, C_in, M, x_pitch);
    if (rc != EML_OK) return rc;
    // the TMA store clips at 16-byte granularity: with C_in = 2 (mod 4) -- block 3 -- it would zero the two channels behind C_in.  The dS
    // map therefore covers whole channel quads: the extra channels are loaded with their true values, left unchanged (k1 = 0) and stored back
    rc = b1_make_map(&tm_ds, dS, (C_in + 3) & ~3, M, ds_pitch);
    if (rc != EML_OK) return rc;
    const bool split = precision == EML_PREC_BF16X3;
    B1Args a{};
    a.dN = dN; a.wpack = static_cast<const unsigned char *>(wpack); a.vec = vec; a.vstride = vstride;
    a.sums = sums; a.sums_stride = sums_stride;
    a.C_in = C_in; a.Cpad = (C_in + 31) & ~31; a.nstg = a.Cpad / B1_SC;
    a.ntiles = M / B1_TILE;
    const size_t wbytes = static_cast<size_t>(a.Cpad) * 256;
    const size_t budget = 227 * 1024 - 12 * 1024;                   // static shared memory: vectors (7 KB), sums (2.8 KB), barriers
    int stages = B1_MAXST, ndn = 2;
    // ring depth even: a ring slot is then always handled by the same epilogue warpgroup (stages alternate), i.e. every phase of its
    // mbarriers has one in-order waiter (a parity wait cannot tell "two completions early" from "done")
    while (stages > 2 && static_cast<size_t>(stages) * B1_STAGE + static_cast<size_t>(ndn) * B1_DNBYTES + wbytes + 1024 > budget) stages -= 2;
    if (static_cast<size_t>(stages) * B1_STAGE + static_cast<size_t>(ndn) * B1_DNBYTES + wbytes + 1024 > budget) ndn = 1;
    if (static_cast<size_t>(stages) * B1_STAGE + static_cast<size_t>(ndn) * B1_DNBYTES + wbytes + 1024 > budget) return EML_E_SHAPE;
    a.stages = stages; a.ndn = ndn;
    const size_t smem = static_cast<size_t>(stages) * B1_STAGE + static_cast<size_t>(ndn) * B1_DNBYTES + wbytes + 1024;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = static_cast<unsigned>(a.ntiles < sms ? a.ntiles : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(dense_bwd1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        dense_bwd1_kernel<true><<<grid, B1_THREADS, smem, st>>>(tm_x, tm_ds, a);
    } else {
        e = cudaFuncSetAttribute(dense_bwd1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        dense_bwd1_kernel<false><<<grid, B1_THREADS, smem, st>>>(tm_x, tm_ds, a);
    }
    return eml_launch_status();
}
