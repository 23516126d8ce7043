// Persistent, warp-specialised 1x1 convolution (norm1 + relu1 + conv1 of RegressionNetwork/DenseNet.py:30-40), sm_100a.
//
// Same math and shared-memory layouts as conv_gemm_kernel<0,SPLIT> (conv_gemm.cu) -- that kernel stays as the
// general path (statistics epilogue, wide N, pooling) -- but organised so that nothing serialises per tile:
//   grid = one CTA per SM, each walking tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...
//   warps 0-15 PRODUCERS  gather 128 x 64 slab elements per K-chunk with coalesced float4 loads, apply the layer's
//                         BN affine + ReLU, split to bf16 hi/lo, store into a 4-stage SWIZZLE_128B ring (full/empty
//                         mbarriers); they run ahead across tile boundaries, so HBM loads are always in flight
//   warp  16   MMA        one lane issues tcgen05.mma against the RESIDENT weights (all K-chunks, <= 72 KB, one
//                         cp.async.bulk per CTA lifetime); tcgen05.commit frees ring stages / publishes accumulators
//   warps 17-20 EPILOGUE  tcgen05.ld the finished accumulator (TMEM double-buffered: 2 x 64 columns), stage it in shared
//                         memory as 16-byte planes and write it out with fully coalesced float4 stores (a thread's own
//                         row is 192 B at a 192..1376 B pitch: storing it directly costs 32 L1 wavefronts per instruction)
// Barrier-init, TMEM allocation and the weight fetch are paid once per SM instead of once per 128 pixels.
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int P_TILE_M = 128;
constexpr int P_CHUNK_K = 64;
constexpr int P_MAX_STAGES = 4;
constexpr int P_PRODUCERS = 512;                // warps 0..15 (4 per scheduler: latency hiding for the gather)
constexpr int P_PWARPS = P_PRODUCERS / 32;
constexpr int P_NROWS = P_TILE_M * 16 / P_PRODUCERS;   // rows per producer thread per chunk (4)
constexpr int P_THREADS = P_PRODUCERS + 32 + 128;   // + MMA warp + 4 epilogue warps
constexpr int P_A_TILE = P_TILE_M * P_CHUNK_K * 2;
constexpr int P_ACC_COLS = 128;                 // TMEM columns per accumulator buffer: [A_hi*B_hi | A_hi*B_lo] = 2 * N_pad <= 128

struct PArgs {
    const float *in;
    const float *scale;
    const float *shift;
    const unsigned char *wpack;
    float *out;
    long M;
    int C_in, in_pitch;
    int C_out, N_pad, out_pitch, out_choff;
    int nchunks;
    int relu;
    double *stats; long stats_stride;   // optional per-output-channel sum / sum of squares (batch-statistic BatchNorm downstream)
    int stages;          // A-ring depth (3 or 4, whatever fits next to the resident weights and the output staging)
    long ntiles;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float p_act(float x, float s, float t, int relu) {
    float v = fmaf(x, s, t);
    return relu ? fmaxf(v, 0.f) : v;
}

template <bool SPLIT, bool RELU>
__global__ void __launch_bounds__(P_THREADS, 1) conv1x1_persist_kernel(const PArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2 * P_MAX_STAGES + 1 + 4];
    __shared__ uint32_t s_tmem;
    __shared__ double s_stat[2][64];

    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * P_A_TILE;
    const int P_STAGES = a.stages;
    unsigned char *w_sm = smem + P_STAGES * STAGE_BYTES;
    const int b_tile_bytes = a.N_pad * 128;
    const int w_chunk_sm = (SPLIT ? 2 : 1) * b_tile_bytes;          // bytes per chunk kept in smem
    const size_t out_stage_off = static_cast<size_t>(P_STAGES) * STAGE_BYTES + static_cast<size_t>(a.nchunks) * w_chunk_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const uint32_t bar_full = smem_u32(&s_bar[0]);
    const uint32_t bar_empty = smem_u32(&s_bar[P_MAX_STAGES]);
    const uint32_t bar_w = smem_u32(&s_bar[2 * P_MAX_STAGES]);
    const uint32_t bar_accfull = smem_u32(&s_bar[2 * P_MAX_STAGES + 1]);    // [2]
    const uint32_t bar_accempty = smem_u32(&s_bar[2 * P_MAX_STAGES + 3]);   // [2]

    if (tid == 0) {
        for (int s = 0; s < P_MAX_STAGES; ++s) { mbar_init(bar_full + 8 * s, P_PRODUCERS / 32); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_w, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accempty + 8 * i, 4); }
        fence_mbar_init();
    }
    if (warp == P_PWARPS) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 2 * P_ACC_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp < P_PWARPS) {
        // =========================================================== PRODUCERS
        const int sub = tid & 15, rgrp = tid >> 4;                   // channel quad, row group (rows rgrp + 32 i, i < 4)
        const uint32_t st_off = static_cast<uint32_t>((rgrp >> 3) * 1024 + (rgrp & 7) * 128 + ((((sub >> 1) ^ rgrp) & 7) << 4) + (sub & 1) * 8);
        constexpr int RSTEP = P_PRODUCERS / 16;                      // 32 rows between a thread's consecutive rows
        constexpr uint32_t SSTEP = RSTEP / 8 * 1024;                 // = 4096 bytes in the swizzled tile
        constexpr unsigned FULLM = (1u << P_NROWS) - 1;
        // Register double-buffering: the loads of chunk g+1 are issued before chunk g is converted and stored, so every
        // producer thread keeps 8-16 float4 (128-256 B) in flight through the transform and the barrier waits.
        auto issue = [&](long tile, int c, float4 (&v)[P_NROWS]) -> unsigned {
            const long m0 = tile * P_TILE_M;
            const int c0 = c * P_CHUNK_K;
            const int ksteps = (min(P_CHUNK_K, a.C_in - c0) + 15) >> 4;
            const int ch = c0 + sub * 4;
            const bool lane_ok = sub * 4 < ksteps * 16 && ch < a.C_in;
            const float *rowp = a.in + (m0 + rgrp) * a.in_pitch + ch;
            const int nv = a.C_in - ch;
            unsigned m = 0;
#pragma unroll
            for (int i = 0; i < P_NROWS; ++i) {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane_ok && (m0 + rgrp + RSTEP * i) < a.M) {
                    const float *src = rowp + static_cast<long>(RSTEP * i) * a.in_pitch;
                    if (nv >= 4) v[i] = __ldg(reinterpret_cast<const float4 *>(src));
                    else {                                           // ragged last quad (C_in % 4 != 0): the channels behind C_in belong to
                        v[i].x = __ldg(src);                         // a layer that has not run yet -- never touch them (initcheck-clean)
                        if (nv > 1) v[i].y = __ldg(src + 1);
                        if (nv > 2) v[i].z = __ldg(src + 2);
                    }
                    m |= 1u << i;
                }
            }
            return m;
        };
        auto process = [&](int c, const float4 (&v)[P_NROWS], unsigned okm, int s) {
            const int c0 = c * P_CHUNK_K;
            const int ksteps = (min(P_CHUNK_K, a.C_in - c0) + 15) >> 4;
            if (sub * 4 >= ksteps * 16) return;                      // this lane's 8-byte slot is never read by the MMA
            const int ch = c0 + sub * 4;
            const int nvalid = min(4, a.C_in - ch);
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nvalid >= 4) {
                if (a.scale) sc = *reinterpret_cast<const float4 *>(a.scale + ch);
                if (a.shift) sh = *reinterpret_cast<const float4 *>(a.shift + ch);
            } else if (nvalid > 0) {
                if (a.scale) { sc.x = a.scale[ch]; if (nvalid > 1) sc.y = a.scale[ch + 1]; if (nvalid > 2) sc.z = a.scale[ch + 2]; }
                if (a.shift) { sh.x = a.shift[ch]; if (nvalid > 1) sh.y = a.shift[ch + 1]; if (nvalid > 2) sh.z = a.shift[ch + 2]; }
            }
            unsigned char *a_hi = smem + static_cast<size_t>(s) * STAGE_BYTES;
            unsigned char *a_lo = a_hi + P_A_TILE;
            if (nvalid >= 4 && okm == FULLM) {
                // fast path (all 8 rows inside M, all 4 channels inside C_in): no predicates, no selects
#pragma unroll
                for (int i = 0; i < P_NROWS; ++i) {
                    float4 o;
                    o.x = fmaf(v[i].x, sc.x, sh.x); o.y = fmaf(v[i].y, sc.y, sh.y);
                    o.z = fmaf(v[i].z, sc.z, sh.z); o.w = fmaf(v[i].w, sc.w, sh.w);
                    if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    store_quad<SPLIT>(a_hi, a_lo, st_off + i * SSTEP, o);
                }
                return;
            }
#pragma unroll
            for (int i = 0; i < P_NROWS; ++i) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((okm >> i) & 1u) {
                    o.x = p_act(v[i].x, sc.x, sh.x, RELU);
                    o.y = nvalid > 1 ? p_act(v[i].y, sc.y, sh.y, RELU) : 0.f;
                    o.z = nvalid > 2 ? p_act(v[i].z, sc.z, sh.z, RELU) : 0.f;
                    o.w = nvalid > 3 ? p_act(v[i].w, sc.w, sh.w, RELU) : 0.f;
                }
                store_quad<SPLIT>(a_hi, a_lo, st_off + i * SSTEP, o);
            }
        };
        float4 cur[P_NROWS], nxt[P_NROWS];
        long tile = blockIdx.x;
        int c = 0;
        bool have = tile < a.ntiles;
        unsigned curm = have ? issue(tile, c, cur) : 0u;
        uint32_t g = 0;
        while (have) {
            long ntile = tile;
            int nc = c + 1;
            if (nc == a.nchunks) { nc = 0; ntile += gridDim.x; }
            const bool nhave = ntile < a.ntiles;
            unsigned nxtm = 0;
            if (nhave) nxtm = issue(ntile, nc, nxt);
            const int s = g % P_STAGES;
            const uint32_t ph = (g / P_STAGES) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);                     // stage free (first lap passes immediately)
            process(c, cur, curm, s);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
#pragma unroll
            for (int i = 0; i < P_NROWS; ++i) cur[i] = nxt[i];
            curm = nxtm; tile = ntile; c = nc; have = nhave; ++g;
        }
    } else if (warp == P_PWARPS) {
        // =========================================================== MMA ISSUER (warp-uniform loops, one elected lane issues)
        const bool leader = elect_one();
        if (leader) {
            // resident weights: every K-chunk's packed image, one bulk copy (SPLIT) or one per chunk (hi halves only)
            if (SPLIT) {
                const uint32_t bytes = static_cast<uint32_t>(a.nchunks) * w_chunk_sm;
                mbar_expect_tx(bar_w, bytes);
                bulk_g2s(smem_u32(w_sm), a.wpack, bytes, bar_w);
            } else {
                mbar_expect_tx(bar_w, static_cast<uint32_t>(a.nchunks) * b_tile_bytes);
                for (int c = 0; c < a.nchunks; ++c)
                    bulk_g2s(smem_u32(w_sm + c * b_tile_bytes), a.wpack + static_cast<size_t>(c) * 2 * b_tile_bytes, b_tile_bytes, bar_w);
            }
        }
        mbar_wait(bar_w, 0);
        // bf16x3 with 2 MMAs per k-step: a chunk's packed weights [hi image | lo image] are 2*N_pad consecutive rows of ONE
        // SWIZZLE_128B tile (N_pad % 8 == 0), so A_hi x [B_hi;B_lo] is a single N = 2*N_pad instruction filling columns
        // [0,N_pad) with A_hi*B_hi and [N_pad,2*N_pad) with A_hi*B_lo; A_lo x B_hi (N = N_pad) adds into [0,N_pad);
        // the epilogue sums the two halves.  A_hi is read from shared memory once instead of twice.
        const uint32_t idesc = make_idesc_bf16(P_TILE_M, a.N_pad);
        const uint32_t idesc_cat = make_idesc_bf16(P_TILE_M, 2 * a.N_pad);
        const uint64_t dA0 = make_sw128_desc(smem_u32(smem));                 // stage 0, hi image
        const uint64_t dB0 = make_sw128_desc(smem_u32(w_sm));                 // chunk 0, hi image
        const uint32_t stage16 = STAGE_BYTES >> 4, alo16 = P_A_TILE >> 4;     // 16-byte units for the address field
        const uint32_t wchunk16 = static_cast<uint32_t>(w_chunk_sm) >> 4;
        uint32_t g = 0, j = 0;
        for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1, aph = (j >> 1) & 1;
            mbar_wait(bar_accempty + 8 * buf, aph ^ 1);                       // epilogue drained this accumulator
            const uint32_t d_tmem = tmem_base + buf * P_ACC_COLS;
            for (int c = 0; c < a.nchunks; ++c, ++g) {
                const uint32_t s = g % P_STAGES;
                const uint32_t ph = (g / P_STAGES) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                if (leader) {
                    const int kvalid = min(P_CHUNK_K, a.C_in - c * P_CHUNK_K);
                    const int ksteps = (kvalid + 15) >> 4;
                    const uint64_t da_hi = dA0 + static_cast<uint64_t>(s * stage16), da_lo = da_hi + alo16;
                    const uint64_t db_hi = dB0 + static_cast<uint64_t>(static_cast<uint32_t>(c) * wchunk16);
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        umma_bf16(d_tmem, da_hi + adv, db_hi + adv, SPLIT ? idesc_cat : idesc, (c | k) != 0 ? 1u : 0u);
                        if (SPLIT) umma_bf16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8 * s);
                    if (c == a.nchunks - 1) umma_commit(bar_accfull + 8 * buf);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================================== EPILOGUE (4 warps -> TMEM lane quarter warp % 4)
        const int q = warp & 3;
        const int et = tid - (P_PRODUCERS + 32);                       // 0..127 within the epilogue group
        const int nq = a.N_pad >> 2;                                   // float4 quads per staged row
        const bool direct = !(((a.out_pitch | a.out_choff) & 3) == 0 && (a.C_out & 3) == 0);
        float *stage = reinterpret_cast<float *>(smem + out_stage_off);      // [nq planes][129 x 16 B]
        constexpr int PLANE_F = 129 * 4;                               // floats per plane (one 16-byte pad slot)
        // statistics epilogue (round 2: the training forward's conv1): in the write-out every thread keeps ONE channel quad -- rows
        // are dealt round-robin in groups of 128 / cq -- so the per-channel sums live in 8 double registers over all the CTA's tiles
        const bool want_stats = a.stats != nullptr && !direct;
        const int cqs = a.C_out >> 2;
        const int rows_per_pass = cqs > 0 ? 128 / cqs : 0;
        const int my_q = cqs > 0 ? et % cqs : 0, my_r = cqs > 0 ? et / cqs : 0;
        double st1[4] = {0.0, 0.0, 0.0, 0.0}, st2[4] = {0.0, 0.0, 0.0, 0.0};
        if (et < 64) { s_stat[0][et] = 0.0; s_stat[1][et] = 0.0; }
        uint32_t j = 0;
        for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1, aph = (j >> 1) & 1;
            mbar_wait(bar_accfull + 8 * buf, aph);
            __syncwarp();
            tc_fence_after();
            const int row = q * 32 + lane;
            const long m0 = tile * P_TILE_M;
            if (direct) {
                // odd channel counts / unaligned destinations: plain per-row stores
                const long m = m0 + row;
                const bool row_ok = m < a.M;
                float *orow = a.out + (row_ok ? m : 0) * a.out_pitch + a.out_choff;
                for (int g16 = 0; g16 < a.N_pad; g16 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + buf * P_ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g16), v);
                    if (SPLIT) {
                        float w2[16];
                        tmem_ld16(tmem_base + buf * P_ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(a.N_pad + g16), w2);
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] += w2[e];
                    }
                    if (row_ok) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (g16 + e < a.C_out) orow[g16 + e] = v[e];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty + 8 * buf);
                continue;
            }
            // 1. TMEM -> registers -> shared planes: plane p holds quad p of all 128 rows, 16 B per row (conflict-free)
            for (int g16 = 0; g16 < a.N_pad; g16 += 16) {
                float v[16];
                tmem_ld16(tmem_base + buf * P_ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g16), v);
                if (SPLIT) {
                    float w2[16];
                    tmem_ld16(tmem_base + buf * P_ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(a.N_pad + g16), w2);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += w2[e];
                }
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
                    *reinterpret_cast<float4 *>(stage + ((g16 >> 2) + qq) * PLANE_F + row * 4) =
                        make_float4(v[qq * 4], v[qq * 4 + 1], v[qq * 4 + 2], v[qq * 4 + 3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty + 8 * buf);            // accumulator drained: the MMA warp may reuse it
            asm volatile("bar.sync 1, 128;" ::: "memory");               // staging complete (epilogue warps only)
            // 2. coalesced write-out: consecutive threads take consecutive float4 of the (row, quad) stream
            const int cq = a.C_out >> 2;                                  // valid quads per row
            if (want_stats) {
                if (my_r < rows_per_pass) {
                    for (int r = my_r; r < P_TILE_M; r += rows_per_pass) {   // same (row, quad) stream as below, fixed quad per thread
                        const long m = m0 + r;
                        if (m < a.M) {
                            const float4 val = *reinterpret_cast<const float4 *>(stage + my_q * PLANE_F + r * 4);
                            *reinterpret_cast<float4 *>(a.out + m * a.out_pitch + a.out_choff + my_q * 4) = val;
                            const double d0 = val.x, d1 = val.y, d2 = val.z, d3 = val.w;
                            st1[0] += d0; st1[1] += d1; st1[2] += d2; st1[3] += d3;
                            st2[0] += d0 * d0; st2[1] += d1 * d1; st2[2] += d2 * d2; st2[3] += d3 * d3;
                        }
                    }
                }
            } else {
                const int total = P_TILE_M * cq;
                for (int f = et; f < total; f += 128) {
                    const int r = f / cq, qd = f - r * cq;
                    const long m = m0 + r;
                    if (m < a.M) {
                        const float4 val = *reinterpret_cast<const float4 *>(stage + qd * PLANE_F + r * 4);
                        *reinterpret_cast<float4 *>(a.out + m * a.out_pitch + a.out_choff + qd * 4) = val;
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");               // staging buffer free for the next tile
        }
        if (want_stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");               // s_stat zeroed by every epilogue thread's store above
            if (my_r < rows_per_pass) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { atomicAdd(&s_stat[0][my_q * 4 + e], st1[e]); atomicAdd(&s_stat[1][my_q * 4 + e], st2[e]); }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et < a.C_out) {
                atomicAdd(a.stats + et, s_stat[0][et]);
                atomicAdd(a.stats + a.stats_stride + et, s_stat[1][et]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == P_PWARPS) {
        __syncwarp();
        tmem_dealloc(tmem_base, 2 * P_ACC_COLS);
    }
}

}  // namespace

static size_t persist_smem(int nchunks, int N_pad, bool split, int stages) {
    return static_cast<size_t>(stages) * (split ? 2 : 1) * P_A_TILE + static_cast<size_t>(nchunks) * (split ? 2 : 1) * N_pad * 128 +
           static_cast<size_t>(N_pad / 4) * 129 * 16 + 1024;
}
static int persist_stages(int nchunks, int N_pad, bool split) {
    for (int st = P_MAX_STAGES; st >= 2; --st)
        if (persist_smem(nchunks, N_pad, split, st) <= 226 * 1024) return st;
    return 0;
}

bool eml_persist_supported(const eml_conv_params *p) {
    if (p->mode != EML_CONV_1x1) return false;
    if (p->stats != nullptr && ((p->C_out & 3) || ((p->out_pitch | p->out_choff) & 3) || p->C_out > 64 || eml_env_flag("EML_NO_PERSIST_STATS")))
        return false;                                 // the statistics epilogue lives in the staged (float4) write-out
    if (p->precision != EML_PREC_BF16 && p->precision != EML_PREC_BF16X3) return false;
    const int N_pad = (p->C_out + 15) & ~15;
    const int nchunks = (p->C_in + P_CHUNK_K - 1) / P_CHUNK_K;
    return 2 * N_pad <= P_ACC_COLS && persist_stages(nchunks, N_pad, p->precision == EML_PREC_BF16X3) >= 3;
}

int eml_persist_forward(const eml_conv_params *p, cudaStream_t st) {
    PArgs a{};
    a.in = p->in; a.scale = p->scale; a.shift = p->shift; a.wpack = static_cast<const unsigned char *>(p->wpack);
    a.out = p->out;
    a.M = static_cast<long>(p->B) * p->H * p->W;
    a.C_in = p->C_in; a.in_pitch = p->in_pitch;
    a.C_out = p->C_out; a.N_pad = (p->C_out + 15) & ~15; a.out_pitch = p->out_pitch; a.out_choff = p->out_choff;
    a.nchunks = (p->C_in + P_CHUNK_K - 1) / P_CHUNK_K;
    a.relu = p->relu;
    a.stats = p->stats; a.stats_stride = p->stats_stride > 0 ? p->stats_stride : p->C_out;
    a.ntiles = (a.M + P_TILE_M - 1) / P_TILE_M;
    const bool split = p->precision == EML_PREC_BF16X3;
    a.stages = persist_stages(a.nchunks, a.N_pad, split);
    const size_t smem = persist_smem(a.nchunks, a.N_pad, split, a.stages);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = static_cast<unsigned>(a.ntiles < sms ? a.ntiles : sms);
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        kern<<<grid, P_THREADS, smem, st>>>(a);
        return EML_OK;
    };
    int rc;
    if (split) rc = p->relu ? go(conv1x1_persist_kernel<true, true>) : go(conv1x1_persist_kernel<true, false>);
    else rc = p->relu ? go(conv1x1_persist_kernel<false, true>) : go(conv1x1_persist_kernel<false, false>);
    if (rc != EML_OK) return rc;
    return eml_launch_status();
}
