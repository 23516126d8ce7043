// DenseNet-BC convolutions as tcgen05 implicit GEMMs (sm_100a).
//
//   out[m, n] = sum_{tap, c} W[n, tap, c] * act(scale[c] * in[src(m, tap), c] + shift[c])
//
// GEMM view: M = output pixels (tile of 128 = one UMMA_M=128 accumulator, one TMEM lane per pixel),
// N = C_out padded to 16 (TMEM columns), K = taps * C_in walked in chunks of 64 bf16 (= one 128-byte
// swizzle-atom row).  Per chunk:
//   * B (weights) is a pre-packed, pre-swizzled image in global memory (eml_conv_pack_weights) that one
//     thread pulls into shared memory with a single cp.async.bulk (TMA bulk engine, mbarrier complete_tx);
//   * A cannot be TMA-loaded as-is: every dense layer applies its OWN BatchNorm(+ReLU) to the shared
//     concatenation buffer (DenseNet.py:30-31,39), so the 128 threads gather the NHWC rows with coalesced
//     float4 loads (16 lanes x 16 B = one 256-byte pixel row segment), apply scale/shift/ReLU, zero the
//     out-of-image taps, split into bf16 hi (+ lo) and store straight into the K-major SWIZZLE_128B
//     canonical layout the UMMA descriptor expects; fence.proxy.async hands the tile to the tensor core;
//   * one thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM): 1 MMA per 16-wide k-step in
//     BF16 mode, 3 (hi*hi + lo*hi + hi*lo) in BF16X3 mode, which carries ~16 mantissa bits per operand and
//     keeps the network within 1e-3 of the fp32 reference (single-pass bf16 does not: DESIGN.md).
//   * tcgen05.commit -> mbarrier releases the stage; a 2-stage ring lets the gather of chunk c+1 overlap
//     the MMAs of chunk c, and 2-5 resident CTAs per SM overlap one CTA's epilogue with another's gather.
// Epilogue: tcgen05.ld 32x32b (thread = pixel row) -> fp32 -> NHWC store at a channel offset (the dense
// block's concat buffer is written in place), optional per-channel sum / sum-of-squares (batch-stat BN)
// reduced through shared memory in double and added to global accumulators.
//
// These kernels are HBM-bound by design (DESIGN.md section "roofline"): per 128-pixel tile the gather moves
// 128*C_in*4 bytes while the MMAs take ~24 cycles per k-step at N=48.
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 64;
constexpr int NTHREADS = 256;
constexpr int ROWS_PER_THREAD = TILE_M / (NTHREADS / 16);   // 8
constexpr int STAGES = 2;
constexpr int A_TILE_BYTES = TILE_M * CHUNK_K * 2;   // 16 KB: 128 rows x 128 B

// ------------------------------------------------------------------------------------------------ kernel
struct GemmArgs {
    const float *in;
    const float *scale;
    const float *shift;
    const unsigned char *wpack;
    float *out;
    double *stats;
    long stats_stride;
    long M;              // output pixels
    int H, W;            // input spatial size
    int C_in, in_pitch;
    int C_out, N_pad, out_pitch, out_choff;
    int cpt;             // chunks per tap
    int nchunks;
    int relu;
    int tmem_cols;
};

__device__ __forceinline__ float act(float x, float s, float t, int relu) {
    float v = fmaf(x, s, t);
    return relu ? fmaxf(v, 0.f) : v;
}
// Channels >= C_in inside the last quad are forced to exactly 0 (they may hold unwritten memory / NaN).
__device__ __forceinline__ float4 act4(float4 v, float4 sc, float4 sh, int relu, int nvalid) {
    float4 o;
    o.x = act(v.x, sc.x, sh.x, relu);
    o.y = nvalid > 1 ? act(v.y, sc.y, sh.y, relu) : 0.f;
    o.z = nvalid > 2 ? act(v.z, sc.z, sh.z, relu) : 0.f;
    o.w = nvalid > 3 ? act(v.w, sc.w, sh.w, relu) : 0.f;
    return o;
}
__device__ __forceinline__ float4 load_quad_guarded(const float *p, int ch, int nvalid, float fill) {
    if (p == nullptr) return make_float4(fill, fill, fill, fill);
    if (nvalid >= 4) return *reinterpret_cast<const float4 *>(p + ch);
    float4 r = make_float4(fill, fill, fill, fill);
    r.x = p[ch];
    if (nvalid > 1) r.y = p[ch + 1];
    if (nvalid > 2) r.z = p[ch + 2];
    return r;
}

// MODE 0: 1x1 | 1: 3x3 pad 1 | 2: act -> 2x2 average pool -> 1x1
template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_kernel(const GemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2 * STAGES + 1];
    __shared__ uint32_t s_tmem;

    // 1024-byte aligned stage ring
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int b_tile_bytes = a.N_pad * 128;
    const int stage_bytes = (SPLIT ? 2 : 1) * (A_TILE_BYTES + b_tile_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long m0 = static_cast<long>(blockIdx.x) * TILE_M;

    const uint32_t bar_full = smem_u32(&s_bar[0]);            // [STAGES] weights landed
    const uint32_t bar_done = smem_u32(&s_bar[STAGES]);       // [STAGES] MMAs that read the stage retired
    const uint32_t bar_acc = smem_u32(&s_bar[2 * STAGES]);    // accumulator complete

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_done + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();                                       // .sync.aligned below needs the warp converged
        tmem_alloc(smem_u32(&s_tmem), static_cast<uint32_t>(a.tmem_cols));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    bool leader = false;                                    // one elected lane of warp 0 issues the MMAs (chosen while converged)
    if (warp == 0) leader = elect_one();

    // ---- gather geometry: thread -> (channel quad `sub`, row group `rgrp` < 16); rows r = rgrp + 16*i, i < 8.
    const int sub = tid & 15, rgrp = tid >> 4;
    // source pixel index per row (-1: row beyond M); for 3x3 also the (x, y) of the centre pixel.
    int pix[ROWS_PER_THREAD];
    int xy[MODE == 1 ? ROWS_PER_THREAD : 1];
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        long m = m0 + rgrp + 16 * i;
        if (m >= a.M) { pix[i] = -1; if (MODE == 1) xy[i] = 0; continue; }
        if (MODE == 0) {
            pix[i] = static_cast<int>(m);
        } else if (MODE == 1) {
            int x = static_cast<int>(m % a.W);
            int y = static_cast<int>((m / a.W) % a.H);
            pix[i] = static_cast<int>(m);
            xy[i] = x | (y << 16);
        } else {
            const int Wo = a.W >> 1, Ho = a.H >> 1;
            int xo = static_cast<int>(m % Wo);
            long t = m / Wo;
            int yo = static_cast<int>(t % Ho);
            long b = t / Ho;
            pix[i] = static_cast<int>((b * a.H + 2 * yo) * a.W + 2 * xo);
        }
    }
    // per-thread constant part of the swizzled store offset: row = rgrp + 16*i  ->  (2i + rgrp/8)*1024 + (rgrp%8)*128 + ...
    const uint32_t st_off = static_cast<uint32_t>((rgrp >> 3) * 1024 + (rgrp & 7) * 128 + ((((sub >> 1) ^ rgrp) & 7) << 4) + (sub & 1) * 8);
    const uint32_t idesc = make_idesc_bf16(TILE_M, a.N_pad);
    const size_t wchunk_bytes = static_cast<size_t>(2) * b_tile_bytes;     // hi image then lo image

    for (int c = 0; c < a.nchunks; ++c) {
        const int s = c % STAGES;
        const int it = c / STAGES;
        unsigned char *st_base = smem + static_cast<size_t>(s) * stage_bytes;
        unsigned char *a_hi = st_base;
        unsigned char *a_lo = st_base + A_TILE_BYTES;                               // only when SPLIT
        unsigned char *b_hi = st_base + (SPLIT ? 2 : 1) * A_TILE_BYTES;
        if (c >= STAGES) mbar_wait(bar_done + 8 * s, static_cast<uint32_t>((it - 1) & 1));
        if (tid == 0) {
            const uint32_t bytes = static_cast<uint32_t>((SPLIT ? 2 : 1) * b_tile_bytes);
            mbar_expect_tx(bar_full + 8 * s, bytes);
            bulk_g2s(smem_u32(b_hi), a.wpack + static_cast<size_t>(c) * wchunk_bytes, bytes, bar_full + 8 * s);
        }
        const int tap = c / a.cpt;
        const int c0 = (c - tap * a.cpt) * CHUNK_K;
        const int kvalid = min(CHUNK_K, a.C_in - c0);
        const int ksteps = (kvalid + 15) >> 4;
        const int ch = c0 + sub * 4;
        const bool ch_ok = ch < a.C_in;
        const int nvalid = min(4, a.C_in - ch);             // C_in need not be a multiple of 4 (block 3: 150 + 12 l)
        if (sub * 4 < ksteps * 16) {                        // this lane's 8-byte slot is read by the MMA
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch_ok) {
                sc = load_quad_guarded(a.scale, ch, nvalid, 1.f);
                sh = load_quad_guarded(a.shift, ch, nvalid, 0.f);
            }
            int dy = 0, dx = 0;
            if (MODE == 1) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
            {
                if (MODE != 2) {
                    float4 v[8];
                    bool ok[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = j;
                        ok[j] = ch_ok && pix[i] >= 0;
                        long src = pix[i];
                        if (MODE == 1) {
                            int x = (xy[i] & 0xffff) + dx, y = (xy[i] >> 16) + dy;
                            ok[j] = ok[j] && x >= 0 && x < a.W && y >= 0 && y < a.H;
                            src += dy * a.W + dx;
                        }
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok[j]) v[j] = __ldg(reinterpret_cast<const float4 *>(a.in + src * a.in_pitch + ch));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);      // zero padding is applied AFTER the affine
                        if (ok[j]) o = act4(v[j], sc, sh, a.relu, nvalid);
                        store_quad<SPLIT>(a_hi, a_lo, st_off + j * 2048, o);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {          // 4 rows x 4 taps in flight
                        float4 v[4][4];
                        bool ok[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = q * 4 + j;
                            ok[j] = ch_ok && pix[i] >= 0;
                            const float *p = a.in + static_cast<long>(pix[i]) * a.in_pitch + ch;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                v[j][t] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (ok[j]) v[j][t] = __ldg(reinterpret_cast<const float4 *>(p + static_cast<long>((t >> 1) * a.W + (t & 1)) * a.in_pitch));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = q * 4 + j;
                            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ok[j]) {
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float4 u = act4(v[j][t], sc, sh, a.relu, nvalid);
                                    o.x += u.x; o.y += u.y; o.z += u.z; o.w += u.w;
                                }
                                o.x *= 0.25f; o.y *= 0.25f; o.z *= 0.25f; o.w *= 0.25f;
                            }
                            store_quad<SPLIT>(a_hi, a_lo, st_off + i * 2048, o);
                        }
                    }
                }
            }
        }
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (warp == 0) {
            // warp-uniform issue: all lanes wait, one elected lane issues (inside `if (tid == 0)` ptxas emits an
            // ELECT/R2UR loop in front of every UTCHMMA because the descriptors look thread-varying)
            mbar_wait(bar_full + 8 * s, static_cast<uint32_t>(it & 1));
            tc_fence_after();
            if (leader) {
                const uint64_t da_hi = make_sw128_desc(smem_u32(a_hi));
                const uint64_t db_hi = make_sw128_desc(smem_u32(b_hi));
                const uint64_t da_lo = da_hi + static_cast<uint64_t>(A_TILE_BYTES >> 4);
                const uint64_t db_lo = db_hi + static_cast<uint64_t>(b_tile_bytes >> 4);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t adv = static_cast<uint64_t>(k * 2);      // +32 bytes per 16-wide k-step, >>4
                    umma_bf16(tmem_base, da_hi + adv, db_hi + adv, idesc, (c | k) != 0 ? 1u : 0u);
                    if (SPLIT) {
                        umma_bf16(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
                        umma_bf16(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
                    }
                }
                umma_commit(bar_done + 8 * s);                 // implies tcgen05.fence::before_thread_sync
                if (c == a.nchunks - 1) umma_commit(bar_acc);
            }
            __syncwarp();
        }
    }

    // ---- epilogue: TMEM lane = tile row; warp w owns lanes [32w, 32w+32)
    mbar_wait(bar_acc, 0);
    __syncwarp();                                           // tcgen05.ld is .sync.aligned
    tc_fence_after();
    const int row = (warp & 3) * 32 + lane;
    const long m = m0 + row;
    const bool row_ok = m < a.M;
    float *orow = a.out + (row_ok ? m : 0) * a.out_pitch + a.out_choff;
    const bool vec_ok = ((a.out_pitch | a.out_choff) & 3) == 0;
    float *tile = reinterpret_cast<float *>(smem);         // [128][N_pad+1] staging for the statistics
    const int tp = a.N_pad + 1;
    // warps w and w+4 share TMEM lanes [32(w%4), +32) and split the 16-column groups between them
    for (int g = (warp >> 2) * 16; g < a.N_pad; g += 32) {
        float v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(g), v);
        if (row_ok) {
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
        }
        if (a.stats != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; ++e) tile[row * tp + g + e] = row_ok ? v[e] : 0.f;
        }
    }
    if (a.stats != nullptr) {
        __syncthreads();
        for (int n = tid; n < a.C_out; n += NTHREADS) {
            double s1 = 0.0, s2 = 0.0;
            for (int r = 0; r < TILE_M; ++r) {
                double x = static_cast<double>(tile[r * tp + n]);
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
        tmem_dealloc(tmem_base, static_cast<uint32_t>(a.tmem_cols));
    }
}

// ------------------------------------------------------------------------------------------------ weights
// OIHW fp32 -> per chunk [hi image | lo image], each N_pad rows x 64 bf16, K-major SWIZZLE_128B.
__global__ void pack_weights_kernel(const float *__restrict__ w, unsigned char *__restrict__ out, int C_out,
                                    int C_in, int taps, int N_pad, int cpt) {
    const int nchunks = taps * cpt;
    const long total = static_cast<long>(nchunks) * N_pad * CHUNK_K;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx % CHUNK_K);
        const int n = static_cast<int>((idx / CHUNK_K) % N_pad);
        const int c = static_cast<int>(idx / (static_cast<long>(CHUNK_K) * N_pad));
        const int tap = c / cpt;
        const int ci = (c % cpt) * CHUNK_K + k;
        float v = 0.f;
        if (n < C_out && ci < C_in) v = w[(static_cast<long>(n) * C_in + ci) * taps + tap];
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        unsigned char *base = out + static_cast<size_t>(c) * 2 * N_pad * 128;
        const uint32_t off = sw128_offset(n, k);
        *reinterpret_cast<__nv_bfloat16 *>(base + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + static_cast<size_t>(N_pad) * 128 + off) = lo;
    }
}

// The same image for a plain (N, K) row-major matrix (taps = 1: GEMM operands of the GenProjector path, ~1350 packs per G+D
// iteration, many of them activations), `nslices` slices of `rows` rows each at a fixed stride: one thread converts one 16-byte
// swizzle chunk (8 consecutive k of one row: two float4 reads, one 16-byte hi store, one 16-byte lo store) instead of one element
// with 2-byte stores.
__global__ void __launch_bounds__(256) pack_matrix_kernel(const float *__restrict__ w, unsigned char *__restrict__ out, int rows, int K,
                                                          int N_pad, int nchunks, int nslices, long slice_bytes) {
    const long per_slice = static_cast<long>(nchunks) * N_pad * 8;
    const long total = per_slice * nslices;
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const int sl = static_cast<int>(idx / per_slice);
        const long r = idx - sl * per_slice;
        const int kc = static_cast<int>(r & 7);
        const int n = static_cast<int>((r >> 3) % N_pad);
        const int c = static_cast<int>((r >> 3) / N_pad);
        const int k0 = c * CHUNK_K + kc * 8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (n < rows && k0 < K) {
            const float *src = w + (static_cast<long>(sl) * rows + n) * K + k0;
            if (vec && k0 + 8 <= K) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(src)), b = __ldg(reinterpret_cast<const float4 *>(src + 4));
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (k0 + i < K) v[i] = src[i];
            }
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - __low2float(hh), v[2 * i + 1] - __high2float(hh));
            h[i] = *reinterpret_cast<const uint32_t *>(&hh);
            l[i] = *reinterpret_cast<const uint32_t *>(&ll);
        }
        unsigned char *base = out + sl * slice_bytes + static_cast<size_t>(c) * 2 * N_pad * 128;
        const uint32_t off = sw128_offset(n, kc * 8);
        *reinterpret_cast<uint4 *>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(base + static_cast<size_t>(N_pad) * 128 + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

inline int pad16(int n) { return (n + 15) & ~15; }
inline int tmem_cols_for(int n) { int c = 32; while (c < n) c <<= 1; return c; }

template <int MODE, bool SPLIT>
int launch(const GemmArgs &a, cudaStream_t st) {
    const int stage_bytes = (SPLIT ? 2 : 1) * (A_TILE_BYTES + a.N_pad * 128);
    size_t smem = static_cast<size_t>(STAGES) * stage_bytes;
    const size_t stats_bytes = a.stats ? static_cast<size_t>(TILE_M) * (a.N_pad + 1) * 4 : 0;
    if (stats_bytes > smem) smem = stats_bytes;
    smem += 1024;    // alignment slack
    if (smem > 227 * 1024) return EML_E_SHAPE;
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const long tiles = (a.M + TILE_M - 1) / TILE_M;
    conv_gemm_kernel<MODE, SPLIT><<<static_cast<unsigned>(tiles), NTHREADS, smem, st>>>(a);
    return eml_launch_status();
}

}  // namespace

int eml_conv_forward_simt(const eml_conv_params *p, cudaStream_t st);   // conv_simt.cu
// conv3x3_rows.cu: planar no-im2col 3x3 kernel
size_t eml_rows_wpack_bytes(int C_out, int C_in);
int eml_rows_pack(const float *w_oihw, unsigned char *dst, int C_out, int C_in, cudaStream_t st);
bool eml_rows_supported(const eml_conv_params *p);
// conv1x1_persist.cu: persistent warp-specialised 1x1 kernel (no statistics epilogue)
bool eml_persist_supported(const eml_conv_params *p);
int eml_persist_forward(const eml_conv_params *p, cudaStream_t st);
int eml_rows_forward(const eml_conv_params *p, const unsigned char *wplanar, cudaStream_t st);
// dense_layer.cu: TMA-fed pipeline with a pooling epilogue (transition with C_out <= 112)
bool eml_dense_pool_supported(const eml_conv_params *p);
int eml_dense_pool_forward(const eml_conv_params *p, cudaStream_t st);

static size_t generic_wpack_bytes(int C_out, int C_in, int taps) {
    const int cpt = (C_in + CHUNK_K - 1) / CHUNK_K;
    return static_cast<size_t>(taps) * cpt * 2 * pad16(C_out) * 128;
}

extern "C" size_t eml_conv_wpack_bytes(int C_out, int C_in, int taps) {
    if (C_out <= 0 || C_in <= 0 || taps <= 0) return 0;
    // generic SWIZZLE_128B chunk images, followed (3x3 only) by the planar image of conv3x3_rows.cu
    return generic_wpack_bytes(C_out, C_in, taps) + (taps == 9 ? eml_rows_wpack_bytes(C_out, C_in) : 0);
}

extern "C" int eml_conv_pack_weights(const float *w_oihw, void *wpack, int C_out, int C_in, int taps, void *stream) {
    EML_CHECK_PTR(w_oihw); EML_CHECK_PTR(wpack);
    EML_CHECK_ALIGN16(wpack);
    if (C_out <= 0 || C_out > 256 || C_in <= 0 || (taps != 1 && taps != 9)) return EML_E_SHAPE;
    const int cpt = (C_in + CHUNK_K - 1) / CHUNK_K;
    if (taps == 1 && !eml_env_flag("EML_PACK_V1")) {
        const long items = static_cast<long>(cpt) * pad16(C_out) * 8;
        const unsigned grid = static_cast<unsigned>(items / 256 + 1 < 148 * 8 ? items / 256 + 1 : 148 * 8);
        pack_matrix_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w_oihw, static_cast<unsigned char *>(wpack), C_out, C_in,
                                                                             pad16(C_out), cpt, 1, 0);
        return eml_launch_status();
    }
    pack_weights_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w_oihw, static_cast<unsigned char *>(wpack), C_out, C_in, taps, pad16(C_out), cpt);
    int rc = eml_launch_status();
    if (rc == EML_OK && taps == 9 && eml_rows_wpack_bytes(C_out, C_in) > 0)
        rc = eml_rows_pack(w_oihw, static_cast<unsigned char *>(wpack) + generic_wpack_bytes(C_out, C_in, taps), C_out, C_in,
                           static_cast<cudaStream_t>(stream));
    return rc;
}


// `nslices` slices of `rows` rows each (rows a multiple of 16, <= 256) of a row-major (nslices*rows, K) matrix, slice s packed at
// wpack + s * slice_bytes exactly as eml_conv_pack_weights(w + s*rows*K, ..., rows, K, 1) would: ONE launch (eml_gemm_bf16_slices).
extern "C" int eml_gemm_pack_slices(const float *w, void *wpack, int nslices, int rows, int K, long slice_bytes, void *stream) {
    EML_CHECK_PTR(w); EML_CHECK_PTR(wpack);
    EML_CHECK_ALIGN16(wpack);
    if (nslices <= 0 || rows <= 0 || rows > 256 || (rows & 15) || K <= 0) return EML_E_SHAPE;
    const int cpt = (K + CHUNK_K - 1) / CHUNK_K;
    if ((slice_bytes & 15) || slice_bytes < static_cast<long>(generic_wpack_bytes(rows, K, 1))) return EML_E_SHAPE;
    const long items = static_cast<long>(cpt) * rows * 8 * nslices;
    const unsigned grid = static_cast<unsigned>(items / 256 + 1 < 148 * 8 ? items / 256 + 1 : 148 * 8);
    pack_matrix_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<unsigned char *>(wpack), rows, K, rows, cpt, nslices,
                                                                         slice_bytes);
    return eml_launch_status();
}

extern "C" int eml_conv_forward(const eml_conv_params *p, void *stream) {
    EML_CHECK_PTR(p); EML_CHECK_PTR(p->in); EML_CHECK_PTR(p->out);
    if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->C_in <= 0 || p->C_out <= 0 || p->C_out > 256) return EML_E_SHAPE;
    if (p->mode < 0 || p->mode > 2) return EML_E_ARG;
    if (p->plane_pixels > 0) {                               // channel-plane input (header): only the TMA transition reads it
        if (p->precision == EML_PREC_FP32 || p->wpack == nullptr || !eml_dense_pool_supported(p)) return EML_E_SHAPE;
    } else if ((p->in_pitch & 3) || p->in_pitch < p->C_in) return EML_E_ALIGN;
    if (p->out_pitch < p->out_choff + p->C_out || p->out_choff < 0) return EML_E_SHAPE;
    if (p->mode == EML_CONV_POOL2 && ((p->H | p->W) & 1)) return EML_E_SHAPE;
    EML_CHECK_ALIGN16(p->in);
    if (p->scale) EML_CHECK_ALIGN16(p->scale);
    if (p->shift) EML_CHECK_ALIGN16(p->shift);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->precision == EML_PREC_FP32) return eml_conv_forward_simt(p, st);
    if (p->precision != EML_PREC_BF16 && p->precision != EML_PREC_BF16X3) return EML_E_ARG;
    EML_CHECK_PTR(p->wpack);
    EML_CHECK_ALIGN16(p->wpack);
    if (static_cast<long>(p->B) * p->H * p->W >= (1L << 31)) return EML_E_SHAPE;
    if (eml_dense_pool_supported(p)) return eml_dense_pool_forward(p, st);
    if (p->plane_pixels > 0) return EML_E_SHAPE;             // channel-plane input: the TMA transition only
    if (eml_persist_supported(p) && !eml_env_flag("EML_NO_PERSIST")) return eml_persist_forward(p, st);
    if (eml_rows_supported(p))
        return eml_rows_forward(p, static_cast<const unsigned char *>(p->wpack) + generic_wpack_bytes(p->C_out, p->C_in, 9), st);

    GemmArgs a{};
    a.in = p->in; a.scale = p->scale; a.shift = p->shift;
    a.wpack = static_cast<const unsigned char *>(p->wpack);
    a.out = p->out; a.stats = p->stats; a.stats_stride = p->stats_stride > 0 ? p->stats_stride : p->C_out;
    a.H = p->H; a.W = p->W;
    a.M = (p->mode == EML_CONV_POOL2) ? static_cast<long>(p->B) * (p->H / 2) * (p->W / 2)
                                      : static_cast<long>(p->B) * p->H * p->W;
    a.C_in = p->C_in; a.in_pitch = p->in_pitch;
    a.C_out = p->C_out; a.N_pad = pad16(p->C_out); a.out_pitch = p->out_pitch; a.out_choff = p->out_choff;
    a.cpt = (p->C_in + CHUNK_K - 1) / CHUNK_K;
    a.nchunks = (p->mode == EML_CONV_3x3 ? 9 : 1) * a.cpt;
    a.relu = p->relu;
    a.tmem_cols = tmem_cols_for(a.N_pad);
    const bool split = p->precision == EML_PREC_BF16X3;
    switch (p->mode) {
        case EML_CONV_1x1: return split ? launch<0, true>(a, st) : launch<0, false>(a, st);
        case EML_CONV_3x3: return split ? launch<1, true>(a, st) : launch<1, false>(a, st);
        default: return split ? launch<2, true>(a, st) : launch<2, false>(a, st);
    }
}
