// Weight gradient of the 1x1 convolutions on tensor cores (sm_100a):   dW[n, c] = sum_m G[m, n] * act(scale[c]*x[m, c] + shift[c])
//
// This is a GEMM whose reduction dimension is the PIXEL index (K = B*H*W ~ 10^6..10^7) and whose output is tiny (48 x C_in).
// In NHWC both operands are naturally "MN-major": for a fixed pixel (k) the channels (the M / N index) are contiguous.  That is
// exactly tcgen05's MN-major SWIZZLE_128B canonical layout -- one 128-byte row (64 bf16 channels) per K index, 8-row groups
// 1024 B apart (SBO), 64-channel blocks LBO apart -- so the producers write the SAME shared-memory image as the forward
// gather (row = pixel, 16-byte chunk index XOR row%8) and only the descriptors (a_major = b_major = MN) differ.
//   D[c, n] (TMEM: lane = input channel c of a 128-channel M tile, column = n)  +=  X^T[c, k] * G[k, n]
// A persistent CTA accumulates over ALL of its 64-pixel K tiles in TMEM and runs one epilogue at the very end (atomicAdd of
// its 128 x 48 partial into dW) -- the SIMT version spent its time re-reading shared memory for 12 FMAs per 7 loads.
// bf16x3: X and G are split hi/lo, 3 MMAs per k-step (hi*hi + lo*hi + hi*lo).
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int W_KT = 64;                       // pixels per K tile (4 UMMA k-steps)
constexpr int W_PRODUCERS = 512;
constexpr int W_PWARPS = W_PRODUCERS / 32;
constexpr int W_THREADS = W_PRODUCERS + 32 + 128;
constexpr int W_BLK = W_KT * 128;              // bytes of one [64 px][64 ch] bf16 block (8 KB)
constexpr int W_STAGES = 2;

struct WArgs {
    const float *G; int g_pitch, N;            // gradient of the conv output, N <= 64 channels
    const float *x; int x_pitch, C;            // conv input (stored slab), C <= 256 channels
    const float *scale, *shift;
    float *dW;                                 // (N, dw_stride): this call fills columns [0, C)
    int dw_stride;
    long M;
    int relu, nblk;                            // nblk = 2 * ceil(C / 128) channel blocks of 64 (zero padded)
    long ntiles;
};

__device__ __forceinline__ void w_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// MN-major SWIZZLE_128B descriptor: LBO = bytes between 64-element MN blocks, SBO = bytes between 8-row K groups.
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {   // a_major = b_major = MN (bits 15, 16)
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

template <bool SPLIT>
__global__ void __launch_bounds__(W_THREADS, 1) wgrad1x1_tc_kernel(const WArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2 * W_STAGES + 1];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_sc[256], s_sh[256];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    // stage: X blocks hi [nblk][8 KB] | X blocks lo | G block hi [8 KB] | G block lo
    const int x_img = a.nblk * W_BLK;
    const int stage_bytes = (SPLIT ? 2 : 1) * (x_img + W_BLK);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[W_STAGES]), bar_done = smem_u32(&s_bar[2 * W_STAGES]);
    const int mt = a.nblk / 2;                                      // 128-channel M tiles
    const int N_pad = (a.N + 15) & ~15;

    for (int i = tid; i < 256; i += W_THREADS) {                    // channels of this launch's range (<= 256); identity behind C
        s_sc[i] = (i < a.C && a.scale) ? a.scale[i] : 1.f;
        s_sh[i] = (i < a.C && a.shift) ? a.shift[i] : 0.f;
    }
    if (tid == 0) {
        for (int s = 0; s < W_STAGES; ++s) { mbar_init(bar_full + 8 * s, W_PWARPS); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == W_PWARPS) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 128);                         // mt (<= 2) accumulators of 64 columns
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp < W_PWARPS) {
        // =========================================================== PRODUCERS: thread -> (channel quad `sub`, rows rgrp, rgrp+32)
        const int sub = tid & 15, rgrp = tid >> 4;
        uint32_t g = 0;
        for (long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++g) {
            const int s = g % W_STAGES;
            const uint32_t ph = (g / W_STAGES) & 1;
            const long m0 = t * W_KT;
            unsigned char *st = smem + static_cast<size_t>(s) * stage_bytes;
            unsigned char *x_hi = st, *x_lo = st + x_img;
            unsigned char *g_hi = st + (SPLIT ? 2 : 1) * x_img, *g_lo = g_hi + W_BLK;
            // loads first (up to 2*nblk + 2 float4 per thread), then wait for the stage, then convert + store
            float4 xv[2][4], gv[2];
            bool rok[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long m = m0 + rgrp + 32 * i;
                rok[i] = m < a.M;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    xv[i][b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int ch = b * 64 + sub * 4;
                    if (b < a.nblk && rok[i] && ch < a.C) {
                        if (ch + 3 < a.C && (a.x_pitch & 3) == 0) {
                            xv[i][b] = __ldg(reinterpret_cast<const float4 *>(a.x + m * a.x_pitch + ch));
                        } else {
                            const float *p = a.x + m * a.x_pitch + ch;
                            xv[i][b].x = p[0];
                            if (ch + 1 < a.C) xv[i][b].y = p[1];
                            if (ch + 2 < a.C) xv[i][b].z = p[2];
                        }
                    }
                }
                gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int n = sub * 4;
                if (rok[i] && n < a.N) {
                    const float *p = a.G + m * a.g_pitch + n;
                    if (n + 3 < a.N && (a.g_pitch & 3) == 0) gv[i] = __ldg(reinterpret_cast<const float4 *>(p));
                    else { gv[i].x = p[0]; if (n + 1 < a.N) gv[i].y = p[1]; if (n + 2 < a.N) gv[i].z = p[2]; }
                }
            }
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = rgrp + 32 * i;
                const uint32_t off = static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((sub >> 1) ^ r) & 7) << 4) + (sub & 1) * 8);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (b >= a.nblk) break;
                    const int ch = b * 64 + sub * 4;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rok[i] && ch < a.C) {
                        // folded BatchNorm affine from shared memory (two broadcast LDS.128; it was 8 predicated scalar global loads per
                        // quad and tile -- the producers were issue / LSU-bound, profiles/r02_ncu_train_kernels_metrics.csv)
                        const float4 sc = *reinterpret_cast<const float4 *>(s_sc + ch), sh = *reinterpret_cast<const float4 *>(s_sh + ch);
                        o.x = fmaf(xv[i][b].x, sc.x, sh.x); o.y = fmaf(xv[i][b].y, sc.y, sh.y);
                        o.z = fmaf(xv[i][b].z, sc.z, sh.z); o.w = fmaf(xv[i][b].w, sc.w, sh.w);
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (ch + 1 >= a.C) o.y = 0.f;
                        if (ch + 2 >= a.C) o.z = 0.f;
                        if (ch + 3 >= a.C) o.w = 0.f;
                    }
                    store_quad<SPLIT>(x_hi + b * W_BLK, x_lo + b * W_BLK, off, o);
                }
                store_quad<SPLIT>(g_hi, g_lo, off, gv[i]);             // channels >= N are zero (sub*4 >= N lanes loaded nothing)
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) w_mbar_arrive(bar_full + 8 * s);
        }
    } else if (warp == W_PWARPS) {
        // =========================================================== MMA ISSUER
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_bf16_mn(128, N_pad);
        const uint64_t d0 = make_mn_sw128_desc(smem_u32(smem), W_BLK, 1024);
        const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4, ximg16 = static_cast<uint32_t>(x_img) >> 4;
        const uint32_t g16off = static_cast<uint32_t>((SPLIT ? 2 : 1) * x_img) >> 4, blk16 = W_BLK >> 4;
        uint32_t g = 0;
        for (long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++g) {
            const uint32_t s = g % W_STAGES, ph = (g / W_STAGES) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            if (leader) {
                const uint64_t dx_hi = d0 + static_cast<uint64_t>(s * stage16), dx_lo = dx_hi + ximg16;
                const uint64_t dg_hi = dx_hi + g16off, dg_lo = dg_hi + blk16;
                for (int j = 0; j < mt; ++j) {
                    const uint32_t d_tmem = tmem_base + j * 64;
                    const uint64_t mo = static_cast<uint64_t>(j * 2 * blk16);          // M tile j = channel blocks 2j, 2j+1
#pragma unroll
                    for (int k = 0; k < W_KT / 16; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * 128);           // 16 pixel rows = 2048 B
                        const uint32_t acc = (g | k) != 0 ? 1u : 0u;
                        umma_bf16(d_tmem, dx_hi + mo + adv, dg_hi + adv, idesc, acc);
                        if (SPLIT) {
                            umma_bf16(d_tmem, dx_lo + mo + adv, dg_hi + adv, idesc, 1u);
                            umma_bf16(d_tmem, dx_hi + mo + adv, dg_lo + adv, idesc, 1u);
                        }
                    }
                }
                umma_commit(bar_empty + 8 * s);
            }
            __syncwarp();
        }
        if (leader) umma_commit(bar_done);
        __syncwarp();
    } else {
        // =========================================================== EPILOGUE (once): D[c, n] -> atomicAdd dW[n, c]
        const int q = warp & 3;
        mbar_wait(bar_done, 0);
        __syncwarp();
        tc_fence_after();
        if (blockIdx.x < a.ntiles) {
            for (int j = 0; j < mt; ++j) {
                const int c = j * 128 + q * 32 + lane;
                for (int g16 = 0; g16 < N_pad; g16 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + j * 64 + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g16), v);
                    if (c < a.C) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (g16 + e < a.N) atomicAdd(a.dW + static_cast<long>(g16 + e) * a.dw_stride + c, v[e]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_PWARPS) {
        __syncwarp();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace

// C > 256 (block 2 layers 14-16, block 3 layers 10-16: up to 342 input channels) runs as two passes over channel ranges of <= 256 -- the
// 48-channel G operand is re-read, the slab columns are not -- instead of falling back to the SIMT kernel (round 1: 10 calls, 10.6 ms).
bool eml_wgrad1x1_tc_supported(int N, int C, int pool, long M) {
    return !pool && N <= 64 && C <= 512 && M >= W_KT && !eml_env_flag("EML_WGRAD_SIMT");
}

static int wgrad1x1_tc_range(const float *G, int g_pitch, int N, const float *x, int x_pitch, int C, const float *scale, const float *shift,
                             int relu, float *dW, int dw_stride, long M, int precision, cudaStream_t st);

int eml_wgrad1x1_tc(const float *G, int g_pitch, int N, const float *x, int x_pitch, int C, const float *scale, const float *shift,
                    int relu, float *dW, long M, int precision, cudaStream_t st) {
    for (int c0 = 0; c0 < C; c0 += 256) {
        const int cc = C - c0 < 256 ? C - c0 : 256;
        const int rc = wgrad1x1_tc_range(G, g_pitch, N, x + c0, x_pitch, cc, scale ? scale + c0 : nullptr, shift ? shift + c0 : nullptr, relu,
                                         dW + c0, C, M, precision, st);
        if (rc != EML_OK) return rc;
    }
    return EML_OK;
}

static int wgrad1x1_tc_range(const float *G, int g_pitch, int N, const float *x, int x_pitch, int C, const float *scale, const float *shift,
                             int relu, float *dW, int dw_stride, long M, int precision, cudaStream_t st) {
    WArgs a{};
    a.G = G; a.g_pitch = g_pitch; a.N = N; a.x = x; a.x_pitch = x_pitch; a.C = C; a.scale = scale; a.shift = shift; a.dW = dW;
    a.dw_stride = dw_stride;
    a.M = M; a.relu = relu;
    a.nblk = 2 * ((C + 127) / 128);
    a.ntiles = (M + W_KT - 1) / W_KT;
    const bool split = precision != EML_PREC_BF16;
    const size_t smem = static_cast<size_t>(W_STAGES) * (split ? 2 : 1) * (a.nblk + 1) * W_BLK + 1024;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = static_cast<unsigned>(a.ntiles < sms ? a.ntiles : sms);
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(wgrad1x1_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        wgrad1x1_tc_kernel<true><<<grid, W_THREADS, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(wgrad1x1_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        wgrad1x1_tc_kernel<false><<<grid, W_THREADS, smem, st>>>(a);
    }
    return eml_launch_status();
}
