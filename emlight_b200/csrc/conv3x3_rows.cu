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

    // ---- 9 taps x 3 k-steps: shifted descriptor windows over the staging buffer
    if (tid == 0) {
        mbar_wait(bar_w, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(ROWS_TILE, a.N_pad);
        const uint32_t wlbo = static_cast<uint32_t>(a.N_pad * 16);         // bytes between weight k-chunks
        const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), w_s = smem_u32(w_sm);
        uint32_t acc = 0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int ks = 0; ks < KC / 2; ++ks) {
                const uint32_t aoff = static_cast<uint32_t>(((dy * KC + 2 * ks) * ROWS_PP + dx) * 16);
                const uint32_t woff = static_cast<uint32_t>((tap * KC + 2 * ks) * a.N_pad * 16);
                const uint64_t dah = make_nosw_desc(a_hi_s + aoff, PLANE, 128);
                const uint64_t dbh = make_nosw_desc(w_s + woff, wlbo, 128);
                umma_bf16(tmem_base, dah, dbh, idesc, acc);
                acc = 1;
                if (SPLIT) {
                    const uint64_t dal = make_nosw_desc(a_lo_s + aoff, PLANE, 128);
                    const uint64_t dbl = make_nosw_desc(w_s + w_img + woff, wlbo, 128);
                    umma_bf16(tmem_base, dal, dbh, idesc, 1u);
                    umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                }
            }
        }
        umma_commit(bar_acc);
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
size_t eml_rows_wpack_bytes(int C_out, int C_in) {
    if (C_in % 8) return 0;
    const int N_pad = (C_out + 15) & ~15;
    return static_cast<size_t>(2) * 9 * (C_in / 8) * N_pad * 16;
}

int eml_rows_pack(const float *w_oihw, unsigned char *dst, int C_out, int C_in, cudaStream_t st) {
    const int N_pad = (C_out + 15) & ~15;
    pack_planar_kernel<<<32, 256, 0, st>>>(w_oihw, dst, C_out, C_in, N_pad);
    return eml_launch_status();
}

bool eml_rows_supported(const eml_conv_params *p) {
    return p->mode == EML_CONV_3x3 && p->C_in == 48 && p->C_out <= 16 && (p->W % ROWS_TILE) == 0 &&
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
