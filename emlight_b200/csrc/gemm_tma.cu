// TMA-fed tcgen05 GEMM for the compute-bound GenProjector convolutions (sm_100a).
//
//   out[m, choff + n] = sum_k A[m, k] * W[n, k]  (+ bias[n])        A = im2col'ed activations, W = 3x3 filters as (N, 9*Cp)
//
// Unlike the DenseNet convolutions (HBM-bound, per-layer BN in the producer) these GEMMs have K = 9*C_in up to 9216 and
// N up to 1024: they are tensor-bound.  The LUT gather (spade_ops.cu) therefore writes A once, already split into
// bf16 hi / lo matrices, and this kernel is a canonical Blackwell pipeline:
//   warp 0   LOADER    per 64-wide K chunk: two cp.async.bulk.tensor.2d (A_hi, A_lo boxes 64 x 128, SWIZZLE_128B, through
//                      CUtensorMaps built on the host) + one cp.async.bulk of the packed weight chunk, all completing on the
//                      stage's mbarrier (expect_tx)
//   warp 1   MMA       one elected lane: per k-step 1 (bf16) or 3 (bf16x3: hi*hi + lo*hi + hi*lo) tcgen05.mma M128 x N<=256 x K16
//                      into a double-buffered TMEM accumulator (2 x 256 columns = all of TMEM); tcgen05.commit frees the stage
//   warps 2-5 EPILOGUE tcgen05.ld -> (+bias) -> row-contiguous float4 stores
// Persistent: one CTA per SM walks the (m-tile, n-slice) list.  At N=256 the MMA (128 cycles of math per instruction)
// outlasts its shared-memory operand reads (96 wavefronts), so the tensor pipe -- not memory -- is the roof.
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int G_TILE_M = 128;
constexpr int G_CHUNK_K = 64;
constexpr int G_THREADS = 32 * 6;
constexpr int G_A_TILE = G_TILE_M * G_CHUNK_K * 2;       // 16 KB

struct GArgs {
    const unsigned char *wpack;     // [chunk][hi | lo] images of N_pad x 128 bytes (eml_conv_pack_weights layout)
    const float *bias;
    float *out;
    long M;
    int N, N_pad, out_pitch, out_choff;
    int nchunks, stages;
    long mtiles;
    int ksplit, cps;                // split-K: work item = (m-tile, K range of cps chunks); partial sums are added with red.global
    int nslices;                    // output-channel slices of N columns each in ONE launch: work item = (slice, K range, m-tile);
    long slice_bytes;               // slice s reads wpack + s * slice_bytes, bias + s * N and writes at out_choff + s * N
};

__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}

template <bool SPLIT>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_tma_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                 const __grid_constant__ CUtensorMap tm_lo, const GArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2 * 6 + 4];
    __shared__ uint32_t s_tmem;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int b_tile = a.N_pad * 128;
    const int stage_bytes = (SPLIT ? 2 : 1) * (G_A_TILE + b_tile);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[6]);
    const uint32_t bar_accfull = smem_u32(&s_bar[12]), bar_accempty = smem_u32(&s_bar[14]);
    const int acc_cols = a.N_pad <= 128 ? 128 : 256;

    if (tid == 0) {
        for (int s = 0; s < 6; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accempty + 8 * i, 4); }
        fence_mbar_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), static_cast<uint32_t>(2 * acc_cols));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ================================================================ LOADER
        const bool leader = elect_one();
        uint32_t g = 0;
        const long per_slice = a.mtiles * a.ksplit;
        for (long ts = blockIdx.x; ts < per_slice * a.nslices; ts += gridDim.x) {
            const long t = ts % per_slice;
            const unsigned char *wp = a.wpack + (ts / per_slice) * a.slice_bytes;
            const int m0 = static_cast<int>((t % a.mtiles) * G_TILE_M);
            const int c_lo = static_cast<int>(t / a.mtiles) * a.cps, c_hi = min(a.nchunks, c_lo + a.cps);
            for (int c = c_lo; c < c_hi; ++c, ++g) {
                const uint32_t s = g % a.stages, ph = (g / a.stages) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                if (leader) {
                    unsigned char *st = smem + static_cast<size_t>(s) * stage_bytes;
                    const uint32_t bar = bar_full + 8 * s;
                    mbar_expect_tx(bar, static_cast<uint32_t>(stage_bytes));
                    tma_load_2d(smem_u32(st), &tm_hi, c * G_CHUNK_K, m0, bar);
                    if (SPLIT) tma_load_2d(smem_u32(st + G_A_TILE), &tm_lo, c * G_CHUNK_K, m0, bar);
                    bulk_g2s(smem_u32(st + (SPLIT ? 2 : 1) * G_A_TILE), wp + static_cast<size_t>(c) * 2 * b_tile,
                             static_cast<uint32_t>((SPLIT ? 2 : 1) * b_tile), bar);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA ISSUER
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_bf16(G_TILE_M, a.N_pad);
        const uint64_t d0 = make_sw128_desc(smem_u32(smem));
        const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4, alo16 = G_A_TILE >> 4;
        const uint32_t bhi16 = static_cast<uint32_t>((SPLIT ? 2 : 1) * G_A_TILE) >> 4, blo16 = static_cast<uint32_t>(b_tile) >> 4;
        uint32_t g = 0, j = 0;
        const long per_slice = a.mtiles * a.ksplit;
        for (long ts = blockIdx.x; ts < per_slice * a.nslices; ts += gridDim.x, ++j) {
            const long t = ts % per_slice;
            const uint32_t buf = j & 1, aph = (j >> 1) & 1;
            mbar_wait(bar_accempty + 8 * buf, aph ^ 1);
            const uint32_t d_tmem = tmem_base + buf * acc_cols;
            const int c_lo = static_cast<int>(t / a.mtiles) * a.cps, c_hi = min(a.nchunks, c_lo + a.cps);
            for (int c = c_lo; c < c_hi; ++c, ++g) {
                const uint32_t s = g % a.stages, ph = (g / a.stages) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                if (leader) {
                    const uint64_t da_hi = d0 + static_cast<uint64_t>(s * stage16), da_lo = da_hi + alo16;
                    const uint64_t db_hi = da_hi + bhi16, db_lo = db_hi + blo16;
#pragma unroll
                    for (int k = 0; k < G_CHUNK_K / 16; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        umma_bf16(d_tmem, da_hi + adv, db_hi + adv, idesc, (c != c_lo || k != 0) ? 1u : 0u);
                        if (SPLIT) {
                            umma_bf16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
                            umma_bf16(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                        }
                    }
                    umma_commit(bar_empty + 8 * s);
                    if (c == c_hi - 1) umma_commit(bar_accfull + 8 * buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ EPILOGUE (warps 2..5 -> TMEM lane quarter warp % 4)
        const int q = warp & 3;
        const bool vec_ok = ((a.out_pitch | a.out_choff) & 3) == 0;
        uint32_t j = 0;
        const long per_slice = a.mtiles * a.ksplit;
        for (long ts = blockIdx.x; ts < per_slice * a.nslices; ts += gridDim.x, ++j) {
            const long t = ts % per_slice;
            const int n_off = static_cast<int>(ts / per_slice) * a.N;
            const float *bias = a.bias ? a.bias + n_off : nullptr;
            const uint32_t buf = j & 1, aph = (j >> 1) & 1;
            mbar_wait(bar_accfull + 8 * buf, aph);
            __syncwarp();
            tc_fence_after();
            const bool first_k = t < a.mtiles;                  // the K range that also adds the bias
            const long m = (t % a.mtiles) * G_TILE_M + q * 32 + lane;
            const bool row_ok = m < a.M;
            float *orow = a.out + (row_ok ? m : 0) * a.out_pitch + a.out_choff + n_off;
            for (int g16 = 0; g16 < a.N_pad; g16 += 16) {
                float v[16];
                tmem_ld16(tmem_base + buf * acc_cols + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g16), v);
                if (row_ok) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int n = g16 + qq * 4;
                        float4 o = make_float4(v[qq * 4], v[qq * 4 + 1], v[qq * 4 + 2], v[qq * 4 + 3]);
                        if (bias != nullptr && first_k) {
                            if (n < a.N) o.x += bias[n];
                            if (n + 1 < a.N) o.y += bias[n + 1];
                            if (n + 2 < a.N) o.z += bias[n + 2];
                            if (n + 3 < a.N) o.w += bias[n + 3];
                        }
                        if (a.ksplit > 1) {                     // partial sum of one K range: accumulate (out zeroed by the caller)
                            if (n < a.N) atomicAdd(orow + n, o.x);
                            if (n + 1 < a.N) atomicAdd(orow + n + 1, o.y);
                            if (n + 2 < a.N) atomicAdd(orow + n + 2, o.z);
                            if (n + 3 < a.N) atomicAdd(orow + n + 3, o.w);
                        } else if (vec_ok && n + 3 < a.N) {
                            *reinterpret_cast<float4 *>(orow + n) = o;
                        } else {
                            if (n < a.N) orow[n] = o.x;
                            if (n + 1 < a.N) orow[n + 1] = o.y;
                            if (n + 2 < a.N) orow[n + 2] = o.z;
                            if (n + 3 < a.N) orow[n + 3] = o.w;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) g_mbar_arrive(bar_accempty + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(2 * acc_cols));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap *map, const void *base, long M, int Kp) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return EML_E_ARG;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(M)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kp) * 2};
    const cuuint32_t box[2] = {G_CHUNK_K, G_TILE_M};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? EML_OK : EML_E_ARG;
}

}  // namespace

static int gemm_impl(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, int N, const float *bias,
                     float *out, int out_pitch, int out_choff, int precision, int ksplit, void *stream, int nslices = 1,
                     long slice_bytes = 0) {
    EML_CHECK_PTR(A_hi); EML_CHECK_PTR(wpack); EML_CHECK_PTR(out);
    EML_CHECK_ALIGN16(A_hi); EML_CHECK_ALIGN16(wpack);
    if (M <= 0 || Kp <= 0 || (Kp % G_CHUNK_K) || N <= 0 || N > 256 || nslices < 1 || out_choff < 0 ||
        out_pitch < out_choff + static_cast<long>(N) * nslices)
        return EML_E_SHAPE;
    if (nslices > 1 && ((N & 15) || (slice_bytes & 15) || slice_bytes < static_cast<long>(Kp / G_CHUNK_K) * 2 * N * 128)) return EML_E_SHAPE;
    const bool split = precision == EML_PREC_BF16X3;
    if (!split && precision != EML_PREC_BF16) return EML_E_ARG;
    if (split) { EML_CHECK_PTR(A_lo); EML_CHECK_ALIGN16(A_lo); }
    CUtensorMap tm_hi, tm_lo;
    int rc = make_map(&tm_hi, A_hi, M, Kp);
    if (rc != EML_OK) return rc;
    rc = make_map(&tm_lo, split ? A_lo : A_hi, M, Kp);
    if (rc != EML_OK) return rc;
    GArgs a{};
    a.wpack = static_cast<const unsigned char *>(wpack); a.bias = bias; a.out = out; a.M = M;
    a.N = N; a.N_pad = (N + 15) & ~15; a.out_pitch = out_pitch; a.out_choff = out_choff;
    a.nslices = nslices; a.slice_bytes = slice_bytes;
    a.nchunks = Kp / G_CHUNK_K;
    a.mtiles = (M + G_TILE_M - 1) / G_TILE_M;
    if (ksplit < 1 || ksplit > a.nchunks) return EML_E_ARG;
    a.cps = (a.nchunks + ksplit - 1) / ksplit;
    a.ksplit = (a.nchunks + a.cps - 1) / a.cps;            // no empty K ranges
    const int stage_bytes = (split ? 2 : 1) * (G_A_TILE + a.N_pad * 128);
    int stages = (225 * 1024) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) return EML_E_SHAPE;
    a.stages = stages;
    const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long items = a.mtiles * a.ksplit * nslices;
    const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(gemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        gemm_tma_kernel<true><<<grid, G_THREADS, smem, st>>>(tm_hi, tm_lo, a);
    } else {
        e = cudaFuncSetAttribute(gemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        gemm_tma_kernel<false><<<grid, G_THREADS, smem, st>>>(tm_hi, tm_lo, a);
    }
    return eml_launch_status();
}

extern "C" int eml_gemm_bf16(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, int N, const float *bias,
                             float *out, int out_pitch, int out_choff, int precision, void *stream) {
    return gemm_impl(A_hi, A_lo, M, Kp, wpack, N, bias, out, out_pitch, out_choff, precision, 1, stream);
}

// `nslices` output slices of exactly N (a multiple of 16, <= 256) columns in ONE launch: slice s multiplies the same A by the
// packed weights at wpack + s * slice_bytes and writes columns [out_choff + s*N, out_choff + (s+1)*N).  Wide layers at low
// resolution (M = B*H*W small, O = 512 / 1024) otherwise run as O/256 launches of ceil(M/128) CTAs each, one after the other.
// ksplit > 1 additionally deals the K chunks to `ksplit` work items per (slice, tile) as eml_gemm_bf16_splitk does (`out` ZERO on entry).
extern "C" int eml_gemm_bf16_slices(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, long slice_bytes, int nslices,
                                    int N, const float *bias, float *out, int out_pitch, int out_choff, int precision, int ksplit,
                                    void *stream) {
    return gemm_impl(A_hi, A_lo, M, Kp, wpack, N, bias, out, out_pitch, out_choff, precision, ksplit, stream, nslices, slice_bytes);
}

// Split-K variant for short-and-deep products (needlet projection: M = 3 B rows, K = 32768 pixels): the K chunks are dealt to
// `ksplit` work items per m-tile and the partial sums are accumulated with float atomics, so `out` must be ZERO on entry and the
// result is reproducible only to fp32 summation order.
extern "C" int eml_gemm_bf16_splitk(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, int N, const float *bias,
                                    float *out, int out_pitch, int out_choff, int precision, int ksplit, void *stream) {
    return gemm_impl(A_hi, A_lo, M, Kp, wpack, N, bias, out, out_pitch, out_choff, precision, ksplit, stream);
}
