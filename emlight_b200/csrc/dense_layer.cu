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
//     three Z accumulators (tcgen05.st / tcgen05.ld, lane-private, no shared-memory traffic);
//   * the completed row U[x, (dx,o)] goes through one shared-memory row buffer, where y[x,o] = U[x-1,0,o] + U[x,1,o] + U[x+1,2,o]
//     + bias is formed and written with coalesced float4 stores into the slab at channel offset C_in.
//
// Operand pipeline (measured on B200 with tools/micro/*: register-staged LDG loads from one CTA per SM top out near 4.2 TB/s
// whatever the depth, TMA boxes of 128-byte rows reach 5.5-6.4 TB/s from a 4-8 stage ring, 64-byte rows only 3.7 TB/s):
//   warp 16     TMA     one cp.async.bulk.tensor.2d per stage: box = 32 channels x 128 pixels of RAW fp32 slab (128-byte rows,
//                       SWIZZLE_128B, channels past C_in zero-filled by the tensor map), ring of 16 KB stages
//   warps 0-15  CONVERT four warpgroups take stages round-robin and convert IN PLACE: a raw 128-byte row (32 fp32) becomes
//                       [32 bf16 hi | 32 bf16 lo] = the same 128 bytes, i.e. the stage turns into a K-major SWIZZLE_128B tile whose
//                       K columns 0-31 are the hi parts and 32-63 the lo parts (norm1 affine, ReLU folded into
//                       cvt.rz.relu.bf16x2).  The 8 lanes that own one row read before any of them writes (__syncwarp).
//   warp 17     MMA     per 16-channel k-step three tcgen05.mma (hi*Bhi, lo*Bhi, hi*Blo): A descriptors are the stage base
//                       + 32k bytes (hi) / + 64 + 32k bytes (lo); composite weights resident in shared memory
//   warps 20-27 EPILOGUE one warpgroup per half row (stencil above)
//
// TS = true (default since round 2): the A operand lives in TENSOR MEMORY instead.  ncu on the smem-A version (profiles/
// r01_ncu_full_dense_layer_v8.md) shows the shared-memory pipe as the roof: per 16 KB stage the converters read 16 KB and write 16 KB
// (+ scale/shift reads) and the three MMAs per k-step re-read A (24 KB) and B (21 KB) -- 93 KB of shared-memory traffic per 16 KB of
// HBM data.  With TS a converter THREAD owns one pixel row: it reads its 128-byte raw row (8 conflict-free LDS.128 through the
// 128-byte swizzle), applies norm1 + ReLU, splits to bf16 hi/lo and writes the packed pairs with tcgen05.st into a ring of four
// 32-column TMEM slots ([hi k0 | hi k1 | lo k0 | lo k1] x 8 columns; lane = pixel, column = channel pair -- layout validated by
// tools/micro/ts_mma_test.cu); the MMAs take A from TMEM (tcgen05.mma [d], [a], b-desc).  Shared-memory traffic per stage drops to
// 16 KB TMA write + 16 KB LDS + 21 KB B reads, the shared-memory stage is released as soon as it has been READ (not after the MMAs),
// and the unused half of a partial last stage is skipped.  TMEM: 2 Z accumulators (224) + 4 partial-row slots (144) + A ring (128).
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace eml;

constexpr int F_TILE_M = 128;
constexpr int F_STAGE_C = 32;                           // channels per ring stage
constexpr int F_STAGE_BYTES = F_TILE_M * F_STAGE_C * 4; // 16 KB: raw fp32 in, [hi | lo] bf16 out
constexpr int F_MAX_STAGES = 12;
// Warp roles for NCW converter warps (16 = four warpgroups; 8 = two, TS only: 640 threads leave 102 registers per thread, which lets
// the epilogue keep 8-column TMEM pieces in flight): converters 0..NCW-1, TMA warp NCW, MMA warp NCW+1, two idle warps (keeps the
// epilogue warpgroups 4-aligned), epilogue NCW+4 .. NCW+11.
constexpr int f_threads(int ncw) { return (ncw + 12) * 32; }
constexpr int F_G = 12;                                 // growth rate (output channels)
constexpr int F_GRP = 3 * F_G;                          // 36 columns per dy group, ordered (dx, o)
constexpr int F_NPAD = 112;
constexpr int F_ZSTRIDE = 112;                          // TMEM columns per Z buffer
constexpr int F_USTRIDE = 36;                           // [slot(2)][half(2)] x 36 columns
constexpr int F_NA = 4;                                 // TS: TMEM A-operand slots (one per converter warpgroup), 32 columns each
constexpr int F_MAX_NZ = 3;
// TMEM map.  smem-A: 3 Z (336) + partial rows (144) = 480.  TS: 2 Z (224) + partial rows (144) + A ring 4 x 32 (128) = 496 of 512.
template <bool TS> struct FTmem {
    static constexpr int NZ = TS ? 2 : 3;               // Z accumulator buffers (tile j uses buffer j % NZ)
    static constexpr int UBASE = NZ * F_ZSTRIDE;        // partial-row slots
    static constexpr int ABASE = UBASE + 4 * F_USTRIDE; // TS only: A ring
};
constexpr int F_MAX_C = 320;                            // 5 weight chunks of 64 channels
constexpr int F_WCHUNK = 2 * F_NPAD * 128;              // POOL (transition) weights: bytes of one eml_conv_pack_weights chunk [hi | lo] (64 channels, SWIZZLE_128B)
// Dense-layer weights (eml_dense_layer_compose): one UNIT per 32-channel stage, [hi plane | lo plane], each plane 120 rows x 32 bf16 in the
// K-major NO-SWIZZLE core-matrix layout: element (n, c) at (c / 8) * (120 * 16) + n * 16 + (c % 8) * 2 bytes (LBO = 1920, SBO = 128).
// 32-channel granularity: C_in = 144 keeps 75 KB resident instead of three 64-channel chunks = 84 KB -- shared memory that goes to the TMA ring.
// The plane has 120 rows: rows 0-107 = (dy, dx, o), 12 zero rows behind them, because the row-sum kernels read one 36-row tap group
// (dy) as an N = 48 operand -- rows are uniformly 16 B apart (SBO = 128), so a group is just a start-address shift of 36 * 16 B.
constexpr int F_WROWS = 120;
constexpr int F_WPLANE = F_WROWS * 64;                  // 7680
constexpr int F_WUNIT = 2 * F_WPLANE;                   // 15360
constexpr int F_RN = 48;                                // row-sum kernels: MMA N (36 used) = TMEM columns per output-row accumulator
constexpr int F_NT = 8;                                 // row-sum kernels: ring of "tile done" barriers
// floats per pixel in the row buffer: template parameter SROW = 36 (packed) or 44 (176 B = 48 B mod 128: the (pixel, quad) stream of the
// output pass and the per-pixel stores become bank-conflict free; chosen by the host whenever it does not cost a ring stage)

struct FArgs {
    const float *scale;
    const float *shift;
    const unsigned char *wpack;
    const float *bias9;          // (3 row classes, 3 column classes, 12)
    float *out;
    int B, H, W, R;              // R = output rows per band (H % R == 0)
    int C_in, out_pitch, out_choff;
    int nwchunks, nstg, stages;  // 64-channel weight chunks (POOL), 32-channel stages per tile, ring depth
    int wbytes;                  // resident weight bytes (dense: nstg units of F_WUNIT; POOL: nwchunks chunks of F_WCHUNK)
    int wide;                    // 0: store the 12 new channels (48 B per pixel); 1: also zero the 4 channels after them (64 B);
                                 // 2: channels start mid-sector (offset = 4 mod 8): re-store the 4 channels IN FRONT with them (64 B, two
                                 //    full sectors) from a stash of this layer's own raw input that the converters fill (TS, in == out)
    int stash_rows;              // wide == 2: image rows in the stash ring
    long nbands;
    int pool;                    // transition mode (eml_transition_forward): tile = a PAIR of image rows accumulated into one Z buffer by
                                 // the tensor core (vertical half of the 2x2 average), epilogue = horizontal pair sum, x 0.25, N channels out
    int n_out;                   // pool mode: output channels
    int pair;                    // W == 64: a tile is row r of image 2k (pixels 0-63) next to row r of image 2k+1 (pixels 64-127)
    long gstride;                // > 0: channel-PLANE input ("grouped" slab): plane g = channels [32g, 32g+32) of every pixel as 128-byte
                                 // rows, planes gstride pixels apart -- a stage's box is one contiguous 16 KB run.  0: NHWC pixel records.
    long ogstride;               // the same for the output (dense mode, wide == 3)
};

__device__ __forceinline__ void f_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void f_tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
// TMEM <-> registers, 32 lanes x N columns.  Loads are issued WITHOUT waiting so that several can be in flight; the values are
// valid after f_tmem_wait_ld(v) on the same array (which also ties the registers to the wait for the compiler).
template <int N> __device__ __forceinline__ void f_tmem_ld(uint32_t taddr, float (&v)[N]);
template <> __device__ __forceinline__ void f_tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void f_tmem_ld<4>(uint32_t taddr, float (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr) : "memory");
}
template <int N> __device__ __forceinline__ void f_tmem_wait_ld(float (&v)[N]);
template <> __device__ __forceinline__ void f_tmem_wait_ld<8>(float (&v)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]) :: "memory");
}
template <> __device__ __forceinline__ void f_tmem_wait_ld<4>(float (&v)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
}
template <int N> __device__ __forceinline__ void f_tmem_st(uint32_t taddr, const float (&v)[N]);
template <> __device__ __forceinline__ void f_tmem_st<8>(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
template <> __device__ __forceinline__ void f_tmem_st<4>(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

__device__ __forceinline__ void f_tmem_ld16(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void f_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void f_tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: A = 128 lanes (rows) x 8 columns (16 bf16, pair (2c, 2c+1) in column c, low half first)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

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

// norm1 affine + ReLU + bf16 hi/lo split of 4 consecutive channels.
template <bool SPLIT>
__device__ __forceinline__ void f_convert_quad(float4 v, float4 sc, float4 sh, uint2 &hv, uint2 &lv) {
    const float o0 = fmaf(v.x, sc.x, sh.x), o1 = fmaf(v.y, sc.y, sh.y), o2 = fmaf(v.z, sc.z, sh.z), o3 = fmaf(v.w, sc.w, sh.w);
    if (SPLIT) {
        hv.x = cvt_rz_relu_bf16x2(o0, o1);
        hv.y = cvt_rz_relu_bf16x2(o2, o3);
        lv.x = cvt_rn_relu_bf16x2(o0 - __uint_as_float(hv.x << 16), o1 - __uint_as_float(hv.x & 0xffff0000u));
        lv.y = cvt_rn_relu_bf16x2(o2 - __uint_as_float(hv.y << 16), o3 - __uint_as_float(hv.y & 0xffff0000u));
    } else {
        hv.x = cvt_rn_relu_bf16x2(o0, o1);
        hv.y = cvt_rn_relu_bf16x2(o2, o3);
        lv = make_uint2(0u, 0u);
    }
}

// One slice of N accumulator columns of the row stencil (see the file header): five TMEM loads in flight, one wait.
//   emit:      srow[off..]  = Z[dy=2] + U[rho-1]           (output row rho-1 complete)
//   init_next: U[rho+1]     = Z[dy=0]                       (overwrites the slot U[rho-1] was read from)
//   upd:       U[rho]       = Z[dy=1] (+ U[rho] when upd_add)
template <int N>
__device__ __forceinline__ void f_stencil_piece(uint32_t zc, uint32_t us0, uint32_t us1, uint32_t off, float *srow, bool emit,
                                                bool init_next, bool upd, bool upd_add) {
    float z2[N], u0[N], z0[N], z1[N], u1[N];
    if (emit) { f_tmem_ld<N>(zc + 2 * F_GRP + off, z2); f_tmem_ld<N>(us0 + off, u0); }
    if (init_next) f_tmem_ld<N>(zc + off, z0);
    if (upd) { f_tmem_ld<N>(zc + F_GRP + off, z1); if (upd_add) f_tmem_ld<N>(us1 + off, u1); }
    if (emit) {
        f_tmem_wait_ld<N>(z2); f_tmem_wait_ld<N>(u0);
#pragma unroll
        for (int e = 0; e < N; e += 4)
            *reinterpret_cast<float4 *>(srow + off + e) = make_float4(z2[e] + u0[e], z2[e + 1] + u0[e + 1], z2[e + 2] + u0[e + 2], z2[e + 3] + u0[e + 3]);
    }
    if (init_next) { f_tmem_wait_ld<N>(z0); f_tmem_st<N>(us0 + off, z0); }
    if (upd) {
        f_tmem_wait_ld<N>(z1);
        if (upd_add) {
            f_tmem_wait_ld<N>(u1);
#pragma unroll
            for (int e = 0; e < N; ++e) z1[e] += u1[e];
        }
        f_tmem_st<N>(us1 + off, z1);
    }
}

// Walks the CTA's bands -> row tiles in the order every role agrees on.
struct BandIter {
    long band;
    long m0;         // first pixel (b*H*W + r*W + x) of the current tile
    int nt, t;       // tiles in the band, current tile
    bool valid;
};
template <bool POOL>
__device__ __forceinline__ void band_init(BandIter &it, const FArgs &a, long band) {
    it.band = band;
    it.valid = band < a.nbands;
    it.t = 0; it.nt = 0; it.m0 = 0;
    if (!it.valid) return;
    const int bpi = a.H / a.R;
    const long img = (band / bpi) * (a.pair ? 2 : 1);     // first image of the band
    const int r0 = static_cast<int>(band % bpi) * a.R;
    const int lo = POOL ? r0 : max(r0 - 1, 0), hi = POOL ? r0 + a.R - 1 : min(r0 + a.R, a.H - 1);
    it.nt = (POOL ? (hi - lo + 1) / 2 : hi - lo + 1) * (a.pair ? 1 : a.W / F_TILE_M);
    it.m0 = (img * a.H + lo) * a.W;
}

template <bool SPLIT, bool POOL, int F_SROW, bool TS, int NCW = 16, int PW = 4, bool RSUM = false>
__global__ void __launch_bounds__(f_threads(NCW), 1) dense_layer_kernel(const __grid_constant__ CUtensorMap tmap, const FArgs a) {
    constexpr int F_NZ = FTmem<TS>::NZ, F_UBASE = FTmem<TS>::UBASE;
    // row-sum kernels: 4 output rows x 2 half rows x 48 columns of accumulators (384), then the A ring (128) = all 512 columns
    constexpr int F_ABASE = (TS && !POOL && RSUM) ? 8 * F_RN : FTmem<TS>::ABASE;
    constexpr int F_CWARPS = NCW, F_TMA_WARP = NCW, F_MMA_WARP = NCW + 1, F_EPI_WARP0 = NCW + 4, F_THREADS = f_threads(NCW);
    constexpr int NCWG = NCW / 4;                          // converter warpgroups
    static_assert(NCW == 16 || (TS && NCW == 8), "converter warps: 16, or 8 with the TMEM-resident A operand");
    static_assert(PW == 4 || PW == 8, "stencil piece width");
    extern __shared__ unsigned char smem_raw[];
    constexpr bool RS = TS && !POOL && RSUM;               // row-sum design: the MMAs add the three vertical taps into per-output-row accumulators
    __shared__ __align__(8) unsigned long long s_bar[3 * F_MAX_STAGES + 1 + 2 * F_MAX_NZ + F_NA + F_NT + 8];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_scale[F_MAX_C], s_shift[F_MAX_C], s_bias[9 * F_G + 4];

    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int NST = a.stages;
    unsigned char *w_sm = smem + NST * F_STAGE_BYTES;
    float *s_row = reinterpret_cast<float *>(w_sm + static_cast<size_t>(a.wbytes));                // [(W + 2) pixels][36]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TPR = a.pair ? 1 : a.W / F_TILE_M;           // tiles per image row (1 or 2)

    const uint32_t bar_full = smem_u32(&s_bar[0]);                          // TMA landed (raw fp32)
    const uint32_t bar_ready = smem_u32(&s_bar[F_MAX_STAGES]);              // converted to bf16 hi/lo
    const uint32_t bar_empty = smem_u32(&s_bar[2 * F_MAX_STAGES]);          // consumed by the MMAs
    const uint32_t bar_w = smem_u32(&s_bar[3 * F_MAX_STAGES]);
    const uint32_t bar_zfull = smem_u32(&s_bar[3 * F_MAX_STAGES + 1]);      // [F_NZ]
    const uint32_t bar_zempty = smem_u32(&s_bar[3 * F_MAX_STAGES + 1 + F_MAX_NZ]);   // [F_NZ]
    const uint32_t bar_afree = smem_u32(&s_bar[3 * F_MAX_STAGES + 1 + 2 * F_MAX_NZ]);    // [F_NA]  TS: MMAs that read TMEM A slot a are done
    const uint32_t bar_tfull = smem_u32(&s_bar[3 * F_MAX_STAGES + 1 + 2 * F_MAX_NZ + F_NA]);          // [F_NT] RS: all MMAs of a tile are done
    const uint32_t bar_rfree = smem_u32(&s_bar[3 * F_MAX_STAGES + 1 + 2 * F_MAX_NZ + F_NA + F_NT]);   // [8]    RS: a row accumulator was drained
    // TS reuses bar_ready[0 .. F_NA) as "A slot written" (4 converter warps arrive) and bar_empty[s] is armed by the converters (4 warps:
    // the raw stage has been read), not by the MMA commits

    if (tid == 0) {
        for (int s = 0; s < F_MAX_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_ready + 8 * s, 4); mbar_init(bar_empty + 8 * s, TS ? 4 : 1); }
        for (int i = 0; i < F_NA; ++i) mbar_init(bar_afree + 8 * i, 1);
        for (int i = 0; i < F_NT; ++i) mbar_init(bar_tfull + 8 * i, 1);
        for (int i = 0; i < 8; ++i) mbar_init(bar_rfree + 8 * i, 4);
        mbar_init(bar_w, 1);
        for (int i = 0; i < F_MAX_NZ; ++i) { mbar_init(bar_zfull + 8 * i, 1); mbar_init(bar_zempty + 8 * i, 4); }
        fence_mbar_init();
    }
    for (int i = tid; i < F_MAX_C; i += F_THREADS) {
        s_scale[i] = i < a.C_in ? a.scale[i] : 0.f;        // channels past C_in arrive as zeros (TMA fill) and stay exact zeros
        s_shift[i] = i < a.C_in ? a.shift[i] : 0.f;
    }
    if (a.bias9 != nullptr)
        for (int i = tid; i < 9 * F_G; i += F_THREADS) s_bias[i] = a.bias9[i];
    if (tid < F_GRP) {                                     // zero pixels left and right of the row (of both rows in pair mode)
        s_row[tid] = 0.f; s_row[(a.W + 1) * F_SROW + tid] = 0.f;
        if (a.pair) { s_row[(a.W + 2) * F_SROW + tid] = 0.f; s_row[(2 * a.W + 3) * F_SROW + tid] = 0.f; }
    }
    if (warp == F_MMA_WARP) {
        __syncwarp();
        tmem_alloc(smem_u32(&s_tmem), 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int grid = static_cast<int>(gridDim.x);

    if (TS && warp < F_CWARPS) {
        // =========================================================== CONVERTERS, TS: raw fp32 row (shared) -> bf16 hi/lo pairs (TMEM A slot)
        // warpgroup cg converts stages g with g % NCWG == cg into A slot g % 4; warp w4 = lane quarter, thread = pixel row of the tile.
        // The host makes the ring depth a multiple of NCWG, so a ring slot (g % NST) and an A slot (g % 4) always belong to the same
        // warpgroup: every phase of their mbarriers is observed in order by one waiter.
        const int cg = warp >> 2, w4 = warp & 3;
        const uint32_t row = static_cast<uint32_t>(w4 * 32 + lane);
        const uint32_t r7 = row & 7u;
        const uint32_t row_off = row * 128u;
        const uint32_t ring = smem_u32(smem);
        const uint32_t a_lane = tmem_base + F_ABASE + (static_cast<uint32_t>(w4 * 32) << 16);
        uint32_t total = 0;
        for (long band = blockIdx.x; band < a.nbands; band += grid) {
            BandIter it;
            band_init<POOL>(it, a, band);
            total += static_cast<uint32_t>(it.nt) * static_cast<uint32_t>(POOL ? 2 * a.nstg : a.nstg);
        }
        const uint32_t nst_u = static_cast<uint32_t>(NST), nstg_u = static_cast<uint32_t>(a.nstg);
        const int c16 = (a.C_in + 15) & ~15;                                     // channels the MMAs read (whole k-steps)
        // wide == 2: the raw values of the layer's last four input channels (= the four slab channels in front of its output) are kept per
        // pixel in a ring of image rows, from where the output pass re-stores them together with the 12 new channels as two full sectors
        const int js = (a.C_in - 4) / F_STAGE_C, qs = ((a.C_in - 4) % F_STAGE_C) >> 2;
        float4 *s_stash = reinterpret_cast<float4 *>(s_row + (a.W + 2) * F_SROW);
        long sb_band = blockIdx.x;
        uint32_t sb_tile0 = 0, sb_row0 = 0, sb_nt = 0;
        if (!POOL && a.wide == 2) { BandIter it; band_init<POOL>(it, a, sb_band); sb_nt = static_cast<uint32_t>(it.nt); }
        for (uint32_t g = static_cast<uint32_t>(cg); g < total; g += NCWG) {
            const uint32_t s = g % nst_u, ph = (g / nst_u) & 1, as = g & 3u, aph = (g >> 2) & 1;
            const uint32_t a_slot = a_lane + as * 32u;
            const int j = static_cast<int>(g % nstg_u);
            const int halves = c16 - j * F_STAGE_C > 16 ? 2 : 1;                  // a partial last stage holds one k-step only
            const uint32_t st = ring + s * F_STAGE_BYTES + row_off;
            const bool stash = !POOL && a.wide == 2 && j == js;
            float4 stash_v = make_float4(0.f, 0.f, 0.f, 0.f);
            mbar_wait(bar_full + 8 * s, ph);
            uint32_t hi[2][8], lo[2][8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h < halves) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 v;
                        const uint32_t chunk = static_cast<uint32_t>(4 * h + q);
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + ((chunk ^ r7) << 4)) : "memory");
                        if (stash && 4 * h + q == qs) stash_v = v;
                        const float4 sc = *reinterpret_cast<const float4 *>(&s_scale[j * F_STAGE_C + 16 * h + 4 * q]);
                        const float4 sh = *reinterpret_cast<const float4 *>(&s_shift[j * F_STAGE_C + 16 * h + 4 * q]);
                        uint2 hv, lv;
                        f_convert_quad<SPLIT>(v, sc, sh, hv, lv);
                        hi[h][2 * q] = hv.x; hi[h][2 * q + 1] = hv.y;
                        lo[h][2 * q] = lv.x; lo[h][2 * q + 1] = lv.y;
                    }
                }
            }
            if (stash) {                                                          // tile ordinal -> (row sequence number, half row)
                const uint32_t tg = g / nstg_u, tpr = static_cast<uint32_t>(TPR);
                while (tg >= sb_tile0 + sb_nt) {
                    sb_tile0 += sb_nt; sb_row0 += sb_nt / tpr; sb_band += grid;
                    BandIter it; band_init<POOL>(it, a, sb_band); sb_nt = static_cast<uint32_t>(it.nt);
                }
                const uint32_t tl = tg - sb_tile0;
                const uint32_t rseq = sb_row0 + tl / tpr, half = tl % tpr;
                s_stash[(rseq % static_cast<uint32_t>(a.stash_rows)) * a.W + half * F_TILE_M + row] = stash_v;
            }
            __syncwarp();                                                         // every lane's loads have been consumed
            if (lane == 0) f_mbar_arrive(bar_empty + 8 * s);                      // raw stage free: the TMA warp may refill it
            mbar_wait(bar_afree + 8 * as, aph ^ 1);                               // MMAs of stage g - 4 have read this A slot
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h < halves) {
                    f_tmem_st8u(a_slot + 8 * h, hi[h]);
                    if (SPLIT) f_tmem_st8u(a_slot + 16 + 8 * h, lo[h]);
                }
            }
            f_tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) f_mbar_arrive(bar_ready + 8 * as);
        }
    } else if (warp < F_CWARPS) {
        // =========================================================== CONVERTERS (in place, raw fp32 -> [bf16 hi | bf16 lo])
        // warpgroup cg owns ring slots s with s % 4 == cg.  lane = (channel octet cp, row sub-index rs): per step a warp converts
        // 8 rows x 32 channels, 4 steps per stage; the 4 lanes of a row finish reading before any of them writes (__syncwarp).
        // Rows are dealt so that the two rows of every quarter-warp differ in bits 0 and 2 of (row & 7): the 16-byte loads
        // (chunks {2cp, 2cp+1} ^ row) and the 16-byte hi / lo stores (chunks cp ^ row, (4+cp) ^ row) are then bank-conflict free.
        const int cg = warp >> 2, w4 = warp & 3;
        const int cp = lane & 3, rs = lane >> 2;
        const uint32_t r7 = static_cast<uint32_t>((0x63724150u >> (4 * rs)) & 7u);     // rs -> 0,5,1,4,2,7,3,6 : pairs (x, x^5)
        const uint32_t lane_row = static_cast<uint32_t>(w4) * 4096u + r7 * 128u;
        const uint32_t rd0 = lane_row + (((2u * cp) ^ r7) << 4), rd1 = lane_row + (((2u * cp + 1u) ^ r7) << 4);
        const uint32_t hi_o = lane_row + ((static_cast<uint32_t>(cp) ^ r7) << 4), lo_o = lane_row + (((4u + cp) ^ r7) << 4);
        const uint32_t ring = smem_u32(smem);
        uint32_t total = 0;                                 // stages this CTA processes (32-bit: the modulo arithmetic below is per stage)
        for (long band = blockIdx.x; band < a.nbands; band += grid) {
            BandIter it;
            band_init<POOL>(it, a, band);
            total += static_cast<uint32_t>(it.nt) * static_cast<uint32_t>(POOL ? 2 * a.nstg : a.nstg);
        }
        // A ring slot always belongs to the same warpgroup (slot % 4): every phase of its mbarriers is then observed in order by
        // one waiter.  (Round-robin over the global stage number would let a warpgroup skip phases of a slot when the ring depth
        // is not a multiple of 4, and a parity wait cannot tell phase n from phase n + 2.)
        const uint32_t nst_u = static_cast<uint32_t>(NST), nstg_u = static_cast<uint32_t>(a.nstg);
        for (uint32_t g = 0; g < total; ++g) {
            const int s = static_cast<int>(g % nst_u);
            if ((s & 3) != cg) continue;
            const uint32_t ph = (g / nst_u) & 1;
            const int j = static_cast<int>(g % nstg_u);                              // (pool mode: 2 nstg stages per tile, same slices twice)                 // channel slice of this stage (pool mode: two rows per tile)
            const float4 sc0 = *reinterpret_cast<const float4 *>(&s_scale[j * F_STAGE_C + 8 * cp]);
            const float4 sc1 = *reinterpret_cast<const float4 *>(&s_scale[j * F_STAGE_C + 8 * cp + 4]);
            const float4 sh0 = *reinterpret_cast<const float4 *>(&s_shift[j * F_STAGE_C + 8 * cp]);
            const float4 sh1 = *reinterpret_cast<const float4 *>(&s_shift[j * F_STAGE_C + 8 * cp + 4]);
            const uint32_t st = ring + static_cast<uint32_t>(s) * F_STAGE_BYTES;
            mbar_wait(bar_full + 8 * s, ph);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t o = st + static_cast<uint32_t>(i) * 1024u;
                float4 v0, v1;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(o + rd0) : "memory");
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(o + rd1) : "memory");
                uint2 h0, l0, h1, l1;
                f_convert_quad<SPLIT>(v0, sc0, sh0, h0, l0);
                f_convert_quad<SPLIT>(v1, sc1, sh1, h1, l1);
                __syncwarp();                                                    // every lane of these rows has read its chunks
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + hi_o), "r"(h0.x), "r"(h0.y), "r"(h1.x), "r"(h1.y) : "memory");
                if (SPLIT) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + lo_o), "r"(l0.x), "r"(l0.y), "r"(l1.x), "r"(l1.y) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) f_mbar_arrive(bar_ready + 8 * s);
        }
    } else if (warp == F_TMA_WARP) {
        // =========================================================== TMA ISSUER
        if (lane == 0) {
            uint32_t g = 0;
            const uint32_t nst_u = static_cast<uint32_t>(NST);
            for (long band = blockIdx.x; band < a.nbands; band += grid) {
                BandIter it;
                band_init<POOL>(it, a, band);
                for (int t = 0; t < it.nt; ++t) {
                    const int TPRt = a.W / F_TILE_M;
                    // pool mode: tile t = (row pair t / TPR, half t % TPR); its first nstg stages are the upper row, the rest the lower one
                    const int m0 = POOL ? static_cast<int>(it.m0 + static_cast<long>(t / TPRt) * 2 * a.W + (t % TPRt) * F_TILE_M)
                                          : static_cast<int>(it.m0 + static_cast<long>(t) * (a.pair ? a.W : F_TILE_M));
                    const int tpt = POOL ? 2 * a.nstg : a.nstg;
                    for (int jj = 0; jj < tpt; ++jj, ++g) {
                        const int j = POOL && jj >= a.nstg ? jj - a.nstg : jj;
                        const int mrow = POOL && jj >= a.nstg ? m0 + a.W : m0;
                        const int s = static_cast<int>(g % nst_u);
                        const uint32_t ph = (g / nst_u) & 1;
                        const uint32_t dst = smem_u32(smem + static_cast<size_t>(s) * F_STAGE_BYTES);
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, F_STAGE_BYTES);
                        if (a.gstride > 0) {          // channel planes: stage j = plane j, rows = pixels
                            f_tma_load_2d(dst, &tmap, 0, static_cast<int>(j * a.gstride + mrow), bar_full + 8 * s);
                        } else if (a.pair) {          // two boxes of 64 pixels: the same row of two consecutive images
                            f_tma_load_2d(dst, &tmap, j * F_STAGE_C, mrow, bar_full + 8 * s);
                            f_tma_load_2d(dst + F_STAGE_BYTES / 2, &tmap, j * F_STAGE_C, mrow + a.H * a.W, bar_full + 8 * s);
                        } else {
                            f_tma_load_2d(dst, &tmap, j * F_STAGE_C, mrow, bar_full + 8 * s);
                        }
                    }
                }
            }
        }
    } else if (warp == F_MMA_WARP) {
        // =========================================================== MMA ISSUER
        const bool leader = elect_one();
        if (leader) {
            const uint32_t bytes = static_cast<uint32_t>(a.wbytes);
            mbar_expect_tx(bar_w, bytes);
            bulk_g2s(smem_u32(w_sm), a.wpack, bytes, bar_w);
        }
        mbar_wait(bar_w, 0);
        const uint32_t idesc = make_idesc_bf16(F_TILE_M, F_NPAD);
        const uint64_t dA0 = make_sw128_desc(smem_u32(smem));
        const uint64_t dB0 = POOL ? make_sw128_desc(smem_u32(w_sm)) : make_nosw_desc(smem_u32(w_sm), F_WROWS * 16, 128);
        const uint32_t stage16 = F_STAGE_BYTES >> 4, wchunk16 = F_WCHUNK >> 4;
        const uint32_t blo16 = POOL ? (F_NPAD * 128) >> 4 : F_WPLANE >> 4;             // hi -> lo image
        const uint64_t kadv = POOL ? 2 : (2 * F_WROWS * 16) >> 4;                      // one 16-channel k-step, in 16-byte units
        const int ksteps_total = (a.C_in + 15) >> 4;
        uint32_t g = 0, jt = 0;
        const uint32_t nst_u = static_cast<uint32_t>(NST);
        if constexpr (RS) {
            // ---- row-sum issue loop.  Z row rho (image row rho) contributes through vertical tap dy to output row r = rho + 1 - dy:
            //   R[r % 4][half] (+)= A(rho) * W[dy]^T      (N = 48: the tap group's 36 rows + 12 rows of the next group, columns never read)
            // so the vertical part of the 3x3 stencil is done by the accumulate flag: the tap that reaches a row first (dy = 0 from the row
            // above; dy = 1 for image row 0) initialises the accumulator, a row is complete when the tile BELOW it is done, and halo rows
            // of a band issue one tap only.  The MMA warp runs up to two rows ahead of the epilogue, which merely drains finished rows.
            const uint32_t idesc_rs = make_idesc_bf16(F_TILE_M, F_RN);
            const uint64_t dBr = make_nosw_desc(smem_u32(w_sm), F_WROWS * 16, 128);
            const uint64_t kadv_r = (2 * F_WROWS * 16) >> 4, blo_r = F_WPLANE >> 4, tap16 = (3 * F_G * 16) >> 4;
            uint32_t rpar = 0;                                    // per accumulator: parity of its next "drained" wait
            for (long band = blockIdx.x; band < a.nbands; band += grid) {
                BandIter it;
                band_init<POOL>(it, a, band);
                const int bpi = a.H / a.R;
                const int r0 = static_cast<int>(band % bpi) * a.R, rlo = max(r0 - 1, 0), rend = r0 + a.R - 1;
                for (int t = 0; t < it.nt; ++t, ++jt) {
                    const int rho = rlo + t / TPR, half = t % TPR;
                    bool valid[3], init[3];
                    uint32_t dcol[3];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int r = rho + 1 - dy;
                        valid[dy] = r >= r0 && r <= rend;
                        init[dy] = valid[dy] && (dy == 0 || (dy == 1 && rho == rlo));    // dy = 1 starts a row only when there is no row above (rho = rlo = r0 = 0)
                        const uint32_t idx = static_cast<uint32_t>((r & 3) * 2 + half);
                        dcol[dy] = tmem_base + idx * F_RN;
                        if (init[dy]) {                                           // the accumulator's previous row must have been drained
                            mbar_wait(bar_rfree + 8 * idx, ((rpar >> idx) & 1u) ^ 1u);
                            rpar ^= 1u << idx;
                        }
                    }
                    tc_fence_after();
                    for (int jj = 0; jj < a.nstg; ++jj, ++g) {
                        const uint32_t s = g & 3u, ph = (g >> 2) & 1;
                        mbar_wait(bar_ready + 8 * s, ph);
                        tc_fence_after();
                        if (leader) {
                            const int ks = min(2, ksteps_total - 2 * jj);
                            const uint64_t db_hi = dBr + static_cast<uint64_t>(static_cast<uint32_t>(jj) * (F_WUNIT >> 4));
                            const uint32_t ta = tmem_base + F_ABASE + s * 32u;
                            for (int k = 0; k < ks; ++k) {
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    if (!valid[dy]) continue;
                                    const uint64_t b = db_hi + static_cast<uint64_t>(k) * kadv_r + static_cast<uint64_t>(dy) * tap16;
                                    umma_bf16_ts(dcol[dy], ta + 8 * k, b, idesc_rs, (init[dy] && (jj | k) == 0) ? 0u : 1u);
                                    if (SPLIT) {
                                        umma_bf16_ts(dcol[dy], ta + 16 + 8 * k, b, idesc_rs, 1u);
                                        umma_bf16_ts(dcol[dy], ta + 8 * k, b + blo_r, idesc_rs, 1u);
                                    }
                                }
                            }
                            umma_commit(bar_afree + 8 * s);
                            if (jj == a.nstg - 1) umma_commit(bar_tfull + 8 * (jt & (F_NT - 1)));
                        }
                        __syncwarp();
                    }
                }
            }
        } else
        for (long band = blockIdx.x; band < a.nbands; band += grid) {
            BandIter it;
            band_init<POOL>(it, a, band);
            for (int t = 0; t < it.nt; ++t, ++jt) {
                const uint32_t zb = jt % F_NZ, zph = (jt / F_NZ) & 1;
                mbar_wait(bar_zempty + 8 * zb, zph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + zb * F_ZSTRIDE;
                const int tpt = POOL ? 2 * a.nstg : a.nstg;
                for (int jj = 0; jj < tpt; ++jj, ++g) {
                    const int j = POOL && jj >= a.nstg ? jj - a.nstg : jj;
                    const uint32_t s = TS ? (g & 3u) : g % nst_u;                    // TS: A slot = warpgroup = g % 4
                    const uint32_t ph = TS ? (g >> 2) & 1 : (g / nst_u) & 1;
                    mbar_wait(bar_ready + 8 * s, ph);
                    tc_fence_after();
                    if (leader) {
                        const int ks = min(2, ksteps_total - 2 * j);
                        const uint64_t db_hi = dB0 + (POOL ? static_cast<uint64_t>(static_cast<uint32_t>(j >> 1) * wchunk16 + static_cast<uint32_t>(j & 1) * 4)
                                                           : static_cast<uint64_t>(static_cast<uint32_t>(j) * (F_WUNIT >> 4)));
                        const uint64_t db_lo = db_hi + blo16;
                        if (TS) {
                            const uint32_t ta = tmem_base + F_ABASE + s * 32u;
                            for (int k = 0; k < ks; ++k) {
                                const uint64_t badv = static_cast<uint64_t>(k) * kadv;
                                umma_bf16_ts(d_tmem, ta + 8 * k, db_hi + badv, idesc, (jj | k) != 0 ? 1u : 0u);
                                if (SPLIT) {
                                    umma_bf16_ts(d_tmem, ta + 16 + 8 * k, db_hi + badv, idesc, 1u);
                                    umma_bf16_ts(d_tmem, ta + 8 * k, db_lo + badv, idesc, 1u);
                                }
                            }
                            umma_commit(bar_afree + 8 * s);
                        } else {
                            const uint64_t da = dA0 + static_cast<uint64_t>(s * stage16);
                            for (int k = 0; k < ks; ++k) {
                                const uint64_t adv = static_cast<uint64_t>(k * 2);       // A: 32 bytes per k-step, in 16-byte units
                                const uint64_t badv = static_cast<uint64_t>(k) * kadv;
                                umma_bf16(d_tmem, da + adv, db_hi + badv, idesc, (jj | k) != 0 ? 1u : 0u);
                                if (SPLIT) {
                                    umma_bf16(d_tmem, da + 4 + adv, db_hi + badv, idesc, 1u);  // lo half of the row: + 64 bytes
                                    umma_bf16(d_tmem, da + adv, db_lo + badv, idesc, 1u);
                                }
                            }
                            umma_commit(bar_empty + 8 * s);
                        }
                        if (jj == tpt - 1) umma_commit(bar_zfull + 8 * zb);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= F_EPI_WARP0) {
        // =========================================================== EPILOGUE: warpgroup wg owns half-row wg; thread = pixel
        const int ew = warp - F_EPI_WARP0;                   // 0..7
        const int wg = ew >> 2, q = warp & 3;                 // TMEM lane quarter = warp % 4
        const int et = tid - F_EPI_WARP0 * 32;                // 0..255
        const int nE = F_TILE_M * TPR;                        // epilogue threads that take part (128 or 256)
        const bool active = wg < TPR;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int px = q * 32 + lane;                         // pixel inside the tile
        if (POOL && active) {
            // ---------------- transition: Z = sum over the two image rows (accumulated by the MMAs); horizontal pair sum by shuffle,
            // x 0.25, the lane pair of a pooled pixel splits the channel quads between its two lanes and stores straight to the output
            uint32_t j = 0;
            const int nquad = (a.n_out + 3) >> 2;
            const int Wp = a.W >> 1, Hp = a.H >> 1;
            for (long band = blockIdx.x; band < a.nbands; band += grid) {
                const int bpi = a.H / a.R;
                const long img = band / bpi;
                const int r0 = static_cast<int>(band % bpi) * a.R;
                for (int rp = 0; rp < a.R / 2; ++rp) {
                    const uint32_t jt = j + static_cast<uint32_t>(wg);
                    j += static_cast<uint32_t>(TPR);
                    const uint32_t zb = jt % F_NZ, zph = (jt / F_NZ) & 1;
                    const uint32_t zc = lane_addr + zb * F_ZSTRIDE;
                    float *op = a.out + ((img * Hp + (r0 >> 1) + rp) * Wp + wg * (F_TILE_M / 2) + (px >> 1)) * a.out_pitch + a.out_choff;
                    mbar_wait(bar_zfull + 8 * zb, zph);
                    __syncwarp();
                    tc_fence_after();
                    for (int q0 = 0; q0 < nquad; q0 += 4) {                  // 16 accumulator columns per round trip
                        float z[16];
                        f_tmem_ld16(zc + q0 * 4, z);
#pragma unroll
                        for (int e = 0; e < 16; ++e) z[e] = 0.25f * (z[e] + __shfl_xor_sync(0xffffffffu, z[e], 1));
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const int qd = q0 + qq;
                            if (qd < nquad && ((qd & 1) == (lane & 1))) {        // even quads from the even lane, odd quads from the odd lane
                                const int n = qd * 4;
                                if (n + 3 < a.n_out) {
                                    *reinterpret_cast<float4 *>(op + n) = make_float4(z[qq * 4], z[qq * 4 + 1], z[qq * 4 + 2], z[qq * 4 + 3]);
                                } else {
                                    for (int e = 0; e < 4; ++e) if (n + e < a.n_out) op[n + e] = z[qq * 4 + e];
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) f_mbar_arrive(bar_zempty + 8 * zb);
                }
            }
        } else if (!POOL && active) {
            uint32_t j = 0;                                   // tile counter (all tiles of this CTA, both halves)
            uint32_t rs = 0;                                  // row sequence number (all Z rows of this CTA): index into the stash ring
            const float4 *s_stash = reinterpret_cast<const float4 *>(s_row + (a.W + 2) * F_SROW);
            for (long band = blockIdx.x; band < a.nbands; band += grid) {
                const int bpi = a.H / a.R;
                const long img = (band / bpi) * (a.pair ? 2 : 1);
                const int r0 = static_cast<int>(band % bpi) * a.R;
                const int rlo = max(r0 - 1, 0), rhi = min(r0 + a.R, a.H - 1), rend = r0 + a.R - 1;    // rend = last output row
                for (int rho = rlo; rho <= rhi; ++rho) {
                    const uint32_t rseq = rs++;
                    // ---- this warpgroup's tile of Z row rho
                    const uint32_t jt = j + static_cast<uint32_t>(wg);
                    j += static_cast<uint32_t>(TPR);
                    const uint32_t zb = jt % F_NZ, zph = (jt / F_NZ) & 1;
                    const bool emit = rho - 1 >= r0;                          // output row rho-1 completes now
                    const bool upd = rho >= r0 && rho <= rend;                // output row rho receives its dy=1 part
                    const bool upd_add = rho > rlo;                           // ... on top of the dy=0 part stored by row rho-1
                    const bool init_next = rho + 1 <= rend;                   // output row rho+1 starts with this row's dy=0 part
                    const bool flush = rho == a.H - 1 && upd;                 // bottom image row: nothing below completes it
                    const uint32_t zc = lane_addr + zb * F_ZSTRIDE;
                    const uint32_t us0 = lane_addr + F_UBASE + ((((rho + 1) & 1) * 2 + wg) * F_USTRIDE);   // U[rho-1] in, U[rho+1] out
                    const uint32_t us1 = lane_addr + F_UBASE + (((rho & 1) * 2 + wg) * F_USTRIDE);         // U[rho]
                    float *srow = s_row + (wg * F_TILE_M + px + 1 + (a.pair ? 2 * (px >> 6) : 0)) * F_SROW;
                    if constexpr (RS) {
                        // the tile of Z row rho is done -> output row rho - 1 is complete in its accumulator: copy it to the row buffer
                        mbar_wait(bar_tfull + 8 * (jt & (F_NT - 1)), (jt / F_NT) & 1);
                        __syncwarp();
                        tc_fence_after();
                        if (emit) {
                            const uint32_t idx = static_cast<uint32_t>(((rho - 1) & 3) * 2 + wg);
                            const uint32_t rc0 = lane_addr + idx * F_RN;
                            float v0[8], v1[8], v2[8], v3[8], v4[4];
                            f_tmem_ld<8>(rc0, v0); f_tmem_ld<8>(rc0 + 8, v1); f_tmem_ld<8>(rc0 + 16, v2); f_tmem_ld<8>(rc0 + 24, v3); f_tmem_ld<4>(rc0 + 32, v4);
                            f_tmem_wait_ld<8>(v0); f_tmem_wait_ld<8>(v1); f_tmem_wait_ld<8>(v2); f_tmem_wait_ld<8>(v3); f_tmem_wait_ld<4>(v4);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) f_mbar_arrive(bar_rfree + 8 * idx);    // the MMA warp may start another row in this accumulator
                            *reinterpret_cast<float4 *>(srow) = make_float4(v0[0], v0[1], v0[2], v0[3]);
                            *reinterpret_cast<float4 *>(srow + 4) = make_float4(v0[4], v0[5], v0[6], v0[7]);
                            *reinterpret_cast<float4 *>(srow + 8) = make_float4(v1[0], v1[1], v1[2], v1[3]);
                            *reinterpret_cast<float4 *>(srow + 12) = make_float4(v1[4], v1[5], v1[6], v1[7]);
                            *reinterpret_cast<float4 *>(srow + 16) = make_float4(v2[0], v2[1], v2[2], v2[3]);
                            *reinterpret_cast<float4 *>(srow + 20) = make_float4(v2[4], v2[5], v2[6], v2[7]);
                            *reinterpret_cast<float4 *>(srow + 24) = make_float4(v3[0], v3[1], v3[2], v3[3]);
                            *reinterpret_cast<float4 *>(srow + 28) = make_float4(v3[4], v3[5], v3[6], v3[7]);
                            *reinterpret_cast<float4 *>(srow + 32) = make_float4(v4[0], v4[1], v4[2], v4[3]);
                        }
                    } else {
                    mbar_wait(bar_zfull + 8 * zb, zph);
                    __syncwarp();
                    tc_fence_after();
                    if (PW == 8) {                                      // 8 columns x 5 arrays = 40 registers in flight per slice (NCW = 8 builds)
#pragma unroll
                        for (int piece = 0; piece < 4; ++piece)
                            f_stencil_piece<8>(zc, us0, us1, piece * 8, srow, emit, init_next, upd, upd_add);
                        f_stencil_piece<4>(zc, us0, us1, 32, srow, emit, init_next, upd, upd_add);
                    } else {
#pragma unroll
                        for (int piece = 0; piece < 9; ++piece)        // 4 columns x 5 arrays = 20 registers in flight per slice
                            f_stencil_piece<4>(zc, us0, us1, piece * 4, srow, emit, init_next, upd, upd_add);
                    }
                    f_tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) f_mbar_arrive(bar_zempty + 8 * zb);        // Z buffer drained: the MMA warp may refill it
                    }
                    // ---- completed rows: U row -> shared row buffer -> 3-tap horizontal sum + bias -> slab
                    for (int pass = 0; pass < 2; ++pass) {
                        int orow;
                        if (pass == 0) { if (!emit) continue; orow = rho - 1; }
                        else {
                            if (!flush) continue;
                            orow = rho;
                            float z4[4];
                            const uint32_t fidx = static_cast<uint32_t>((rho & 3) * 2 + wg);
                            const uint32_t fsrc = RS ? lane_addr + fidx * F_RN : us1;   // RS: the bottom row's own accumulator (complete: nothing below it)
#pragma unroll
                            for (int piece = 0; piece < 4; ++piece) {            // rare (once per image): no need to batch
                                float z[8];
                                f_tmem_ld<8>(fsrc + piece * 8, z);
                                f_tmem_wait_ld<8>(z);
                                *reinterpret_cast<float4 *>(srow + piece * 8) = make_float4(z[0], z[1], z[2], z[3]);
                                *reinterpret_cast<float4 *>(srow + piece * 8 + 4) = make_float4(z[4], z[5], z[6], z[7]);
                            }
                            f_tmem_ld<4>(fsrc + 32, z4);
                            f_tmem_wait_ld<4>(z4);
                            *reinterpret_cast<float4 *>(srow + 32) = make_float4(z4[0], z4[1], z4[2], z4[3]);
                            if constexpr (RS) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) f_mbar_arrive(bar_rfree + 8 * fidx);
                            }
                        }
                        asm volatile("bar.sync 1, %0;" ::"r"(nE) : "memory");     // the whole row is staged
                        const int rc = orow == 0 ? 0 : (orow == a.H - 1 ? 2 : 1);
                        float *obase = a.out + ((img * a.H + orow) * a.W) * a.out_pitch + a.out_choff;
                        if (a.wide == 3) {
                            // channel-plane output: quad cq = out_choff / 4 + qd of pixel m lives at plane cq / 8, row m, 16-byte slot cq % 8
                            // (the 12 new channels may straddle two planes); consecutive pixels = consecutive 128-byte rows
                            const long mbase = (img * a.H + orow) * static_cast<long>(a.W);
                            const int cq0 = a.out_choff >> 2;
                            for (int f = et; f < a.W * 3; f += nE) {
                                const int x = f / 3, qd = f - x * 3;
                                const float *s0 = s_row + x * F_SROW + qd * 4;
                                const float4 v0 = *reinterpret_cast<const float4 *>(s0);
                                const float4 v1 = *reinterpret_cast<const float4 *>(s0 + F_SROW + F_G);
                                const float4 v2 = *reinterpret_cast<const float4 *>(s0 + 2 * F_SROW + 2 * F_G);
                                const int cc = x == 0 ? 0 : (x == a.W - 1 ? 2 : 1);
                                const float4 bb = *reinterpret_cast<const float4 *>(&s_bias[(rc * 3 + cc) * F_G + qd * 4]);
                                float4 o;
                                o.x = v0.x + v1.x + v2.x + bb.x; o.y = v0.y + v1.y + v2.y + bb.y;
                                o.z = v0.z + v1.z + v2.z + bb.z; o.w = v0.w + v1.w + v2.w + bb.w;
                                const int cq = cq0 + qd;
                                *reinterpret_cast<float4 *>(a.out + ((cq >> 3) * a.ogstride + mbase + x) * 32 + (cq & 7) * 4) = o;
                            }
                        } else if (a.wide == 2) {
                            // 64 bytes per pixel starting 4 channels IN FRONT of the new ones: [4 stashed raw channels | 12 new channels]
                            const float4 *srow4 = s_stash + ((pass == 0 ? rseq - 1 : rseq) % static_cast<uint32_t>(a.stash_rows)) * a.W;
                            for (int f = et; f < a.W * 4; f += nE) {
                                const int x = f >> 2, qd = f & 3;                          // qd 0: the stashed quad; 1..3: quads of the new channels
                                float *dst = obase + static_cast<long>(x) * a.out_pitch + (qd - 1) * 4;
                                float4 o;
                                if (qd == 0) {
                                    o = srow4[x];
                                } else {
                                    const float *s0 = s_row + x * F_SROW + (qd - 1) * 4;
                                    const float4 v0 = *reinterpret_cast<const float4 *>(s0);
                                    const float4 v1 = *reinterpret_cast<const float4 *>(s0 + F_SROW + F_G);
                                    const float4 v2 = *reinterpret_cast<const float4 *>(s0 + 2 * F_SROW + 2 * F_G);
                                    const int cc = x == 0 ? 0 : (x == a.W - 1 ? 2 : 1);
                                    const float4 bb = *reinterpret_cast<const float4 *>(&s_bias[(rc * 3 + cc) * F_G + (qd - 1) * 4]);
                                    o.x = v0.x + v1.x + v2.x + bb.x; o.y = v0.y + v1.y + v2.y + bb.y;
                                    o.z = v0.z + v1.z + v2.z + bb.z; o.w = v0.w + v1.w + v2.w + bb.w;
                                }
                                *reinterpret_cast<float4 *>(dst) = o;
                            }
                        } else if (a.wide == 0) {
                            const int npx = a.pair ? 2 * a.W : a.W;                    // pixels staged per row (pair mode: two images)
                            const bool al16 = (a.out_choff & 3) == 0;                  // block 3 starts its channels on an 8-byte boundary only
                            for (int f = et; f < npx * 3; f += nE) {
                                const int p = f / 3, qd = f - p * 3;
                                const int half = a.pair ? p >> 6 : 0, x = a.pair ? p & 63 : p;
                                const float *s0 = s_row + (p + 2 * half) * F_SROW + qd * 4;   // pixel x-1 (buffer index of x, minus one), dx = 0
                                const float4 v0 = *reinterpret_cast<const float4 *>(s0);
                                const float4 v1 = *reinterpret_cast<const float4 *>(s0 + F_SROW + F_G);
                                const float4 v2 = *reinterpret_cast<const float4 *>(s0 + 2 * F_SROW + 2 * F_G);
                                const int cc = x == 0 ? 0 : (x == a.W - 1 ? 2 : 1);
                                const float4 bb = *reinterpret_cast<const float4 *>(&s_bias[(rc * 3 + cc) * F_G + qd * 4]);
                                float4 o;
                                o.x = v0.x + v1.x + v2.x + bb.x; o.y = v0.y + v1.y + v2.y + bb.y;
                                o.z = v0.z + v1.z + v2.z + bb.z; o.w = v0.w + v1.w + v2.w + bb.w;
                                float *dst = obase + (static_cast<long>(half) * a.H * a.W + x) * a.out_pitch + qd * 4;
                                if (al16) {
                                    *reinterpret_cast<float4 *>(dst) = o;
                                } else {
                                    *reinterpret_cast<float2 *>(dst) = make_float2(o.x, o.y);
                                    *reinterpret_cast<float2 *>(dst + 2) = make_float2(o.z, o.w);
                                }
                            }
                        } else {
                            // 64 bytes per pixel = two FULL 32-byte sectors when the new channels start on a sector boundary: the fourth
                            // quad zeroes the 4 channels behind them (the next layer overwrites those), so no sector is left half
                            // written and the (pixel, quad) index needs no division.  Measured: 1.23 -> 0.95 ms at C_in = 24.  (The
                            // mirror image for layers that start mid-sector -- re-storing the 4 channels in front -- was tried with a
                            // cp.async prefetch of those values and lost 0.15 ms per layer to the extra traffic; they store 48 bytes.)
                            for (int f = et; f < a.W * 4; f += nE) {
                                const int x = f >> 2, qd = f & 3;                          // qd: quad of the 12 new channels, 3 = filler
                                float *dst = obase + static_cast<long>(x) * a.out_pitch + qd * 4;
                                float4 o;
                                if (qd > 2) {
                                    o = make_float4(0.f, 0.f, 0.f, 0.f);
                                } else {
                                    const float *s0 = s_row + x * F_SROW + qd * 4;
                                    const float4 v0 = *reinterpret_cast<const float4 *>(s0);
                                    const float4 v1 = *reinterpret_cast<const float4 *>(s0 + F_SROW + F_G);
                                    const float4 v2 = *reinterpret_cast<const float4 *>(s0 + 2 * F_SROW + 2 * F_G);
                                    const int cc = x == 0 ? 0 : (x == a.W - 1 ? 2 : 1);
                                    const float4 bb = *reinterpret_cast<const float4 *>(&s_bias[(rc * 3 + cc) * F_G + qd * 4]);
                                    o.x = v0.x + v1.x + v2.x + bb.x; o.y = v0.y + v1.y + v2.y + bb.y;
                                    o.z = v0.z + v1.z + v2.z + bb.z; o.w = v0.w + v1.w + v2.w + bb.w;
                                }
                                *reinterpret_cast<float4 *>(dst) = o;
                            }
                        }
                        asm volatile("bar.sync 1, %0;" ::"r"(nE) : "memory");     // row buffer free again
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == F_MMA_WARP) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

size_t fused_smem(size_t wbytes, int W, int stages, int srow = F_GRP, size_t extra = 0) {
    return static_cast<size_t>(stages) * F_STAGE_BYTES + wbytes + static_cast<size_t>(W == 64 ? 2 * (W + 2) : W + 2) * srow * 4 + extra + 1024;
}
int fused_stages(size_t wbytes, int W, int srow = F_GRP, size_t extra = 0, int min_stages = 4) {
    for (int st = F_MAX_STAGES; st >= min_stages; --st)
        if (fused_smem(wbytes, W, st, srow, extra) <= 227 * 1024 - 3700) return st;   // static shared memory (barriers, affine tables: ~3.6 KB) counts too
    return 0;
}
// Smallest ring the pipeline accepts: one slot per converter warpgroup.  The default build (TMEM-resident A, two converter warpgroups)
// runs with 2 slots when the resident weights leave no more (block 3's widest layers, C_in up to 318: 150 KB of weights).
inline int dense_min_stages() {
    if (eml_env_flag("EML_DENSE_SMEM_A")) return 4;
    const char *cw_env = getenv("EML_DENSE_CW");
    return (cw_env && atoi(cw_env) == 16) ? 4 : 2;
}
inline size_t dense_wbytes(int C_in) { return static_cast<size_t>((C_in + F_STAGE_C - 1) / F_STAGE_C) * F_WUNIT; }
inline size_t pool_wbytes(int C_in) { return static_cast<size_t>((C_in + 63) / 64) * F_WCHUNK; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn f_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

extern "C" int eml_dense_layer_supported(int H, int W, int C_in, int growth, int precision) {
    if (growth != F_G || (W != 64 && W != 128 && W != 256) || H < 2 || C_in <= 0 || C_in > F_MAX_C) return 0;
    if (W == 64 ? (C_in & 1) : (C_in & 3)) return 0;          // output channel offset: 8-byte (W = 64, float2 stores) or 16-byte aligned
    if (precision != EML_PREC_BF16 && precision != EML_PREC_BF16X3) return 0;
    return fused_stages(dense_wbytes(C_in), W, F_GRP, 0, dense_min_stages()) >= dense_min_stages() ? 1 : 0;
}

extern "C" size_t eml_dense_layer_wpack_bytes(int C_in) { return C_in > 0 ? dense_wbytes(C_in) : 0; }

extern "C" int eml_dense_layer_forward(const eml_dense_layer_params *p, void *stream) {
    EML_CHECK_PTR(p); EML_CHECK_PTR(p->in); EML_CHECK_PTR(p->out); EML_CHECK_PTR(p->scale); EML_CHECK_PTR(p->shift);
    EML_CHECK_PTR(p->wpack); EML_CHECK_PTR(p->bias9);
    EML_CHECK_ALIGN16(p->in); EML_CHECK_ALIGN16(p->out); EML_CHECK_ALIGN16(p->wpack);
    if (p->B <= 0 || !eml_dense_layer_supported(p->H, p->W, p->C_in, p->growth, p->precision)) return EML_E_SHAPE;
    const bool pair = p->W == 64;                            // two images side by side in one 128-pixel tile
    if (pair && (p->B & 1)) return EML_E_SHAPE;
    const long npix = static_cast<long>(p->B) * p->H * p->W;
    const bool grouped = p->plane_pixels > 0;                // channel-plane slab (header): in == out, planes of 32 channels
    if (grouped) {
        if (pair || p->in != p->out || p->plane_pixels < npix || (p->out_choff & 3) || p->out_choff < 0) return EML_E_SHAPE;
        if (((p->C_in + 31) / 32) * p->plane_pixels >= (1L << 31)) return EML_E_SHAPE;
    } else if ((p->in_pitch & 3) || p->in_pitch < p->C_in || (p->out_pitch & 3) || (p->out_choff & (pair ? 1 : 3)) || p->out_choff < 0 ||
               p->out_pitch < p->out_choff + F_G)
        return EML_E_ALIGN;
    if (npix >= (1L << 31)) return EML_E_SHAPE;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // rows per band: minimise (waves of bands over the SMs) x (rows of Z per band, R + 2 halo rows)
    int R = p->H;
    long best = -1;
    for (int r = 1; r <= p->H; ++r) {
        if (p->H % r) continue;
        const long nb = static_cast<long>(pair ? p->B / 2 : p->B) * (p->H / r);
        const long cost = ((nb + sms - 1) / sms) * (r + 2);
        if (best < 0 || cost < best || (cost == best && r > R)) { best = cost; R = r; }
    }
    if (const char *env = getenv("EML_DENSE_ROWS")) {     // debug / test switch: force the band height
        const int r = atoi(env);
        if (r > 0 && p->H % r == 0) R = r;
    }
    // the slab as a 2-D fp32 tensor (channels [0, C_in) x pixels, pixel stride in_pitch): boxes of 32 channels x 128 pixels
    EncodeTiledFn enc = f_get_encode();
    if (enc == nullptr) return EML_E_ARG;
    CUtensorMap tmap;
    {
        // grouped: ceil(C_in / 32) planes of (plane_pixels x 32 floats) seen as ONE 2-D tensor of 128-byte rows
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(grouped ? F_STAGE_C : p->C_in),
                                    static_cast<cuuint64_t>(grouped ? ((p->C_in + 31) / 32) * p->plane_pixels : npix)};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(grouped ? F_STAGE_C : p->in_pitch) * 4};
        const cuuint32_t box[2] = {F_STAGE_C, static_cast<cuuint32_t>(pair ? 64 : F_TILE_M)};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p->in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return EML_E_ARG;
    }
    FArgs a{};
    a.scale = p->scale; a.shift = p->shift; a.wpack = static_cast<const unsigned char *>(p->wpack);
    a.bias9 = p->bias9; a.out = p->out;
    a.B = p->B; a.H = p->H; a.W = p->W; a.R = R;
    a.C_in = p->C_in; a.out_pitch = p->out_pitch; a.out_choff = p->out_choff;
    a.nwchunks = (p->C_in + 63) / 64;
    a.nstg = (p->C_in + F_STAGE_C - 1) / F_STAGE_C;
    a.pair = pair ? 1 : 0;
    a.nbands = static_cast<long>(pair ? p->B / 2 : p->B) * (p->H / R);
    // full-sector stores (see the output pass): possible when every pixel record starts on a 32-byte boundary
    a.wide = 0;
    if (grouped) { a.gstride = p->plane_pixels; a.ogstride = p->plane_pixels; a.wide = 3; }
    else if ((reinterpret_cast<uintptr_t>(p->out) & 31u) == 0 && (p->out_pitch & 7) == 0 && !eml_env_flag("EML_DENSE_NARROW_STORE")) {
        if (!pair && (p->out_choff & 7) == 0 && p->out_choff + F_G + 4 <= p->out_pitch) a.wide = 1;
    }
    a.wbytes = static_cast<int>(dense_wbytes(p->C_in));
    const char *cw_env = getenv("EML_DENSE_CW");          // converter warps: 8 (default: two warpgroups, 8-column stencil pieces, ring depth in steps of 2) or 16
    const int cw = (cw_env && atoi(cw_env) == 16) ? 16 : 8;
    const bool ts = !eml_env_flag("EML_DENSE_SMEM_A");
    const int gran = ts ? (cw == 8 ? 2 : 4) : 1;          // ring-depth granularity (see below)
    // Layers whose channels start mid-sector (offset = 4 mod 8; every other layer of blocks 1 and 2): 48-byte stores leave a half-written
    // 32-byte sector per pixel -- a read-modify-write in the memory system that costs 0.3-0.4 ms per block-1 layer at B = 256 (tools/
    // layer_times.py).  With the TMEM-resident A operand the converters hold the layer's raw input in registers, so the 4 channels in front
    // of the output are stashed (16 B per pixel, a ring of image rows) and re-stored with the new ones: 64 B = two full sectors.
    size_t stash = 0;
    if (ts && a.wide == 0 && !pair && (reinterpret_cast<uintptr_t>(p->out) & 31u) == 0 && (p->out_pitch & 7) == 0 && (p->out_choff & 7) == 4 &&
        p->in == p->out && p->in_pitch == p->out_pitch && p->out_choff == p->C_in && a.nstg >= 2 && !eml_env_flag("EML_DENSE_NARROW_STORE") &&
        !eml_env_flag("EML_DENSE_NO_STASH")) {
        a.stash_rows = 1024 / p->W;                            // 4 rows (W = 256) / 8 rows (W = 128): more than the converters can run ahead
        stash = static_cast<size_t>(a.stash_rows) * p->W * 16;
        if (fused_stages(a.wbytes, p->W, F_GRP, stash) >= 4) a.wide = 2; else { stash = 0; a.stash_rows = 0; }
    }
    const int minst = dense_min_stages();
    const int st36 = fused_stages(a.wbytes, p->W, F_GRP, stash, minst), st44 = fused_stages(a.wbytes, p->W, 44, stash, minst);
    a.stages = st36;
    // 44-float row-buffer records (bank-conflict-free output pass) whenever that does not cost ring depth
    const bool wide_rows = (st44 - st44 % gran) == (st36 - st36 % gran) && st44 >= minst && !eml_env_flag("EML_DENSE_PACKED_ROWS");
    // TS: the ring depth must be a multiple of the number of converter warpgroups, so that a ring slot is always converted by the same
    // warpgroup and every phase of its mbarrier has one in-order waiter.  (A parity wait cannot tell "two completions early" from "done":
    // with e.g. 5 slots and 4 warpgroups the waiter of round k on a slot may arrive before round k - 1 has even landed -- TMA loads of
    // neighbouring stages complete out of order -- pass spuriously and read stale data: seen on B200 as run-to-run differences and hangs.)
    a.stages -= a.stages % gran;
    const bool split = p->precision == EML_PREC_BF16X3;
    const size_t smem = fused_smem(a.wbytes, p->W, a.stages, wide_rows ? 44 : F_GRP, stash);
    const unsigned grid = static_cast<unsigned>(a.nbands < sms ? a.nbands : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto kern, int threads) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        kern<<<grid, threads, smem, st>>>(tmap, a);
        return EML_OK;
    };
    int rc;
    if (!ts) {                                            // round-1 pipeline (A operand converted in place in shared memory), kept for A/B runs
        if (split) rc = wide_rows ? go(dense_layer_kernel<true, false, 44, false>, f_threads(16)) : go(dense_layer_kernel<true, false, F_GRP, false>, f_threads(16));
        else rc = wide_rows ? go(dense_layer_kernel<false, false, 44, false>, f_threads(16)) : go(dense_layer_kernel<false, false, F_GRP, false>, f_threads(16));
    } else if (cw == 8) {
        // Two epilogue designs on the same TMA -> convert -> TMEM-A pipeline, chosen per layer from an interleaved in-process A/B at B = 256
        // (tools/layer_ab.py, profiles/r02_layer_ab.txt):
        //  row-sum   the MMAs add the vertical taps into per-row accumulators (18 MMAs of N = 48 per stage): the epilogue only drains rows,
        //            so the fixed cost per tile is lowest -- wins while a tile has few stages: C_in <= 60 at W = 256 (block 1, layers 1-4:
        //            0.97-1.28 ms against 1.16-1.38 ms), C_in <= 156 at W = 128 (block 2, layers 1-5);
        //  Z stencil 6 MMAs of N = 112 per stage into a Z tile, vertical sum by the epilogue through TMEM: fewer, larger MMAs -- wins for the
        //            wide layers, where the 18 small MMAs per stage hold the converters back (ncu: converters wait for a free A slot).
        const char *rs_env = getenv("EML_DENSE_RS_MAX_C");
        const int rs_max_c = rs_env ? atoi(rs_env) : (p->W == 256 ? 60 : (p->W == 128 ? 156 : 0));
        if (p->C_in <= rs_max_c) {
            if (split) rc = wide_rows ? go(dense_layer_kernel<true, false, 44, true, 8, 8, true>, f_threads(8)) : go(dense_layer_kernel<true, false, F_GRP, true, 8, 8, true>, f_threads(8));
            else rc = wide_rows ? go(dense_layer_kernel<false, false, 44, true, 8, 8, true>, f_threads(8)) : go(dense_layer_kernel<false, false, F_GRP, true, 8, 8, true>, f_threads(8));
        } else {
            if (split) rc = wide_rows ? go(dense_layer_kernel<true, false, 44, true, 8, 8>, f_threads(8)) : go(dense_layer_kernel<true, false, F_GRP, true, 8, 8>, f_threads(8));
            else rc = wide_rows ? go(dense_layer_kernel<false, false, 44, true, 8, 8>, f_threads(8)) : go(dense_layer_kernel<false, false, F_GRP, true, 8, 8>, f_threads(8));
        }
    } else {
        if (split) rc = wide_rows ? go(dense_layer_kernel<true, false, 44, true>, f_threads(16)) : go(dense_layer_kernel<true, false, F_GRP, true>, f_threads(16));
        else rc = wide_rows ? go(dense_layer_kernel<false, false, 44, true>, f_threads(16)) : go(dense_layer_kernel<false, false, F_GRP, true>, f_threads(16));
    }
    if (rc != EML_OK) return rc;
    return eml_launch_status();
}

// ------------------------------------------------------------------------------------------------ composite filter (once per parameter version)
// Weff[(dy,dx,o), c] = sum_b W2[o,b,dy,dx] * scale2[b] * W1[b,c]  and  bias9[rc][cc][o] = sum over the taps inside the image of
// sum_b W2[o,b,dy,dx] * shift2[b]  (file header; RegressionNetwork/DenseNet.py:30-43 has no nonlinearity between conv1 and conv2),
// accumulated in double and written STRAIGHT into the packed operand image the kernel above reads (per 32-channel unit [hi | lo], 112 rows,
// K-major no-swizzle core matrices: F_WUNIT).  One launch per layer instead of a dozen float64 ATen kernels plus a pack launch.
namespace {
__global__ void __launch_bounds__(256) dense_compose_kernel(const float *__restrict__ w1, const float *__restrict__ w2, const float *__restrict__ s2,
                                                            const float *__restrict__ t2, int nb, int C_in, unsigned char *__restrict__ out,
                                                            float *__restrict__ bias9) {
    const int nunits = (C_in + F_STAGE_C - 1) / F_STAGE_C;
    const long total = static_cast<long>(nunits) * F_WROWS * F_STAGE_C;
    for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx % F_STAGE_C), n = static_cast<int>((idx / F_STAGE_C) % F_WROWS), u = static_cast<int>(idx / (static_cast<long>(F_STAGE_C) * F_WROWS));
        const int ci = u * F_STAGE_C + k;
        double acc = 0.0;
        if (n < 9 * F_G && ci < C_in) {
            const int tap = n / F_G, o = n - tap * F_G;
            for (int b = 0; b < nb; ++b)
                acc += static_cast<double>(w2[(static_cast<long>(o) * nb + b) * 9 + tap]) * static_cast<double>(s2[b]) *
                       static_cast<double>(w1[static_cast<long>(b) * C_in + ci]);
        }
        const float v = static_cast<float>(acc);
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        unsigned char *base = out + static_cast<size_t>(u) * F_WUNIT;
        const uint32_t off = static_cast<uint32_t>((k >> 3) * (F_WROWS * 16) + n * 16 + (k & 7) * 2);      // no-swizzle core-matrix layout
        *reinterpret_cast<__nv_bfloat16 *>(base + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + F_WPLANE + off) = lo;
    }
    if (blockIdx.x == 0 && threadIdx.x < 9 * F_G) {       // (row class, column class, o): taps that fall inside the image
        const int o = threadIdx.x % F_G, cc = (threadIdx.x / F_G) % 3, rc = threadIdx.x / (3 * F_G);
        double acc = 0.0;
        for (int dy = (rc == 0 ? 1 : 0); dy <= (rc == 2 ? 1 : 2); ++dy)
            for (int dx = (cc == 0 ? 1 : 0); dx <= (cc == 2 ? 1 : 2); ++dx)
                for (int b = 0; b < nb; ++b)
                    acc += static_cast<double>(w2[(static_cast<long>(o) * nb + b) * 9 + dy * 3 + dx]) * static_cast<double>(t2[b]);
        bias9[threadIdx.x] = static_cast<float>(acc);
    }
}
}  // namespace

extern "C" int eml_dense_layer_compose(const float *w1, const float *w2, const float *scale2, const float *shift2, int nb, int C_in,
                                       int growth, void *wpack, float *bias9, void *stream) {
    EML_CHECK_PTR(w1); EML_CHECK_PTR(w2); EML_CHECK_PTR(scale2); EML_CHECK_PTR(shift2); EML_CHECK_PTR(wpack); EML_CHECK_PTR(bias9);
    EML_CHECK_ALIGN16(wpack);
    if (growth != F_G || nb <= 0 || nb > 256 || C_in <= 0 || C_in > F_MAX_C) return EML_E_SHAPE;
    dense_compose_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(w1, w2, scale2, shift2, nb, C_in, static_cast<unsigned char *>(wpack), bias9);
    return eml_launch_status();
}

// ------------------------------------------------------------------------------------------------ transition (POOL2) on the same pipeline
// norm + relu + conv1x1 + avg_pool2d(2) of RegressionNetwork/DenseNet.py:14-21 for C_out <= 112 (transition1: 216 -> 108): the TMA ring,
// the in-place conversion and the MMA loop are the dense layer's; the 2x2 average is split between the tensor core (the two image rows
// of a pooled row accumulate into ONE accumulator) and the epilogue (adjacent pixels = adjacent lanes, one shuffle).  Reached through
// eml_conv_forward(EML_CONV_POOL2); every other shape keeps conv_gemm_kernel<2>.
bool eml_dense_pool_supported(const eml_conv_params *p) {
    if (p->mode != EML_CONV_POOL2 || p->stats != nullptr || !p->relu) return false;
    if (p->precision != EML_PREC_BF16 && p->precision != EML_PREC_BF16X3) return false;
    if ((p->W != 128 && p->W != 256) || (p->H & 1) || p->C_in > F_MAX_C || (p->C_in & 3) || ((p->C_out + 15) & ~15) != F_NPAD) return false;
    if ((p->out_pitch & 3) || (p->out_choff & 3) || p->scale == nullptr || p->shift == nullptr) return false;
    if (eml_env_flag("EML_NO_TMA_TRANSITION")) return false;
    return fused_stages(pool_wbytes(p->C_in), p->W) >= 4;
}

// Can transition (C_in -> C_out, 2x2 average) at H x W read the block's slab as channel planes (eml_conv_params.plane_pixels)?
extern "C" int eml_transition_planes_supported(int H, int W, int C_in, int C_out, int precision) {
    eml_conv_params p{};
    static const float one = 1.f;
    p.mode = EML_CONV_POOL2; p.relu = 1; p.precision = precision; p.H = H; p.W = W; p.C_in = C_in; p.C_out = C_out;
    p.out_pitch = (C_out + 3) & ~3; p.out_choff = 0; p.scale = &one; p.shift = &one;
    return eml_dense_pool_supported(&p) ? 1 : 0;
}

int eml_dense_pool_forward(const eml_conv_params *p, cudaStream_t st) {
    const long npix = static_cast<long>(p->B) * p->H * p->W;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int R = p->H;
    long best = -1;
    for (int r = 2; r <= p->H; r += 2) {                       // bands of whole row pairs, no halo
        if (p->H % r) continue;
        const long nb = static_cast<long>(p->B) * (p->H / r);
        const long cost = ((nb + sms - 1) / sms) * r;
        if (best < 0 || cost < best || (cost == best && r > R)) { best = cost; R = r; }
    }
    EncodeTiledFn enc = f_get_encode();
    if (enc == nullptr) return EML_E_ARG;
    const bool grouped = p->plane_pixels > 0;                // the block's slab as channel planes (eml_dense_layer_forward)
    if (grouped && (p->plane_pixels < npix || ((p->C_in + 31) / 32) * p->plane_pixels >= (1L << 31))) return EML_E_SHAPE;
    CUtensorMap tmap;
    {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(grouped ? F_STAGE_C : p->C_in),
                                    static_cast<cuuint64_t>(grouped ? ((p->C_in + 31) / 32) * p->plane_pixels : npix)};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(grouped ? F_STAGE_C : p->in_pitch) * 4};
        const cuuint32_t box[2] = {F_STAGE_C, F_TILE_M};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p->in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return EML_E_ARG;
    }
    FArgs a{};
    a.scale = p->scale; a.shift = p->shift; a.wpack = static_cast<const unsigned char *>(p->wpack);
    a.bias9 = nullptr; a.out = p->out;
    a.B = p->B; a.H = p->H; a.W = p->W; a.R = R;
    a.C_in = p->C_in; a.out_pitch = p->out_pitch; a.out_choff = p->out_choff;
    a.nwchunks = (p->C_in + 63) / 64;
    a.nstg = (p->C_in + F_STAGE_C - 1) / F_STAGE_C;
    a.pool = 1; a.n_out = p->C_out;
    a.gstride = grouped ? p->plane_pixels : 0;
    a.wbytes = static_cast<int>(pool_wbytes(p->C_in));
    a.stages = fused_stages(a.wbytes, p->W);
    if (!eml_env_flag("EML_DENSE_SMEM_A")) a.stages &= ~3;          // TS: ring depth a multiple of the 4 converter warpgroups (see eml_dense_layer_forward)
    a.nbands = static_cast<long>(p->B) * (p->H / R);
    const bool split = p->precision == EML_PREC_BF16X3;
    const size_t smem = fused_smem(a.wbytes, p->W, a.stages);
    const unsigned grid = static_cast<unsigned>(a.nbands < sms ? a.nbands : sms);
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        kern<<<grid, f_threads(16), smem, st>>>(tmap, a);
        return EML_OK;
    };
    int rc;
    if (eml_env_flag("EML_DENSE_SMEM_A")) rc = split ? go(dense_layer_kernel<true, true, F_GRP, false>) : go(dense_layer_kernel<false, true, F_GRP, false>);
    else rc = split ? go(dense_layer_kernel<true, true, F_GRP, true>) : go(dense_layer_kernel<false, true, F_GRP, true>);
    if (rc != EML_OK) return rc;
    return eml_launch_status();
}
