// Micro-benchmark: HBM read bandwidth of TMA box loads [KW channels x 128 pixels] out of an NHWC slab (pitch 216 floats) into an
// shared-memory ring, consumers only touch the data (4 x LDS.128 per thread).   One CTA per SM, persistent over tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_read_bw tma_read_bw.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <int KW>   // floats per box row: 16 (64 B, SWIZZLE_64B) or 32 (128 B, SWIZZLE_128B)
__global__ void __launch_bounds__(512 + 32, 1) tma_read(const __grid_constant__ CUtensorMap tm, long ntiles, int c_in, int ns, float *sink) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[64];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int STAGE = KW * 4 * 128;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[32]);
    if (tid == 0) {
        for (int s = 0; s < ns; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nk = (c_in + KW - 1) / KW;
    const long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long t0 = blockIdx.x * per, t1 = min(ntiles, t0 + per);
    const long total = (t1 - t0) * nk;
    float acc = 0.f;
    if (warp == 16) {
        if ((tid & 31) == 0) {
            for (long g = 0; g < total; ++g) {
                const int s = g % ns; const uint32_t ph = (g / ns) & 1;
                mbar_wait(empty + 8 * s, ph ^ 1);
                mbar_expect_tx(full + 8 * s, STAGE);
                const long t = t0 + g / nk; const int k = (int)(g % nk);
                tma_load_2d(smem_u32(smem + (size_t)s * STAGE), &tm, k * KW, (int)(t * 128), full + 8 * s);
            }
        }
    } else {
        const int wg = warp >> 2, row = tid & 127;
        for (long g = wg; g < total; g += 4) {
            const int s = g % ns; const uint32_t ph = (g / ns) & 1;
            mbar_wait(full + 8 * s, ph);
            const float4 *p = reinterpret_cast<const float4 *>(smem + (size_t)s * STAGE + row * KW * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float4 v = p[i ^ ((row >> (KW == 16 ? 1 : 0)) & 3)]; acc += v.x + v.y + v.z + v.w; }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(empty + 8 * s);
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const long npix = 256L * 192 * 256;
    const int pitch = 216;
    float *buf, *sink;
    cudaMalloc(&buf, npix * pitch * 4);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, npix * pitch * 4);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int cs[] = {24, 48, 96, 144, 204};
    for (int kw = 16; kw <= 32; kw *= 2)
    for (int ci = 0; ci < 5; ++ci) {
        const int c = cs[ci];
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)npix};
        const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
        const cuuint32_t box[2] = {(cuuint32_t)kw, 128};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         kw == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int nss[] = {4, 8, 12, 16, 24};
        for (int ni = 0; ni < 5; ++ni) {
            const int ns = nss[ni];
            const size_t smem = (size_t)ns * kw * 4 * 128 + 1024;
            if (smem > 227 * 1024 || ns > 32) continue;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (kw == 16) { cudaFuncSetAttribute(tma_read<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tma_read<16><<<148, 544, smem>>>(tm, npix / 128, c, ns, sink); }
                else { cudaFuncSetAttribute(tma_read<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tma_read<32><<<148, 544, smem>>>(tm, npix / 128, c, ns, sink); }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("KW %2d C_in %3d stages %2d (%3zu KB): %.3f ms  %.0f GB/s algorithmic  [%s]\n", kw, c, ns, smem / 1024, best, npix * c * 4.0 / best / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
