// Micro-benchmark: tcgen05.ld / tcgen05.st throughput and round-trip latency (one CTA, 4 or 8 warps).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int NLD>
__global__ void tmem_bench(long long *out, int iters, int mode) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t r[16];
    float acc = 0.f;
    // initialise
    for (int c = 0; c < 512; c += 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(base + c), "r"(0) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int l = 0; l < NLD; ++l) {
            if (mode == 0 || mode == 1) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                               "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(base + ((it * NLD + l) * 16) % 496) : "memory");
                if (mode == 1) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += __uint_as_float(r[0]); }
            } else {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(base + ((it * NLD + l) * 16) % 496), "r"(it) : "memory");
            }
        }
        if (mode == 0) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += __uint_as_float(r[0]); }
        if (mode == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 1234.5f) out[1] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512) : "memory");
}
int main() {
    long long *d; cudaMalloc(&d, 16);
    const int iters = 2000;
    for (int threads = 128; threads <= 256; threads *= 2)
        for (int mode = 0; mode < 3; ++mode) {
            tmem_bench<8><<<1, threads>>>(d, iters, mode);
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double bytes = double(iters) * 8 * 16 * 4 * threads;
            printf("threads %d mode %s: %lld cycles, %.1f B/cycle, %.1f cycles per x16 instruction per warp  [%s]\n", threads,
                   mode == 0 ? "ld x8 then wait" : (mode == 1 ? "ld+wait each" : "st x8 then wait"), h[0], bytes / h[0], double(h[0]) / (iters * 8),
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
