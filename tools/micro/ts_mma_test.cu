// Micro-test: tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form), written there by tcgen05.st from registers.
// Validates the layout assumption behind the dense-layer kernel's TMEM-resident A ring:
//   A (M=128 x K bf16, K-major): row m = TMEM lane m, 32-bit column c holds the bf16 pair (k = 2c in the low half, k = 2c+1 in the
//   high half); one MMA of K = 16 reads 8 columns starting at the given TMEM address.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I emlight_b200/csrc -o /tmp/ts_mma_test tools/micro/ts_mma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"
using namespace eml;

constexpr int M = 128, N = 112, K = 32;      // two k-steps of 16
constexpr int ACOL = 256;                    // TMEM column of the A tile (D at column 0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) ts_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int mode) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    // B: N rows x 64 bf16 (only the first K columns are used), K-major SWIZZLE_128B
    for (int i = tid; i < N * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16 *>(smem + sw128_offset(n, k)) = k < K ? B[n * K + k] : __float2bfloat16(0.f);
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(&s_tmem), 512); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = s_tmem;
    // A row = thread (TMEM lane = 32 * warp + lane): 16 packed registers
    uint32_t r[16];
    for (int c = 0; c < 16; ++c) {
        const uint16_t lo = *reinterpret_cast<const uint16_t *>(&A[tid * K + 2 * c]);
        const uint16_t hi = *reinterpret_cast<const uint16_t *>(&A[tid * K + 2 * c + 1]);
        r[c] = mode == 0 ? (static_cast<uint32_t>(hi) << 16) | lo : (static_cast<uint32_t>(lo) << 16) | hi;
    }
    const uint32_t ta = tb + ACOL + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(M, N);
            const uint64_t db = make_sw128_desc(smem_u32(smem));
            for (int k = 0; k < K / 16; ++k)
                umma_bf16_ts(tb, tb + ACOL + 8 * k, db + 2 * k, idesc, k ? 1u : 0u);
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t td = tb + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(td + c0, v);
        for (int e = 0; e < 16; ++e) D[tid * N + c0 + e] = v[e];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tb, 512); }
}

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 700.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float *dD;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    const int smem = N * 128 + 2048;
    cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int ok_mode = -1;
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dD, 0, M * N * 4);
        ts_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(M * N);
        cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += static_cast<double>(fA[m * K + k]) * fB[n * K + k];
                maxerr = fmax(maxerr, fabs(s - hD[m * N + n])); maxref = fmax(maxref, fabs(s));
            }
        printf("mode %d (%s): max |err| %.3e (max |ref| %.3e) -> %s\n", mode, mode == 0 ? "k even in low half" : "k even in high half", maxerr, maxref,
               maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
        if (maxerr < 1e-3 * maxref) ok_mode = mode;
    }
    printf("TS MMA layout %s\n", ok_mode == 0 ? "CONFIRMED (lane = row, column c = bf16 pair (2c, 2c+1), low half first)" : ok_mode == 1 ? "is the swapped packing" : "NOT UNDERSTOOD");
    return ok_mode == 0 ? 0 : 2;
}
