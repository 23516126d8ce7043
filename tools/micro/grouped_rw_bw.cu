// Micro-benchmark for a CHANNEL-GROUP-MAJOR slab: buf[group][pixel][32 floats] (128-byte rows, one plane per 32-channel group).
// A "layer" with C_in input channels reads the first ceil(C_in/32) planes completely (contiguous) and writes its 12 new channels
// (48 bytes) at channel offset p = C_in: inside one plane's rows, or split over two planes when p % 32 > 20.  Compare with
// record_rw_bw.cu (interleaved 864-byte pixel records: the 64-byte scattered write costs 0.71 ms per layer at 12.58 M pixels).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/grouped_rw_bw tools/micro/grouped_rw_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

// 8 threads per 128-byte row (float4 each); a warp covers 4 consecutive rows of one plane
__global__ void __launch_bounds__(256) grouped_kernel(float *buf, long npix, int ngroups_read, int p, float *sink) {
    const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x, nt = static_cast<long>(gridDim.x) * blockDim.x;
    const int q = static_cast<int>(t & 7);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long plane4 = npix * 8;                              // float4 per plane
    float4 *b4 = reinterpret_cast<float4 *>(buf);
    for (long row = t >> 3; row < npix; row += nt >> 3) {
        for (int g = 0; g < ngroups_read; ++g) {
            const float4 v = __ldg(b4 + g * plane4 + row * 8 + q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        // write: 3 float4 at channel offset p (p % 4 == 0): lanes q = 0..2 write quad (p/4 + q)
        if (p >= 0 && q < 3) {
            const int cq = p / 4 + q;                          // absolute channel quad
            b4[(cq >> 3) * plane4 + row * 8 + (cq & 7)] = make_float4(1.f, 2.f, 3.f, 4.f);
        }
    }
    if (acc.x == 12345.f) sink[0] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
    const long npix = 256L * 192 * 256;
    const int ngroups = 7;                                     // 216 channels -> 7 planes of 32
    float *buf, *sink;
    cudaMalloc(&buf, npix * 128 * ngroups);
    cudaMalloc(&sink, 16);
    cudaMemset(buf, 0, npix * 128 * ngroups);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // (C_in read, write offset p or -1)
    const int cases[][2] = {{24, 24}, {24, -1}, {0, 24}, {0, 32}, {0, 36}, {0, 48}, {0, 60}, {36, 36}, {48, 48}, {72, 72}, {144, 144}, {144, -1}, {204, 204}, {204, -1}};
    for (auto &c : cases) {
        const int C = c[0], p = c[1];
        const int gr = (C + 31) / 32;
        float ms = 0.f;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            grouped_kernel<<<148 * 8, 256>>>(buf, npix, gr, p, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        const double useful = npix * (4.0 * C + (p >= 0 ? 48.0 : 0.0)), moved = npix * (128.0 * gr + (p >= 0 ? 48.0 : 0.0));
        printf("grouped: read C_in %3d (%d planes)  write 48 B @ ch %3d : %.3f ms  useful %.0f GB/s  (moved %.0f GB/s)\n", C, gr, p, ms, useful / ms / 1e6, moved / ms / 1e6);
    }
    return 0;
}
