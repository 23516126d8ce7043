// Micro-benchmark: what HBM delivers for the dense layer's ACCESS PATTERN, independent of any compute -- per pixel record (pitch P bytes)
// read a prefix of R bytes and write W bytes behind it, records visited in order by a grid of 148 x k CTAs with plain coalesced
// LDG.128 / STG.128 (16 lanes per 256 B).  Prints GB/s of useful bytes for the (R, W) pairs of block 1 (P = 864).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/record_rw_bw tools/micro/record_rw_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) rw_kernel(float *buf, long nrec, int pitch4, int r4, int w4, int woff4, float *sink) {
    // pitch4, r4, w4, woff4 in float4 units; a group of G = 64 threads handles one record at a time (up to 1 KB prefix in one pass)
    const int G = 64;
    const long group = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) / G, ngroups = static_cast<long>(gridDim.x) * blockDim.x / G;
    const int l = threadIdx.x % G;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long rec = group; rec < nrec; rec += ngroups) {
        float4 *p = reinterpret_cast<float4 *>(buf) + rec * pitch4;
        if (l < r4) { const float4 v = __ldg(p + l); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        if (l < w4) p[woff4 + l] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
    if (acc.x == 12345.f) sink[0] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
    const long nrec = 256L * 192 * 256;          // 12.58 M pixels (B = 256 block 1)
    const int pitch = 864;
    float *buf, *sink;
    cudaMalloc(&buf, nrec * pitch);
    cudaMalloc(&sink, 16);
    cudaMemset(buf, 0, nrec * pitch);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int cases[][3] = {{96, 64, 96}, {96, 48, 96}, {96, 0, 0}, {0, 64, 96}, {0, 48, 112}, {144, 48, 144}, {144, 64, 128}, {288, 64, 288},
                            {576, 64, 576}, {576, 0, 0}, {816, 48, 816}, {816, 0, 0}};
    for (auto &c : cases) {
        const int R = c[0], W = c[1], off = c[2];
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            rw_kernel<<<148 * 8, 256>>>(buf, nrec, pitch / 16, R / 16, W / 16, off / 16, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("pitch %d  read %3d B  write %2d B @ %3d : %.3f ms  %.0f GB/s useful\n", pitch, R, W, off, ms, nrec * double(R + W) / ms / 1e6);
    }
    return 0;
}
