// Micro-benchmark: achievable HBM read bandwidth for "channel prefix of an NHWC record" reads (the dense-block access pattern).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prefix_read_bw prefix_read_bw.cu && ./prefix_read_bw
#include <cstdio>
#include <cuda_runtime.h>

// variant 0: thread = one float4, consecutive threads walk the channels of a pixel then the next pixel (fully parallel, grid-stride)
__global__ void read_prefix(const float4 *__restrict__ in, long npix, int pitch4, int c4, float *sink, int unroll_dummy) {
    float acc = 0.f;
    const long total = npix * c4;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const long p = i / c4; const int c = (int)(i - p * c4);
        const float4 v = __ldg(in + p * pitch4 + c);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 12345.678f) *sink = acc;
}
// variant 1: like the fused kernel's producers: each CTA owns contiguous tiles of 128 pixels; thread = (row, sub) loads 64-channel chunks
// with `depth` chunks in flight (registers)
template <int DEPTH>
__global__ void __launch_bounds__(512, 1) read_tiles(const float4 *__restrict__ in, long ntiles, int pitch4, int c4, float *sink) {
    float acc = 0.f;
    const int row = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const int nch = (c4 + 15) / 16;
    const long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long t0 = blockIdx.x * per, t1 = min(ntiles, t0 + per);
    const long total = (t1 - t0) * nch;
    float4 buf[DEPTH][4];
    auto issue = [&](long g, float4 (&v)[4]) {
        const long t = t0 + g / nch; const int c = (int)(g % nch);
        const float4 *p = in + (t * 128 + row) * pitch4 + c * 16 + sub;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = make_float4(0, 0, 0, 0); if (c * 16 + k * 4 + sub < c4) v[k] = __ldg(p + k * 4); }
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) if (d < total) issue(d, buf[d]);
    for (long g = 0; g < total; g += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (g + d + DEPTH - 1 < total) issue(g + d + DEPTH - 1, buf[(d + DEPTH - 1) % DEPTH]);
            if (g + d < total) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += buf[d][k].x + buf[d][k].y + buf[d][k].z + buf[d][k].w;
            }
            __syncthreads();     // lockstep like the ring hand-off
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

// variant: the old conv1x1 producer mapping -- 16 lanes x 16 B = 256 B contiguous per row, rows rgrp + (T/16) i
template <int DEPTH, int T>
__global__ void __launch_bounds__(T, 1) read_tiles16(const float4 *__restrict__ in, long ntiles, int pitch4, int c4, float *sink) {
    float acc = 0.f;
    const int sub = threadIdx.x & 15, rgrp = threadIdx.x >> 4;
    constexpr int NR = 128 * 16 / T;
    const int nch = (c4 + 15) / 16;
    const long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long t0 = blockIdx.x * per, t1 = min(ntiles, t0 + per);
    const long total = (t1 - t0) * nch;
    float4 buf[DEPTH][NR];
    auto issue = [&](long g, float4 (&v)[NR]) {
        const long t = t0 + g / nch; const int c = (int)(g % nch);
        const float4 *p = in + (t * 128 + rgrp) * pitch4 + c * 16 + sub;
#pragma unroll
        for (int k = 0; k < NR; ++k) { v[k] = make_float4(0, 0, 0, 0); if (c * 16 + sub < c4) v[k] = __ldg(p + (long)k * (T / 16) * pitch4); }
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) if (d < total) issue(d, buf[d]);
    for (long g = 0; g < total; g += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (g + d + DEPTH - 1 < total) issue(g + d + DEPTH - 1, buf[(d + DEPTH - 1) % DEPTH]);
            if (g + d < total) {
#pragma unroll
                for (int k = 0; k < NR; ++k) acc += buf[d][k].x + buf[d][k].y + buf[d][k].z + buf[d][k].w;
            }
            __syncthreads();
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

int main() {
    const long npix = 256L * 192 * 256;
    const int pitch = 216;
    float *buf, *sink;
    cudaMalloc(&buf, npix * pitch * 4);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, npix * pitch * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int cs[] = {64, 128, 192};
    for (int ci = 0; ci < 3; ++ci) {
        const int c = cs[ci];
        for (int variant = 0; variant < 10; ++variant) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (variant == 0) read_prefix<<<148 * 8, 512>>>((const float4 *)buf, npix, pitch / 4, c / 4, sink, 0);
                if (variant == 1) read_tiles<2><<<148, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 2) read_tiles<3><<<148, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 3) read_tiles<4><<<148, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 4) read_tiles<2><<<148 * 2, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 5) read_tiles16<2, 512><<<148, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 6) read_tiles16<3, 512><<<148, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 7) read_tiles16<2, 1024><<<148, 1024>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 8) read_tiles16<4, 1024><<<148, 1024>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                if (variant == 9) read_tiles16<1, 512><<<148, 512>>>((const float4 *)buf, npix / 128, pitch / 4, c / 4, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("C_in %3d variant %d: %.3f ms  %.0f GB/s algorithmic\n", c, variant, best, npix * c * 4.0 / best / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
