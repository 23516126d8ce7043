// Spherical-Gaussian -> equirectangular panorama render (forward + backward), sm_100a.
//
// Replaces the reference's Python loop over lights (RegressionNetwork/util.py:222-245): N iterations of
// ~6 ATen kernels each re-reading and re-writing a (B,3,128,256) accumulator.  Here every output pixel is
// produced once: one thread owns 4 consecutive pixels of a row, the N lights of the sample sit in shared
// memory as two float4 each (direction + log2(e)/size, colour), and the only HBM traffic is the coalesced
// float4 store of the panorama -- 393,216 B written per map against ~7N floats read.
// Bound: MUFU.EX2 (N*32768 exp per map) balanced with ~8 FP32 ops per (light,pixel); see DESIGN.md.
#include "common.cuh"

namespace {

constexpr int PANO_H = 128;
constexpr int PANO_W = 256;
constexpr int PIX = PANO_H * PANO_W;
constexpr int MAX_LIGHTS = 512;
constexpr float LOG2E = 1.4426950408889634f;

struct RenderArgs {
    const float *dirs; long dirs_bs;
    const float *sizes; long sizes_bs;
    const float *colors;                 // (B,3N) or nullptr when composing from params
    const float *dist; long dist_bs;
    const float *intensity; long int_bs;
    const float *rgb; long rgb_bs;
    float gain;
    const float *ambient; long amb_bs;
    float *out;
    int B, N;
};

// Pixel direction exactly as util.py:223-233 builds it: fp32 (r+0.5)*(pi/128), fp32 sin/cos.
__device__ __forceinline__ void lat_terms(int r, float &s, float &c) {
    const float k = 0.02454369260617026f;     // float32(pi/128)
    float lat = (static_cast<float>(r) + 0.5f) * k;
    s = sinf(lat); c = cosf(lat);
}

__global__ void __launch_bounds__(256) sg_render_fwd_kernel(RenderArgs a) {
    __shared__ float4 s_dir[MAX_LIGHTS];   // (dx, dy, dz, log2e/size)
    __shared__ float4 s_col[MAX_LIGHTS];   // (r, g, b, -)
    const int b = blockIdx.y;
    const int N = a.N;
    const float *dirs = a.dirs + static_cast<long>(b) * a.dirs_bs;
    const float *sizes = a.sizes + static_cast<long>(b) * a.sizes_bs;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        float inv = LOG2E / sizes[k];
        s_dir[k] = make_float4(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2], inv);
        float cr, cg, cb;
        if (a.colors != nullptr) {
            const float *c = a.colors + static_cast<long>(b) * 3 * N + 3 * k;
            cr = c[0]; cg = c[1]; cb = c[2];
        } else {   // train.py:117-121 composition, same association order: (dist * (intensity*gain)) * rgb
            float di = a.dist[static_cast<long>(b) * a.dist_bs + k] * (a.intensity[static_cast<long>(b) * a.int_bs] * a.gain);
            const float *rgb = a.rgb + static_cast<long>(b) * a.rgb_bs;
            cr = di * rgb[0]; cg = di * rgb[1]; cb = di * rgb[2];
        }
        s_col[k] = make_float4(cr, cg, cb, 0.f);
    }
    __syncthreads();

    // 256 threads x 4 pixels = 1024 pixels = 4 panorama rows per block.
    const int p0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const int r = p0 / PANO_W, c0 = p0 % PANO_W;
    float sl, cl;
    lat_terms(r, sl, cl);
    float px[4], py[4];
    const float k = 0.02454369260617026f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lon = (static_cast<float>(c0 + i) + 0.5f) * k;
        px[i] = sl * cosf(lon);
        py[i] = sl * sinf(lon);
    }
    float acc[3][4];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[ch][i] = 0.f;

#pragma unroll 4
    for (int l = 0; l < N; ++l) {
        const float4 d = s_dir[l];
        const float4 col = s_col[l];
        const float dz = d.z * cl;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float t = fmaf(d.x, px[i], fmaf(d.y, py[i], dz));   // dirs . p
            float e = exp2f((t - 1.0f) * d.w);                   // exp((t-1)/size)
            acc[0][i] = fmaf(col.x, e, acc[0][i]);
            acc[1][i] = fmaf(col.y, e, acc[1][i]);
            acc[2][i] = fmaf(col.z, e, acc[2][i]);
        }
    }
    float amb[3] = {0.f, 0.f, 0.f};
    if (a.ambient != nullptr) {
        const float *am = a.ambient + static_cast<long>(b) * a.amb_bs;
        amb[0] = am[0]; amb[1] = am[1]; amb[2] = am[2];
    }
    float *o = a.out + static_cast<long>(b) * 3 * PIX + p0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float4 v = make_float4(acc[ch][0] + amb[ch], acc[ch][1] + amb[ch], acc[ch][2] + amb[ch], acc[ch][3] + amb[ch]);
        __stcs(reinterpret_cast<float4 *>(o + static_cast<long>(ch) * PIX), v);   // streaming store: written once, never re-read here
    }
}

// Backward.  grid = (pixel chunks, light groups of 8, B); one warp per light, lanes stride over the chunk's
// pixels keeping the 7 partial sums of that light in registers; one warp reduction + 7 atomics per warp.
constexpr int BWD_CHUNK = 4096;      // pixels per block (16 rows)
__global__ void __launch_bounds__(256) sg_render_bwd_kernel(const float *__restrict__ dirs_, long dirs_bs,
                                                            const float *__restrict__ sizes_, long sizes_bs,
                                                            const float *__restrict__ colors,
                                                            const float *__restrict__ go, float *g_dirs,
                                                            float *g_sizes, float *g_colors, int N) {
    __shared__ float s_slon[PANO_W], s_clon[PANO_W];
    const int b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float k = 0.02454369260617026f;
    for (int c = threadIdx.x; c < PANO_W; c += blockDim.x) {
        float lon = (static_cast<float>(c) + 0.5f) * k;
        s_slon[c] = sinf(lon); s_clon[c] = cosf(lon);
    }
    __syncthreads();
    const int l = blockIdx.y * 8 + warp;
    if (l >= N) return;
    const float *dirs = dirs_ + static_cast<long>(b) * dirs_bs + 3 * l;
    const float dx = dirs[0], dy = dirs[1], dz = dirs[2];
    const float size = sizes_[static_cast<long>(b) * sizes_bs + l];
    const float inv = LOG2E / size;
    const float *col = colors + static_cast<long>(b) * 3 * N + 3 * l;
    const float cr = col[0], cg = col[1], cb = col[2];
    const float *gob = go + static_cast<long>(b) * 3 * PIX;

    float a_c0 = 0.f, a_c1 = 0.f, a_c2 = 0.f, a_dx = 0.f, a_dy = 0.f, a_dz = 0.f, a_s = 0.f;
    const int pbeg = blockIdx.x * BWD_CHUNK;
    for (int row = 0; row < BWD_CHUNK / PANO_W; ++row) {
        const int r = pbeg / PANO_W + row;
        float sl, cl;
        lat_terms(r, sl, cl);
#pragma unroll 2
        for (int c = lane; c < PANO_W; c += 32) {
            const int p = r * PANO_W + c;
            const float pxv = sl * s_clon[c], pyv = sl * s_slon[c];
            const float t = fmaf(dx, pxv, fmaf(dy, pyv, dz * cl));
            const float e = exp2f((t - 1.0f) * inv);
            const float g0 = __ldg(gob + p), g1 = __ldg(gob + PIX + p), g2 = __ldg(gob + 2 * PIX + p);
            a_c0 = fmaf(g0, e, a_c0); a_c1 = fmaf(g1, e, a_c1); a_c2 = fmaf(g2, e, a_c2);
            const float w = (g0 * cr + g1 * cg + g2 * cb) * e;     // dL/d(exponent)
            a_dx = fmaf(w, pxv, a_dx); a_dy = fmaf(w, pyv, a_dy); a_dz = fmaf(w, cl, a_dz);
            a_s = fmaf(w, t - 1.0f, a_s);
        }
    }
    a_c0 = warp_sum(a_c0); a_c1 = warp_sum(a_c1); a_c2 = warp_sum(a_c2);
    a_dx = warp_sum(a_dx); a_dy = warp_sum(a_dy); a_dz = warp_sum(a_dz); a_s = warp_sum(a_s);
    if (lane == 0) {
        const long o3 = static_cast<long>(b) * 3 * N + 3 * l;
        if (g_colors) { atomicAdd(g_colors + o3, a_c0); atomicAdd(g_colors + o3 + 1, a_c1); atomicAdd(g_colors + o3 + 2, a_c2); }
        if (g_dirs) { const float is = 1.0f / size; atomicAdd(g_dirs + o3, a_dx * is); atomicAdd(g_dirs + o3 + 1, a_dy * is); atomicAdd(g_dirs + o3 + 2, a_dz * is); }
        if (g_sizes) atomicAdd(g_sizes + static_cast<long>(b) * N + l, -a_s / (size * size));
    }
}

int launch_fwd(const RenderArgs &a, cudaStream_t st) {
    if (a.B <= 0) return EML_OK;
    if (a.N < 1 || a.N > MAX_LIGHTS) return EML_E_SHAPE;
    dim3 grid(PIX / 1024, a.B);
    sg_render_fwd_kernel<<<grid, 256, 0, st>>>(a);
    return eml_launch_status();
}

}  // namespace

extern "C" int eml_sg_render_fwd(const float *dirs, long dirs_bstride, const float *sizes, long sizes_bstride,
                                 const float *colors, const float *ambient, float *out, int B, int N,
                                 void *stream) {
    if (B < 0) return EML_E_SHAPE;
    if (B == 0) return EML_OK;
    EML_CHECK_PTR(dirs); EML_CHECK_PTR(sizes); EML_CHECK_PTR(colors); EML_CHECK_PTR(out);
    EML_CHECK_ALIGN16(out);
    RenderArgs a{};
    a.dirs = dirs; a.dirs_bs = dirs_bstride; a.sizes = sizes; a.sizes_bs = sizes_bstride; a.colors = colors;
    a.ambient = ambient; a.amb_bs = 3; a.out = out; a.B = B; a.N = N;
    return launch_fwd(a, static_cast<cudaStream_t>(stream));
}

extern "C" int eml_sg_render_params_fwd(const float *dirs, long dirs_bstride, const float *sizes,
                                        long sizes_bstride, const float *dist, long dist_bstride,
                                        const float *intensity, long int_bstride, const float *rgb_ratio,
                                        long rgb_bstride, float gain, const float *ambient, long amb_bstride,
                                        float *out, int B, int N, void *stream) {
    if (B < 0) return EML_E_SHAPE;
    if (B == 0) return EML_OK;
    EML_CHECK_PTR(dirs); EML_CHECK_PTR(sizes); EML_CHECK_PTR(dist); EML_CHECK_PTR(intensity);
    EML_CHECK_PTR(rgb_ratio); EML_CHECK_PTR(out);
    EML_CHECK_ALIGN16(out);
    RenderArgs a{};
    a.dirs = dirs; a.dirs_bs = dirs_bstride; a.sizes = sizes; a.sizes_bs = sizes_bstride; a.colors = nullptr;
    a.dist = dist; a.dist_bs = dist_bstride; a.intensity = intensity; a.int_bs = int_bstride;
    a.rgb = rgb_ratio; a.rgb_bs = rgb_bstride; a.gain = gain; a.ambient = ambient; a.amb_bs = amb_bstride;
    a.out = out; a.B = B; a.N = N;
    return launch_fwd(a, static_cast<cudaStream_t>(stream));
}

extern "C" int eml_sg_render_bwd(const float *dirs, long dirs_bstride, const float *sizes, long sizes_bstride,
                                 const float *colors, const float *grad_out, float *g_dirs, float *g_sizes,
                                 float *g_colors, int B, int N, void *stream) {
    if (B < 0) return EML_E_SHAPE;
    if (B == 0) return EML_OK;
    EML_CHECK_PTR(dirs); EML_CHECK_PTR(sizes); EML_CHECK_PTR(colors); EML_CHECK_PTR(grad_out);
    if (N < 1 || N > MAX_LIGHTS) return EML_E_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g_dirs) cudaMemsetAsync(g_dirs, 0, sizeof(float) * 3 * N * B, st);
    if (g_sizes) cudaMemsetAsync(g_sizes, 0, sizeof(float) * N * B, st);
    if (g_colors) cudaMemsetAsync(g_colors, 0, sizeof(float) * 3 * N * B, st);
    dim3 grid(PIX / BWD_CHUNK, (N + 7) / 8, B);
    sg_render_bwd_kernel<<<grid, 256, 0, st>>>(dirs, dirs_bstride, sizes, sizes_bstride, colors, grad_out,
                                               g_dirs, g_sizes, g_colors, N);
    return eml_launch_status();
}
