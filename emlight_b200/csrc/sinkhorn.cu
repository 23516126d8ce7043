// Debiased Sinkhorn divergence over N sphere anchors: forward value and d/dx in ONE launch, sm_100a.
//
// Reference path (RegressionNetwork/geomloss): 4 (B,N,N) cost matrices in HBM, ~40 logsumexp passes of ~7
// ATen kernels each, and a .item() host sync for the eps schedule (sinkhorn_divergence.py:15).  Here one CTA
// owns one sample: the four cost matrices are never materialised (C_ij is rebuilt from x_i, y_j and the shared
// anchor-distance matrix M held in shared memory), the four dual potentials live in shared memory across the
// whole eps-scaling loop, and the eps schedule is derived on the device from the batch-wide min/max.
// One thread per (problem, row): rows of M are read through the symmetric transpose so a warp touches 32
// consecutive words (no bank conflicts); h_j / b_j reads are warp broadcasts.
//
// Bound: MUFU.EX2 -- (4*n_eps + 8) * N^2 exponentials per sample; HBM traffic is 2N+N floats per sample.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int MAX_N = 160;
constexpr int MAX_EPS = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// workspace layout: [0] min, [1] max (floats, as ordered ints for atomics)
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__global__ void sinkhorn_minmax_init(int *ws) {
    ws[0] = float_to_ordered(INFINITY);
    ws[1] = float_to_ordered(-INFINITY);
}

__global__ void __launch_bounds__(256) sinkhorn_minmax_kernel(const float *__restrict__ x,
                                                              const float *__restrict__ y, long n, int *ws) {
    float lo = INFINITY, hi = -INFINITY;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float a = x[i], b = y[i];
        lo = fminf(lo, fminf(a, b));
        hi = fmaxf(hi, fmaxf(a, b));
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(ws, float_to_ordered(lo));
        atomicMax(ws + 1, float_to_ordered(hi));
    }
}

// One softmin row:  -eps * logsumexp_j( h_j - C_ij/eps ),  C_ij = 0.5*(0.1*(a_i-b_j)^2 + M_ji).
// Everything is carried in base-2 units: u_j = (h_j - C_ij/eps) * log2(e).
// If WITH_MEAN, also returns E_P[b] = sum_j softmax_j * b_j (the only term the x-gradient needs).
template <bool WITH_MEAN>
__device__ __forceinline__ float softmin_row(const float *__restrict__ Mt_col,   // &Mt[0][i], stride ldm
                                             int ldm, const float *__restrict__ bvec,
                                             const float *__restrict__ h2,      // h_j * log2e (already scaled)
                                             float ai, int N, float eps, float inv_eps2 /* log2e/eps */,
                                             float *mean_out) {
    float m = -INFINITY;
    for (int j = 0; j < N; ++j) {
        float d = ai - bvec[j];
        float c = 0.5f * fmaf(0.1f * d, d, Mt_col[j * ldm]);
        m = fmaxf(m, fmaf(-c, inv_eps2, h2[j]));
    }
    float s = 0.f, sb = 0.f;
    for (int j = 0; j < N; ++j) {
        float bj = bvec[j];
        float d = ai - bj;
        float c = 0.5f * fmaf(0.1f * d, d, Mt_col[j * ldm]);
        float e = exp2f(fmaf(-c, inv_eps2, h2[j]) - m);
        s += e;
        if (WITH_MEAN) sb = fmaf(e, bj, sb);
    }
    if (WITH_MEAN) *mean_out = sb / s;
    return -eps * LN2 * (m + log2f(s));
}

struct SinkArgs {
    const float *x, *y, *M;
    float *loss, *grad;
    const int *ws;
    int N;
    float blur, scaling, diameter;
};

// dynamic smem: Mt[N][ldm] | xs[N] ys[N] | pot[2][4][N] | h2[4][N] | red[..]
__global__ void __launch_bounds__(640) sinkhorn_kernel(SinkArgs a) {
    extern __shared__ float smem[];
    __shared__ float s_eps[MAX_EPS];
    __shared__ int s_neps;
    __shared__ float s_red[32];

    const int N = a.N;
    const int ldm = N | 1;                       // odd pitch: column walks stay conflict-free for any N
    float *Mt = smem;
    float *xs = Mt + N * ldm;
    float *ys = xs + N;
    float *pot = ys + N;                          // [2][4][N]: a_x, b_y, a_y, b_x
    float *h2 = pot + 8 * N;                      // [4][N]: scaled log-weights for the current pass
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;

    for (int i = tid; i < N * N; i += nthr) {
        int r = i / N, c = i - r * N;
        Mt[c * ldm + r] = a.M[i];                 // transpose while staging (M is symmetric; this keeps it exact even if not)
    }
    for (int i = tid; i < N; i += nthr) {
        xs[i] = a.x[static_cast<long>(b) * N + i];
        ys[i] = a.y[static_cast<long>(b) * N + i];
    }
    if (tid == 0) {
        // eps schedule, sinkhorn_divergence.py:21-25 evaluated in double like numpy does.
        float df = a.diameter > 0.f ? a.diameter : fabsf(ordered_to_float(a.ws[1]) - ordered_to_float(a.ws[0]));
        double d = static_cast<double>(df), blur = static_cast<double>(a.blur), sc = static_cast<double>(a.scaling);
        int n = 0;
        s_eps[n++] = static_cast<float>(d * d);
        double start = 2.0 * log(d), stop = 2.0 * log(blur), step = 2.0 * log(sc);
        double cnt = ceil((stop - start) / step);
        int steps = cnt > 0 ? (cnt > MAX_EPS - 2 ? MAX_EPS - 2 : static_cast<int>(cnt)) : 0;
        for (int k = 0; k < steps; ++k) s_eps[n++] = static_cast<float>(exp(start + k * step));
        s_eps[n++] = static_cast<float>(blur * blur);
        s_neps = n;
    }
    __syncthreads();

    const float lw = logf(1.0f / static_cast<float>(N));
    const int neps = s_neps;
    // task t in [0, 4N): problem q = t / N (0: a_x<-C_xx,h=a_x | 1: b_y<-C_yy,h=b_y | 2: a_y<-C_yx,h=b_x | 3: b_x<-C_xy,h=a_y), row i.
    // row vector a / column vector bvec per problem:
    //   q0: a=x, b=x ; q1: a=y, b=y ; q2: a=y, b=x ; q3: a=x, b=y     (samples_loss.py:85-86)
    // potential feeding h per problem (sinkhorn_divergence.py:91-94): q0<-a_x(0), q1<-b_y(1), q2<-b_x(3), q3<-a_y(2)
    int cur = 0;
    // ---- initialisation at eps_s[0] with plain log-weights (sinkhorn_divergence.py:82-85)
    {
        const float eps = s_eps[0];
        for (int i = tid; i < 4 * N; i += nthr) h2[i] = lw * LOG2E;
        __syncthreads();
        for (int t = tid; t < 4 * N; t += nthr) {
            int q = t / N, i = t - q * N;
            const float *av = (q == 0 || q == 3) ? xs : ys;
            const float *bv = (q == 0 || q == 2) ? xs : ys;
            pot[(cur * 4 + q) * N + i] = softmin_row<false>(Mt + i, ldm, bv, h2 + q * N, av[i], N, eps, LOG2E / eps, nullptr);
        }
        __syncthreads();
    }
    // ---- eps-scaling descent with symmetrised updates (:87-97)
    for (int it = 0; it < neps; ++it) {
        const float eps = s_eps[it];
        const float inv_eps = 1.0f / eps;
        for (int t = tid; t < 4 * N; t += nthr) {
            int q = t / N, i = t - q * N;
            int src = (q == 2) ? 3 : (q == 3 ? 2 : q);
            h2[t] = (lw + pot[(cur * 4 + src) * N + i] * inv_eps) * LOG2E;
        }
        __syncthreads();
        for (int t = tid; t < 4 * N; t += nthr) {
            int q = t / N, i = t - q * N;
            const float *av = (q == 0 || q == 3) ? xs : ys;
            const float *bv = (q == 0 || q == 2) ? xs : ys;
            float nv = softmin_row<false>(Mt + i, ldm, bv, h2 + q * N, av[i], N, eps, LOG2E * inv_eps, nullptr);
            pot[((cur ^ 1) * 4 + q) * N + i] = 0.5f * (pot[(cur * 4 + q) * N + i] + nv);
        }
        __syncthreads();
        cur ^= 1;
    }
    // ---- last extrapolation at the final eps (:102-107) + cost (:65-69) + gradient
    {
        const float eps = s_eps[neps - 1];
        const float inv_eps = 1.0f / eps;
        for (int t = tid; t < 4 * N; t += nthr) {
            int q = t / N, i = t - q * N;
            int src = (q == 2) ? 3 : (q == 3 ? 2 : q);
            h2[t] = (lw + pot[(cur * 4 + src) * N + i] * inv_eps) * LOG2E;
        }
        __syncthreads();
        float part = 0.f;
        for (int t = tid; t < 4 * N; t += nthr) {
            int q = t / N, i = t - q * N;
            const float *av = (q == 0 || q == 3) ? xs : ys;
            const float *bv = (q == 0 || q == 2) ? xs : ys;
            float mean = 0.f;
            float f;
            if (q == 0 || q == 3) f = softmin_row<true>(Mt + i, ldm, bv, h2 + q * N, av[i], N, eps, LOG2E * inv_eps, &mean);
            else f = softmin_row<false>(Mt + i, ldm, bv, h2 + q * N, av[i], N, eps, LOG2E * inv_eps, nullptr);
            // loss = mean_i (b_x - a_x)_i + mean_j (a_y - b_y)_j
            part += (q == 3 || q == 2) ? f : -f;
            // stash E_P[b] for the gradient: q0 -> E_{P^xx}[x], q3 -> E_{P^xy}[y]
            if (q == 0) pot[((cur ^ 1) * 4 + 0) * N + i] = mean;
            if (q == 3) pot[((cur ^ 1) * 4 + 1) * N + i] = mean;
        }
        part = warp_sum(part);
        if ((tid & 31) == 0) s_red[tid >> 5] = part;
        __syncthreads();
        if (tid < 32) {
            float v = tid < ((nthr + 31) >> 5) ? s_red[tid] : 0.f;
            v = warp_sum(v);
            if (tid == 0) a.loss[b] = v / static_cast<float>(N);
        }
        if (a.grad != nullptr) {
            // d loss_b / d x_i = (0.1/N) * (E_{P^xx_i}[x] - E_{P^xy_i}[y])
            for (int i = tid; i < N; i += nthr)
                a.grad[static_cast<long>(b) * N + i] =
                    (0.1f / static_cast<float>(N)) * (pot[((cur ^ 1) * 4 + 0) * N + i] - pot[((cur ^ 1) * 4 + 1) * N + i]);
        }
    }
}

size_t smem_bytes(int N) { return sizeof(float) * (static_cast<size_t>(N) * (N | 1) + 2 * N + 8 * N + 4 * N); }

}  // namespace

extern "C" size_t eml_sinkhorn_workspace_bytes(int B, int N) { (void)B; (void)N; return 16; }

extern "C" int eml_sinkhorn_fwdbwd(const float *x, const float *y, const float *M, float *loss, float *grad_x,
                                   int B, int N, float blur, float scaling, float diameter, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    if (B < 0) return EML_E_SHAPE;
    if (B == 0) return EML_OK;
    EML_CHECK_PTR(x); EML_CHECK_PTR(y); EML_CHECK_PTR(M); EML_CHECK_PTR(loss); EML_CHECK_PTR(workspace);
    if (N < 8 || N > MAX_N) return EML_E_SHAPE;
    if (!(blur > 0.f) || !(scaling > 0.f && scaling < 1.f)) return EML_E_ARG;
    if (workspace_bytes < 16) return EML_E_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int *ws = static_cast<int *>(workspace);
    if (!(diameter > 0.f)) {
        sinkhorn_minmax_init<<<1, 1, 0, st>>>(ws);
        long n = static_cast<long>(B) * N;
        int blocks = static_cast<int>((n + 255) / 256);
        if (blocks > 148) blocks = 148;
        sinkhorn_minmax_kernel<<<blocks, 256, 0, st>>>(x, y, n, ws);
    }
    size_t sm = smem_bytes(N);
    cudaError_t e = cudaFuncSetAttribute(sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
    if (e != cudaSuccess) return static_cast<int>(e);
    int threads = ((4 * N + 31) / 32) * 32;
    if (threads > 640) threads = 640;
    SinkArgs a{x, y, M, loss, grad_x, ws, N, blur, scaling, diameter};
    sinkhorn_kernel<<<B, threads, sm, st>>>(a);
    return eml_launch_status();
}
