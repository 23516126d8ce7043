/*
 * emlight_b200 -- C ABI of the B200 (sm_100a) implementation of EMLight's illumination hot path.
 *
 * One shared object (libemlight_b200.so), plain `extern "C"` entry points, raw DEVICE pointers,
 * explicit sizes, a CUDA stream passed as `void*` (a cudaStream_t; NULL = legacy default stream).
 * No torch types, no allocation, no implicit synchronisation, no global mutable state: every call
 * is re-entrant and enqueues work on the caller's stream.
 *
 * Return value: 0 = ok; negative = argument error (EML_E_*); positive = cudaError_t of the launch.
 * `eml_error_string()` renders either.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef EMLIGHT_B200_H
#define EMLIGHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EML_OK 0
#define EML_E_NULL (-1)      /* required pointer is NULL */
#define EML_E_SHAPE (-2)     /* unsupported size / shape */
#define EML_E_ALIGN (-3)     /* pointer or pitch not aligned as documented */
#define EML_E_ARG (-4)       /* invalid scalar argument */
#define EML_E_WORKSPACE (-5) /* workspace too small */

int eml_version(void);                 /* ABI version, bumped on any signature change */
const char *eml_error_string(int code); /* static string, never NULL */
int eml_device_ok(void);               /* 0 iff the current device is compute capability 10.x */

/* ------------------------------------------------------------------------------------------------
 * R2/R3 -- spherical-Gaussian -> 128x256 equirect panorama.
 * Replaces  RegressionNetwork/util.py:222-245 `convert_to_panorama(dirs, sizes, colors)`
 * (copies: representation/util.py:205-228, GenProjector/util.py:346-369, Needlets/utils.py:10-33).
 *   out[b,ch,r,c] = sum_k colors[b,3k+ch] * exp((dirs[b,3k:3k+3] . p(r,c) - 1) / sizes[b,k])  (+ ambient[b,ch])
 * dirs   (B,3N) fp32, row stride `dirs_bstride` floats (0 => one (3N) vector shared by the batch)
 * sizes  (B,N)  fp32, row stride `sizes_bstride` floats (0 => shared)
 * colors (B,3N) fp32 contiguous, k-major / channel-minor
 * ambient NULL or (B,3): added to every pixel (GenProjector/data.py:100)
 * out    (B,3,128,256) fp32 contiguous NCHW, 16-byte aligned.   1 <= N <= 512.
 */
int eml_sg_render_fwd(const float *dirs, long dirs_bstride, const float *sizes, long sizes_bstride,
                      const float *colors, const float *ambient, float *out, int B, int N, void *stream);

/* Same render with the light colours composed on the fly from the regression heads
 * (RegressionNetwork/train.py:117-121; GenProjector/data.py:86-98):
 *   colors[b,k,ch] = dist[b,k] * intensity[b] * gain * rgb_ratio[b,ch]
 * dist (B,N) row stride dist_bstride; intensity (B) stride int_bstride; rgb_ratio (B,3) stride rgb_bstride
 * (strides in floats, so the three can alias one packed head-output matrix). */
int eml_sg_render_params_fwd(const float *dirs, long dirs_bstride, const float *sizes, long sizes_bstride,
                             const float *dist, long dist_bstride, const float *intensity, long int_bstride,
                             const float *rgb_ratio, long rgb_bstride, float gain, const float *ambient,
                             long amb_bstride, float *out, int B, int N, void *stream);

/* Backward of eml_sg_render_fwd w.r.t. dirs, sizes, colors (any of the three outputs may be NULL).
 * Outputs are (B,3N),(B,N),(B,3N) contiguous and are OVERWRITTEN (zeroed inside the call, then accumulated). */
int eml_sg_render_bwd(const float *dirs, long dirs_bstride, const float *sizes, long sizes_bstride,
                      const float *colors, const float *grad_out, float *g_dirs, float *g_sizes,
                      float *g_colors, int B, int N, void *stream);

/* ------------------------------------------------------------------------------------------------
 * S1-S5 -- debiased Sinkhorn divergence over N anchors, forward value and d/dx in one launch.
 * Replaces  RegressionNetwork/geomloss/samples_loss.py:35-46,79-93 `SamplesLoss.forward(x, y)` with
 * loss="sinkhorn", p=2, reach=None, uniform weights (samples_loss.py:62-70), cost
 * C_ij = (0.1 (a_i-b_j)^2 + M_ij)/2 (geomloss/utils.py:85-99), the eps-scaling loop of
 * sinkhorn_divergence.py:72-109 and the cost of :65-69; gmloss/* only changes M.
 *   x, y   (B,N) fp32 contiguous (the reference's (B,N,1))
 *   M      (N,N) fp32 symmetric anchor distance matrix shared by the batch
 *   loss   (B)   out
 *   grad_x (B,N) out or NULL:  d loss[b] / d x[b,i]
 *   diameter <= 0  => computed on the device as |max - min| over all of x and y
 *                     (sinkhorn_divergence.py:9-18,28-31; replaces the reference's .item() host sync)
 *   workspace: eml_sinkhorn_workspace_bytes(B,N) bytes of device scratch, 16-byte aligned.
 * 8 <= N <= 160.
 */
size_t eml_sinkhorn_workspace_bytes(int B, int N);
int eml_sinkhorn_fwdbwd(const float *x, const float *y, const float *M, float *loss, float *grad_x, int B,
                        int N, float blur, float scaling, float diameter, void *workspace,
                        size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * D1-D5 -- DenseNet-BC convolutions as implicit GEMMs on NHWC fp32 activations.
 * Replaces the ATen conv2d/batch_norm/relu/cat/avg_pool2d chain issued by
 * RegressionNetwork/DenseNet.py:14-21 (_Transition), :26-55 (_DenseLayer), :88-93 (stem).
 *
 *   out[m, n] = sum_{tap, c} W[n, tap, c] * act( scale[c] * in[src(m, tap), c] + shift[c] )
 *
 * with act = ReLU or identity, src() the 1x1 / 3x3(pad 1) / 2x2-average-pool gather, out-of-image taps
 * contributing exactly 0 (zero padding is applied AFTER the affine, like the reference's BN -> conv).
 * `in` is an NHWC tensor whose pixel pitch (`in_pitch` floats) may exceed C_in: the dense block's
 * concatenation buffer is read in place and the 12 new channels are written in place at `out_choff`
 * (this is what removes the reference's torch.cat, DenseNet.py:55).
 */
typedef struct eml_conv_params {
    const float *in;        /* (B,H,W,in_pitch) fp32 NHWC, 16-byte aligned */
    const float *scale;     /* (C_in) per-channel multiplier of the fused BatchNorm (NULL => 1), 16-byte aligned */
    const float *shift;     /* (C_in) per-channel offset (NULL => 0), 16-byte aligned */
    const float *w_oihw;    /* (C_out,C_in,kh,kw) fp32 weights, reference layout (read by EML_PREC_FP32) */
    const void *wpack;      /* the same weights packed by eml_conv_pack_weights() (read by the tensor-core modes) */
    float *out;             /* (B,Ho,Wo,out_pitch) fp32 NHWC; channels [out_choff, out_choff+C_out) written */
    double *stats;          /* NULL or accumulators: stats[n] += sum, stats[stats_stride + n] += sum of squares */
    long stats_stride;      /* doubles between the sum row and the sum-of-squares row (0 => C_out) */
    int B, H, W;            /* input spatial size */
    int C_in, in_pitch;     /* channels read / pixel pitch in floats (multiple of 4) */
    int C_out, out_pitch, out_choff;
    int mode;               /* EML_CONV_1x1, EML_CONV_3x3, EML_CONV_POOL2 (act -> 2x2 average -> 1x1) */
    int relu;               /* 1: act = ReLU, 0: identity */
    int precision;          /* EML_PREC_BF16 (1 MMA pass), EML_PREC_BF16X3 (hi/lo split, fp32-grade), EML_PREC_FP32 (SIMT) */
    long plane_pixels;      /* 0: `in` is NHWC.  > 0 (EML_CONV_POOL2 on the TMA pipeline only, i.e. transition 1): `in` is the channel-plane
                             * slab of eml_dense_layer_params.plane_pixels; in_pitch is ignored.  Other modes reject it. */
} eml_conv_params;

#define EML_CONV_1x1 0
#define EML_CONV_3x3 1
#define EML_CONV_POOL2 2
#define EML_PREC_BF16 0
#define EML_PREC_BF16X3 1
#define EML_PREC_FP32 2

/* Bytes of the packed weight image for (C_out, C_in, taps); w is OIHW fp32 on the DEVICE. */
size_t eml_conv_wpack_bytes(int C_out, int C_in, int taps);
int eml_conv_pack_weights(const float *w_oihw, void *wpack, int C_out, int C_in, int taps, void *stream);
int eml_conv_forward(const eml_conv_params *p, void *stream);

/* D2 in ONE kernel for inference-mode BatchNorm -- norm1, relu1, conv1, norm2, conv2 of RegressionNetwork/DenseNet.py:26-55.
 * There is no nonlinearity between conv1 and conv2 (DenseNet.py:30-43), so with running statistics the layer is a single 3x3
 * convolution of a = relu(scale*x + shift) with the composite filter
 *     Weff[(dy,dx,o), c] = sum_b W2[o,b,dy,dx] * scale2[b] * W1[b,c]        (9*growth rows ordered dy, dx, o)
 * plus the position-dependent bias  sum_{(dy,dx) inside the image} sum_b W2[o,b,dy,dx] * shift2[b]  (zero padding is applied
 * to the norm2 OUTPUT).  The 4*growth-channel bottleneck is never written to memory.
 *   in     (B,H,W,in_pitch) fp32 NHWC slab, channels [0,C_in) read
 *   scale, shift  (C_in) folded norm1 (eml_bn_fold)
 *   wpack  eml_conv_pack_weights(Weff viewed as a (9*growth, C_in, 1, 1) filter, taps = 1)
 *   bias9  (3,3,growth) fp32: bias for [row class][column class], classes 0 = first row/column, 1 = interior, 2 = last
 *   out    (B,H,W,out_pitch); channels [out_choff, out_choff+growth) written (normally the same slab at out_choff = C_in)
 * Full-sector stores: when `out` is 32-byte aligned, out_pitch % 8 == 0, out_choff % 8 == 0 and out_choff + 16 <= out_pitch the
 * kernel writes 64 bytes per pixel instead of 48 -- it ZEROES channels [out_choff+12, out_choff+16) (in a dense block these belong
 * to the next layer, which writes them later) so that no 32-byte DRAM sector is left half-written.
 * Supported (eml_dense_layer_supported != 0): growth == 12, C_in <= 320, and W in {128, 256} with C_in % 4 == 0, or W == 64 with
 * C_in % 2 == 0 and an even B (a 128-pixel tile is then the same row of two consecutive images; float2 stores),
 * precision EML_PREC_BF16 / EML_PREC_BF16X3.  Other shapes: eml_conv_forward twice (conv1 then conv2). */
typedef struct eml_dense_layer_params {
    const float *in;
    const float *scale;
    const float *shift;
    const void *wpack;
    const float *bias9;
    float *out;
    int B, H, W;
    int C_in, in_pitch;
    int growth, out_pitch, out_choff;
    int precision;
    long plane_pixels;      /* 0: NHWC pixel records (above).  > 0: CHANNEL-PLANE slab, in == out: plane g holds channels [32g, 32g+32) of
                             * every pixel as 128-byte rows, consecutive planes plane_pixels rows apart (>= B*H*W); in_pitch / out_pitch are
                             * ignored.  A stage's TMA box is then one contiguous 16 KB run and the 12 new channels are written into
                             * consecutive rows instead of 64 bytes every in_pitch*4 bytes.  Channels the block has not produced yet must
                             * hold FINITE values (allocate the slab zeroed): they are multiplied by zero weights, not skipped.  W = 64
                             * (image-pair tiles) is not supported in this layout. */
} eml_dense_layer_params;
int eml_dense_layer_supported(int H, int W, int C_in, int growth, int precision);
/* 1 when eml_conv_forward(EML_CONV_POOL2, relu) for this transition runs on the TMA pipeline and therefore accepts a channel-plane
 * input (eml_conv_params.plane_pixels > 0) -- the caller's test before laying a block's slab out as planes. */
int eml_transition_planes_supported(int H, int W, int C_in, int C_out, int precision);
int eml_dense_layer_forward(const eml_dense_layer_params *p, void *stream);
/* The composite operands of eml_dense_layer_forward, built on the device once per parameter version (fp64 accumulation):
 *   w1 (nb, C_in) = conv1.weight (DenseNet.py:37), w2 (growth, nb, 3, 3) = conv2.weight (:42), scale2 / shift2 (nb) = folded norm2 (:41)
 *   wpack <- Weff[(dy,dx,o), c] = sum_b w2[o,b,dy,dx] scale2[b] w1[b,c] as bf16 hi / lo in the layout eml_dense_layer_forward reads
 *            (eml_dense_layer_wpack_bytes(C_in) bytes; the only producer of that operand);  bias9 (3,3,growth) <- the norm2 shift
 *            through the in-image taps. */
size_t eml_dense_layer_wpack_bytes(int C_in);
int eml_dense_layer_compose(const float *w1, const float *w2, const float *scale2, const float *shift2, int nb, int C_in, int growth,
                            void *wpack, float *bias9, void *stream);

/* Stem: conv0 3x3 (3 -> C_out<=32) on the NCHW input image, fused affine (+ReLU when relu != 0), NHWC output at channel 0.
 * Replaces DenseNet.py:89-92 (conv0, norm0, relu0).  scale/shift NULL => raw convolution output.
 * stats_raw: NULL or (2,C_out) double accumulators of the pre-affine values; stats_out: NULL or accumulators of
 * the written values with the sum-of-squares row `stats_out_stride` doubles after the sum row (0 => C_out). */
int eml_stem_forward(const float *x_nchw, const float *w_oihw, const float *scale, const float *shift,
                     float *out, int out_pitch, double *stats_raw, double *stats_out, long stats_out_stride,
                     int B, int H, int W, int C_out, int write_out, int relu, void *stream);

/* BatchNorm bookkeeping (DenseNet.py:17,30,41,91,122; eps 1e-5).
 * Given per-channel accumulators stats=(sum, sumsq) over `count` values of the STORED tensor t, an optional
 * upstream affine u = pre_scale*t + pre_shift (the folded last_norm of the previous transition), and this
 * norm's gamma/beta, writes scale/shift such that  BN_batchstat(u) = scale*t + shift.
 * stats[c] is the sum and stats[stats_stride + c] the sum of squares (stats_stride 0 => C).
 * With stats == NULL uses running_mean/running_var (eval mode) instead of batch statistics. */
int eml_bn_fold(const double *stats, long stats_stride, double count, const float *running_mean, const float *running_var,
                const float *gamma, const float *beta, const float *pre_scale, const float *pre_shift,
                float *scale, float *shift, float *batch_mean, float *batch_var, int C, float eps,
                void *stream);

/* Head: relu(scale*t+shift) -> 4x4 average pool -> (B, Hp*Wp*C) in NHWC order (DenseNet.py:136-138;
 * the fc weight columns are permuted to this order at pack time). */
int eml_head_pool(const float *in, int in_pitch, const float *scale, const float *shift, float *out, int B,
                  int H, int W, int C, int pool, void *stream);

/* out (M,N) = a (M,K) @ w (N,K)^T + bias (N), fp32 FFMA (DenseNet.py:139-150 fc and the four heads). */
int eml_linear_fp32(const float *a, const float *w, const float *bias, float *out, int M, int N, int K,
                    void *stream);

/* ------------------------------------------------------------------------------------------------
 * G1-G4 -- GenProjector SPADE / SphereNet generator building blocks (all activations NHWC fp32).
 *
 * eml_im2col_lut: the resampling half of SphereConv2D (GenProjector/models/networks/spherenet/sphere_cnn.py:111-124,
 * grid_sample on the tangent-plane pattern of :31-84) and of the ConvEncoder's stride-2 convolutions
 * (models/networks/generator.py:100-104):
 *     A[m, tap*Cp + c] = sum_{t<4} lut_w[p,tap,t] * act(x[b, lut_idx[p,tap,t], c] + bias[c])      m = b*out_pixels + p
 * lut_idx (out_pixels,9,4) int32 source pixel (or -1 = zero padding), lut_w (out_pixels,9,4) fp32, shared by the batch.
 * act: 0 none, 1 ReLU, 2 LeakyReLU(0.2).  Cp = C rounded up to 4 (padding columns are written as 0).
 * The convolution itself is eml_conv_forward(EML_CONV_1x1) on A with weights laid out (O, 9*Cp).
 */
int eml_im2col_lut(const float *x, int x_pitch, int C, int Cp, const int *lut_idx, const float *lut_w, const float *bias,
                   int act, float *A, int B, long out_pixels, long in_pixels, void *stream);

/* Same gather emitting the operand split into bf16 hi / lo matrices (row length Kp = multiple of 64 >= 9*Cp, zero padded)
 * for eml_gemm_bf16.  A_lo may be NULL (single-pass bf16). */
int eml_im2col_lut_bf16(const float *x, int x_pitch, int C, int Cp, const int *lut_idx, const float *lut_w, const float *bias,
                        int act, void *A_hi, void *A_lo, int Kp, int B, long out_pixels, long in_pixels, void *stream);

/* TMA-fed tcgen05 GEMM (the convolution half of SphereConv2D, sphere_cnn.py:123, and of the ConvEncoder convs):
 *   out[m, out_choff + n] = sum_k A[m,k] * W[n,k] (+ bias[n]),  A_hi/A_lo (M,Kp) bf16 row-major from eml_im2col_lut_bf16,
 *   wpack = eml_conv_pack_weights(W as (N, Kp, 1, 1)), N <= 256 per call, precision EML_PREC_BF16 or EML_PREC_BF16X3. */
int eml_gemm_bf16(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, int N, const float *bias, float *out,
                  int out_pitch, int out_choff, int precision, void *stream);
/* Split-K form for short-and-deep products (needlet projection, gt_gen_j3.py:39-43: M = 3 B rows, K = 32768 pixels): the K chunks are
 * dealt to `ksplit` work items per 128-row tile and accumulated with float atomics -- `out` must be ZERO on entry (1 <= ksplit <= Kp/64). */
int eml_gemm_bf16_splitk(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, int N, const float *bias, float *out,
                         int out_pitch, int out_choff, int precision, int ksplit, void *stream);

/* `nslices` output slices of exactly N columns (N a multiple of 16, <= 256) in ONE launch: slice s multiplies the same A by the packed
 * weights at wpack + s * slice_bytes (each slice packed by eml_conv_pack_weights as its own (N, Kp) matrix), adds bias[s*N + n] and writes
 * columns [out_choff + s*N, out_choff + (s+1)*N).  The wide low-resolution SphereConv layers (O = 512 / 1024 at 4x8 .. 16x32,
 * generator.py:40-52) otherwise run as O/256 launches of ceil(M/128) CTAs each; results are bit-identical to per-slice eml_gemm_bf16. */
int eml_gemm_bf16_slices(const void *A_hi, const void *A_lo, long M, int Kp, const void *wpack, long slice_bytes, int nslices, int N,
                         const float *bias, float *out, int out_pitch, int out_choff, int precision, int ksplit, void *stream);
/* Packs `nslices` slices of `rows` rows (a multiple of 16, <= 256) of the row-major (nslices*rows, K) fp32 matrix w in one launch:
 * slice s lands at wpack + s * slice_bytes (slice_bytes >= eml_conv_wpack_bytes(rows, K, 1)) in eml_conv_pack_weights' layout.
 * ksplit of eml_gemm_bf16_slices: 1, or the split-K factor of eml_gemm_bf16_splitk (then `out` must be ZERO on entry). */
int eml_gemm_pack_slices(const float *w, void *wpack, int nslices, int rows, int K, long slice_bytes, void *stream);

/* torch.nn.utils.spectral_norm's sigma for a weight viewed as (O, K) row-major (architecture.py:37-40, normalization.py:29 wrap every
 * generator / discriminator convolution): training != 0 runs ONE power iteration first, in place on the module's buffers --
 * v <- normalize(W^T u), u <- normalize(W v), eps as in torch (1e-12) -- then *sigma = u . (W v); training == 0 uses the stored u, v.
 * scratch: K + O floats.  Deterministic (fixed summation order). */
int eml_spectral_norm(const float *w, int O, int K, float *u, float *v, int training, float eps, float *scratch, float *sigma, void *stream);

/* SPADE.forward (models/networks/normalization.py:101-115) after the gamma/beta convolutions:
 *   out = ((x - mean[c]) * inv_std[c]) * (1 + gamma + bias_gamma[c]) + (beta + bias_beta[c]), optional LeakyReLU(0.2);
 * gamma_beta (M, gb_pitch) holds gamma in channels [0,C) and beta in [C,2C) (one GEMM with concatenated weights). */
int eml_spade_modulate(const float *x, int x_pitch, const float *mean, const float *inv_std, const float *gamma_beta,
                       int gb_pitch, const float *bias_gamma, const float *bias_beta, float *out, int out_pitch, long M,
                       int C, int leaky_relu, void *stream);

/* Training-mode statistics of SPADE's parameter-free (Sync)BatchNorm (normalization.py:80,104): sums[c] += sum_m x[m,c],
 * sums[C + c] += sum_m x[m,c]^2 over the M = B*H*W rows of an NHWC tensor (the caller zeroes `sums`; with several processes the
 * 2C doubles are all-reduced before the mean / variance are formed, which is what SynchronizedBatchNorm2d's master does). */
int eml_channel_stats(const float *x, int x_pitch, long M, int C, double *sums, void *stream);

/* out = a + bias_a (+ r + bias_r): SPADEResnetBlock's x_s + dx (models/networks/architecture.py:51-58). r may be NULL. */
int eml_bias_residual(const float *a, int a_pitch, const float *bias_a, const float *r, int r_pitch, const float *bias_r,
                      float *out, int out_pitch, long M, int C, void *stream);

/* F.interpolate(mode='nearest') to (Ho,Wo); source NHWC (pitch x_pitch) or NCHW (src_is_nchw=1); output NHWC
 * (normalization.py:107 guide resize, generator.py:42,70 upsampling / latent expansion). */
int eml_resize_nearest(const float *x, int x_pitch, int Hi, int Wi, float *out, int out_pitch, int Ho, int Wo, int C, int B,
                       int src_is_nchw, void *stream);

/* F.interpolate(mode='bilinear', align_corners=False): NCHW in, NHWC out (generator.py:116). */
int eml_resize_bilinear_nchw(const float *x, int Hi, int Wi, float *out, int out_pitch, int Ho, int Wo, int C, int B, void *stream);

/* nn.InstanceNorm2d(affine=False) + optional LeakyReLU(0.2) (normalization.py:45, generator.py:118-123). */
int eml_instance_norm(const float *x, int x_pitch, float *out, int out_pitch, int B, int HW, int C, float eps, int leaky_relu,
                      void *stream);

/* out_nchw = (tanh(x + bias[c]) + 1) * scale (generator.py:85-86). */
int eml_tanh_to_nchw(const float *x, int x_pitch, const float *bias, float *out, int B, int HW, int C, float scale, void *stream);

/* ------------------------------------------------------------------------------------------------
 * D1-D5 backward -- autograd of RegressionNetwork/DenseNet.py under training-mode BatchNorm (what train.py:82-102 runs).
 * The data-gradient convolutions are eml_conv_forward with transposed / flipped weights; these entry points add the rest.
 *
 * BatchNorm backward in two passes over a BN whose input is the stored tensor x (u = pre_a*x + pre_b, xhat = (u-mean)*inv_std,
 * z = gamma*xhat + beta, optional ReLU after it) and whose output gradient is `grad` (pool=1: grad is indexed by the 2x2-pooled
 * pixel and scaled by 1/4 -- the transition's avg_pool2d):
 *   reduce: sums[c] += sum g,  sums[sums_stride + c] += sum g*xhat      (g = grad masked by z > 0 when relu)   => d beta, d gamma
 *   apply : du = gamma*inv_std*(g - sums0/M - xhat*sums1/M);  out (=|+=) du   (times pre_a when to_stored)
 */
int eml_bn_bwd_reduce(const float *grad, int g_pitch, const float *x, int x_pitch, const float *pre_a, const float *pre_b,
                      const float *mean, const float *inv_std, const float *gamma, const float *beta, int relu, int pool, int H,
                      int W, long M, int C, double *sums, long sums_stride, void *stream);
int eml_bn_bwd_apply(const float *grad, int g_pitch, const float *x, int x_pitch, const float *pre_a, const float *pre_b,
                     const float *mean, const float *inv_std, const float *gamma, const float *beta, int relu, int pool, int H,
                     int W, long M, int C, const double *sums, long sums_stride, float *out, int out_pitch, int accumulate,
                     int to_stored, void *stream);
/* Weight gradients (accumulated into dW with atomics; zero dW first):
 *   1x1 : dW[n,c]      += sum_m G[m,n] * act(scale[c]*x[m,c] + shift[c])       (pool=1: act, then 2x2 average; m over pooled pixels)
 *         precision EML_PREC_FP32: fp32 FFMA; otherwise (N <= 64, C <= 256, no pool) a tcgen05 GEMM over K = pixels with MN-major operands
 *   3x3 : dW[n,c,tap]  += sum_m dY[m,n] * (scale[c]*b[m+tap,c] + shift[c])      (zero outside the image; N <= 16, C <= 64;
 *         precision other than EML_PREC_FP32 with C == 48: nine tcgen05 GEMMs over K = pixels on the rolling halo ring)
 *   stem: dW[o,ci,ky,kx] += sum_m dZ[m,o] * x_nchw[b,ci,y+ky-1,x+kx-1] */
int eml_wgrad_1x1(const float *G, int g_pitch, int N, const float *x, int x_pitch, int C, const float *scale, const float *shift,
                  int relu, int pool, int H, int W, float *dW, long M, int precision, void *stream);
int eml_wgrad_3x3(const float *dY, int dy_pitch, int N, const float *b, int b_pitch, int C, const float *scale, const float *shift,
                  float *dW, int B, int H, int W, int precision, void *stream);
int eml_wgrad_stem(const float *dZ, int dz_pitch, int O, const float *x_nchw, float *dW, int B, int H, int W, void *stream);
/* The transition's pooled activation (DenseNet.py:14-21, pooling commuted in front of the 1x1 conv), NHWC:
 *   out[b, y, x, c] = mean_{2x2} relu(scale[c] in[b, 2y+dy, 2x+dx, c] + shift[c]);  out_pitch >= round_up4(C), channels C..round_up4(C) zero.
 * Materialised once per transition so that its weight gradient runs on the tensor-core eml_wgrad_1x1 path (pool = 0 on `out`). */
int eml_pool_act(const float *in, int in_pitch, const float *scale, const float *shift, int B, int H, int W, int C, float *out, int out_pitch,
                 void *stream);

/* Fused backward of norm1 -> relu1 -> conv1 (DenseNet.py:30-37) into the block's gradient slab (csrc/dense_bwd1.cu), training-mode BN:
 *   dA = dN W1 on tensor cores, never written to memory;  g = dA * [sc x + sh > 0];  dS[:, c] += k1[c] g  (in place);
 *   sums[c] += sum g (= d beta),  sums[sums_stride + c] += sum g xhat (= d gamma),  xhat = e x + f.
 * BatchNorm's two mean terms are affine in the stored x with per-channel coefficients that add up over the layers reading a channel:
 * eml_dense_bwd1_accum folds a layer's sums into coefA / coefB (and emits d gamma / d beta), eml_dense_bwd1_gather applies them when a
 * channel range's gradient is read:  out[m, i] = dS[m, c0+i] + coefA[c0+i] + coefB[c0+i] x[m, c0+i]  (i < n; zero for n <= i < out_n).
 *   dN (M, 48) contiguous; x / dS (M, pitch) with channels [0, C_in); wpack from eml_dense_bwd1_pack(conv1.weight viewed (48, C_in));
 *   vec (5, vstride) from eml_dense_bwd1_prep: sc, sh (the forward's folded norm1 affine), e = pre_a inv, f = (pre_b - mean) inv,
 *   k1 = gamma inv;  M % 128 == 0, C_in <= 352;  precision BF16 / BF16X3.  sums: double, zeroed by the caller. */
size_t eml_dense_bwd1_wpack_bytes(int C_in);
int eml_dense_bwd1_supported(int C_in, long M, int precision);
int eml_dense_bwd1_pack(const float *w1, void *wpack, int C_in, void *stream);
int eml_dense_bwd1_prep(const float *sc, const float *sh, const float *pre_a, const float *pre_b, const float *mean, const float *inv_std,
                        const float *gamma, int C, int vstride, float *vec, void *stream);
int eml_dense_bwd1(const float *dN, const float *x, int x_pitch, float *dS, int ds_pitch, const void *wpack, const float *vec, int vstride,
                   int C_in, long M, double *sums, long sums_stride, int precision, void *stream);
int eml_dense_bwd1_accum(const double *sums, long sums_stride, const float *vec, int vstride, double n, int C, float *coefA, float *coefB,
                         float *dgamma, float *dbeta, void *stream);
int eml_dense_bwd1_gather(const float *dS, int ds_pitch, const float *x, int x_pitch, const float *coefA, const float *coefB, int c0, int n,
                          float *out, int out_pitch, int out_n, long M, void *stream);

/* ------------------------------------------------------------------------------------------------
 * G1-G6 backward -- adjoints used by the GenProjector training steps (pix2pix_model.py:92-141 -> trainers' loss.backward()).
 * The contractions (data gradient dA = dY Wk, weight gradient dWk^T = A^T dY) are eml_gemm_bf16[_splitk] on operands prepared with
 * eml_split_bf16 / eml_conv_pack_weights / eml_im2col_lut; these entry points are the HBM-bound passes around them.  Per-channel
 * sums are accumulated in double into caller-zeroed buffers.
 *
 * eml_col2im_lut: adjoint of the eml_im2col_lut gather (sphere_cnn.py:111-124 grid_sample backward):
 *     dx[b, lut_idx[p,tap,t], c] += lut_w[p,tap,t] * dA[b*out_pixels + p, tap*Cp + c]      (float atomics; dx ZERO on entry)
 * eml_col2im_csr: the same adjoint in gather form, without atomics: the table inverted on the host into CSR over input pixels
 *     (offs (in_pixels+1), src = p*9 + tap, w), dx[b,q,c] = sum_{e in [offs[q], offs[q+1])} w[e] * dA[b*out_pixels + src[e]/9, (src[e]%9)*Cp + c];
 *     every element of dx[..., :Cp] is written (no zeroing needed), deterministic summation order.
 * eml_act_bwd: dx *= act'(x + bias) in place (the activation eml_im2col_lut applies before gathering), bias_sums[c] += sum_m dx
 *     (act 0 with bias_sums != NULL: only the sums).
 * eml_bias_act_bwd: adjoint of eml_bias_act from its OUTPUT: dx = g * act'(out), bias_sums[c] += sum_m dx.
 * eml_spade_bwd: adjoint of eml_spade_modulate (normalization.py:101-115): with g' = g * lrelu'(out), xhat = (x - mean) * inv_std
 *     d_gb[:, c] = g' * xhat, d_gb[:, C + c] = g' (same pitch as gamma_beta), d_xhat = g' * (1 + gamma + bias_gamma),
 *     sums (4, C) += [sum g' xhat (d bias_gamma), sum g' (d bias_beta), sum d_xhat, sum d_xhat * xhat].
 * eml_bn_free_bwd: parameter-free BatchNorm backward (normalization.py:80,104): dx = inv_std * (d_xhat - s0/count - xhat * s1/count)
 *     with sums = (2, C) [sum d_xhat, sum d_xhat * xhat] (all-reduced by the caller when several processes share the batch);
 *     sums == NULL: running-statistics mode, dx = inv_std * d_xhat.
 * eml_instance_norm_bwd: adjoint of eml_instance_norm (InstanceNorm2d(affine=False) + optional LeakyReLU) from its output and raw
 *     input; `sums` is a caller-zeroed (B, 4, C) double scratch.
 * eml_im2col_lut_bf16_t: the eml_im2col_lut_bf16 gather written TRANSPOSED -- At[(tap*Cp + c), m], row length Mp >= B*out_pixels a
 *     multiple of 64, bf16 hi / lo (At_lo may be NULL), caller-zeroed -- the A operand (rows = filter taps x channels, K = pixels) of the
 *     weight-gradient GEMM dWk^T = A^T dY, so that no fp32 im2col matrix and no transpose pass exist.
 * eml_upsample2_bwd: adjoint of the nearest x2 upsampling (generator.py:42,70): dx[b,h,w,c] = sum of the 2x2 block of g.
 * eml_tanh_nchw_bwd: adjoint of eml_tanh_to_nchw from its NCHW output: d_raw (NHWC) = g * scale * (1 - tanh^2), bias_sums[c] += sum.
 * eml_pool2d_bwd: adjoint of eml_pool2d; mode 0 (3x3 average, stride 2, pad 1, count_include_pad=False) needs only g, mode 1 (2x2 max)
 *     routes g to the FIRST maximum of each window in row-major order (what ATen's max_pool2d backward does) and reads the input x.
 * eml_loss_seed: da = coef * (coef_dev ? *coef_dev : 1) * d(eml_loss_reduce(mode))/da for modes 0..5 (coef_dev: the upstream scalar
 *     gradient read on the device, so no host synchronisation is needed; mode 5: cosine distance with ATen's clamping,
 *     1 - (a/max(|a|,eps)).(b/max(|b|,eps)), eps 1e-20 as in pix2pix_model.py:95); b / mask as for eml_loss_reduce. */
int eml_col2im_lut(const float *dA, int Cp, const int *lut_idx, const float *lut_w, float *dx, int dx_pitch, int B, long out_pixels,
                   long in_pixels, void *stream);
int eml_col2im_csr(const float *dA, int Cp, const int *offs, const int *src, const float *w, float *dx, int dx_pitch, int B,
                   long out_pixels, long in_pixels, void *stream);
int eml_act_bwd(float *dx, int dx_pitch, const float *x, int x_pitch, const float *bias, int act, long M, int C, double *bias_sums,
                void *stream);
int eml_bias_act_bwd(const float *g, int g_pitch, const float *out, int out_pitch, int act, float *dx, int dx_pitch, long M, int C,
                     double *bias_sums, void *stream);
int eml_spade_bwd(const float *g, int g_pitch, const float *out, int out_pitch, const float *x, int x_pitch, const float *mean,
                  const float *inv_std, const float *gamma_beta, int gb_pitch, const float *bias_gamma, float *d_gb, float *d_xhat,
                  int dxh_pitch, long M, int C, int leaky_relu, double *sums, void *stream);
int eml_bn_free_bwd(const float *d_xhat, int dxh_pitch, const float *x, int x_pitch, const float *mean, const float *inv_std,
                    const double *sums, double count, float *dx, int dx_pitch, long M, int C, void *stream);
int eml_instance_norm_bwd(const float *g, int g_pitch, const float *out, int out_pitch, const float *raw, int raw_pitch, int B, long HW,
                          int C, float eps, int leaky_relu, double *sums, float *dx, int dx_pitch, void *stream);
int eml_im2col_lut_bf16_t(const float *x, int x_pitch, int C, int Cp, const int *lut_idx, const float *lut_w, const float *bias, int act,
                          void *At_hi, void *At_lo, long Mp, int B, long out_pixels, long in_pixels, void *stream);
int eml_upsample2_bwd(const float *g, int g_pitch, float *dx, int dx_pitch, int B, int H, int W, int C, void *stream);
int eml_tanh_nchw_bwd(const float *g_nchw, const float *out_nchw, float scale, float *d_raw, int pitch, int B, long HW, int C,
                      double *bias_sums, void *stream);
int eml_pool2d_bwd(const float *g, int g_pitch, const float *x, int x_pitch, float *dx, int dx_pitch, int Hi, int Wi, int C, int B,
                   int mode, void *stream);
int eml_loss_seed(const float *a, int a_pitch, const float *b, int b_pitch, const float *mask, long M, int C, int mode, float coef,
                  const float *coef_dev, float *da, int da_pitch, void *stream);

/* G6-G7 -- discriminator / loss building blocks (NHWC fp32).
 * eml_bias_act: out = act(x + bias[c]) (act 0 none, 1 ReLU, 2 LeakyReLU(0.2)); discriminator.py:91-92, VGG conv+ReLU; also the input
 *   transform of a SphereConv (relu(mlp_shared(seg)) in SPADE.forward, normalization.py:104-106) applied ONCE per value so that
 *   eml_im2col_lut_bf16 takes its bias-free fast path instead of re-applying it for each of the 36 (filter tap, bilinear tap) reads.
 * eml_pool2d : mode 0 = avg_pool2d(3, stride 2, pad 1, count_include_pad=False) (discriminator.py:48-51), mode 1 = max_pool2d(2,2) (VGG19).
 * eml_loss_reduce: *acc += sum of  0: a | 1: min(a-1,0) | 2: min(-a-1,0) | 3: |a-b| | 4: |a-b|*(m+(1-m)*50), m = mask[pixel] |
 *                  5: per pixel 1 - cos(a,b) over channels   (loss.py:57-82,109-114; pix2pix_model.py:101-122); caller divides by the count. */
int eml_bias_act(const float *x, int x_pitch, const float *bias, int act, float *out, int out_pitch, long M, int C, void *stream);
int eml_pool2d(const float *x, int x_pitch, int Hi, int Wi, float *out, int out_pitch, int C, int B, int mode, void *stream);
int eml_loss_reduce(const float *a, int a_pitch, const float *b, int b_pitch, const float *mask, long M, int C, int mode, double *acc,
                    void *stream);

/* ------------------------------------------------------------------------------------------------
 * N1-N2 -- spherical needlets (Needlets/sphere_needlets.py, mat_gen2.py, gt_gen_j3.py).
 *
 * eml_needlet_basis: SN_matrix of sphere_needlets.py:196-238 `SNvertex` ([Y_00 | psi_0 | ... | psi_jmax], float64) through the
 * addition theorem,  psi_jk(x) = sum_l coef[j][l] P_l(x . xi_jk),  coef[j][l] = sqrt(lambda_j) b(l/B^j) (2l+1)/(4 pi)  (zero outside
 * the level's band; the host builds it from sphere_needlets.py:10-29,39-50,73-74).
 *   xyz (P,3) unit vectors of the evaluation grid, centres (K,3) HEALPix cubature points of all levels concatenated
 *   (sphere_needlets.py:109-116), level (K) int32 level index of each centre, coef (nlev, lmax+1), out (P, out_pitch >= K+1).
 * eml_split_bf16: x (rows, cols) fp32, row stride ld -> bf16 hi / lo (rows, Kp >= cols, zero padded) = operands of eml_gemm_bf16,
 * which computes the projection coef = (SN*omega)^T pano (gt_gen_j3.py:39-43) and the reconstruction rec = SN coef (mat_gen2.py:55).
 * eml_needlet_sparsify: coef (B,n,ch) in place; for every range [ranges[2i], ranges[2i+1]) of rows, per image, zero the entries
 * with |v| <= frac * max|v| over the range (mat_gen2.py:43-51: the j=3 and j=2 blocks, frac 0.1). */
int eml_needlet_basis(const double *xyz, long P, const double *centres, const int *level, int K, const double *coef, int nlev,
                      int lmax, double *out, long out_pitch, void *stream);
int eml_split_bf16(const float *x, long rows, int cols, long ld, void *hi, void *lo, int Kp, void *stream);
int eml_needlet_sparsify(float *coef, int B, int n, int ch, const int *ranges, int nranges, float frac, void *stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY 8(f) rank 1 -- ground-truth light parameters from an HDR panorama, the inverse of the SG render.
 * Replaces RegressionNetwork/representation/distribution_representation.py:89-119 `extract_mesh.compute` for a batch:
 *   hdr (B,H,W,3) fp32; idx (H*W) int32 nearest-anchor LUT (host, :77-86); ster (H) float64 row weights sin((r+.5)/H*pi) (:69-74);
 *   lit = weighted intensity > 5 % of the image's maximum; anchors[k] = sum of lit weighted pixels with idx == k; ambient = the rest;
 *   dist (B,ln) = anchor energy / total, intensity (B) = |sum_k anchors[k]|, rgb_ratio (B,3), ambient (B,3), map (B,H,W) u8 or NULL.
 * Accumulation in float64 like the reference's numpy; ln <= 512. */
int eml_extract_params(const float *hdr, const int *idx, const double *ster, int B, int H, int W, int ln, float *dist,
                       float *intensity, float *rgb_ratio, float *ambient, unsigned char *map, void *stream);

/* SURVEY 8(f) rank 2 -- `TonemapHDR.__call__` (RegressionNetwork/util.py:36-66; applied to every crop by data.py:62-73):
 *   p = x^(1/gamma) (use_gamma) ; r = percentile_q of the strictly positive p, numpy's linear interpolation (exact radix selection per
 *   image, no sort) ; alpha[b] = max_mapping / (r + 1e-10) unless alpha_given ; out = alpha[b] * p, clipped to [0,1] when clip.
 * x, out (B, per_image) fp32 (out may alias x); alpha (B) fp32 in (alpha_given) or out. */
int eml_tonemap_hdr(const float *x, float *out, float *alpha, int B, long per_image, float gamma, float percentile, float max_mapping,
                    int use_gamma, int clip, int alpha_given, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Training step epilogue -- torch.optim.Adam over ONE flat fp32 buffer holding all parameters of a network
 * (RegressionNetwork/train.py:55-57 `Adam(lr=1e-4, betas=(0.9, 0.999))`, GenProjector/models/pix2pix_model.py:56-70 `betas=(0, 0.9)`),
 * with the 1/world_size of the gradient all-reduce (replaces nn.DataParallel's gather, model_trainer.py:20-24) folded in:
 *   g' = grad_scale * g ;  m = b1 m + (1-b1) g' ;  v = b2 v + (1-b2) g'^2 ;  p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 * p, g, m, v: n fp32 each, 16-byte aligned; step >= 1 is the 1-based update count (torch's state['step'] after increment). */
int eml_adam_step(float *p, const float *g, float *m, float *v, long n, float lr, float beta1, float beta2, float eps, int step,
                  float grad_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EMLIGHT_B200_H */
