/*
 * stereoscene_b200 -- C ABI of the B200 (sm_100a) volumetric hot path of StereoScene / BRGScene.
 *
 * The reference has no native code and no FFI: its hot path is PyTorch module code that reaches
 * ATen / cuDNN / cuBLAS and ONE external compiled operator (mmdet3d.ops.bev_pool).  The drop-in
 * boundary is therefore (a) the mmdet3d registry names + module signatures, mirrored in Python by
 * stereoscene_b200/plugin (every module there), and (b) this C ABI, which is what those modules (or a maintainer's
 * ctypes / pybind stub) bind.  Every entry point names the reference code it
 * replaces (paths relative to /root/reference/projects/mmdet3d_plugin/occupancy/).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless a parameter is documented as host.  The caller owns
 *     every buffer, including workspaces.  No global state; re-entrant; one stream per call.
 *   - Volumes are channels-last fp32: x[b][d][h][w][c], c fastest, `ldc` floats between voxels
 *     (ldc >= C lets a layer write into a channel slice of a concatenation buffer).  This is the
 *     memory of a torch tensor of logical shape [B,C,D,H,W] in torch.channels_last_3d format.
 *   - "Pending affine": GroupNorm / BatchNorm(eval) / SE gates are never applied in a pass of
 *     their own.  A producer writes the raw tensor plus per-(batch,channel) sums; the consumer
 *     receives per-(batch,channel) scale/shift arrays (float[B*C], NULL = identity) and an
 *     activation code and applies act(x*scale+shift) while it loads x.
 *   - Return value: 0 on success, a negative SS_ERR_* code otherwise.  Nothing throws.
 *   - Streams are passed as void* (cudaStream_t).
 */
#ifndef STEREOSCENE_B200_H
#define STEREOSCENE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_OK 0
#define SS_ERR_INVALID_ARGUMENT (-1)   /* bad shape / unsupported configuration */
#define SS_ERR_CUDA (-2)               /* a CUDA runtime call failed; see ss_last_error_string() */
#define SS_ERR_WORKSPACE (-3)          /* caller-provided workspace too small */

#define SS_ACT_NONE 0
#define SS_ACT_RELU 1
#define SS_ACT_GELU 2                  /* exact erf GELU (nn.GELU default) */
#define SS_ACT_SWISH 3                 /* x * sigmoid(x): the image encoder's activation (efficientnet.py:374) */
#define SS_ACT_SIGMOID 4               /* ss_se_fc_fwd only (the SE gate) */

#define SS_MATH_TF32 0                 /* tensor-core TF32 multiply, fp32 accumulate */
#define SS_MATH_3XTF32 1               /* error-compensated split (hi/lo) TF32 on the mma.sync kernels: ~fp32 accuracy */
#define SS_MATH_TF32X3 2               /* the same compensation on the tcgen05 kernels (ss_conv3d_tc_fwd / _join_fwd): three TF32
                                          launches lo(x)*hi(w) + hi(x)*lo(w) + hi(x)*hi(w) accumulated in fp32 */

#define SS_MATH_F16X3 3                /* single-launch compensated mode of ss_conv3d_tc_fwd: both operands split into fp16 hi/lo halves
                                          (22 significand bits), six kind::f16 MMAs per 32-channel chunk = 1.5x the TF32 tensor work */

#define SS_MATH_F16 4                  /* fp16 operands (the 11-bit significand of TF32 in fp16's range: activations saturate at 65504, the
                                          weights are pre-scaled by a power of two), fp32 accumulate: kind::f16 runs at twice the TF32 MMA
                                          rate.  Same packing and availability as SS_MATH_F16X3 (only the hi halves are multiplied) */

/* ABI version of this header; ss_abi_version() of the library must match. */
#define SS_ABI_VERSION 6
int ss_abi_version(void);
/* Text of the last CUDA error seen by the calling thread (host pointer, never NULL). */
const char* ss_last_error_string(void);
/* Number of kernel launches issued through this library by the calling process so far. */
long long ss_launch_count(void);
/* Per-kernel launch census of the calling process: writes "kernel_name=count\n" lines (one per kernel family that has
 * launched at least once) into buf (HOST, NUL-terminated, truncated to cap bytes) and returns the number of families. */
int ss_kernel_census(char* buf, size_t cap);

/* ---------------------------------------------------------------------------------------------
 * Dense 3-D / 2-D convolution family (implicit GEMM on tensor cores).
 * Replaces every nn.Conv3d / nn.ConvTranspose3d / nn.Conv2d on the path:
 *   image2bev/ViewTransformerLSSVoxel.py:38-58 (stereofeature_net), :66-69 (convbn_3d),
 *   :73-88 (hourglass), :167-187 (dres0/1, classif3_*), :239-241 (MIE redir1/2),
 *   image2bev/attention.py:93-112 (CA3D), backbones/resnet3d.py:18-32, 143-148, 196-198,
 *   necks/second_fpn_3d.py:53-59, dense_heads/occhead.py:100-107.
 * ------------------------------------------------------------------------------------------- */
typedef struct ss_conv3d_desc {
    int32_t B;                        /* batch */
    int32_t Din, Hin, Win, Cin;       /* input volume (2-D conv: Din = 1, kd = 1) */
    int32_t Dout, Hout, Wout, Cout;   /* output volume */
    int32_t kd, kh, kw;               /* kernel extent, each in {1,2,3,4} */
    int32_t sd, sh, sw;               /* stride */
    int32_t pd, ph, pw;               /* padding */
    int32_t dd, dh, dw;               /* dilation (ordinary convolution only) */
    int32_t transposed;               /* 1 = ConvTranspose semantics (output_padding is implied by Dout) */
    int32_t in_ldc, out_ldc;          /* floats between consecutive voxels of x / y */
    int32_t in_act;                   /* SS_ACT_NONE | SS_ACT_RELU, applied after the pending affine of x */
    int32_t out_act;                  /* SS_ACT_*, applied after bias, before the statistics and the store */
    int32_t math;                     /* SS_MATH_* */
    int32_t cout_packed;              /* Cout rounded up to a multiple of 8: row length of w_packed */
    int32_t stats_d0, stats_d1;       /* only output planes d in [stats_d0, stats_d1) contribute to `stats` (stats_d1 <= stats_d0: all
                                         planes).  Lets a rank of the X-slab sharded mode run a layer on its slab + halo planes while the
                                         GroupNorm sums cover the slab only (the halo outputs are overwritten by the next exchange). */
    float acc_scale;                  /* SS_MATH_F16X3 only: power of two the accumulator is multiplied by (the packed weights were
                                         multiplied by its inverse so that their lo halves stay out of the fp16 subnormals) */
    int32_t accumulate;               /* ss_conv3d_tc_fwd only: y = act(conv + bias + y_old) -- the identity shortcut of a residual block whose
                                         branch ends in this convolution, taken in place on the block's input (efficientnet.py:219-222);
                                         not combined with `stats` */
    void* splitk_ws;                  /* ss_conv3d_tc_fwd only, optional: device workspace the layer may use to split its K range over several
                                         CTAs per output tile (partial tiles float[k][voxels][Cout], then one reduce kernel that adds bias /
                                         shortcut / activation in a fixed order).  Taken when the tiles alone fill less than half of the SMs and
                                         K is long: the 1x1 projections of the image encoder's late stages.  NULL: never split. */
    int64_t splitk_ws_bytes;          /* size of splitk_ws */
} ss_conv3d_desc;

/* w_packed: float[taps][Cin][cout_packed], taps = kd*kh*kw in (kd,kh,kw) row-major order, i.e.
 *   Conv:          w_packed[t][ci][co] = weight[co][ci][kd][kh][kw]
 *   ConvTranspose: w_packed[t][ci][co] = weight[ci][co][kd][kh][kw]      (columns >= Cout are zero)
 * bias: float[Cout] or NULL.  in_scale / in_shift: float[B*Cin] or NULL (pending affine of x).
 * stats: double[B][Cout][2] or NULL; the kernel ADDS sum(y) and sum(y*y) over the voxels of each
 *        (batch, output channel); the caller zeroes it beforehand. */
int ss_conv3d_fwd(const ss_conv3d_desc* desc, const float* x, const float* in_scale, const float* in_shift,
                  const float* w_packed, const float* bias, float* y, double* stats, void* stream);

/* Same contract on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM,
 * mbarrier-pipelined shared-memory staging with the pending affine applied by the producer warps).
 * Requirements: Cin % 32 == 0, in_ldc % 4 == 0, x 16-byte aligned, desc->math == SS_MATH_TF32 or SS_MATH_TF32X3.
 * w_kmajor: float[taps][cout_packed][Cin] (K-major: the 32-channel chunk of one output channel is one
 * 128-byte shared-memory row), values pre-rounded to TF32 (round-to-nearest, ties away from zero):
 *   Conv:          w_kmajor[t][co][ci] = tf32(weight[co][ci][kd][kh][kw])
 *   ConvTranspose: w_kmajor[t][co][ci] = tf32(weight[ci][co][kd][kh][kw])     (rows >= Cout are zero)
 * SS_MATH_TF32X3: w_kmajor holds TWO such arrays back to back, hi = tf32(w) followed by lo = tf32(w - hi); the activations
 * are split the same way on the fly and the three partial products are accumulated into y by three launches (y is read
 * back by the second and third), so the result has ~fp32 accuracy at ~3x the tensor work.
 * SS_MATH_F16X3 / SS_MATH_F16 (only where ss_conv3d_tc_f16x3_supported(desc) == 1: the halo-resident and per-tap box kernels): every
 * 128-byte row of w_kmajor, i.e. one output channel's 32-channel chunk, holds 64 fp16 values instead of 32 floats:
 * hi = fp16(w / acc_scale) of the 32 channels followed by lo = fp16(w / acc_scale - hi); same array extent as in TF32 mode. */
int ss_conv3d_tc_f16x3_supported(const ss_conv3d_desc* desc);
int ss_conv3d_tc_fwd(const ss_conv3d_desc* desc, const float* x, const float* in_scale, const float* in_shift,
                     const float* w_kmajor, const float* bias, float* y, double* stats, void* stream);

/* Convolution with the residual join fused into its epilogue:
 *   y = out_act( (conv(x) + bias) * out_scale + out_shift + res_act(res * res_scale + res_shift) )
 * out_scale / out_shift: float[B*Cout] (eval BatchNorm of the convolution result) or NULL; res: channels-last volume of
 * the output shape with voxel stride res_ldc, or NULL; res_scale / res_shift: its pending affine, float[B*Cout] or NULL.
 * Replaces conv5 / conv6 + the two joins of every hourglass (ViewTransformerLSSVoxel.py:92-95), i.e. saves writing and
 * re-reading the up-convolved tensor.  Served by the stride-2 transposed k3 kernel only: ask
 * ss_conv3d_tc_join_supported(desc) first (1 = yes); otherwise use ss_conv3d_tc_fwd + ss_affine_join_fwd. */
typedef struct ss_conv3d_join {
    const float* out_scale;
    const float* out_shift;
    const float* res;
    const float* res_scale;
    const float* res_shift;
    int32_t res_ldc;
    int32_t res_act;                  /* SS_ACT_NONE | SS_ACT_RELU, applied to the residual after its affine */
} ss_conv3d_join;
int ss_conv3d_tc_join_supported(const ss_conv3d_desc* desc);
int ss_conv3d_tc_join_fwd(const ss_conv3d_desc* desc, const float* x, const float* in_scale, const float* in_shift,
                          const float* w_kmajor, const float* bias, const ss_conv3d_join* join, float* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Normalisation bookkeeping on [B,C] vectors (the volume itself is never touched).
 * ------------------------------------------------------------------------------------------- */
/* GroupNorm (torch.nn.GroupNorm, eps inside the sqrt) from the sums a producer accumulated:
 * scale[b,c] = gamma[c]*rstd[b,g], shift[b,c] = beta[c] - mean[b,g]*scale[b,c].
 * Replaces the nn.GroupNorm layers built by build_norm_layer (ViewTransformerLSSVoxel.py:31,69;
 * attention.py:96,111; stereoscene.py:55).  count = voxels per (batch, channel).
 * scale/shift are written with row stride ld_out (>= C) at column offset 0. */
int ss_gn_finalize(const double* stats, const float* gamma, const float* beta, int B, int C, int groups,
                   double count, float eps, float* scale, float* shift, int ld_out, void* stream);

/* Same, with the result multiplied by up to two per-(batch,channel) gates > 0 (the SE gates of stereofeature_net / DepthNet,
 * ViewTransformerLSSBEVDepth.py:442-454, 507-511: relu(gn(y)) * g == relu(gn(y) * g)): scale_k = scale * gate_k,
 * shift_k = shift * gate_k, contiguous float[B*C] each; gate2 may be NULL. */
int ss_gn_finalize_gated(const double* stats, const float* gamma, const float* beta, int B, int C, int groups,
                         double count, float eps, const float* gate1, float* scale1, float* shift1,
                         const float* gate2, float* scale2, float* shift2, void* stream);

/* ASPP image-pooling branch (ViewTransformerLSSBEVDepth.py:373-379, 394-406) folded into the pending shift of the fusing conv's
 * BatchNorm: shift_out[b][j] = bn_shift[b][j] + bn_scale[b][j] * (Wp relu(GroupNorm(W1 mean_x)))[j], where mean_x comes from the
 * per-channel sums of the ASPP input (stats: double[B][C][2] as written by ss_channel_sums_fwd, count voxels each).
 * w1: float[mid][C] (the branch's 1x1 conv), w_pool: float[mid][mid] (conv1's columns that multiply the pooled branch);
 * t_ws: workspace float[B*mid]. */
int ss_aspp_pool_shift(const double* stats, double count, const float* w1, const float* gamma, const float* beta, int groups,
                       float eps, const float* w_pool, const float* bn_scale, const float* bn_shift, float* shift_out,
                       float* t_ws, int B, int C, int mid, void* stream);

/* CA3D squeeze-excite folded into the pending affine (attention.py:98-107, 113-118):
 * pool[b,c] = scale*mean_raw + shift (mean of the GroupNorm output), s = GELU(W2*GELU(W1*pool+b1)+b2),
 * then scale *= sigmoid(s), shift *= sigmoid(s) in place.  W1: [Cmid][C], W2: [C][Cmid]. */
int ss_ca3d_gate(const double* stats, double count, float* scale, float* shift, const float* w1, const float* b1,
                 const float* w2, const float* b2, int B, int C, int Cmid, void* stream);

/* out = act( alpha * A(x) + A(r) ), A(t) = act_t(t*scale_t + shift_t), on [B][V][C] channels-last
 * volumes.  r may be NULL (no second operand); alpha is a DEVICE pointer to one float or NULL (=1).
 * Replaces the residual joins ViewTransformerLSSVoxel.py:94-95, 215, 234 and resnet3d.py:62-63. */
int ss_affine_join_fwd(const float* x, const float* x_scale, const float* x_shift, int x_act,
                       const float* r, const float* r_scale, const float* r_shift, int r_act,
                       const float* alpha, int out_act, int B, long long V, int C,
                       int x_ldc, int r_ldc, int out_ldc, float* out, void* stream);

/* Per-(batch,channel) sums of a pending volume: stats[b][c][0] += sum_v a, stats[b][c][1] += sum_v a*a with
 * a = act(x*scale+shift) (double[B][C][2], caller zeroes it).  Serves the global-average-pool branch of
 * DepthNet's ASPP (image2bev/ViewTransformerLSSBEVDepth.py:373-379, 394) without a pass through ATen. */
int ss_channel_sums_fwd(const float* x, const float* x_scale, const float* x_shift, int x_act, int B,
                        long long V, int C, int x_ldc, double* stats, void* stream);

/* ---- 2-D image encoder (CustomEfficientNet-B7, backbones/efficientnet.py:113-231, 274-534).
 * Its pointwise convolutions run through ss_conv3d_tc_fwd / ss_conv3d_fwd as depth-1 volumes; these three entries are the rest.
 *
 * Stem (efficientnet.py:396-405): K x K stride-S convolution of a few-channel image with TensorFlow "SAME" padding
 * (mmcv Conv2dAdaptivePadding: total = max((ceil(H/S)-1)*S + K - H, 0), the smaller half in front), bias (the folded BatchNorm)
 * and activation.  x: float[N][Cin][H][W] (Cin <= 4); w: float[K*K*Cin][Cout] ordered (ky, kx, ci); y: channels-last
 * float[N][ceil(H/S)][ceil(W/S)][Cout]. */
int ss_stem_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int H, int W, int Cout,
                       int K, int S, int out_act, void* stream);

/* Depthwise K x K convolution (K = 3 | 5, stride 1 | 2, "SAME" padding as above) of a channels-last image
 * float[N][H][W][C] (pixel stride in_ldc) with bias and activation (InvertedResidual.depthwise_conv, efficientnet.py:181-190);
 * w: float[K*K][C].  pool (optional): double[N][C][2], slot 0 += the per-image sum of the activated output -- the global
 * average pool of the squeeze-excite block that follows (mmdet SELayer), taken in the same pass. */
int ss_dwconv2d_fwd(const float* x, const float* w, const float* bias, float* y, double* pool, int N, int H, int W, int C,
                    int in_ldc, int out_ldc, int K, int S, int out_act, void* stream);

/* One fully connected layer of the squeeze-excite block: out[n][o] = act(bias[o] + in_mul * sum_c in[n][c] * w[o][c]),
 * w: float[Cout][Cin]; act may be SS_ACT_SIGMOID.  in_is_stats != 0: in = double[N][Cin][2] channel sums (slot 0 read), so
 * in_mul = 1 / pixels makes it the pooled mean; otherwise in = float[N][Cin]. */
int ss_se_fc_fwd(const void* in, int in_is_stats, const float* w, const float* bias, float* out, int N, int Cin, int Cout,
                 float in_mul, int act, void* stream);

/* Softmax over the depth axis of a [B][D][P] volume (P = H*W pixels, contiguous), batch strides
 * in floats.  Replaces F.softmax(dim=1) at ViewTransformerLSSVoxel.py:222, 267 and
 * ViewTransformerLSSBEVDepth.py:107-108. */
int ss_softmax_d_fwd(const float* x, long long x_batch_stride, float* y, long long y_batch_stride,
                     int B, int D, int P, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stereo cost volume: group-wise correlation fused with the disparity->depth-bin resampling.
 * Replaces build_gwc_volume + groupwise_correlation + warp (ViewTransformerLSSVoxel.py:97-114,
 * 128-156); the [B,G,maxdisp,H,W] disparity volume is never materialised.
 *   fea:  float[2B][H][W][C] channels-last features, left = batches [0,B), right = [B,2B)
 *   i0:   int32[B][K], w0/w1: float[B][K]: per depth bin k the lower disparity index floor(p) and
 *         the two linear-interpolation weights (host code mirrors the reference's coordinate
 *         arithmetic; indices outside [0, maxdisp-1] contribute zero)
 *   out:  float[B][K][H][W][G] channels-last cost volume
 * ------------------------------------------------------------------------------------------- */
int ss_gwc_warp_fwd(const float* fea, const int32_t* i0, const float* w0, const float* w1, float* out,
                    int B, int C, int G, int H, int W, int K, int maxdisp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * BRI: confidence-weighted cross-volume attention (attention.py:58-86), flash-style: the
 * [N,N] energy / attention matrices are never materialised.
 *   q, kv: float[B][D][N] (N = H*W tokens);  params: DEVICE float[7] = wq,bq,wk,bk,wv,bv,gamma
 *   out[b][d][n*out_ld + 0] = gamma * sum_j V[d,j]*softmax_j(E[n,:])[j]*conf[j] + kv[b][d][n]
 *   ws: workspace of ss_bri_workspace_bytes(B, D, N) bytes (conf[j] = max_d softmax_d q[:,j] and, when
 *       the keys are split across CTAs to fill the GPU, the per-split partial softmax states).
 * ------------------------------------------------------------------------------------------- */
size_t ss_bri_workspace_bytes(int B, int D, int N);
int ss_bri_attn_fwd(const float* q, const float* kv, const float* params, float* ws, size_t ws_bytes, float* out,
                    int out_ld, int B, int D, int N, int math, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSS lift (x) splat.
 * ------------------------------------------------------------------------------------------- */
/* Voxel-index path of voxel_pooling (ViewTransformerLSSVoxel.py:441-451), bit-exact:
 * idx = trunc((geom - (bx - dx/2)) / dx) per axis in fp32, keep iff 0 <= idx < n per axis.
 *   geom: float[B][P][3] ego-frame points (P = D*H*W frustum points per sample)
 *   dx3, bx3: HOST float[3] voxel size and first-voxel centre (the module's dx / bx buffers)
 *   coords: int32[B*P][4] = (ix,iy,iz, kept ? 1 : 0), written for every point (may be NULL)
 *   order: int32[B*P] point ids sorted by voxel rank ((b*nx+ix)*ny+iy)*nz+iz, stable in point id;
 *          dropped points are sorted to the end
 *   voxel_start: int32[B*nx*ny*nz + 1] CSR offsets into `order`
 *   ws / ws_bytes: workspace (query the size with ss_splat_index_workspace_bytes). */
size_t ss_splat_index_workspace_bytes(long long n_points);
int ss_splat_build_index(const float* geom, const float* dx3, const float* bx3, int nx, int ny, int nz,
                         int B, long long P, int32_t* coords, int32_t* order, int32_t* voxel_start,
                         void* ws, size_t ws_bytes, void* stream);

/* Fused lift (outer product, ViewTransformerLSSVoxel.py:517-519) + per-voxel sum (bev_pool):
 *   out[b][x][y][z][c] = sum over the voxel's points n=(d,h,w), in ascending point id, of
 *                        depth_prob[b][d][h][w] * img_feat[b][h][w][c]
 * (fp32 multiply then add, no fma, so the sum reproduces a sequential CPU accumulation bit for bit).
 * The [B,N,D,H,W,C] lifted volume is never materialised; every voxel is written exactly once. */
int ss_lift_splat_fwd(const float* depth_prob, const float* img_feat, const int32_t* order,
                      const int32_t* voxel_start, float* out, int B, int D, int H, int W, int C,
                      int nx, int ny, int nz, void* stream);

/* Exact operator-level drop-in for mmdet3d.ops.bev_pool.bev_pool (call site
 * ViewTransformerLSSVoxel.py:473): feats float[N][C], coords int64[N][4] = (x,y,z,b) all inside the
 * grid; out float[B][C][D][H][W] with D=nz, H=nx, W=ny (NCDHW, contiguous); empty voxels are 0.
 * ws sized by ss_bev_pool_workspace_bytes(N, B*D*H*W). */
size_t ss_bev_pool_workspace_bytes(long long n_points, long long n_voxels);
int ss_bev_pool_fwd(const float* feats, const int64_t* coords, long long N, int C, int B, int D, int H, int W,
                    float* out, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Trilinear resize of channels-last logits, align_corners=False (F.interpolate semantics,
 * detectors/bevdepth_occupancy.py:293-294).  x: [B][Di][Hi][Wi][C] -> y: [B][Do][Ho][Wo][C].
 * labels (optional, may be NULL): uint8[B][Do][Ho][Wo] = argmax over C of y (first maximum).
 * ------------------------------------------------------------------------------------------- */
int ss_trilinear_fwd(const float* x, float* y, uint8_t* labels, int B, int C, int Di, int Hi, int Wi,
                     int Do, int Ho, int Wo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Semantic-scene-completion scores (the step after the path: utils/ssc_metric.py:62-85, 109-168; call site
 * occupancy/apis/test.py:113-115).  One pass over a predicted label volume and its ground truth:
 *   counts[t*C + p]   += number of voxels selected by `nonempty` (NULL = all) whose remapped target / prediction
 *                        are (t, p); where target == ignore_label both are remapped to 0, as the reference does
 *                        in place (ssc_metric.py:147-148)
 *   counts[C*C + 0..2] += completion tp, fp, fn over voxels with target != ignore_label (and nonempty, nonsurface),
 *                        occupied = label > 0
 *   counts[C*C + 3 + t] += voxels of target t whose prediction lies outside [0, C): misses of class t that are nobody's false positive
 * Per-class tp = counts[j*C+j], fp = column sum - tp, fn = row sum - tp + counts[C*C+3+j].  counts: int64[C*C + 3 + C], the caller
 * zeroes it (or keeps accumulating across samples).  pred: uint8[n]; target: uint8[n] or int64[n]
 * (target_elem_bytes = 1 / 8); nonempty / nonsurface: uint8[n] or NULL.  C <= 32.
 * ------------------------------------------------------------------------------------------- */
int ss_ssc_confusion_fwd(const uint8_t* pred, const void* target, int target_elem_bytes, const uint8_t* nonempty,
                         const uint8_t* nonsurface, long long n, int C, int ignore_label, long long* counts,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Deformable-convolution sampling (mmcv DeformConv2dPack / torchvision deform_conv2d semantics,
 * deform_groups = 1), DepthNet's DCN layer (image2bev/ViewTransformerLSSBEVDepth.py:490-498).
 *   x: [B][H][W][C] channels-last;  offsets: [B][2*kh*kw][Ho][Wo] (NCHW: 2t = dy, 2t+1 = dx)
 *   out: [B][Ho][Wo][groups][kh*kw][C/groups]  -- the im2col rows of each group, ready for a 1x1
 *        ss_conv3d_tc_fwd per group with K = kh*kw*C/groups.
 * ------------------------------------------------------------------------------------------- */
int ss_deform_sample_fwd(const float* x, const float* offsets, float* out, int B, int H, int W, int C,
                         int groups, int kh, int kw, int stride, int pad, int dil, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NVLink peer-memory collectives of the X-slab sharded mode (no reference counterpart: the reference only has DDP,
 * occupancy/apis/mmdet_train.py:75-79).  Every rank allocates one pool, exports it through CUDA IPC and opens every peer's
 * pool; exchanged tensors, flag words and slot areas live at the same offset in every pool.  `epoch` is a device counter
 * bumped once per forward (flags are never reset: a call waits for flag >= *epoch).  All calls are stream-ordered kernels
 * (capturable in a CUDA graph); they return 0 / SS_ERR_*.
 *   ss_peer_halo_push: buf = [n + 2][plane_floats] of this rank.  Copies plane 1 into the lower neighbour's plane n + 1 and
 *     plane n into the upper neighbour's plane 0 (peer_*_buf = the neighbour's buffer, NULL at the two ends of the grid, where
 *     the outer halo plane is zeroed or, with edge_replicate, filled with a copy of the edge plane), raises the neighbours' flags
 *     and waits until both neighbours have pushed theirs (a rank pushes only after the neighbour's stream has reached the same call, so
 *     the neighbour's producer kernel can no longer overwrite the halo).  my_flags / peer_*_flags: int[4] at the same pool offset on every rank;
 *     ticket: one zeroed unsigned int per call site.
 *   ss_peer_stats_allreduce: stats[n] (double) += the same array of every other rank, summed in rank order on every rank.
 *     slots[r] / flags[r]: rank r's slot area (world * n doubles) / flag array (world ints) for this call site, HOST arrays of
 *     device pointers (own pool for r == rank, peer mappings otherwise).
 * ------------------------------------------------------------------------------------------- */
int ss_peer_pool_alloc(size_t bytes, void** ptr);
int ss_peer_pool_free(void* ptr);
int ss_peer_ipc_export(void* ptr, void* handle64);                          /* handle64: HOST buffer of 64 bytes */
int ss_peer_ipc_open(const void* handle64, int peer_device, void** ptr);
int ss_peer_ipc_close(void* ptr);
int ss_peer_epoch_bump(int* epoch, void* stream);
int ss_peer_halo_push(float* buf, float* peer_lo_buf, float* peer_hi_buf, long long plane_floats, int n, int edge_replicate,
                      int* my_flags, int* peer_lo_flags, int* peer_hi_flags, unsigned int* ticket, const int* epoch, void* stream);
int ss_peer_stats_allreduce(double* stats, int n, int world, int rank, double* const* slots, int* const* flags,
                            const int* epoch, void* stream);

/* Layout helpers: NCHW/NCDHW <-> channels-last copies ([B][C][V] <-> [B][V][C]). */
int ss_nchw_to_nhwc(const float* x, float* y, int B, int C, long long V, int out_ldc, void* stream);
int ss_nhwc_to_nchw(const float* x, float* y, int B, int C, long long V, int in_ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STEREOSCENE_B200_H */
