// The CUDA-core kernels of the 2-D image encoder (CustomEfficientNet-B7, reference
// projects/mmdet3d_plugin/occupancy/backbones/efficientnet.py:113-231, 274-534).  The pointwise convolutions of the
// MBConv blocks (expand 1x1, linear 1x1, the 1x1 head, the SECONDFPN deblocks) are GEMMs and run on the tcgen05 kernels
// of conv3d_tc.cu as depth-1 volumes; what is left is HBM-bound and lives here:
//
//   stem_conv_kernel      3 -> Cout, 3x3 stride 2, TF-"SAME" padding, NCHW fp32 images in, channels-last out, folded
//                         BatchNorm bias + Swish (efficientnet.py:396-405).  K = 27: CUDA cores, weights in shared memory.
//   dwconv2d_kernel<K,S>  depthwise K x K (3 or 5), stride S (1 or 2), TF-"SAME" padding, channels-last, folded BatchNorm
//                         bias + Swish (efficientnet.py:181-190) and, in the same pass, the per-(image, channel) sums that
//                         the squeeze-excite block's global average pool needs (mmdet SELayer) -- the activation is read
//                         once and written once.  One thread owns 4 channels x TW neighbouring output columns, so an input
//                         value is loaded once per (TW-1)*S+K columns instead of once per tap.
//   se_fc_kernel          the two tiny fully connected layers of the SE block (C -> C/24 Swish, C/24 -> C sigmoid): one warp
//                         per output, coalesced weight rows.  The resulting gate [image, channel] is NOT multiplied into the
//                         activation: it travels as the pending per-(batch, channel) scale of the linear 1x1 convolution's
//                         input and is applied by that kernel's operand fix-up (no extra pass).
#include "common.cuh"

namespace ss {

// ---------------------------------------------------------------------------------------------------------------------
// stem: block = (Cout/4 channel quads) x PX output pixels of one output row
// ---------------------------------------------------------------------------------------------------------------------
constexpr int STEM_PX = 16;

__global__ void __launch_bounds__(512)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                 int Cin, int H, int W, int Ho, int Wo, int Cout, int K, int S, int pt, int pl, int out_act) {
    extern __shared__ float stem_smem[];
    const int taps = K * K * Cin;
    float* sw = stem_smem;                                   // [taps][Cout]
    float* sx = stem_smem + (size_t)taps * Cout;             // [Cin][K][span]
    const int span = (STEM_PX - 1) * S + K;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const int n = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * STEM_PX;
    for (int i = tid; i < taps * Cout; i += nthr) sw[i] = __ldg(w + i);
    const int ix0 = ox0 * S - pl, iy0 = oy * S - pt;
    for (int i = tid; i < Cin * K * span; i += nthr) {
        const int c = i / (K * span), r = (i / span) % K, j = i % span;
        const int iy = iy0 + r, ix = ix0 + j;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((size_t)n * Cin + c) * H + iy) * W + ix);
        sx[i] = v;
    }
    __syncthreads();
    const int cq = threadIdx.x, px = threadIdx.y, ox = ox0 + px;
    if (ox >= Wo || cq * 4 >= Cout) return;
    float4 acc = ldg_f4(bias + cq * 4);
    for (int ky = 0; ky < K; ++ky)
        for (int kx = 0; kx < K; ++kx)
            for (int c = 0; c < Cin; ++c) {
                const float v = sx[(c * K + ky) * span + px * S + kx];
                const float4 wv = *reinterpret_cast<const float4*>(sw + (size_t)((ky * K + kx) * Cin + c) * Cout + cq * 4);
                acc.x = fmaf(v, wv.x, acc.x); acc.y = fmaf(v, wv.y, acc.y); acc.z = fmaf(v, wv.z, acc.z); acc.w = fmaf(v, wv.w, acc.w);
            }
    acc.x = apply_act_sw(acc.x, out_act); acc.y = apply_act_sw(acc.y, out_act); acc.z = apply_act_sw(acc.z, out_act); acc.w = apply_act_sw(acc.w, out_act);
    *reinterpret_cast<float4*>(y + (((size_t)n * Ho + oy) * Wo + ox) * Cout + cq * 4) = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// depthwise: block = qx channel quads x gy column groups (qx = the largest divisor of C/4 that is <= 32, so no lane idles on
// 288- or 1344-channel tensors); a block walks ``rows`` output rows.  Loads use clamped addresses and a zero mask instead of
// branches, so the (TW-1)*S+K loads of a row are issued back to back.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DW_MAX_THREADS = 320;

template <int K, int S, int TW>
__global__ void __launch_bounds__(DW_MAX_THREADS)
dwconv2d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                double* __restrict__ pool, int H, int W, int C, int in_ldc, int Ho, int Wo, int out_ldc, int pt, int pl, int out_act,
                int row_blocks, int rows) {
    constexpr int SPAN = (TW - 1) * S + K;
    __shared__ float4 red[DW_MAX_THREADS];
    const int qx = blockDim.x, gy = blockDim.y;
    const int cq = blockIdx.x * qx + threadIdx.x;            // channel quad (always < C/4: qx divides C/4)
    const int n = blockIdx.z / row_blocks, rb = blockIdx.z % row_blocks;
    const int ox0 = (blockIdx.y * gy + threadIdx.y) * TW;
    const int c = cq * 4;
    float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ox0 < Wo) {
        const float4 bv = ldg_f4(bias + c);
        const int ix0 = ox0 * S - pl;
        int xoff[SPAN];
        float xm[SPAN];
#pragma unroll
        for (int j = 0; j < SPAN; ++j) {
            const int ix = ix0 + j;
            const bool ok = ix >= 0 && ix < W;
            xoff[j] = (ok ? ix : 0) * in_ldc;
            xm[j] = ok ? 1.f : 0.f;
        }
        // K = 3: the nine weight vectors stay in registers over the block's rows; K = 5 (25 vectors) reloads the row's five per ky
        constexpr bool HOIST = (K == 3);
        float4 wall[HOIST ? K * K : 1];
        if constexpr (HOIST) {
#pragma unroll
            for (int i = 0; i < K * K; ++i) wall[i] = ldg_f4(w + (size_t)i * C + c);
        }
        for (int r = 0; r < rows; ++r) {
            const int oy = rb * rows + r;
            if (oy >= Ho) break;
            float4 acc[TW];
#pragma unroll
            for (int t = 0; t < TW; ++t) acc[t] = bv;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const int iy = oy * S - pt + ky;
                if (iy < 0 || iy >= H) continue;              // uniform over the block
                const float* xr = x + (((size_t)n * H + iy) * W) * in_ldc + c;
                float4 v[SPAN];
#pragma unroll
                for (int j = 0; j < SPAN; ++j) v[j] = ldg_f4(xr + xoff[j]);
                float4 wv[K];
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    if constexpr (HOIST) wv[kx] = wall[ky * K + kx];
                    else wv[kx] = ldg_f4(w + (size_t)(ky * K + kx) * C + c);
                }
#pragma unroll
                for (int j = 0; j < SPAN; ++j) {
                    const float m = xm[j];
                    const float4 u = make_float4(v[j].x * m, v[j].y * m, v[j].z * m, v[j].w * m);
#pragma unroll
                    for (int t = 0; t < TW; ++t) {
                        const int kx = j - t * S;             // compile-time after unrolling
                        if (kx >= 0 && kx < K) {
                            acc[t].x = fmaf(u.x, wv[kx].x, acc[t].x); acc[t].y = fmaf(u.y, wv[kx].y, acc[t].y);
                            acc[t].z = fmaf(u.z, wv[kx].z, acc[t].z); acc[t].w = fmaf(u.w, wv[kx].w, acc[t].w);
                        }
                    }
                }
            }
            float* yr = y + (((size_t)n * Ho + oy) * Wo) * out_ldc + c;
#pragma unroll
            for (int t = 0; t < TW; ++t) {
                if (ox0 + t < Wo) {
                    float4 o = acc[t];
                    o.x = apply_act_sw(o.x, out_act); o.y = apply_act_sw(o.y, out_act); o.z = apply_act_sw(o.z, out_act); o.w = apply_act_sw(o.w, out_act);
                    *reinterpret_cast<float4*>(yr + (size_t)(ox0 + t) * out_ldc) = o;
                    psum.x += o.x; psum.y += o.y; psum.z += o.z; psum.w += o.w;
                }
            }
        }
    }
    if (pool == nullptr) return;
    red[threadIdx.y * qx + threadIdx.x] = psum;
    __syncthreads();
    if (threadIdx.y == 0) {
        float4 s = red[threadIdx.x];
        for (int g = 1; g < gy; ++g) {
            const float4 t = red[g * qx + threadIdx.x];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        double* dst = pool + ((size_t)n * C + c) * 2;         // double[B][C][2] like the GroupNorm sums: slot 0 = sum
        atomicAdd(dst + 0, (double)s.x); atomicAdd(dst + 2, (double)s.y); atomicAdd(dst + 4, (double)s.z); atomicAdd(dst + 6, (double)s.w);
    }
}

template <int K, int S, int TW>
static int launch_dw(const float* x, const float* w, const float* bias, float* y, double* pool, int N, int H, int W, int C, int in_ldc,
                     int Ho, int Wo, int out_ldc, int pt, int pl, int out_act, cudaStream_t st) {
    const int Q = C / 4;
    int qx = 1;
    for (int d = 1; d <= 32; ++d)
        if (Q % d == 0) qx = d;
    const int groups = (Wo + TW - 1) / TW;
    int gy = min(max(256 / qx, 1), groups);
    gy = (groups + (groups + gy - 1) / gy - 1) / ((groups + gy - 1) / gy);       // same block count, least idle groups
    // rows per block: enough blocks for ~4 waves of 148 SMs x 2 resident blocks, at most 4 rows (L1 re-use of the K-row window)
    const long long per_row = (long long)(Q / qx) * ((groups + gy - 1) / gy) * N;
    int rows = 4;
    while (rows > 1 && per_row * ((Ho + rows - 1) / rows) < 148LL * 8) rows >>= 1;
    const int rbk = (Ho + rows - 1) / rows;
    if ((long long)N * rbk > 65535) return set_arg_error("ss_dwconv2d_fwd: too many row blocks");
    dim3 grid((unsigned)(Q / qx), (unsigned)((groups + gy - 1) / gy), (unsigned)(N * rbk)), block(qx, gy);
    dwconv2d_kernel<K, S, TW><<<grid, block, 0, st>>>(x, w, bias, y, pool, H, W, C, in_ldc, Ho, Wo, out_ldc, pt, pl, out_act, rbk, rows);
    return check_launch("dwconv2d_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// SE fully connected layer: out[n][o] = act(bias[o] + in_mul * sum_c in[n][c] * w[o][c]).  WIDE = one 256-thread block per
// output (the squeeze layer: few outputs, thousands of inputs), else one warp per output (the excite layer).
// ---------------------------------------------------------------------------------------------------------------------
template <bool IN_DOUBLE, bool WIDE>
__global__ void __launch_bounds__(256)
se_fc_kernel(const void* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
             int Cin, int Cout, float in_mul, int act) {
    __shared__ float part[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = WIDE ? blockIdx.x : blockIdx.x * 8 + warp, n = blockIdx.y;
    if (o >= Cout) return;
    const float* wr = w + (size_t)o * Cin;
    float acc = 0.f;
    for (int c = WIDE ? threadIdx.x : lane; c < Cin; c += WIDE ? 256 : 32) {
        float v;
        if constexpr (IN_DOUBLE) v = (float)(reinterpret_cast<const double*>(in)[((size_t)n * Cin + c) * 2] * (double)in_mul);
        else v = reinterpret_cast<const float*>(in)[(size_t)n * Cin + c] * in_mul;
        acc = fmaf(v, __ldg(wr + c), acc);
    }
    acc = warp_sum(acc);
    if constexpr (WIDE) {
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (threadIdx.x != 0) return;
        acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += part[i];
    } else if (lane != 0) {
        return;
    }
    float v = acc + (bias ? __ldg(bias + o) : 0.f);
    if (act == SS_ACT_SIGMOID) v = sigmoid_f(v);
    else v = apply_act_sw(v, act);
    out[(size_t)n * Cout + o] = v;
}

}  // namespace ss

// x: float[N][Cin][H][W] (the reference's image layout); w: float[K*K*Cin][Cout] ordered (ky, kx, ci); y: channels-last
// float[N][Ho][Wo][Cout] with Ho = ceil(H/S), Wo = ceil(W/S) (TF "SAME": pad_top = total/2, the rest at the bottom).
extern "C" int ss_stem_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int H, int W, int Cout,
                                  int K, int S, int out_act, void* stream) {
    using namespace ss;
    SS_REQUIRE(x && w && bias && y, "ss_stem_conv2d_fwd: null pointer");
    SS_REQUIRE(N > 0 && N <= 65535 && Cin > 0 && Cin <= 4 && H > 0 && W > 0 && K >= 1 && K <= 7 && S >= 1 && S <= 4,
               "ss_stem_conv2d_fwd: shape");
    SS_REQUIRE(Cout % 4 == 0 && Cout >= 4 && Cout <= 128, "ss_stem_conv2d_fwd: Cout must be a multiple of 4, at most 128");
    const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
    SS_REQUIRE(Ho <= 65535, "ss_stem_conv2d_fwd: too many rows");
    const int th = max((Ho - 1) * S + K - H, 0), tw = max((Wo - 1) * S + K - W, 0);
    const int span = (STEM_PX - 1) * S + K;
    const size_t smem = ((size_t)K * K * Cin * Cout + (size_t)Cin * K * span) * sizeof(float);
    SS_REQUIRE(smem <= 48 * 1024, "ss_stem_conv2d_fwd: weights do not fit in shared memory");
    dim3 grid((unsigned)((Wo + STEM_PX - 1) / STEM_PX), (unsigned)Ho, (unsigned)N), block((unsigned)(Cout / 4), STEM_PX);
    stem_conv_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(x, w, bias, y, Cin, H, W, Ho, Wo, Cout, K, S, th / 2, tw / 2, out_act);
    return check_launch("stem_conv_kernel");
}

// x: channels-last float[N][H][W][C] (pixel stride in_ldc); w: float[K*K][C]; y: float[N][Ho][Wo][C] (pixel stride out_ldc);
// pool (optional): double[N][C][2], slot 0 += sum over the image of the activated output (slot 1 untouched).
extern "C" int ss_dwconv2d_fwd(const float* x, const float* w, const float* bias, float* y, double* pool, int N, int H, int W, int C,
                               int in_ldc, int out_ldc, int K, int S, int out_act, void* stream) {
    using namespace ss;
    SS_REQUIRE(x && w && bias && y, "ss_dwconv2d_fwd: null pointer");
    SS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "ss_dwconv2d_fwd: shape (C must be a multiple of 4)");
    SS_REQUIRE(in_ldc >= C && out_ldc >= C && in_ldc % 4 == 0 && out_ldc % 4 == 0, "ss_dwconv2d_fwd: ldc");
    SS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w) |
                 reinterpret_cast<uintptr_t>(bias)) & 15) == 0, "ss_dwconv2d_fwd: 16-byte alignment");
    SS_REQUIRE((K == 3 || K == 5) && (S == 1 || S == 2), "ss_dwconv2d_fwd: kernel 3 or 5, stride 1 or 2");
    const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
    const int th = max((Ho - 1) * S + K - H, 0), tw = max((Wo - 1) * S + K - W, 0);
    const int pt = th / 2, pl = tw / 2;
    SS_REQUIRE(N <= 65535, "ss_dwconv2d_fwd: batch");
    cudaStream_t st = (cudaStream_t)stream;
    // 5 x 5 on mid-sized maps (24 x 80 x 1344 channels): 8 output columns per thread re-use an input vector for up to 5 outputs
    // (measured 36 -> 29 us); on the large maps of stages 2-3 the 4-column variant's extra parallelism wins (57 vs 70 us)
    if (K == 5 && S == 1 && Wo >= 64 && Wo < 128) return launch_dw<5, 1, 8>(x, w, bias, y, pool, N, H, W, C, in_ldc, Ho, Wo, out_ldc, pt, pl, out_act, st);
    if (K == 3 && S == 1) return launch_dw<3, 1, 4>(x, w, bias, y, pool, N, H, W, C, in_ldc, Ho, Wo, out_ldc, pt, pl, out_act, st);
    if (K == 3 && S == 2) return launch_dw<3, 2, 2>(x, w, bias, y, pool, N, H, W, C, in_ldc, Ho, Wo, out_ldc, pt, pl, out_act, st);
    if (K == 5 && S == 1) return launch_dw<5, 1, 4>(x, w, bias, y, pool, N, H, W, C, in_ldc, Ho, Wo, out_ldc, pt, pl, out_act, st);
    return launch_dw<5, 2, 2>(x, w, bias, y, pool, N, H, W, C, in_ldc, Ho, Wo, out_ldc, pt, pl, out_act, st);
}

// out[n][o] = act(bias[o] + in_mul * sum_c in[n][c] * w[o][c]).  in_is_stats != 0: ``in`` is a double[N][Cin][2] block of
// channel sums (slot 0 is read) -- with in_mul = 1 / pixels that is the SE block's global average pool.
extern "C" int ss_se_fc_fwd(const void* in, int in_is_stats, const float* w, const float* bias, float* out, int N, int Cin, int Cout,
                            float in_mul, int act, void* stream) {
    using namespace ss;
    SS_REQUIRE(in && w && out, "ss_se_fc_fwd: null pointer");
    SS_REQUIRE(N > 0 && N <= 65535 && Cin > 0 && Cout > 0, "ss_se_fc_fwd: shape");
    SS_REQUIRE(Cout <= 65535 * 8, "ss_se_fc_fwd: too many outputs");
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = Cin >= 512 && Cout <= 1024;          // squeeze: a block per output; excite: a warp per output
    dim3 grid((unsigned)(wide ? Cout : (Cout + 7) / 8), (unsigned)N);
    if (in_is_stats) {
        if (wide) se_fc_kernel<true, true><<<grid, 256, 0, st>>>(in, w, bias, out, Cin, Cout, in_mul, act);
        else se_fc_kernel<true, false><<<grid, 256, 0, st>>>(in, w, bias, out, Cin, Cout, in_mul, act);
    } else {
        if (wide) se_fc_kernel<false, true><<<grid, 256, 0, st>>>(in, w, bias, out, Cin, Cout, in_mul, act);
        else se_fc_kernel<false, false><<<grid, 256, 0, st>>>(in, w, bias, out, Cin, Cout, in_mul, act);
    }
    return check_launch("se_fc_kernel");
}
