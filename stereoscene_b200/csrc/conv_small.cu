// Single-output-channel 3-D convolution (Cout == 1, Cin % 32 == 0, stride 1): the two 32 -> 1
// k3 layers that turn an aggregated cost volume into depth logits (classif3_2,
// ViewTransformerLSSVoxel.py:186-187, and the MIE redir2, :241).  With one output channel there is
// no GEMM to speak of (864 MACs per voxel): it is a bandwidth problem, so it runs on the FMA pipe.
// One warp produces 8 consecutive voxels along W: lanes = input channels, the 10-voxel input window
// of every (kd,kh) row is loaded once (coalesced 128-byte rows, pending affine + ReLU applied on load,
// zero padding kept zero) and reused by the three kw taps; a shuffle tree reduces over channels.
#include "common.cuh"

namespace ss {

constexpr int C1_NV = 8;         // output voxels per warp
constexpr int C1_WARPS = 8;

struct C1Params {
    int B, D, H, W, Cin, in_ldc, out_ldc, in_act, out_act;
    int kd, kh, kw, pd, ph, pw;
    const float* x; const float* in_scale; const float* in_shift; const float* w; const float* bias; float* y;
};

// w: [taps][Cin] (the kernel's packed layout with cout_packed = 8 -> column 0 of [taps][Cin][8])
__global__ void __launch_bounds__(C1_WARPS * 32)
conv_cout1_kernel(const C1Params p, int w_ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int segs_w = (p.W + C1_NV - 1) / C1_NV;
    const long long seg = (long long)blockIdx.x * C1_WARPS + warp;
    const long long nseg = (long long)p.B * p.D * p.H * segs_w;
    if (seg >= nseg) return;
    const int sw = (int)(seg % segs_w);
    const int h = (int)((seg / segs_w) % p.H);
    const int d = (int)((seg / ((long long)segs_w * p.H)) % p.D);
    const int b = (int)(seg / ((long long)segs_w * p.H * p.D));
    const int w0 = sw * C1_NV;
    const bool has_aff = p.in_scale != nullptr;
    const bool relu = p.in_act == SS_ACT_RELU;
    float acc[C1_NV];
#pragma unroll
    for (int i = 0; i < C1_NV; ++i) acc[i] = 0.f;
    for (int c0 = 0; c0 < p.Cin; c0 += 32) {
        const int c = c0 + lane;
        float sc = 1.f, sh = 0.f;
        if (has_aff) { sc = __ldg(p.in_scale + (size_t)b * p.Cin + c); sh = __ldg(p.in_shift + (size_t)b * p.Cin + c); }
        for (int a = 0; a < p.kd; ++a) {
            const int id = d + a - p.pd;
            if ((unsigned)id >= (unsigned)p.D) continue;
            for (int e = 0; e < p.kh; ++e) {
                const int ih = h + e - p.ph;
                if ((unsigned)ih >= (unsigned)p.H) continue;
                const float* row = p.x + (((size_t)(b * p.D + id) * p.H + ih) * p.W) * p.in_ldc + c;
                float xv[C1_NV + 2];
#pragma unroll
                for (int i = 0; i < C1_NV + 2; ++i) {
                    const int iw = w0 + i - p.pw;
                    float v = 0.f;
                    if (i < C1_NV + p.kw - 1 && (unsigned)iw < (unsigned)p.W) {
                        v = __ldg(row + (size_t)iw * p.in_ldc);
                        if (has_aff) v = fmaf(v, sc, sh);
                        if (relu) v = fmaxf(v, 0.f);
                    }
                    xv[i] = v;
                }
                const float* wrow = p.w + (size_t)((a * p.kh + e) * p.kw) * p.Cin * w_ld + (size_t)c * w_ld;
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (f < p.kw) {
                        const float wt = __ldg(wrow + (size_t)f * p.Cin * w_ld);
#pragma unroll
                        for (int i = 0; i < C1_NV; ++i) acc[i] = fmaf(xv[i + f], wt, acc[i]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < C1_NV; ++i) acc[i] = warp_sum(acc[i]);
    if (lane < C1_NV && w0 + lane < p.W) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < C1_NV; ++i) if (lane == i) v = acc[i];
        if (p.bias) v += __ldg(p.bias);
        v = apply_act(v, p.out_act);
        p.y[((((size_t)b * p.D + d) * p.H + h) * p.W + w0 + lane) * p.out_ldc] = v;
    }
}

// used by ss_conv3d_fwd for eligible layers; returns 1 if the layer was handled here
int try_conv_cout1(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                   const float* w_packed, const float* bias, float* y, double* stats, cudaStream_t st, int* rc) {
    if (d->Cout != 1 || d->transposed || stats || d->Cin % 32 != 0 || d->math != SS_MATH_TF32 || d->out_act == SS_ACT_SWISH) return 0;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1 || d->kw > 3) return 0;
    if (d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win) return 0;
    C1Params p;
    p.B = d->B; p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.Cin = d->Cin; p.in_ldc = d->in_ldc; p.out_ldc = d->out_ldc;
    p.in_act = d->in_act; p.out_act = d->out_act; p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
    p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w_packed; p.bias = bias; p.y = y;
    const long long nseg = (long long)p.B * p.D * p.H * ((p.W + C1_NV - 1) / C1_NV);
    conv_cout1_kernel<<<(unsigned)((nseg + C1_WARPS - 1) / C1_WARPS), C1_WARPS * 32, 0, st>>>(p, d->cout_packed);
    *rc = check_launch("conv_cout1_kernel");
    return 1;
}


// ------------------------------------------------------------------------------------------------
// Few-input-channel 3x3x3 convolution (Cin <= 2 -> Cout <= 32, stride 1, pad 1): the MIE redir1 layer
// (Conv3d 2 -> 32 + bias + ReLU on the two BRI outputs, ViewTransformerLSSVoxel.py:239, 259).  K = 27*Cin = 54
// is far too short for the generic implicit-GEMM tiles (one K step of 32 per tap, mostly padding): here a
// CTA stages the (3 x 10 x 34 x Cin) halo of an 8 x 32 voxel tile in shared memory once (TF32-rounded) and
// every warp runs mma.sync m16n8k8 over K = 56 with the whole weight matrix held in registers as B
// fragments; A fragments are shared-memory reads at per-lane precomputed tap offsets.  The layer is then
// bound by writing its 110 MB output.
// ------------------------------------------------------------------------------------------------
constexpr int CS_TH = 8, CS_TW = 32, CS_THREADS = 256;

struct CsParams {
    int B, D, H, W, Cout, CoutP, in_ldc, out_ldc, out_act, nTH, nTW;
    const float* x; const float* w; const float* bias; float* y; double* stats;
};

// PRECISE: error-compensated split TF32 (SS_MATH_3XTF32 / the Cin <= 2 layer of an SS_MATH_TF32X3 forward): the halo tile is
// kept in fp32, A and B fragments are split hi/lo in registers and every k-step issues lo*hi + hi*lo + hi*hi.
// (its B fragments -- hi and lo halves of the whole weight matrix, 112 registers per thread -- live in shared memory instead, one
// conflict-free 64-bit load per MMA: the kernel keeps two CTAs per SM like the TF32 variant; with the fragments in registers it
// ran one CTA per SM and took 0.20 ms against 0.10 ms)
template <int CIN, bool PRECISE>
__global__ void __launch_bounds__(CS_THREADS, 2)
conv_cin_small_kernel(const CsParams p) {
    constexpr int K = 27 * CIN, KSTEPS = (K + 7) / 8;
    constexpr int HH = CS_TH + 2, HW = CS_TW + 2;
    constexpr int TILE = 3 * HH * HW * CIN;
    __shared__ float xs[TILE];
    __shared__ float sst[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int bi = blockIdx.x;
    const int tw = bi % p.nTW; bi /= p.nTW;
    const int th = bi % p.nTH; bi /= p.nTH;
    const int d = bi % p.D;
    const int b = bi / p.D;
    const int h0 = th * CS_TH, w0 = tw * CS_TW;

    if (tid < 64) sst[tid] = 0.f;
    for (int idx = tid; idx < TILE; idx += CS_THREADS) {
        const int c = idx % CIN, ww = (idx / CIN) % HW, hh = (idx / (CIN * HW)) % HH, kd = idx / (CIN * HW * HH);
        const int gd = d + kd - 1, gh = h0 + hh - 1, gw = w0 + ww - 1;
        float v = 0.f;
        if ((unsigned)gd < (unsigned)p.D && (unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W)
            v = __ldg(p.x + ((((size_t)b * p.D + gd) * p.H + gh) * p.W + gw) * p.in_ldc + c);
        xs[idx] = PRECISE ? v : __uint_as_float(f2tf32(v));
    }
    // B fragments of the whole [K x 32] weight matrix and the tap offsets of this lane's two K columns
    uint32_t breg[PRECISE ? 1 : KSTEPS][4][2];
    __shared__ uint2 sbh[PRECISE ? KSTEPS * 4 * 32 : 1], sbl[PRECISE ? KSTEPS * 4 * 32 : 1];     // [ks][nt][lane]: (e = 0, e = 1)
    int koff[KSTEPS][2];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = ks * 8 + t + 4 * e;
            int off = 0;
            if (k < K) {
                const int tap = k / CIN, c = k % CIN;
                off = (((tap / 9) * HH + (tap / 3) % 3) * HW + tap % 3) * CIN + c;
            }
            koff[ks][e] = off;
            if constexpr (!PRECISE) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int n = nt * 8 + g;
                    const float wv = (k < K) ? __ldg(p.w + (size_t)k * p.CoutP + n) : 0.f;
                    breg[ks][nt][e] = f2tf32(wv);
                }
            }
        }
    if constexpr (PRECISE) {        // every warp would build the same fragments: warp w fills the K steps w, w + 8, ...
        for (int ks = warp; ks < KSTEPS; ks += CS_THREADS / 32)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                uint32_t hi[2], lo[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = ks * 8 + t + 4 * e, n = nt * 8 + g;
                    const float wv = (k < K) ? __ldg(p.w + (size_t)k * p.CoutP + n) : 0.f;
                    hi[e] = f2tf32(wv);
                    lo[e] = f2tf32(wv - __uint_as_float(hi[e]));
                }
                sbh[(ks * 4 + nt) * 32 + lane] = make_uint2(hi[0], hi[1]);
                sbl[(ks * 4 + nt) * 32 + lane] = make_uint2(lo[0], lo[1]);
            }
    }
    float bias_r[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = nt * 8 + 2 * t + e;
            bias_r[nt][e] = (p.bias && n < p.Cout) ? __ldg(p.bias + n) : 0.f;
        }
    __syncthreads();

    const int hl = warp;                                   // one tile row per warp, two 16-voxel M tiles along w
    const int oh = h0 + hl;
    const bool pair_ok = ((p.out_ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 7) == 0);
    float ss[4][2], sq[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { ss[nt][0] = ss[nt][1] = sq[nt][0] = sq[nt][1] = 0.f; }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int rb0 = (hl * HW + j * 16 + g) * CIN, rb1 = rb0 + 8 * CIN;
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            if constexpr (PRECISE) {
                const float af[4] = {xs[rb0 + koff[ks][0]], xs[rb1 + koff[ks][0]], xs[rb0 + koff[ks][1]], xs[rb1 + koff[ks][1]]};
                uint32_t a[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = f2tf32(af[i]); al[i] = f2tf32(af[i] - __uint_as_float(a[i])); }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const uint2 h2 = sbh[(ks * 4 + nt) * 32 + lane], l2 = sbl[(ks * 4 + nt) * 32 + lane];
                    const uint32_t bh[2] = {h2.x, h2.y}, bl[2] = {l2.x, l2.y};
                    mma_tf32(acc[nt], al, bh);
                    mma_tf32(acc[nt], a, bl);
                    mma_tf32(acc[nt], a, bh);
                }
            } else {
                const uint32_t a[4] = {__float_as_uint(xs[rb0 + koff[ks][0]]), __float_as_uint(xs[rb1 + koff[ks][0]]),
                                       __float_as_uint(xs[rb0 + koff[ks][1]]), __float_as_uint(xs[rb1 + koff[ks][1]])};
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], a, breg[ks][nt]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ow = w0 + j * 16 + g + 8 * r;
            if (oh < p.H && ow < p.W) {
                float* dst = p.y + ((((size_t)b * p.D + d) * p.H + oh) * p.W + ow) * p.out_ldc;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int n = nt * 8 + 2 * t;
                    const float v0 = apply_act(acc[nt][2 * r + 0] + bias_r[nt][0], p.out_act);
                    const float v1 = apply_act(acc[nt][2 * r + 1] + bias_r[nt][1], p.out_act);
                    if (pair_ok && n + 1 < p.Cout) *reinterpret_cast<float2*>(dst + n) = make_float2(v0, v1);
                    else {
                        if (n < p.Cout) dst[n] = v0;
                        if (n + 1 < p.Cout) dst[n + 1] = v1;
                    }
                    ss[nt][0] += v0; sq[nt][0] = fmaf(v0, v0, sq[nt][0]);
                    ss[nt][1] += v1; sq[nt][1] = fmaf(v1, v1, sq[nt][1]);
                }
            }
        }
    }
    if (p.stats) {                                          // reduce over the 8 row groups of the warp, then over warps
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float s = ss[nt][e], q = sq[nt][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
                if (g == 0) { atomicAdd(&sst[2 * (nt * 8 + 2 * t + e)], s); atomicAdd(&sst[2 * (nt * 8 + 2 * t + e) + 1], q); }
            }
        __syncthreads();
        if (tid < 32 && tid < p.Cout) {
            atomicAdd(p.stats + ((size_t)b * p.Cout + tid) * 2 + 0, (double)sst[2 * tid]);
            atomicAdd(p.stats + ((size_t)b * p.Cout + tid) * 2 + 1, (double)sst[2 * tid + 1]);
        }
    }
}

// used by ss_conv3d_fwd for eligible layers; returns 1 if the layer was handled here
int try_conv_cin_small(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                       const float* w_packed, const float* bias, float* y, double* stats, cudaStream_t st, int* rc) {
    if (d->transposed || d->Cin > 2 || d->cout_packed != 32 || in_scale || d->in_act != SS_ACT_NONE || d->out_act == SS_ACT_SWISH) return 0;
    if (d->math != SS_MATH_TF32 && d->math != SS_MATH_3XTF32) return 0;
    const bool precise = d->math == SS_MATH_3XTF32;
    if (d->kd != 3 || d->kh != 3 || d->kw != 3 || d->sd != 1 || d->sh != 1 || d->sw != 1 || d->pd != 1 || d->ph != 1 || d->pw != 1) return 0;
    if (d->dd != 1 || d->dh != 1 || d->dw != 1 || d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win) return 0;
    CsParams p;
    p.B = d->B; p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.Cout = d->Cout; p.CoutP = d->cout_packed; p.in_ldc = d->in_ldc;
    p.out_ldc = d->out_ldc; p.out_act = d->out_act; p.nTH = (p.H + CS_TH - 1) / CS_TH; p.nTW = (p.W + CS_TW - 1) / CS_TW;
    p.x = x; p.w = w_packed; p.bias = bias; p.y = y; p.stats = stats;
    const long long blocks = (long long)p.B * p.D * p.nTH * p.nTW;
    if (blocks > 0x7fffffffLL) return 0;
    if (d->Cin == 1) {
        if (precise) conv_cin_small_kernel<1, true><<<(unsigned)blocks, CS_THREADS, 0, st>>>(p);
        else conv_cin_small_kernel<1, false><<<(unsigned)blocks, CS_THREADS, 0, st>>>(p);
    } else {
        if (precise) conv_cin_small_kernel<2, true><<<(unsigned)blocks, CS_THREADS, 0, st>>>(p);
        else conv_cin_small_kernel<2, false><<<(unsigned)blocks, CS_THREADS, 0, st>>>(p);
    }
    *rc = check_launch("conv_cin_small_kernel");
    return 1;
}

}  // namespace ss

