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
    if (d->Cout != 1 || d->transposed || stats || d->Cin % 32 != 0 || d->math != SS_MATH_TF32) return 0;
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

}  // namespace ss

