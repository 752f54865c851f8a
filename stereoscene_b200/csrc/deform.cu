// Deformable-convolution sampling stage (mmcv DeformConv2dPack / torchvision deform_conv2d
// semantics, deform_groups = 1) for the DCN layer of DepthNet
// (reference: image2bev/ViewTransformerLSSBEVDepth.py:490-498).
//
//   S[b, y, x, g, t, cc] = bilinear(in[b, :, :, g*Cg + cc], y*stride - pad + i*dil + dy_t(y,x),
//                                                           x*stride - pad + j*dil + dx_t(y,x))
// with t = i*kw + j, offsets stored NCHW as [b][2t] = dy, [b][2t+1] = dx, zero outside the image and
// per-corner masking exactly as torchvision's bilinear_interpolate.  The grouped GEMM over
// K = kh*kw*Cg that follows runs on the tcgen05 conv kernel (one 1x1 "conv" per group on S).
// One warp per (pixel, tap): lanes sweep the channels with 128-bit loads/stores.
#include "common.cuh"

namespace ss {

__global__ void __launch_bounds__(256)
deform_sample_kernel(const float* __restrict__ x, const float* __restrict__ off, float* __restrict__ out, int B, int H,
                     int W, int C, int G, int kh, int kw, int stride, int pad, int dil, int Ho, int Wo) {
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int T = kh * kw;
    const long long total = (long long)B * Ho * Wo * T;
    if (wid >= total) return;
    const int t = (int)(wid % T);
    const long long pix = wid / T;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    const int i = t / kw, j = t % kw;
    const size_t plane = (size_t)Ho * Wo;
    const float dy = __ldg(off + ((size_t)b * 2 * T + 2 * t) * plane + (size_t)oy * Wo + ox);
    const float dx = __ldg(off + ((size_t)b * 2 * T + 2 * t + 1) * plane + (size_t)oy * Wo + ox);
    const float py = (float)(oy * stride - pad + i * dil) + dy;
    const float px = (float)(ox * stride - pad + j * dil) + dx;
    const int Cg = C / G;
    float* dst = out + ((size_t)pix * G * T) * Cg;          // [pix][g][t][cc]
    const bool inside = !(py <= -1.f || (float)H <= py || px <= -1.f || (float)W <= px);
    const int h_low = (int)floorf(py), w_low = (int)floorf(px);
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = py - (float)h_low, lw = px - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
    const bool m1 = inside && h_low >= 0 && w_low >= 0, m2 = inside && h_low >= 0 && w_high <= W - 1;
    const bool m3 = inside && h_high <= H - 1 && w_low >= 0, m4 = inside && h_high <= H - 1 && w_high <= W - 1;
    const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    const float* base = x + (size_t)b * H * W * C;
    const float* p1 = base + ((size_t)(m1 ? h_low : 0) * W + (m1 ? w_low : 0)) * C;
    const float* p2 = base + ((size_t)(m2 ? h_low : 0) * W + (m2 ? w_high : 0)) * C;
    const float* p3 = base + ((size_t)(m3 ? h_high : 0) * W + (m3 ? w_low : 0)) * C;
    const float* p4 = base + ((size_t)(m4 ? h_high : 0) * W + (m4 ? w_high : 0)) * C;
    for (int c = lane * 4; c < C; c += 128) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m1) { const float4 v = ldg_f4(p1 + c); r.x += w1 * v.x; r.y += w1 * v.y; r.z += w1 * v.z; r.w += w1 * v.w; }
        if (m2) { const float4 v = ldg_f4(p2 + c); r.x += w2 * v.x; r.y += w2 * v.y; r.z += w2 * v.z; r.w += w2 * v.w; }
        if (m3) { const float4 v = ldg_f4(p3 + c); r.x += w3 * v.x; r.y += w3 * v.y; r.z += w3 * v.z; r.w += w3 * v.w; }
        if (m4) { const float4 v = ldg_f4(p4 + c); r.x += w4 * v.x; r.y += w4 * v.y; r.z += w4 * v.z; r.w += w4 * v.w; }
        const int g = c / Cg, cc = c % Cg;
        *reinterpret_cast<float4*>(dst + ((size_t)g * T + t) * Cg + cc) = r;
    }
}

}  // namespace ss

extern "C" int ss_deform_sample_fwd(const float* x, const float* offsets, float* out, int B, int H, int W, int C,
                                    int groups, int kh, int kw, int stride, int pad, int dil, void* stream) {
    using namespace ss;
    SS_REQUIRE(x && offsets && out, "ss_deform_sample_fwd: null pointer");
    SS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && groups > 0 && C % groups == 0 && (C / groups) % 4 == 0,
               "ss_deform_sample_fwd: channels per group must be a multiple of 4");
    SS_REQUIRE(kh > 0 && kw > 0 && stride > 0 && dil > 0, "ss_deform_sample_fwd: kernel geometry");
    const int Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    const long long total = (long long)B * Ho * Wo * kh * kw;
    const int warps = 8;
    deform_sample_kernel<<<(unsigned)((total + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        x, offsets, out, B, H, W, C, groups, kh, kw, stride, pad, dil, Ho, Wo);
    return check_launch("deform_sample_kernel");
}
