// Trilinear resize of channels-last logits (F.interpolate, mode='trilinear', align_corners=False)
// with an optional fused per-voxel argmax.  Reference: detectors/bevdepth_occupancy.py:293-294
// (upsample to gt_occ size) followed by argmax in apis/test.py:113.
// HBM-bound: reads the small logits volume (L2 resident), writes B*Do*Ho*Wo*C floats once.
#include "common.cuh"

namespace ss {

struct AxisLerp { int i0, i1; float w0, w1; };

// PyTorch area_pixel_compute_source_index (align_corners=False): src = scale*(dst+0.5)-0.5, clamped at 0
__device__ __forceinline__ AxisLerp axis_lerp(int o, int in_size, float scale) {
    float src = scale * (o + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    AxisLerp a;
    a.i0 = (int)src;
    a.i1 = a.i0 + ((a.i0 < in_size - 1) ? 1 : 0);
    a.w1 = src - (float)a.i0;
    a.w0 = 1.0f - a.w1;
    return a;
}

// one thread per output voxel, loops over channels in float4 (C % 4 == 0) or scalars
template <bool VEC>
__global__ void __launch_bounds__(256)
trilinear_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ labels, int C, int Di,
                 int Hi, int Wi, int Do, int Ho, int Wo, float sd, float sh, float sw, long long total) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int ow = (int)(o % Wo);
    const int oh = (int)((o / Wo) % Ho);
    const int od = (int)((o / ((long long)Wo * Ho)) % Do);
    const int b = (int)(o / ((long long)Wo * Ho * Do));
    const AxisLerp ad = axis_lerp(od, Di, sd), ah = axis_lerp(oh, Hi, sh), aw = axis_lerp(ow, Wi, sw);
    const float* base = x + (size_t)b * Di * Hi * Wi * C;
    const size_t o000 = (((size_t)ad.i0 * Hi + ah.i0) * Wi + aw.i0) * C, o001 = (((size_t)ad.i0 * Hi + ah.i0) * Wi + aw.i1) * C;
    const size_t o010 = (((size_t)ad.i0 * Hi + ah.i1) * Wi + aw.i0) * C, o011 = (((size_t)ad.i0 * Hi + ah.i1) * Wi + aw.i1) * C;
    const size_t o100 = (((size_t)ad.i1 * Hi + ah.i0) * Wi + aw.i0) * C, o101 = (((size_t)ad.i1 * Hi + ah.i0) * Wi + aw.i1) * C;
    const size_t o110 = (((size_t)ad.i1 * Hi + ah.i1) * Wi + aw.i0) * C, o111 = (((size_t)ad.i1 * Hi + ah.i1) * Wi + aw.i1) * C;
    float best = -INFINITY;
    int arg = 0;
    float* dst = y + (size_t)o * C;
    // same association as ATen's upsample_trilinear3d: d0*(h0*(w0*a+w1*b)+h1*(..)) + d1*(..)
    auto blend = [&](float v000, float v001, float v010, float v011, float v100, float v101, float v110, float v111) {
        return ad.w0 * (ah.w0 * (aw.w0 * v000 + aw.w1 * v001) + ah.w1 * (aw.w0 * v010 + aw.w1 * v011)) +
               ad.w1 * (ah.w0 * (aw.w0 * v100 + aw.w1 * v101) + ah.w1 * (aw.w0 * v110 + aw.w1 * v111));
    };
    if (VEC) {
        for (int c = 0; c < C; c += 4) {
            const float4 a = ldg_f4(base + o000 + c), bq = ldg_f4(base + o001 + c), cq = ldg_f4(base + o010 + c),
                         dq = ldg_f4(base + o011 + c), e = ldg_f4(base + o100 + c), f = ldg_f4(base + o101 + c),
                         gq = ldg_f4(base + o110 + c), h = ldg_f4(base + o111 + c);
            float4 r;
            r.x = blend(a.x, bq.x, cq.x, dq.x, e.x, f.x, gq.x, h.x);
            r.y = blend(a.y, bq.y, cq.y, dq.y, e.y, f.y, gq.y, h.y);
            r.z = blend(a.z, bq.z, cq.z, dq.z, e.z, f.z, gq.z, h.z);
            r.w = blend(a.w, bq.w, cq.w, dq.w, e.w, f.w, gq.w, h.w);
            st_cs_f4(dst + c, r);
            if (r.x > best) { best = r.x; arg = c; }
            if (r.y > best) { best = r.y; arg = c + 1; }
            if (r.z > best) { best = r.z; arg = c + 2; }
            if (r.w > best) { best = r.w; arg = c + 3; }
        }
    } else {
        for (int c = 0; c < C; ++c) {
            const float r = blend(__ldg(base + o000 + c), __ldg(base + o001 + c), __ldg(base + o010 + c),
                                  __ldg(base + o011 + c), __ldg(base + o100 + c), __ldg(base + o101 + c),
                                  __ldg(base + o110 + c), __ldg(base + o111 + c));
            dst[c] = r;
            if (r > best) { best = r; arg = c; }
        }
    }
    if (labels) labels[o] = (uint8_t)arg;
}

}  // namespace ss

extern "C" int ss_trilinear_fwd(const float* x, float* y, uint8_t* labels, int B, int C, int Di, int Hi, int Wi,
                                int Do, int Ho, int Wo, void* stream) {
    using namespace ss;
    SS_REQUIRE(x && y, "ss_trilinear_fwd: null pointer");
    SS_REQUIRE(B > 0 && C > 0 && C <= 256 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "ss_trilinear_fwd: shape");
    const long long total = (long long)B * Do * Ho * Wo;
    const float sd = (float)Di / (float)Do, sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (vec)
        trilinear_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(x, y, labels, C, Di, Hi, Wi, Do, Ho, Wo, sd, sh, sw, total);
    else
        trilinear_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(x, y, labels, C, Di, Hi, Wi, Do, Ho, Wo, sd, sh, sw, total);
    return check_launch("trilinear_kernel");
}
