// Group-wise correlation cost volume fused with the disparity -> depth-bin resampling.
//
// Reference (projects/mmdet3d_plugin/occupancy/image2bev/ViewTransformerLSSVoxel.py):
//   build_gwc_volume / groupwise_correlation (:97-114) builds vol[b,g,i,h,w] =
//   mean_c ref[b,g*cpg+c,h,w]*tgt[b,g*cpg+c,h,w-i] (0 for w<i) with a 112-iteration Python loop,
//   warp (:128-156) then resamples the disparity axis i to depth bins k with grid_sample
//   (1-D linear interpolation at p_k = calib/(4(k+1)), zero padding).
// Here out[b,k,h,w,g] = w0[b,k]*vol[..i0..] + w1[b,k]*vol[..i0+1..] is produced directly from the
// two feature rows, which are staged once per CTA in shared memory; the disparity volume is never
// written.  HBM-bound: reads 2*C*H*W floats, writes G*K*H*W floats (coalesced 128 B per voxel).
#include "common.cuh"

namespace ss {

constexpr int GW_THREADS = 256;
constexpr int GW_KCHUNK = 8;

// fea: [2B][H][W][C]; CTA = (k-chunk, h, b)
template <int CPG>
__global__ void __launch_bounds__(GW_THREADS)
gwc_warp_kernel(const float* __restrict__ fea, const int32_t* __restrict__ i0p, const float* __restrict__ w0p,
                const float* __restrict__ w1p, float* __restrict__ out, int B, int C, int G, int H, int W, int K,
                int maxdisp) {
    extern __shared__ __align__(16) float sm[];
    float* sref = sm;                 // [W][C]
    float* stgt = sm + (size_t)W * C; // [W][C]
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * GW_KCHUNK;
    const int kn = min(GW_KCHUNK, K - k0);
    const float* rrow = fea + ((size_t)b * H + h) * W * C;
    const float* trow = fea + ((size_t)(B + b) * H + h) * W * C;
    const int n4 = W * C / 4;
    for (int i = threadIdx.x; i < n4; i += GW_THREADS) {
        reinterpret_cast<float4*>(sref)[i] = ldg_f4(rrow + 4 * i);
        reinterpret_cast<float4*>(stgt)[i] = ldg_f4(trow + 4 * i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr float inv_cpg = 1.0f / CPG;
    // each warp takes whole depth bins: the two disparity taps are loop constants, lanes = groups
    for (int kk = warp; kk < kn; kk += GW_THREADS / 32) {
        const int k = k0 + kk;
        const int i0 = __ldg(i0p + b * K + k);
        const float a0 = __ldg(w0p + b * K + k), a1 = __ldg(w1p + b * K + k);
        const int i1 = i0 + 1;
        const bool t0 = (i0 >= 0 && i0 < maxdisp), t1 = (i1 >= 0 && i1 < maxdisp);
        float* orow = out + ((((size_t)b * K + k) * H + h) * W) * G;
        for (int g = lane; g < G; g += 32) {
            const float* rg = sref + g * CPG;
            const float* tg = stgt + g * CPG;
#pragma unroll 4
            for (int w = 0; w < W; ++w) {
                float r[CPG];
#pragma unroll
                for (int c = 0; c < CPG; ++c) r[c] = rg[w * C + c];
                float acc = 0.f;
                bool any = false;
                if (t0 && w >= i0) {
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < CPG; ++c) s += r[c] * tg[(w - i0) * C + c];
                    acc = (s * inv_cpg) * a0;
                    any = true;
                }
                if (t1 && w >= i1) {
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < CPG; ++c) s += r[c] * tg[(w - i1) * C + c];
                    const float tv = (s * inv_cpg) * a1;
                    acc = any ? acc + tv : tv;
                }
                __stcs(orow + (size_t)w * G + g, acc);
            }
        }
    }
}

}  // namespace ss

extern "C" int ss_gwc_warp_fwd(const float* fea, const int32_t* i0, const float* w0, const float* w1, float* out,
                               int B, int C, int G, int H, int W, int K, int maxdisp, void* stream) {
    using namespace ss;
    SS_REQUIRE(fea && i0 && w0 && w1 && out, "ss_gwc_warp_fwd: null pointer");
    SS_REQUIRE(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0 && K > 0 && G > 0 && C % G == 0 && C % 4 == 0,
               "ss_gwc_warp_fwd: shape");
    const int cpg = C / G;
    const size_t smem = 2 * (size_t)W * C * sizeof(float);
    SS_REQUIRE(smem <= 200 * 1024, "ss_gwc_warp_fwd: feature row too large for shared memory");
    dim3 grid((K + GW_KCHUNK - 1) / GW_KCHUNK, H, B);
    cudaStream_t st = (cudaStream_t)stream;
#define SS_GW_LAUNCH(CPG)                                                                                   \
    do {                                                                                                    \
        static thread_local size_t configured = 0;                                                          \
        if (smem > configured) {                                                                            \
            SS_CUDA(cudaFuncSetAttribute(gwc_warp_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            configured = smem;                                                                              \
        }                                                                                                   \
        gwc_warp_kernel<CPG><<<grid, GW_THREADS, smem, st>>>(fea, i0, w0, w1, out, B, C, G, H, W, K, maxdisp); \
    } while (0)
    switch (cpg) {
        case 1: SS_GW_LAUNCH(1); break;
        case 2: SS_GW_LAUNCH(2); break;
        case 4: SS_GW_LAUNCH(4); break;
        case 8: SS_GW_LAUNCH(8); break;
        default: return set_arg_error("ss_gwc_warp_fwd: channels per group must be 1, 2, 4 or 8");
    }
#undef SS_GW_LAUNCH
    return check_launch("gwc_warp_kernel");
}
