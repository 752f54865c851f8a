// Shared device/host helpers for the stereoscene_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/stereoscene_b200.h"

namespace ss {

extern thread_local char g_last_error[256];
extern std::atomic<long long> g_launches;

int set_cuda_error(cudaError_t e, const char* where);
void count_kernel(const char* name);
int set_arg_error(const char* msg);

inline int check_launch(const char* where) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    count_kernel(where);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, where);
    return SS_OK;
}

#define SS_REQUIRE(cond, msg)                         \
    do {                                              \
        if (!(cond)) return ss::set_arg_error(msg);   \
    } while (0)

#define SS_CUDA(call)                                               \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ss::set_cuda_error(_e, #call);\
    } while (0)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// output planes that contribute to the GroupNorm sums (ss_conv3d_desc.stats_d0 / stats_d1; empty range = all planes)
struct StatsRange {
    int d0, d1;
    __host__ __device__ bool has(int d) const { return d >= d0 && d < d1; }
};
inline StatsRange stats_range_of(const ss_conv3d_desc* d) {
    if (d->stats_d1 <= d->stats_d0) return StatsRange{0, 0x7fffffff};
    return StatsRange{d->stats_d0, d->stats_d1};
}

// One pass of the error-compensated mode (SS_MATH_TF32X3) on the tcgen05 kernels.  x = hi + lo with hi = tf32(x),
// lo = tf32(x - hi), likewise for the weights (split on the host); conv(x, w) ~= lo_x*hi_w + hi_x*lo_w + hi_x*hi_w, each
// term one ordinary TF32 launch whose products are exact in fp32.  a_lo selects which part of the A operand the fix-up
// warps write; accumulate makes the epilogue add the output already in memory (bias, activation, statistics and a
// fused join belong to the last pass only).
struct ConvPass {
    int a_lo;
    int accumulate;
    int f16 = 0;              // SS_MATH_F16X3 / SS_MATH_F16: single launch on fp16 operands (see split_f16x4)
    int f16_n = 6;            //   MMAs per 32-channel chunk: 6 = hi*hi + lo*hi + hi*lo (F16X3), 2 = hi*hi only (F16)
    float acc_scale = 1.0f;   //   the accumulator is multiplied by this power of two (the weights were pre-scaled by its inverse)
};
// SS_MATH_F16X3: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 significand bits); one 128-byte shared-memory row
// that held 32 fp32 channels holds [hi(32 ch) | lo(32 ch)] as 64 fp16 K elements, the weight rows hold [hi(w) | lo(w)] the
// same way, and every 32-channel chunk issues six kind::f16 MMAs (K = 16): hi*hi, lo*hi, hi*lo -- 1.5x the MMA count of a
// plain TF32 pass (four K = 8 MMAs), in ONE launch with the accumulator staying in TMEM.
__device__ __forceinline__ void split_f16x4(float4 v, uint2& hi, uint2& lo) {
    v.x = fminf(fmaxf(v.x, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
    v.z = fminf(fmaxf(v.z, -65504.f), 65504.f); v.w = fminf(fmaxf(v.w, -65504.f), 65504.f);
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}
// two floats -> packed fp16 pair (lo half = a), round-to-nearest, saturating to +-65504 (one F2FP instruction)
__device__ __forceinline__ uint32_t f2h2_sat(float a, float b) {
    uint32_t r;
    asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint2 hi_f16x4(const float4& v) { return make_uint2(f2h2_sat(v.x, v.y), f2h2_sat(v.z, v.w)); }

// cute::UMMA::InstrDescriptor for kind::f16 with fp16 operands: c_format F32 (1) at [4,6), a/b format F16 (0), K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// descriptor-word offsets (16-byte units inside the 128-byte row) of the six MMAs of one chunk: A part, B part
__device__ constexpr int kF16A[6] = {0, 2, 4, 6, 0, 2};
__device__ constexpr int kF16B[6] = {0, 2, 0, 2, 4, 6};
template <bool LO>
__device__ __forceinline__ uint32_t f2tf32_part(float x) {
    const uint32_t h = f2tf32(x);
    if constexpr (LO) return f2tf32(x - __uint_as_float(h));
    else return h;
}
// The fix-up loops are unswitched on the (kernel-uniform) a_lo flag: `SS_UNSWITCH_LO(p.a_lo, body)` runs `body(tag)` with
// decltype(tag)::value == a_lo as a compile-time constant, so the plain-TF32 path carries no extra instructions.
struct LoTrue { static constexpr bool value = true; };
struct LoFalse { static constexpr bool value = false; };
#define SS_UNSWITCH_LO(flag, body) do { if (flag) body(ss::LoTrue{}); else body(ss::LoFalse{}); } while (0)

// D(16x8,f32) += A(16x8,tf32,row) * B(8x8,tf32,col)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// x * sigmoid(x) on the fast-math SFU path (ex2.approx + rcp.approx, ~4e-7 relative): an IEEE division here costs a ~100-instruction
// slow path per element and made the Swish epilogues of the image encoder 5x slower than the GEMMs in front of them
__device__ __forceinline__ float swish_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == SS_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == SS_ACT_GELU) return gelu_erf(v);
    return v;
}
// + Swish: only the kernels the image encoder reaches call this one (the extra branch cost the elementwise join kernel of the
// volumetric path 19 % when it sat in apply_act)
__device__ __forceinline__ float apply_act_sw(float v, int act) {
    if (act == SS_ACT_SWISH) return swish_f(v);
    return apply_act(v, act);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (evict-first) 128-bit store for write-once outputs
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- warp-uniform issue helpers for the tcgen05 / TMA kernels ----------------------------------------------
// tcgen05.mma, tcgen05.commit and cp.async.bulk.tensor are issued by ONE thread but execute on the uniform
// datapath (UTCHMMA / UTCBAR / UTMALDG take uniform registers).  If the issuing code sits inside an
// `if (lane == 0)` region ptxas cannot keep the descriptors in uniform registers: every instruction is wrapped
// in an ELECT / R2UR.BROADCAST loop (~90-250 clk per MMA measured on B200), which makes layers with N <= 128
// issue-bound.  So the producer and MMA warps run their loops warp-uniformly (all 32 lanes, role index taken
// through __shfl_sync so the branch is provably uniform) and only the instruction is predicated on elect.sync.
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) {      // K-major SWIZZLE_128B, version 1
    return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, descriptors given as varying low word + compile-time high word
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void umma_ss_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(A_HI), "n"(B_HI) : "memory");
}
// same with A read from tensor memory (TS form)
template <uint32_t B_HI>
__device__ __forceinline__ void umma_ts_tf32(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(B_HI) : "memory");
}
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void umma_ss_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(A_HI), "n"(B_HI) : "memory");
}
template <uint32_t B_HI>
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(B_HI) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_5d_elect(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t"
        "}\n" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_2d_elect(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t"
        "}\n" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

}  // namespace ss

