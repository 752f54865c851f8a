// "Marching" tcgen05 convolution for the HBM-bound 32-channel 3x3x3 layers of the frustum volume
// (dres0/dres1/classif3_1, CA3D, the 32->1 logit convs: ViewTransformerLSSVoxel.py:167-187, 239-241,
// attention.py:93-112).  Algorithmic intensity of these layers is ~216 FLOP/B, i.e. they are bound by
// reading the input once and writing the output once -- IF the 27-fold stencil reuse happens on chip.
//
// Design (persistent, one CTA per SM):
//   * weights of all 27 taps stay RESIDENT in shared memory (27 x 32 x 128 B = 108 KB, K-major
//     SWIZZLE_128B, loaded once per CTA by TMA);
//   * a CTA owns a column of 16(h) x 8(w) output voxels and MARCHES along d.  Input planes (one TMA 5-D
//     box of 18 x 10 halo voxels x 32 channels = 22.5 KB, out-of-bounds zero fill = conv padding) live
//     in a 4-slot ring; every plane is loaded once and used by the 3 output planes that touch it;
//   * the A operand of tap (kd,kh,kw) is NOT re-staged: it is the same plane slot addressed through a
//     UMMA shared-memory descriptor whose start is shifted by (kh*10 + kw) rows and whose
//     stride-byte-offset is the halo line pitch (10 rows = 1280 B).  The 128-byte swizzle is a function
//     of the absolute shared-memory address (verified on B200 by tools/probes/umma_shift_probe.cu), so
//     row-shifted starts and a non-1024 SBO are legal;
//   * the three kd taps of one (kh,kw) shift are ONE MMA: an input plane p feeds the output planes p-1, p, p+1
//     (kd = 2, 1, 0), whose accumulators sit side by side in a 4-slot TMEM ring, and the resident weights
//     are stored [kh,kw][kd=2|kd=1|kd=0][cout] so that B is one 96-row operand.  N = 96 instead of 32 cuts
//     the shared-memory operand traffic per FLOP 2.1x (A, the 4 KB plane window, is the dominant read: at
//     N = 32 the layer is bound by the 128 B/clk tensor-core shared-memory port, not by the MMA rate) and every
//     plane is consumed by 36 MMAs right after it lands, so the 4-slot plane ring is 3 planes of prefetch;
//   * the epilogue of output plane d (tcgen05.ld, bias, activation, GroupNorm sums, stores) overlaps the
//     MMAs of the following planes (ring slot d % 4);
//   * a pending affine / ReLU of the producer layer is applied once per plane, in place, by 4 fix-up
//     warps (padding stays zero) -- amortised over the 27 taps x 3 planes that read it.
// L2->SM traffic per output tile: one 22.5 KB plane (vs 27 x 20 KB for the per-tap box kernel).
#include <cuda.h>
#include "common.cuh"

namespace ss {

constexpr int MR_TH = 16, MR_TW = 8;                 // output tile (h, w); M = 128 rows
constexpr int MR_NP = 4;                             // plane ring slots
constexpr int MR_BN = 32;
constexpr int MR_ACC = 4;                            // TMEM accumulator ring (32 columns each)
constexpr int MR_THREADS = 8 * 32 + 64;              // 4 epilogue warps, 4 fix-up warps, producer warp, MMA warp

// KS = kernel extent per axis (3: 3x3x3 pad 1; 1: 1x1x1, e.g. the hourglass redir1 layers)
template <int KS>
struct MarchCfg {
    static constexpr int PAD = (KS - 1) / 2;
    static constexpr int HH = MR_TH + KS - 1, HW = MR_TW + KS - 1;     // halo plane
    static constexpr int PLANE_ROWS = HH * HW;                        // 180 / 128
    static constexpr int PLANE_BYTES = (PLANE_ROWS * 128 + 1023) / 1024 * 1024;
    static constexpr int TAPS = KS * KS * KS;
    static constexpr int W_BYTES = TAPS * MR_BN * 128;                // 110592 / 4096
    static constexpr int W_BOX_ROWS = 32;                             // weight rows per TMA box (one tap)
    static constexpr int W_BOXES = TAPS * MR_BN / W_BOX_ROWS;
};

struct MarchParams {
    int B, D, H, W, Cout, out_ldc, in_act, out_act;
    int nTH, nTW;
    long long total_tiles;                           // B * nTH * nTW * D
    int a_lo, accumulate;                            // ConvPass (common.cuh)
    StatsRange sr;                                   // output planes that contribute to stats
    float acc_scale;                                 // F16 variant: accumulator scale (power of two)
    int f16_n;                                       // F16 variant: MMAs per chunk (6 = compensated, 2 = fp16 single pass)
    const float* in_scale;
    const float* in_shift;
    const float* bias;
    float* y;
    double* stats;
};

// ---- PTX wrappers (same as conv3d_tc.cu; kept local to this translation unit) ------------------
__device__ __forceinline__ uint32_t m_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void m_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void m_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MWAIT_DONE;\n\t"
        "bra MWAIT_LOOP;\n\t"
        "MWAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void m_tma_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void m_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// The MMA warp runs its loops warp-uniformly (all 32 lanes compute the same descriptors, so ptxas keeps them in
// uniform registers) and only the instruction itself is predicated on the elected lane.  Issuing from inside an
// `if (lane == 0)` region instead makes ptxas wrap every UTCHMMA in an ELECT / R2UR.BROADCAST loop (~90 clk per MMA),
// which bounds small-N layers by instruction issue.
// warp-uniform variants for the producer warp: every lane runs the loop, one elected lane issues
__device__ __forceinline__ void m_mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m_tma_5d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t"
        "}\n" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void m_tma_2d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t"
        "}\n" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void m_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t leader) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void m_umma_commit(uint32_t bar, uint32_t leader) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(bar), "r"(leader) : "memory");
}
// same, with the descriptors given as their (varying) low words; the high words are compile-time constants
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void m_umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
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
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void m_umma_lo_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
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
__device__ __forceinline__ void m_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B descriptor with an explicit stride-byte-offset (8-row group pitch)
__device__ __forceinline__ uint64_t m_desc(uint32_t saddr, uint32_t sbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

// flat tile id -> (b, column, d); tiles of one column are consecutive in d (the marching axis)
struct TileCoord { int b, th, tw, d; };
__device__ __forceinline__ TileCoord decode_tile(long long t, const MarchParams& p) {
    TileCoord c;
    c.d = (int)(t % p.D);
    long long col = t / p.D;
    c.tw = (int)(col % p.nTW);
    c.th = (int)((col / p.nTW) % p.nTH);
    c.b = (int)(col / ((long long)p.nTW * p.nTH));
    return c;
}

// Epilogue of one worker warp: NCOL accumulator columns of its 32 voxel rows, for every tile of the CTA.
// The per-tile body is the critical path of the HBM-bound layers (every tile passes through all worker warps), so
// everything tile-invariant is hoisted: bias values live in registers, the tile coordinate is advanced
// incrementally (no 64-bit divisions), the activation switch sits outside the element loops, and the
// GroupNorm sums are per-thread running sums that are transposed / reduced once per (CTA, sample).
template <int NCOL, bool F16>
__device__ __forceinline__ void march_epilogue(const MarchParams& p, long long t_begin, long long t_end, int q, int lane, int col0,
                                               uint32_t tmem_base, uint32_t t_full0, uint32_t t_empty0, float* tr) {
    // tr: this warp's 32 x (NCOL + 4) float transposition tile.  A lane holds NCOL columns of ITS row; stored directly, every
    // STG.128 touches 32 different rows (32 LSU wavefronts per instruction).  Through the tile NCOL/4 neighbouring lanes write one
    // row's NCOL * 4 bytes, and the rows of a tile line are adjacent in memory: an instruction covers 512 contiguous bytes.
    constexpr int TS = NCOL + 4, LPR = NCOL / 4, RPI = 32 / LPR;
    if (t_begin >= t_end) return;
    const int row = q * 32 + lane;
    const int lh = row / MR_TW, lw = row % MR_TW;
    const bool vec_ok = ((p.out_ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) && p.Cout == 32;
    const bool want_stats = p.stats != nullptr;
    const int act = p.out_act;
    float bias_r[NCOL], acc_s[NCOL], acc_q[NCOL];
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
        bias_r[k] = (p.bias && col0 + k < p.Cout) ? __ldg(p.bias + col0 + k) : 0.f;
        acc_s[k] = 0.f; acc_q[k] = 0.f;
    }
    TileCoord c = decode_tile(t_begin, p);
    int run_b = c.b;
    auto flush = [&]() {
        if (want_stats) {
            // transposing butterfly over the 32 rows of the warp: lane L ends up with column L of the slice
            float s[32], qq[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) { s[k] = k < NCOL ? acc_s[k % NCOL] : 0.f; qq[k] = k < NCOL ? acc_q[k % NCOL] : 0.f; }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float send_s = up ? s[i] : s[i + off], keep_s = up ? s[i + off] : s[i];
                    const float send_q = up ? qq[i] : qq[i + off], keep_q = up ? qq[i + off] : qq[i];
                    s[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
                    qq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
                }
            }
            const int ch = col0 + lane;
            if (lane < NCOL && ch < p.Cout) {
                atomicAdd(p.stats + ((size_t)run_b * p.Cout + ch) * 2 + 0, (double)s[0]);
                atomicAdd(p.stats + ((size_t)run_b * p.Cout + ch) * 2 + 1, (double)qq[0]);
            }
        }
#pragma unroll
        for (int k = 0; k < NCOL; ++k) { acc_s[k] = 0.f; acc_q[k] = 0.f; }
    };
    uint32_t tile_n = 0;
    for (long long t = t_begin; t < t_end; ++t, ++tile_n) {
        if (c.b != run_b) { flush(); run_b = c.b; }
        const uint32_t acc = tile_n % MR_ACC;
        m_mbar_wait(t_full0 + 8 * acc, (tile_n / MR_ACC) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[NCOL];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * MR_BN + (uint32_t)col0;
        if constexpr (NCOL == 32) {
            m_tmem_ld32(taddr, r);
        } else {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        m_mbar_arrive(t_empty0 + 8 * acc);                     // accumulator may be overwritten
        float v[NCOL];
#pragma unroll
        for (int k = 0; k < NCOL; ++k) v[k] = F16 ? fmaf(__uint_as_float(r[k]), p.acc_scale, bias_r[k]) : __uint_as_float(r[k]) + bias_r[k];
        if (p.accumulate) {                                     // later pass of the compensated mode: add the partial result
            const int ah = c.th * MR_TH + lh, aw = c.tw * MR_TW + lw;
            if (ah < p.H && aw < p.W) {
                const float* src = p.y + ((((size_t)c.b * p.D + c.d) * p.H + ah) * p.W + aw) * p.out_ldc + col0;
                if (vec_ok) {
#pragma unroll
                    for (int k = 0; k < NCOL; k += 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(src + k);
                        v[k] += t4.x; v[k + 1] += t4.y; v[k + 2] += t4.z; v[k + 3] += t4.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NCOL; ++k)
                        if (col0 + k < p.Cout) v[k] += src[k];
                }
            }
        }
        if (act == SS_ACT_RELU) {
#pragma unroll
            for (int k = 0; k < NCOL; ++k) v[k] = fmaxf(v[k], 0.f);
        } else if (act == SS_ACT_GELU) {
#pragma unroll
            for (int k = 0; k < NCOL; ++k) v[k] = gelu_erf(v[k]);
        } else if (act == SS_ACT_SWISH) {
#pragma unroll
            for (int k = 0; k < NCOL; ++k) v[k] = swish_f(v[k]);
        }
        const int oh = c.th * MR_TH + lh, ow = c.tw * MR_TW + lw;
        if (vec_ok) {
#pragma unroll
            for (int k = 0; k < NCOL; k += 4) *reinterpret_cast<float4*>(tr + lane * TS + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
            __syncwarp();
            float* plane = p.y + (((size_t)c.b * p.D + c.d) * p.H) * p.W * p.out_ldc + col0 + (lane % LPR) * 4;
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                const int rr = RPI * j + lane / LPR;
                const int r2 = q * 32 + rr;
                const int oh2 = c.th * MR_TH + r2 / MR_TW, ow2 = c.tw * MR_TW + r2 % MR_TW;
                const float4 t4 = *reinterpret_cast<const float4*>(tr + rr * TS + (lane % LPR) * 4);
                if (oh2 < p.H && ow2 < p.W) *reinterpret_cast<float4*>(plane + ((size_t)oh2 * p.W + ow2) * p.out_ldc) = t4;
            }
            __syncwarp();
        }
        if (oh < p.H && ow < p.W) {
            if (!vec_ok) {
                float* dst = p.y + ((((size_t)c.b * p.D + c.d) * p.H + oh) * p.W + ow) * p.out_ldc + col0;
#pragma unroll
                for (int k = 0; k < NCOL; ++k)
                    if (col0 + k < p.Cout) dst[k] = v[k];
            }
            if (want_stats && p.sr.has(c.d)) {
#pragma unroll
                for (int k = 0; k < NCOL; ++k) { acc_s[k] += v[k]; acc_q[k] = fmaf(v[k], v[k], acc_q[k]); }
            }
        }
        if (++c.d == p.D) {                                     // next flat tile: d fastest, then tw, th, b
            c.d = 0;
            if (++c.tw == p.nTW) {
                c.tw = 0;
                if (++c.th == p.nTH) { c.th = 0; ++c.b; }
            }
        }
    }
    flush();
}

// F16 = the single-launch fp16-split variant (SS_MATH_F16X3 / SS_MATH_F16, common.cuh:split_f16x4): resident weight rows and
// plane rows hold [hi | lo] fp16 halves, six kind::f16 MMAs per shift instead of four kind::tf32 ones.
template <int KS, bool F16>
__global__ void __launch_bounds__(MR_THREADS, 1)
conv_march32_kernel(const MarchParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW) {
    using Cfg = MarchCfg<KS>;
    constexpr int MR_HW = Cfg::HW, MR_PLANE_ROWS = Cfg::PLANE_ROWS, MR_PLANE_BYTES = Cfg::PLANE_BYTES, MR_W_BYTES = Cfg::W_BYTES;
    constexpr int PAD = Cfg::PAD;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment as an OFFSET from the extern __shared__ array: pointers derived this way keep the shared state space
    // (ld/st.shared); rounding a uintptr_t instead turns every access through them into a generic load / store
    unsigned char* base = smem_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    unsigned char* wres = base;                                   // resident weights
    unsigned char* planes = base + MR_W_BYTES;                    // MR_NP plane slots (MR_W_BYTES is a multiple of 1024)
    unsigned char* aux = planes + MR_NP * MR_PLANE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
    // barriers: w_full, p_full[NP], p_ready[NP], p_empty[NP], t_full[ACC], t_empty[ACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 3 * MR_NP + 2 * MR_ACC);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // provably warp-uniform role index
    const uint32_t leader = (lane == 0) ? 1u : 0u;
    const uint32_t w_full = m_smem_u32(bars), p_full0 = m_smem_u32(bars + 1), p_ready0 = m_smem_u32(bars + 1 + MR_NP),
                   p_empty0 = m_smem_u32(bars + 1 + 2 * MR_NP), t_full0 = m_smem_u32(bars + 1 + 3 * MR_NP),
                   t_empty0 = m_smem_u32(bars + 1 + 3 * MR_NP + MR_ACC);
    const bool has_aff = (p.in_scale != nullptr);
    const bool in_relu = (p.in_act == SS_ACT_RELU);
    const bool fixup = F16 || has_aff || in_relu || p.a_lo;

    // this CTA's contiguous range of flat tiles
    const long long t_begin = p.total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = p.total_tiles * (blockIdx.x + 1) / gridDim.x;

    if (tid == 0) {
        m_mbar_init(w_full, 1);
        for (int s = 0; s < MR_NP; ++s) {
            m_mbar_init(p_full0 + 8 * s, 1);
            m_mbar_init(p_ready0 + 8 * s, 128);
            m_mbar_init(p_empty0 + 8 * s, 1);
        }
        for (int a = 0; a < MR_ACC; ++a) {
            m_mbar_init(t_full0 + 8 * a, 1);
            m_mbar_init(t_empty0 + 8 * a, fixup ? 128 : 256);   // plain inputs: all 8 worker warps drain TMEM
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(m_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t planes_u32 = m_smem_u32(planes), wres_u32 = m_smem_u32(wres);

    // The plane stream: for every maximal run of consecutive tiles of one column [d0, d1) the planes
    // d0-1 .. d1 are loaded in order; tile d uses stream planes (d-d0), +1, +2.  All roles walk the
    // same runs, so a global plane counter L gives slot = L % NP and the barrier parity.
    if (warp == 8) {
        // ======================= TMA PRODUCER (warp-uniform; one elected lane issues) ===========
        if (t_begin < t_end) {
            m_mbar_expect_tx_elect(w_full, MR_W_BYTES);
            for (int i = 0; i < Cfg::TAPS; ++i) {           // tap (kd, ce) lands at slot ce*3 + (2-kd): [kd=2|kd=1|kd=0] per shift
                const int kd = i / (KS * KS), ce = i % (KS * KS);
                const int pos = KS == 3 ? ce * 3 + (2 - kd) : i;
                m_tma_2d_elect(wres_u32 + pos * 32 * 128, &tmW, w_full, 0, i * 32);
            }
            uint32_t L = 0;
            long long t = t_begin;
            while (t < t_end) {
                const TileCoord c = decode_tile(t, p);
                const int run = (int)min((long long)(p.D - c.d), t_end - t);     // tiles of this column handled here
                for (int s = 0; s < run + KS - 1; ++s, ++L) {
                    const uint32_t slot = L % MR_NP;
                    m_mbar_wait(p_empty0 + 8 * slot, ((L / MR_NP) & 1u) ^ 1u);
                    const uint32_t bar = p_full0 + 8 * slot;
                    m_mbar_expect_tx_elect(bar, MR_PLANE_ROWS * 128);
                    m_tma_5d_elect(planes_u32 + slot * MR_PLANE_BYTES, &tmA, bar, 0, c.tw * MR_TW - PAD, c.th * MR_TH - PAD, c.d - PAD + s, c.b);
                    __syncwarp();
                }
                t += run;
            }
        }
    } else if (warp == 9) {
        // ======================= MMA ISSUER ======================================================
        if (t_begin < t_end) {
            constexpr uint32_t idesc0 = F16 ? ((1u << 4) | ((uint32_t)(128 >> 4) << 24))
                                            : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24));
            constexpr int NMMA = F16 ? 6 : 4;                // MMAs per (kh,kw) shift
            m_mbar_wait(w_full, 0);
            const uint32_t rdy0 = fixup ? p_ready0 : p_full0;
            uint32_t L = 0, tile_n = 0;
            long long t = t_begin;
            while (t < t_end) {
                const TileCoord c = decode_tile(t, p);
                const int run = (int)min((long long)(p.D - c.d), t_end - t);
                if constexpr (KS == 1) {
                    for (int i = 0; i < run; ++i, ++tile_n, ++L) {
                        const uint32_t slot = L % MR_NP, acc = tile_n % MR_ACC;
                        m_mbar_wait(rdy0 + 8 * slot, (L / MR_NP) & 1u);
                        m_mbar_wait(t_empty0 + 8 * acc, ((tile_n / MR_ACC) & 1u) ^ 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        {
                            const uint64_t adesc = m_desc(planes_u32 + (uint32_t)slot * MR_PLANE_BYTES, MR_HW * 128);
                            const uint64_t bdesc = m_desc(wres_u32, 1024);
                            if constexpr (F16) {
                                constexpr uint32_t A1_HI = ((uint32_t)(MR_HW * 128) >> 4) | (1u << 14) | (2u << 29);
                                constexpr uint32_t B1_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
#pragma unroll
                                for (int k = 0; k < 6; ++k)
                                    if (k < p.f16_n)
                                        m_umma_lo_f16<A1_HI, B1_HI>(tmem_base + (uint32_t)(acc * MR_BN), (uint32_t)adesc + kF16A[k], (uint32_t)bdesc + kF16B[k],
                                                                    idesc0 | ((uint32_t)(MR_BN >> 3) << 17), k ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    m_umma_tf32(tmem_base + (uint32_t)(acc * MR_BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                                idesc0 | ((uint32_t)(MR_BN >> 3) << 17), k ? 1u : 0u, leader);
                            }
                            m_umma_commit(t_full0 + 8 * acc, leader);
                            m_umma_commit(p_empty0 + 8 * slot, leader);
                        }
                        __syncwarp();
                    }
                } else {
                    // stream plane s (depth c.d - 1 + s) feeds the output tiles j = s-2, s-1, s of this run through
                    // kd = 2, 1, 0; tile j (global number tile_n + j) accumulates in TMEM ring slot (tile_n + j) % 4.
                    // The issue loop is the critical path of this kernel (one thread, ~1 MMA per 50 clk needed):
                    // everything that varies per MMA is a 32-bit add of a compile-time constant to a descriptor word.
                    constexpr uint32_t A_HI = ((uint32_t)(MR_HW * 128) >> 4) | (1u << 14) | (2u << 29);
                    constexpr uint32_t B_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
                    const uint32_t w_lo = ((wres_u32 >> 4) & 0x3FFFu) | (1u << 16);
                    for (int s = 0; s < run + 2; ++s, ++L) {
                        const uint32_t slot = L & (MR_NP - 1);
                        m_mbar_wait(rdy0 + 8 * slot, (L / MR_NP) & 1u);
                        const int jlo = max(0, s - 2), jhi = min(run - 1, s);
                        const bool fresh = s <= run - 1;                        // tile s gets its first contribution here
                        if (fresh) {
                            const uint32_t n = tile_n + (uint32_t)s;
                            m_mbar_wait(t_empty0 + 8 * (n & (MR_ACC - 1)), ((n / MR_ACC) & 1u) ^ 1u);
                        }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_lo = (((planes_u32 + slot * MR_PLANE_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
                        // window [jlo, jhi] -> at most two runs of consecutive ring slots (split at the wrap)
                        const uint32_t r0 = (tile_n + (uint32_t)jlo) & (MR_ACC - 1);
                        const int cnt = jhi - jlo + 1;
                        const int n0 = min(cnt, MR_ACC - (int)r0), n1 = cnt - n0;
                        const uint32_t d0 = tmem_base + r0 * MR_BN, d1 = tmem_base;
                        const uint32_t b0 = w_lo + (uint32_t)(2 - (s - jlo)) * 256u, b1 = b0 + (uint32_t)n0 * 256u;   // 32 rows = 256 x 16 B
                        const uint32_t id0 = idesc0 | ((uint32_t)(n0 * MR_BN >> 3) << 17), id1 = idesc0 | ((uint32_t)(n1 * MR_BN >> 3) << 17);
                        // first MMA of the plane: tile s (the last of the window) starts from zero, the others accumulate
                        // one MMA of the plane: TF32 k-step or fp16 (A part, B part) pair number j of shift ce
                        auto issue = [&](uint32_t dcol, uint32_t aw, uint32_t bw, uint32_t idw, uint32_t accum) {
                            if constexpr (F16) m_umma_lo_f16<A_HI, B_HI>(dcol, aw, bw, idw, accum);
                            else m_umma_lo<A_HI, B_HI>(dcol, aw, bw, idw, accum);
                        };
                        if (fresh) {
                            const uint32_t rs = (tile_n + (uint32_t)jhi) & (MR_ACC - 1);
                            if (cnt > 1) {
                                const int m0 = min(cnt - 1, MR_ACC - (int)r0), m1 = cnt - 1 - m0;
                                issue(d0, a_lo, b0, idesc0 | ((uint32_t)(m0 * MR_BN >> 3) << 17), 1u);
                                if (m1 > 0) issue(d1, a_lo, b0 + (uint32_t)m0 * 256u, idesc0 | ((uint32_t)(m1 * MR_BN >> 3) << 17), 1u);
                            }
                            issue(tmem_base + rs * MR_BN, a_lo, b0 + (uint32_t)(cnt - 1) * 256u, idesc0 | ((uint32_t)(MR_BN >> 3) << 17), 0u);
                        }
#pragma unroll
                        for (int i = 0; i < 9 * NMMA; ++i) {
                            if (i == 0 && fresh) continue;
                            const int ce = i / NMMA, j = i % NMMA;
                            if (F16 && j >= p.f16_n) continue;
                            const uint32_t ao = ((uint32_t)(((ce / 3) * MR_HW + (ce % 3)) * 128) >> 4) + (uint32_t)(F16 ? kF16A[j] : 2 * j);
                            const uint32_t bo = ((uint32_t)(ce * 96 * 128) >> 4) + (uint32_t)(F16 ? kF16B[j] : 2 * j);
                            issue(d0, a_lo + ao, b0 + bo, id0, 1u);
                            if (n1 != 0) issue(d1, a_lo + ao, b1 + bo, id1, 1u);
                        }
                        if (s >= 2) m_umma_commit(t_full0 + 8 * ((tile_n + (uint32_t)s - 2u) & (MR_ACC - 1)), leader);   // tile s-2 is complete
                        m_umma_commit(p_empty0 + 8 * slot, leader);
                        __syncwarp();
                    }
                    tile_n += (uint32_t)run;
                }
                t += run;
            }
        }
    } else if (warp >= 4 && fixup) {
        // ======================= FIX-UP WARPS (4..7): pending affine / ReLU once per plane, in place ==
        if (t_begin < t_end) {
            const int ft = tid - 128;                  // 0..127
            long long L = 0;
            long long t = t_begin;
            int cur_b = -1;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            const int chunk = ft & 7;                  // 16-byte chunk (4 channels) handled by this thread
            while (t < t_end) {
                const TileCoord c = decode_tile(t, p);
                const int run = (int)min((long long)(p.D - c.d), t_end - t);
                if (has_aff && c.b != cur_b) {
                    cur_b = c.b;
                    sc = ldg_f4(p.in_scale + (size_t)c.b * 32 + chunk * 4);
                    sh = ldg_f4(p.in_shift + (size_t)c.b * 32 + chunk * 4);
                }
                // item `it` of this thread = row (ft >> 3) + 16 it, chunk ft & 7; row & 7 is the same for all its items, so the swizzled
                // offsets inside a row are thread constants and the row validity (a bit mask) only depends on the tile column
                constexpr int ITERS = (MR_PLANE_ROWS + 15) / 16;
                const int r0 = ft >> 3, rsw = r0 & 7;
                const uint32_t rd = (uint32_t)(r0 * 128 + ((chunk ^ rsw) << 4));
                const uint32_t wr_hi = (uint32_t)(r0 * 128 + (((chunk >> 1) ^ rsw) << 4) + ((chunk & 1) << 3));
                const uint32_t wr_lo = (uint32_t)(r0 * 128 + (((4 + (chunk >> 1)) ^ rsw) << 4) + ((chunk & 1) << 3));
                uint32_t vmask = 0;
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const int r = r0 + 16 * it;
                    const int hh = c.th * MR_TH - PAD + r / MR_HW, ww = c.tw * MR_TW - PAD + r % MR_HW;
                    if (r < MR_PLANE_ROWS && (unsigned)hh < (unsigned)p.H && (unsigned)ww < (unsigned)p.W) vmask |= 1u << it;
                }
                auto fixv = [&](float4 v) {
                    if (has_aff) {
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                    }
                    if (in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    return v;
                };
                for (int s = 0; s < run + KS - 1; ++s, ++L) {
                    const int slot = (int)(L % MR_NP);
                    m_mbar_wait(p_full0 + 8 * slot, (uint32_t)(L / MR_NP) & 1u);
                    const int dpl = c.d - PAD + s;
                    if ((unsigned)dpl < (unsigned)p.D) {           // planes outside the volume are all padding
                        unsigned char* pl = planes + slot * MR_PLANE_BYTES;
                        if constexpr (F16) {
                            // the 8 lanes of a row read their chunks of a batch of rows, sync, then overwrite the rows with [hi | lo]
                            constexpr int HALF = (ITERS + 1) / 2;
#pragma unroll
                            for (int b0 = 0; b0 < ITERS; b0 += HALF) {
                                uint2 hi[HALF], lo[HALF];
#pragma unroll
                                for (int k = 0; k < HALF; ++k) {
                                    const int it = b0 + k;
                                    if (it < ITERS && ((vmask >> it) & 1u))
                                        split_f16x4(fixv(*reinterpret_cast<const float4*>(pl + rd + it * 2048)), hi[k], lo[k]);
                                }
                                __syncwarp();
#pragma unroll
                                for (int k = 0; k < HALF; ++k) {
                                    const int it = b0 + k;
                                    if (it < ITERS && ((vmask >> it) & 1u)) {
                                        *reinterpret_cast<uint2*>(pl + wr_hi + it * 2048) = hi[k];
                                        *reinterpret_cast<uint2*>(pl + wr_lo + it * 2048) = lo[k];
                                    }
                                }
                            }
                            __syncwarp();
                        } else {
                            auto fix = [&](auto lo_tag) {
                                constexpr bool LO = decltype(lo_tag)::value;
#pragma unroll
                                for (int it = 0; it < ITERS; ++it)
                                    if ((vmask >> it) & 1u) {
                                        float4* ptr = reinterpret_cast<float4*>(pl + rd + it * 2048);
                                        const float4 v = fixv(*ptr);
                                        uint4 o;
                                        o.x = f2tf32_part<LO>(v.x); o.y = f2tf32_part<LO>(v.y); o.z = f2tf32_part<LO>(v.z); o.w = f2tf32_part<LO>(v.w);
                                        *reinterpret_cast<uint4*>(ptr) = o;
                                    }
                            };
                            SS_UNSWITCH_LO(p.a_lo, fix);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    m_mbar_arrive(p_ready0 + 8 * slot);
                }
                t += run;
            }
        }
    } else {
        // ======================= EPILOGUE WARPS (0..3, plus 4..7 when there is no fix-up work) =======
        // warps sharing a TMEM lane quarter split the 32 output columns when all 8 worker warps drain
        float* trbase = reinterpret_cast<float*>(aux + 512);
        if (fixup) march_epilogue<32, F16>(p, t_begin, t_end, warp & 3, lane, 0, tmem_base, t_full0, t_empty0, trbase + warp * (32 * 36));
        else march_epilogue<16, F16>(p, t_begin, t_end, warp & 3, lane, (warp >> 2) * 16, tmem_base, t_full0, t_empty0, trbase + warp * (32 * 20));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

typedef CUresult (*MEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int KS, bool F16>
static int launch_march(const MarchParams& p, const ss_conv3d_desc* d, const float* x, const float* w_kmajor, bool fixup,
                        MEncodeTiledFn encode, cudaStream_t st) {
    using Cfg = MarchCfg<KS>;
    alignas(64) CUtensorMap tmA, tmW;
    {
        cuuint64_t gdim[5] = {32, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.B};
        cuuint64_t gstr[4] = {(cuuint64_t)d->in_ldc * 4, (cuuint64_t)p.W * d->in_ldc * 4, (cuuint64_t)p.H * p.W * d->in_ldc * 4,
                              (cuuint64_t)p.D * p.H * p.W * d->in_ldc * 4};
        cuuint32_t box[5] = {32, (cuuint32_t)Cfg::HW, (cuuint32_t)Cfg::HH, 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (encode(&tmA, fixup ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, const_cast<float*>(x), gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return set_arg_error("conv_march32: tensor map A");
    }
    {
        cuuint64_t gdim[2] = {32, (cuuint64_t)Cfg::TAPS * 32};
        cuuint64_t gstr[1] = {32 * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)Cfg::W_BOX_ROWS};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w_kmajor), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return set_arg_error("conv_march32: tensor map W");
    }
    const size_t smem = 1024 + Cfg::W_BYTES + MR_NP * Cfg::PLANE_BYTES + 512 + 8 * 32 * 20 * sizeof(float);   // barriers, then the epilogue warps' transposition tiles
    static thread_local bool configured = false;
    if (!configured) {
        SS_CUDA(cudaFuncSetAttribute(conv_march32_kernel<KS, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)min((long long)sms, p.total_tiles);
    conv_march32_kernel<KS, F16><<<grid, MR_THREADS, smem, st>>>(p, tmA, tmW);
    return check_launch(F16 ? "conv_march32_f16x3_kernel" : "conv_march32_kernel");
}

int conv_march32_eligible(const ss_conv3d_desc* d) {
    if (d->transposed || d->Cin != 32 || d->cout_packed != 32) return 0;
    const bool k3 = d->kd == 3 && d->kh == 3 && d->kw == 3 && d->pd == 1 && d->ph == 1 && d->pw == 1;
    const bool k1 = d->kd == 1 && d->kh == 1 && d->kw == 1 && d->pd == 0 && d->ph == 0 && d->pw == 0;
    if (!k3 && !k1) return 0;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return 0;
    if (d->Win < 8 || d->Hin < 8 || d->Din < 3) return 0;
    if (d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win || d->math != SS_MATH_TF32) return 0;
    return 1;
}

// returns 1 if the layer was handled by the marching kernel
int try_conv_march32(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                     const float* w_kmajor, const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ConvPass& ps) {
    if (d->transposed || d->Cin != 32 || d->cout_packed != 32) return 0;      // Cout < 32 layers arrive padded to 32 weight rows
    const bool k3 = d->kd == 3 && d->kh == 3 && d->kw == 3 && d->pd == 1 && d->ph == 1 && d->pw == 1;
    const bool k1 = d->kd == 1 && d->kh == 1 && d->kw == 1 && d->pd == 0 && d->ph == 0 && d->pw == 0;
    if (!k3 && !k1) return 0;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return 0;
    if (d->Win < 8 || d->Hin < 8 || d->Din < 3) return 0;
    if (d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win || d->math != SS_MATH_TF32) return 0;
    static MEncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) return 0;
        encode = reinterpret_cast<MEncodeTiledFn>(ptr);
    }
    MarchParams p;
    p.B = d->B; p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.Cout = d->Cout; p.out_ldc = d->out_ldc;
    p.in_act = d->in_act; p.out_act = d->out_act;
    p.nTH = (p.H + MR_TH - 1) / MR_TH; p.nTW = (p.W + MR_TW - 1) / MR_TW;
    p.total_tiles = (long long)p.B * p.nTH * p.nTW * p.D;
    p.in_scale = in_scale; p.in_shift = in_shift; p.bias = bias; p.y = y; p.stats = stats;
    p.a_lo = ps.a_lo; p.accumulate = ps.accumulate; p.sr = stats_range_of(d); p.acc_scale = ps.acc_scale; p.f16_n = ps.f16_n;
    const bool fixup = (in_scale != nullptr) || (d->in_act == SS_ACT_RELU) || ps.a_lo || ps.f16;
    if (ps.f16) *rc = k3 ? launch_march<3, true>(p, d, x, w_kmajor, fixup, encode, st) : launch_march<1, true>(p, d, x, w_kmajor, fixup, encode, st);
    else *rc = k3 ? launch_march<3, false>(p, d, x, w_kmajor, fixup, encode, st) : launch_march<1, false>(p, d, x, w_kmajor, fixup, encode, st);
    return 1;
}

}  // namespace ss
