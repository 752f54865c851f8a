// Transposed 3-D convolution k3 s2 p1 (output_padding 1) on tcgen05: the hourglass up-convolutions
// conv5 / conv6 (ConvTranspose3d 128->64 and 64->32, ViewTransformerLSSVoxel.py:81-86).
//
// A transposed conv with stride 2 splits into 8 output parity classes (od%2, oh%2, ow%2); class r
// along an axis uses kernel tap 1 at input offset 0 (r = 0) or taps 0 / 2 at offsets +1 / 0 (r = 1),
// 27 (class, tap) pairs in total.  The per-tap box kernel runs 8 grids of tiny CTAs (2..16 K-steps
// each, prologue/epilogue dominated).  Here ONE CTA owns a 16(h) x 8(w) INPUT tile at input plane q and
// produces all 8 classes (1024 output voxels): per 32-channel chunk it loads the planes q and q+1 once
// (TMA 5-D box 17 x 9 voxels, OOB zero fill), every (class, tap) A operand is that plane through a
// shifted UMMA descriptor (start + (off_h*9 + off_w) rows, SBO = 9 rows), weight tiles stream through
// a TMA ring, and the 8 class accumulators (8 x BN fp32 columns) live side by side in TMEM.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace ss {

constexpr int TP_TH = 16, TP_TW = 8;
constexpr int TP_HH = TP_TH + 1, TP_HW = TP_TW + 1;
constexpr int TP_PLANE_ROWS = TP_HH * TP_HW;           // 153
constexpr int TP_PLANE_BYTES = 20 * 1024;              // 153*128 = 19584 -> padded
constexpr int TP_NPL = 4;                              // plane ring: 2 planes per chunk, one chunk of prefetch
constexpr int TP_SB_BYTES = 24 * 1024;                 // weight-tile ring (small, so that two CTAs fit on an SM when BN = 32)
constexpr int TP_WORKERS = 256;
constexpr int TP_THREADS = TP_WORKERS + 96;

struct TposeParams {
    int B, Din, Hin, Win, Cin, Dout, Hout, Wout, Cout, CoutP, out_ldc, in_act, out_act;
    int nTH, nTW;
    const float* in_scale;
    const float* in_shift;
    const float* bias;
    float* y;
    double* stats;
    // fused join (ss_conv3d_tc_join_fwd): y = out_act((conv + bias) * o_scale + o_shift + res_act(res * r_scale + r_shift))
    const float* o_scale;
    const float* o_shift;
    const float* res;
    const float* r_scale;
    const float* r_shift;
    int res_ldc, res_act, has_join;
    int a_lo, accumulate;            // ConvPass (common.cuh)
    StatsRange sr;                   // output planes that contribute to stats
    int gather;                      // residual read through the transposition tile (default; STEREOSCENE_B200_TPOSE_GATHER=0: every lane reads its own row)
    float acc_scale;                 // F16 variant: accumulator scale (power of two)
    int f16_n;                       // F16 variant: MMAs per chunk and tap (6 = compensated, 2 = fp16 single pass)
};

__device__ __forceinline__ uint32_t t_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TWAIT_DONE;\n\t"
        "bra TWAIT_LOOP;\n\t"
        "TWAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void t_tma_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void t_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void t_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void t_umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ uint64_t t_desc(uint32_t saddr, uint32_t sbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

// The 27 (class, tap) pairs in issue order.  Per axis: parity 0 -> (kernel 1, offset 0); parity 1 ->
// (kernel 0, offset 1) then (kernel 2, offset 0).  entry = cls | off_d<<3 | off_h<<4 | off_w<<5 | wtap<<8
struct TposeTable { int v[27]; };
__host__ __device__ constexpr TposeTable tpose_table() {
    TposeTable t{};
    int n = 0;
    for (int cls = 0; cls < 8; ++cls) {
        const int rd = cls >> 2, rh = (cls >> 1) & 1, rw = cls & 1;
        for (int a = 0; a < (rd ? 2 : 1); ++a)
            for (int c = 0; c < (rh ? 2 : 1); ++c)
                for (int e = 0; e < (rw ? 2 : 1); ++e) {
                    const int kd = rd ? (a ? 2 : 0) : 1, kh = rh ? (c ? 2 : 0) : 1, kw = rw ? (e ? 2 : 0) : 1;
                    const int od = rd ? (a ? 0 : 1) : 0, oh = rh ? (c ? 0 : 1) : 0, ow = rw ? (e ? 0 : 1) : 0;
                    t.v[n++] = cls | (od << 3) | (oh << 4) | (ow << 5) | (((kd * 3 + kh) * 3 + kw) << 8);
                }
    }
    return t;   // 27 entries
}

// F16 = the single-launch fp16-split variant (SS_MATH_F16X3 / SS_MATH_F16, common.cuh:split_f16x4)
template <int BN, bool F16>
__global__ void __launch_bounds__(TP_THREADS, BN <= 32 ? 2 : 1)
conv_tpose_kernel(const TposeParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    constexpr int B_BYTES = BN * 128;
    constexpr int TP_SB = TP_SB_BYTES / B_BYTES;
    constexpr int TMEM_COLS = 8 * BN <= 256 ? 256 : 512;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment as an OFFSET from the extern __shared__ array: pointers derived this way keep the shared state space
    // (ld/st.shared); rounding a uintptr_t instead turns every access through them into a generic load / store
    unsigned char* base = smem_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    unsigned char* planes = base;
    unsigned char* bring = base + TP_NPL * TP_PLANE_BYTES;
    unsigned char* aux = bring + TP_SB * B_BYTES;
    double* sstat = reinterpret_cast<double*>(aux);                          // [BN][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sstat + 2 * BN);            // pa_full[4] pa_ready[4] pa_empty[4] pb_full[8] pb_empty[8] accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TP_NPL + 2 * TP_SB + 1);
    int* tab = reinterpret_cast<int*>(tmem_slot + 4);                        // [27]
    float* ssc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tab + 32) + 15) & ~(uintptr_t)15);
    float* jsc = ssc + 2 * p.Cin;                                            // [4][BN]: o_scale, o_shift, r_scale, r_shift of batch b

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = uniform_warp_index();
    constexpr TposeTable TAB = tpose_table();
    const int tw_i = blockIdx.x % p.nTW, th_i = (blockIdx.x / p.nTW) % p.nTH;
    const int q = (blockIdx.x / (p.nTW * p.nTH)) % p.Din, b = blockIdx.x / (p.nTW * p.nTH * p.Din);
    const int h0 = th_i * TP_TH, w0 = tw_i * TP_TW;

    const uint32_t pa_full0 = t_smem_u32(bars), pa_ready0 = t_smem_u32(bars + TP_NPL), pa_empty0 = t_smem_u32(bars + 2 * TP_NPL),
                   pb_full0 = t_smem_u32(bars + 3 * TP_NPL), pb_empty0 = t_smem_u32(bars + 3 * TP_NPL + TP_SB),
                   accum_bar = t_smem_u32(bars + 3 * TP_NPL + 2 * TP_SB);
    const bool has_aff = (p.in_scale != nullptr);
    const bool in_relu = (p.in_act == SS_ACT_RELU);
    const bool fixup = F16 || has_aff || in_relu || p.a_lo;
    const int kchunks = p.Cin / 32;

    if (tid == 0) {
        for (int i = 0; i < 27; ++i) tab[i] = TAB.v[i];
        for (int s = 0; s < TP_NPL; ++s) {
            t_mbar_init(pa_full0 + 8 * s, 1);
            t_mbar_init(pa_ready0 + 8 * s, TP_WORKERS);
            t_mbar_init(pa_empty0 + 8 * s, 1);
        }
        for (int s = 0; s < TP_SB; ++s) {
            t_mbar_init(pb_full0 + 8 * s, 1);
            t_mbar_init(pb_empty0 + 8 * s, 1);
        }
        t_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    for (int i = tid; i < 2 * BN; i += TP_THREADS) sstat[i] = 0.0;
    if (p.has_join)
        for (int i = tid; i < BN; i += TP_THREADS) {
            const bool in = i < p.Cout;
            jsc[i] = (in && p.o_scale) ? __ldg(p.o_scale + (size_t)b * p.Cout + i) : 1.f;
            jsc[BN + i] = (in && p.o_shift) ? __ldg(p.o_shift + (size_t)b * p.Cout + i) : 0.f;
            jsc[2 * BN + i] = (in && p.r_scale) ? __ldg(p.r_scale + (size_t)b * p.Cout + i) : 1.f;
            jsc[3 * BN + i] = (in && p.r_shift) ? __ldg(p.r_shift + (size_t)b * p.Cout + i) : 0.f;
        }
    if (has_aff)
        for (int i = tid; i < p.Cin; i += TP_THREADS) {
            ssc[i] = __ldg(p.in_scale + (size_t)b * p.Cin + i);
            ssc[p.Cin + i] = __ldg(p.in_shift + (size_t)b * p.Cin + i);
        }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t_smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t planes_u32 = t_smem_u32(planes), bring_u32 = t_smem_u32(bring);

    if (warp == 8) {
        // ======================= A PRODUCER: planes q, q+1 per chunk (warp-uniform, elected issue) ===
        for (int L = 0; L < kchunks * 2; ++L) {
            const int slot = L % TP_NPL;
            const uint32_t use = (uint32_t)(L / TP_NPL);
            t_mbar_wait(pa_empty0 + 8 * slot, (use & 1u) ^ 1u);
            const uint32_t bar = pa_full0 + 8 * slot;
            mbar_expect_tx_elect(bar, TP_PLANE_ROWS * 128);
            tma_5d_elect(planes_u32 + slot * TP_PLANE_BYTES, &tmA, bar, (L / 2) * 32, w0, h0, q + (L & 1), b);
            __syncwarp();
        }
    } else if (warp == 10) {
        // ======================= B PRODUCER: one weight tile per (chunk, class-tap) =================
        int L = 0;
        for (int ch = 0; ch < kchunks; ++ch) {
#pragma unroll
            for (int e = 0; e < 27; ++e, ++L) {
                const int slot = L % TP_SB;
                const uint32_t use = (uint32_t)(L / TP_SB);
                t_mbar_wait(pb_empty0 + 8 * slot, (use & 1u) ^ 1u);
                const uint32_t bar = pb_full0 + 8 * slot;
                mbar_expect_tx_elect(bar, B_BYTES);
                tma_2d_elect(bring_u32 + slot * B_BYTES, &tmB, bar, ch * 32, (TAB.v[e] >> 8) * p.CoutP);
                __syncwarp();
            }
        }
    } else if (warp == 9) {
        // ======================= MMA ISSUER (warp-uniform; see common.cuh) ==========================
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t A_HI = umma_desc_hi(TP_HW * 128), B_HI = umma_desc_hi(1024);
        const uint32_t rdy0 = fixup ? pa_ready0 : pa_full0;
        int Lb = 0;
        for (int ch = 0; ch < kchunks; ++ch) {
            const int s0 = (2 * ch) % TP_NPL, s1 = (2 * ch + 1) % TP_NPL;
            t_mbar_wait(rdy0 + 8 * s0, (uint32_t)((2 * ch) / TP_NPL) & 1u);
            t_mbar_wait(rdy0 + 8 * s1, (uint32_t)((2 * ch + 1) / TP_NPL) & 1u);
            const uint32_t a0_lo = umma_desc_lo(planes_u32 + (uint32_t)s0 * TP_PLANE_BYTES);
            const uint32_t a1_lo = umma_desc_lo(planes_u32 + (uint32_t)s1 * TP_PLANE_BYTES);
#pragma unroll
            for (int e = 0; e < 27; ++e, ++Lb) {
                const int bslot = Lb % TP_SB;
                t_mbar_wait(pb_full0 + 8 * bslot, (uint32_t)(Lb / TP_SB) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ent = TAB.v[e];
                const int cls = ent & 7, od = (ent >> 3) & 1, oh = (ent >> 4) & 1, ow = (ent >> 5) & 1;
                const uint32_t a_lo = (od ? a1_lo : a0_lo) + ((uint32_t)((oh * TP_HW + ow) * 128) >> 4);
                const uint32_t b_lo = umma_desc_lo(bring_u32 + bslot * B_BYTES);
                // first MMA into a class accumulator: its first tap of chunk 0 (taps of a class are consecutive)
                const bool first_tap = (e == 0) || ((TAB.v[e > 0 ? e - 1 : 0] & 7) != cls);
                if constexpr (F16) {
                    constexpr uint32_t idesc16 = make_idesc_f16(128, BN);
#pragma unroll
                    for (int i = 0; i < 6; ++i)
                        if (i < p.f16_n)
                            umma_ss_f16<A_HI, B_HI>(tmem_base + (uint32_t)(cls * BN), a_lo + kF16A[i], b_lo + kF16B[i], idesc16,
                                                    (ch == 0 && first_tap && i == 0) ? 0u : 1u);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_ss_tf32<A_HI, B_HI>(tmem_base + (uint32_t)(cls * BN), a_lo + 2 * k, b_lo + 2 * k, idesc,
                                                 (ch == 0 && first_tap && k == 0) ? 0u : 1u);
                }
                umma_commit_elect(pb_empty0 + 8 * bslot);
                if (e == 26) {
                    umma_commit_elect(pa_empty0 + 8 * s0);
                    umma_commit_elect(pa_empty0 + 8 * s1);
                }
                __syncwarp();
            }
        }
        umma_commit_elect(accum_bar);
        __syncwarp();
    } else if (fixup) {
        // ======================= WORKERS: pending affine / ReLU once per landed plane, in place =====
        for (int L = 0; L < kchunks * 2; ++L) {
            const int slot = L % TP_NPL;
            t_mbar_wait(pa_full0 + 8 * slot, (uint32_t)(L / TP_NPL) & 1u);
            const int dpl = q + (L & 1), c0 = (L / 2) * 32;
            if constexpr (F16) {
                if (dpl < p.Din) {
                    unsigned char* pl = planes + slot * TP_PLANE_BYTES;
                    constexpr int ITERS = (TP_PLANE_ROWS * 8 + TP_WORKERS - 1) / TP_WORKERS;
                    for (int it = 0; it < ITERS; ++it) {          // the 8 lanes of a row read, sync, then overwrite it with [hi | lo]
                        const int idx = tid + it * TP_WORKERS;
                        const int r = idx >> 3, chunk = idx & 7;
                        const int hh = h0 + r / TP_HW, ww = w0 + r % TP_HW;
                        const bool act = idx < TP_PLANE_ROWS * 8 && hh < p.Hin && ww < p.Win;
                        unsigned char* row = pl + r * 128;
                        uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
                        if (act) {
                            float4 v = *reinterpret_cast<const float4*>(row + ((chunk ^ (r & 7)) << 4));
                            if (has_aff) {
                                const float4 sc = *reinterpret_cast<const float4*>(ssc + c0 + chunk * 4);
                                const float4 sh = *reinterpret_cast<const float4*>(ssc + p.Cin + c0 + chunk * 4);
                                v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                            }
                            if (in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            split_f16x4(v, hi, lo);
                        }
                        __syncwarp();
                        if (act) {
                            *reinterpret_cast<uint2*>(row + (((chunk >> 1) ^ (r & 7)) << 4) + ((chunk & 1) << 3)) = hi;
                            *reinterpret_cast<uint2*>(row + (((4 + (chunk >> 1)) ^ (r & 7)) << 4) + ((chunk & 1) << 3)) = lo;
                        }
                        __syncwarp();
                    }
                }
            } else if (dpl < p.Din) {
                unsigned char* pl = planes + slot * TP_PLANE_BYTES;
                auto fix = [&](auto lo_tag) {
                    constexpr bool LO = decltype(lo_tag)::value;
                    for (int idx = tid; idx < TP_PLANE_ROWS * 8; idx += TP_WORKERS) {
                        const int r = idx >> 3, chunk = idx & 7;
                        const int hh = h0 + r / TP_HW, ww = w0 + r % TP_HW;
                        if (hh < p.Hin && ww < p.Win) {
                            float4* ptr = reinterpret_cast<float4*>(pl + r * 128 + ((chunk ^ (r & 7)) << 4));
                            float4 v = *ptr;
                            if (has_aff) {
                                const float4 sc = *reinterpret_cast<const float4*>(ssc + c0 + chunk * 4);
                                const float4 sh = *reinterpret_cast<const float4*>(ssc + p.Cin + c0 + chunk * 4);
                                v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                            }
                            if (in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            uint4 o;
                            o.x = f2tf32_part<LO>(v.x); o.y = f2tf32_part<LO>(v.y); o.z = f2tf32_part<LO>(v.z); o.w = f2tf32_part<LO>(v.w);
                            *reinterpret_cast<uint4*>(ptr) = o;
                        }
                    }
                };
                SS_UNSWITCH_LO(p.a_lo, fix);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t_mbar_arrive(pa_ready0 + 8 * slot);
        }
    }

    // ======================= EPILOGUE: 8 warps, (lane quarter, class half) ==========================
    if (warp < TP_WORKERS / 32) {
        t_mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int qq = warp & 3, chalf = warp >> 2;
        const int row = qq * 32 + lane;
        const int ih = h0 + row / TP_TW, iw = w0 + row % TP_TW;
        const bool vec_ok = ((p.out_ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        // The residual of the fused join does not depend on the MMAs and is the epilogue's longest wait (ncu: 38 % of the stall samples
        // were long-scoreboard on these loads): the 32 rows x 32 columns of the NEXT (class, chunk) item are copied global -> shared
        // memory with cp.async (8 neighbouring lanes per 128-byte row: full lines) while the current item is computed and stored.
        // Two per-warp 32 x 36-float tiles in the idle plane ring alternate; an item's tile is then re-used to transpose its stores.
        constexpr int CH = BN / 32, ITEMS = 4 * CH;
        float* tiles = reinterpret_cast<float*>(planes) + warp * (2 * 32 * 36);
        const bool res_async = p.gather && p.has_join && p.res != nullptr && ((p.res_ldc & 3) == 0) &&
                               ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0) && p.Cout >= BN;
        auto issue_res = [&](int item, float* buf) {
            const int cls2 = chalf * 4 + item / CH, cb = (item % CH) * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r2 = 4 * j + (lane >> 3), rowq = qq * 32 + r2;
                const int ih2 = h0 + rowq / TP_TW, iw2 = w0 + rowq % TP_TW;
                const int od2 = 2 * q + (cls2 >> 2), oh2 = 2 * ih2 + ((cls2 >> 1) & 1), ow2 = 2 * iw2 + (cls2 & 1);
                float* dst = buf + r2 * 36 + (lane & 7) * 4;
                if (ih2 < p.Hin && iw2 < p.Win && od2 < p.Dout && oh2 < p.Hout && ow2 < p.Wout) {
                    const size_t ov2 = (((size_t)b * p.Dout + od2) * p.Hout + oh2) * p.Wout + ow2;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(t_smem_u32(dst)), "l"(p.res + ov2 * p.res_ldc + cb + (lane & 7) * 4) : "memory");
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (res_async) issue_res(0, tiles);
#pragma unroll 1
        for (int item = 0; item < ITEMS; ++item) {
            const int cls = chalf * 4 + item / CH, ci = item % CH;
            const int od = 2 * q + (cls >> 2), oh = 2 * ih + ((cls >> 1) & 1), ow = 2 * iw + (cls & 1);
            const bool valid = ih < p.Hin && iw < p.Win && od < p.Dout && oh < p.Hout && ow < p.Wout;
            const size_t ov = (((size_t)b * p.Dout + od) * p.Hout + oh) * p.Wout + ow;
            const long long ov_ll = valid ? (long long)ov : -1;
            float* trw = tiles + (item & 1) * (32 * 36);
            if (res_async) {
                if (item + 1 < ITEMS) {
                    issue_res(item + 1, tiles + ((item + 1) & 1) * (32 * 36));
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncwarp();
            }
            {
                uint32_t r[32];
                t_tmem_ld32(tmem_base + ((uint32_t)(qq * 32) << 16) + (uint32_t)(cls * BN + ci * 32), r);
                const int cbase = ci * 32;
                float v[32];
                if constexpr (F16) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * p.acc_scale);
                }
                if (p.accumulate) {                            // later pass of the compensated mode: add the partial result
                    const float* src = p.y + ov * p.out_ldc + cbase;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (valid && cbase + k < p.Cout) r[k] = __float_as_uint(__uint_as_float(r[k]) + src[k]);
                }
                if (p.has_join) {
                    // residual-fused epilogue: the join that would re-read this tensor and the residual runs here
                    const bool rvec = valid && p.res && ((p.res_ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0) &&
                                      cbase + 32 <= p.Cout;
                    const float* rsrc = p.res ? p.res + ov * p.res_ldc + cbase : nullptr;
                    const bool rvec_w = res_async;                                 // warp-uniform: this item's residual rows are in trw
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        float rr[4] = {0.f, 0.f, 0.f, 0.f};
                        if (rvec_w) {
                            const float4 t4 = *reinterpret_cast<const float4*>(trw + lane * 36 + 4 * k4);
                            rr[0] = t4.x; rr[1] = t4.y; rr[2] = t4.z; rr[3] = t4.w;
                        } else if (rvec) {
                            const float4 t4 = ldg_f4(rsrc + 4 * k4);
                            rr[0] = t4.x; rr[1] = t4.y; rr[2] = t4.z; rr[3] = t4.w;
                        } else if (valid && p.res) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (cbase + 4 * k4 + e < p.Cout) rr[e] = __ldg(rsrc + 4 * k4 + e);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k = 4 * k4 + e, c = cbase + k;
                            float f = __uint_as_float(r[k]);
                            if (p.bias && c < p.Cout) f += __ldg(p.bias + c);
                            f = fmaf(f, jsc[c], jsc[BN + c]);
                            float rv = fmaf(rr[e], jsc[2 * BN + c], jsc[3 * BN + c]);
                            if (p.res_act == SS_ACT_RELU) rv = fmaxf(rv, 0.f);
                            if (p.res) f += rv;
                            v[k] = apply_act(f, p.out_act);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        float f = __uint_as_float(r[k]);
                        const int c = cbase + k;
                        if (p.bias && c < p.Cout) f += __ldg(p.bias + c);
                        v[k] = apply_act(f, p.out_act);
                    }
                }
                if (vec_ok && cbase + 32 <= p.Cout) {
                    // transposed through a per-warp 32 x 36 tile in the (idle) plane ring: 8 neighbouring lanes store one output
                    // voxel's 128 bytes instead of every lane storing its own row (32 wavefronts per STG.128)
                    __syncwarp();                                  // every lane has read its residual row from this tile
#pragma unroll
                    for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(trw + lane * 36 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r2 = 4 * j + (lane >> 3);
                        const long long ovr = __shfl_sync(0xffffffffu, ov_ll, r2);
                        const float4 t4 = *reinterpret_cast<const float4*>(trw + r2 * 36 + (lane & 7) * 4);
                        if (ovr >= 0) *reinterpret_cast<float4*>(p.y + (size_t)ovr * p.out_ldc + cbase + (lane & 7) * 4) = t4;
                    }
                    __syncwarp();
                } else if (valid) {
                    float* dst = p.y + ov * p.out_ldc + cbase;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (cbase + k < p.Cout) dst[k] = v[k];
                }
                if (p.stats) {
                    float s[32], sq[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) { s[k] = (valid && p.sr.has(od)) ? v[k] : 0.f; sq[k] = s[k] * s[k]; }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float send_s = up ? s[i] : s[i + off], keep_s = up ? s[i + off] : s[i];
                            const float send_q = up ? sq[i] : sq[i + off], keep_q = up ? sq[i + off] : sq[i];
                            s[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
                            sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
                        }
                    }
                    atomicAdd(&sstat[2 * (ci * 32 + lane) + 0], (double)s[0]);
                    atomicAdd(&sstat[2 * (ci * 32 + lane) + 1], (double)sq[0]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (p.stats) {
        for (int i = tid; i < BN; i += TP_THREADS)
            if (i < p.Cout) {
                atomicAdd(p.stats + ((size_t)b * p.Cout + i) * 2 + 0, sstat[2 * i + 0]);
                atomicAdd(p.stats + ((size_t)b * p.Cout + i) * 2 + 1, sstat[2 * i + 1]);
            }
    }
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*TEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BN, bool F16>
static int launch_tpose(const TposeParams& p, const CUtensorMap& tmA, const float* wk, TEncodeTiledFn encode, cudaStream_t st) {
    alignas(64) CUtensorMap tmB;
    cuuint64_t gdim[2] = {(cuuint64_t)p.Cin, (cuuint64_t)27 * p.CoutP};
    cuuint64_t gstr[1] = {(cuuint64_t)p.Cin * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wk), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return set_arg_error("conv_tpose: tensor map B");
    constexpr int TP_SB = TP_SB_BYTES / (BN * 128);
    const size_t smem = 1024 + (size_t)TP_NPL * TP_PLANE_BYTES + (size_t)TP_SB * BN * 128 + 2 * BN * sizeof(double) +
                        (3 * TP_NPL + 2 * TP_SB + 1) * sizeof(uint64_t) + 16 + 32 * sizeof(int) + 32 + (2 * (size_t)p.Cin + 4 * BN) * sizeof(float);
    static thread_local size_t configured = 0;
    if (smem > configured) {
        SS_CUDA(cudaFuncSetAttribute(conv_tpose_kernel<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)((long long)p.B * p.Din * p.nTH * p.nTW), 1, 1);
    conv_tpose_kernel<BN, F16><<<grid, TP_THREADS, smem, st>>>(p, tmA, tmB);
    return check_launch(F16 ? "conv_tpose_f16x3_kernel" : "conv_tpose_kernel");
}

// returns 1 if the layer was handled here
static bool tpose_eligible(const ss_conv3d_desc* d) {
    if (!d->transposed || d->Cin % 32 != 0 || d->kd != 3 || d->kh != 3 || d->kw != 3) return false;
    if (d->sd != 2 || d->sh != 2 || d->sw != 2 || d->pd != 1 || d->ph != 1 || d->pw != 1) return false;
    if (d->math != SS_MATH_TF32 || (d->cout_packed != 32 && d->cout_packed != 64)) return false;
    if (d->Dout > 2 * d->Din || d->Hout > 2 * d->Hin || d->Wout > 2 * d->Win) return false;
    const int nTH = (d->Hin + TP_TH - 1) / TP_TH, nTW = (d->Win + TP_TW - 1) / TP_TW;
    return (double)d->Hin * d->Win / ((double)nTH * TP_TH * nTW * TP_TW) >= 0.6;
}

int conv_tpose_join_supported(const ss_conv3d_desc* d) { return tpose_eligible(d) ? 1 : 0; }

int try_conv_tpose(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                   const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ss_conv3d_join* join, const ConvPass& ps) {
    if (!d->transposed || d->Cin % 32 != 0 || d->kd != 3 || d->kh != 3 || d->kw != 3) return 0;
    if (d->out_act == SS_ACT_SWISH) return 0;             // Swish epilogues live in the box / pointwise / plane kernels only
    if (d->sd != 2 || d->sh != 2 || d->sw != 2 || d->pd != 1 || d->ph != 1 || d->pw != 1) return 0;
    if (d->math != SS_MATH_TF32 || (d->cout_packed != 32 && d->cout_packed != 64)) return 0;
    if (d->Dout > 2 * d->Din || d->Hout > 2 * d->Hin || d->Wout > 2 * d->Win) return 0;
    const int nTH = (d->Hin + TP_TH - 1) / TP_TH, nTW = (d->Win + TP_TW - 1) / TP_TW;
    const double eff = (double)d->Hin * d->Win / ((double)nTH * TP_TH * nTW * TP_TW);
    if (eff < 0.6) return 0;
    static TEncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) return 0;
        encode = reinterpret_cast<TEncodeTiledFn>(ptr);
    }
    TposeParams p;
    p.B = d->B; p.Din = d->Din; p.Hin = d->Hin; p.Win = d->Win; p.Cin = d->Cin; p.Dout = d->Dout; p.Hout = d->Hout; p.Wout = d->Wout;
    p.Cout = d->Cout; p.CoutP = d->cout_packed; p.out_ldc = d->out_ldc; p.in_act = d->in_act; p.out_act = d->out_act;
    p.nTH = nTH; p.nTW = nTW;
    p.in_scale = in_scale; p.in_shift = in_shift; p.bias = bias; p.y = y; p.stats = stats;
    p.has_join = join ? 1 : 0;
    p.o_scale = join ? join->out_scale : nullptr; p.o_shift = join ? join->out_shift : nullptr;
    p.res = join ? join->res : nullptr; p.r_scale = join ? join->res_scale : nullptr; p.r_shift = join ? join->res_shift : nullptr;
    p.res_ldc = join ? join->res_ldc : 0; p.res_act = join ? join->res_act : 0;
    p.a_lo = ps.a_lo; p.accumulate = ps.accumulate; p.sr = stats_range_of(d); p.acc_scale = ps.acc_scale; p.f16_n = ps.f16_n;
    static const int gather = [] { const char* e = getenv("STEREOSCENE_B200_TPOSE_GATHER"); return (e && e[0] == '0') ? 0 : 1; }();
    p.gather = gather;
    const bool fixup = (in_scale != nullptr) || (d->in_act == SS_ACT_RELU) || ps.a_lo || ps.f16;
    alignas(64) CUtensorMap tmA;
    cuuint64_t gdim[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.Din, (cuuint64_t)p.B};
    cuuint64_t gstr[4] = {(cuuint64_t)d->in_ldc * 4, (cuuint64_t)p.Win * d->in_ldc * 4, (cuuint64_t)p.Hin * p.Win * d->in_ldc * 4,
                          (cuuint64_t)p.Din * p.Hin * p.Win * d->in_ldc * 4};
    cuuint32_t box[5] = {32, TP_HW, TP_HH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (encode(&tmA, fixup ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, const_cast<float*>(x), gdim, gstr, box,
               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { *rc = set_arg_error("conv_tpose: tensor map A"); return 1; }
    if (ps.f16) *rc = (d->cout_packed == 32) ? launch_tpose<32, true>(p, tmA, w_kmajor, encode, st) : launch_tpose<64, true>(p, tmA, w_kmajor, encode, st);
    else *rc = (d->cout_packed == 32) ? launch_tpose<32, false>(p, tmA, w_kmajor, encode, st) : launch_tpose<64, false>(p, tmA, w_kmajor, encode, st);
    return 1;
}

}  // namespace ss
