// Halo-resident tcgen05 convolution for the tensor-bound 3x3x3 stride-1 layers with wide channels
// (3-D encoder blocks, occupancy head, hourglass conv2: resnet3d.py:18-32, occhead.py:100-107,
// ViewTransformerLSSVoxel.py:75-76).
//
// The per-tap box kernel (conv3d_tc.cu) re-fetches a 16 KB A box per tap and is bound by L2->SM
// bandwidth.  Here a CTA owns 256 output voxels (32 h x 8 w at one depth plane d, two M=128 tiles that
// share every weight tile) and, per 32-channel chunk, loads the three input planes d-1, d, d+1 ONCE
// (TMA 5-D box of 34 x 10 halo voxels, out-of-bounds zero fill = conv padding, 42.5 KB) into a 2-slot
// ring.  The A operand of tap (kd,kh,kw) / M-tile mt is the same plane addressed through a UMMA
// descriptor whose start is shifted by ((16 mt + kh) * 10 + kw) rows with stride-byte-offset = the
// halo line pitch (1280 B) -- legal because the 128-byte swizzle is a function of the absolute smem
// address (tools/probes/umma_shift_probe.cu).  Weight tiles (BN x 128 B per tap) stream through their
// own TMA ring and are used by both M-tiles, so L2->SM traffic per MMA drops ~4-6x and the kernel
// becomes tensor-pipe bound.  Pending affine / ReLU: applied once per landed plane, in place, by the 8
// worker warps (padding stays zero).  Accumulators: 2 x BN fp32 columns in TMEM.
#include <cuda.h>
#include "common.cuh"

namespace ss {

constexpr int HL_TH = 32, HL_TW = 8;
constexpr int HL_HH = HL_TH + 2, HL_HW = HL_TW + 2;
constexpr int HL_PLANE_ROWS = HL_HH * HL_HW;          // 340
constexpr int HL_PLANE_BYTES = 43 * 1024;             // 340*128 = 43520 -> padded to a multiple of 1024
// plane ring: 2 slots (a plane lasts 9 taps x 8 MMAs: one slot of prefetch suffices); 3 in the fp16 single-pass mode, whose planes
// are consumed twice as fast
constexpr int HL_WORKERS = 256;
constexpr int HL_THREADS = HL_WORKERS + 96;           // + A producer, MMA, B producer warps

struct HaloParams {
    int B, D, H, W, Cin, Cout, CoutP, out_ldc, in_act, out_act;
    int nTH, nTW;
    int sH, sW, swap;             // voxel strides of the kernel's h / w axes; swap = 1: the kernel's (h,w) are the tensor's (w,h)
    int KD;                       // depth taps: 3 (3x3x3, pad 1) or 1 (2-D 3x3 layers, D == 1 planes)
    int a_lo, accumulate;         // ConvPass (common.cuh)
    float acc_scale;              // F16 variant: accumulator scale (power of two)
    int f16_n;                    // F16 variant: MMAs per chunk (6 = compensated, 2 = fp16 single pass)
    StatsRange sr;                // output planes that contribute to stats
    const float* in_scale;
    const float* in_shift;
    const float* bias;
    float* y;
    double* stats;
};

__device__ __forceinline__ uint32_t h_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void h_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void h_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void h_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void h_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "HWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HWAIT_DONE;\n\t"
        "bra HWAIT_LOOP;\n\t"
        "HWAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void h_tma_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void h_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void h_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void h_umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void h_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ uint64_t h_desc(uint32_t saddr, uint32_t sbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

// (Measured and dropped in round 2: a persistent variant of MODE 2 with double-buffered TMEM accumulators and dedicated epilogue
// warps -- 0.273 vs 0.274 ms on the 128 -> 128 encoder layers: after the fix-up loops were trimmed the kernel streams 1.4 GB of TMA
// loads in 0.24 ms, i.e. it is bound by the L2 -> SM stream of weight tiles every CTA re-fetches, not by prologue / epilogue.  Also
// dropped: a "depth pair" variant in which a CTA owns two output planes (4 M-tiles share every weight tile, 4 input planes instead of
// 2 x 3 per chunk: -44 % L2 -> SM bytes) -- 0.30 vs 0.22 ms on the same layer: with two of three plane slots pinned by the current
// step only one plane prefetches, and 512 double-size CTAs leave the last of 3.5 waves half empty.  Also dropped: two-CTA clusters
// in which each CTA fetches half of every weight tile and multicasts it into both rings (cp.async.bulk.tensor .multicast::cluster,
// slot release through tcgen05.commit .multicast::cluster) -- correct, and no faster (head 0.725 vs 0.715 ms, encoder 0.29 vs 0.27):
// the L2 -> SM stream is not the limit either.  What is: the shared-memory port.  An fp16 MMA of M = 128, N = 128, K = 16 reads
// 4 KB of A and 4 KB of B in its 64 cycles = the port's 128 B/clk, and the TMA writes (115 KB per plane and chunk) and the fix-up
// (64 KB) share that port: 467 KB per plane = 3650 clk against 2304 clk of MMA.  The next step is therefore cta_group::2 (M = 256
// across an SM pair, each SM reading half of B) or fp16 storage (no fix-up pass, half the plane bytes), not more pipelining.)
// MODE 0: TF32.  MODE 1: fp16 hi/lo split, six MMAs per chunk (SS_MATH_F16X3).  MODE 2: fp16 single pass (SS_MATH_F16): only the
// hi halves are multiplied, so the weight tiles are the FIRST 64 bytes of every 128-byte row (TMA box of 16 floats, SWIZZLE_64B
// in shared memory): half the L2->SM bytes and twice as many tiles in flight for the same shared memory, and a third plane slot.
template <int BN, int MODE>
struct HaloCfg {
    static constexpr bool SINGLE = MODE == 2;
    static constexpr int NPL = SINGLE ? 3 : 2;
    // weight-tile ring depth: a tile is consumed in 8 MMAs (~64 BN/128 x 8 cycles), far less than the TMA
    // round trip, so the ring has to hold several microseconds of tiles
    static constexpr int SB = SINGLE ? (BN >= 256 ? 5 : BN >= 192 ? 7 : BN >= 160 ? 8 : BN >= 128 ? 10 : 16)
                                     : (BN >= 256 ? 4 : BN >= 192 ? 5 : BN >= 160 ? 6 : BN >= 128 ? 8 : 12);
    static constexpr int B_BYTES = SINGLE ? BN * 64 : BN * 128;
    static constexpr uint32_t B_HI = SINGLE ? ((512u >> 4) | (1u << 14) | (4u << 29))       // K-major SWIZZLE_64B, 8-row groups of 512 B
                                            : ((1024u >> 4) | (1u << 14) | (2u << 29));
    static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
};

// ---- fix-up of a landed plane by the 256 worker threads ---------------------------------------------------------------------
// Item `it` of a thread is row (tid >> 3) + 32 it, 16-byte chunk tid & 7 (4 channels); row & 7 == (tid >> 3) & 7 for every item, so
// the swizzled read / write offsets inside the row are per-thread constants and only the validity of the rows depends on the tile.
constexpr int HL_FIX_ITERS = (HL_PLANE_ROWS * 8 + HL_WORKERS - 1) / HL_WORKERS;      // 11
struct FixMap { uint32_t rd, wr_hi, wr_lo, vmask; };
__device__ __forceinline__ FixMap make_fixmap(int tid, int h0, int w0, int H, int W) {
    FixMap m;
    const int r0 = tid >> 3, chunk = tid & 7, rsw = r0 & 7;
    m.rd = (uint32_t)(r0 * 128 + ((chunk ^ rsw) << 4));
    m.wr_hi = (uint32_t)(r0 * 128 + (((chunk >> 1) ^ rsw) << 4) + ((chunk & 1) << 3));
    m.wr_lo = (uint32_t)(r0 * 128 + (((4 + (chunk >> 1)) ^ rsw) << 4) + ((chunk & 1) << 3));
    m.vmask = 0;
#pragma unroll
    for (int it = 0; it < HL_FIX_ITERS; ++it) {
        const int r = r0 + 32 * it;
        const int hh = h0 - 1 + r / HL_HW, ww = w0 - 1 + r % HL_HW;
        if (r < HL_PLANE_ROWS && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W) m.vmask |= 1u << it;
    }
    return m;
}
__device__ __forceinline__ float4 fix_value(float4 v, const float4& sc, const float4& sh, bool has_aff, bool in_relu) {
    if (has_aff) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w); }
    if (in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
}
// MODE 0: TF32 (hi, or the lo part when LO) written back in place.  MODE 1 / 2: the row becomes [hi(32 fp16) | lo(32 fp16)] (MODE 2: hi
// only); the 8 lanes of a row read their chunks of a batch of rows, sync, then overwrite -- two batches, two warp syncs per plane.
template <int MODE, bool LO>
__device__ __forceinline__ void fix_plane(unsigned char* pl, const FixMap& m, const float4& sc, const float4& sh, bool has_aff, bool in_relu) {
    if constexpr (MODE == 0) {
#pragma unroll
        for (int it = 0; it < HL_FIX_ITERS; ++it)
            if ((m.vmask >> it) & 1u) {
                float4* ptr = reinterpret_cast<float4*>(pl + m.rd + it * 4096);
                const float4 v = fix_value(*ptr, sc, sh, has_aff, in_relu);
                uint4 o;
                o.x = f2tf32_part<LO>(v.x); o.y = f2tf32_part<LO>(v.y); o.z = f2tf32_part<LO>(v.z); o.w = f2tf32_part<LO>(v.w);
                *reinterpret_cast<uint4*>(ptr) = o;
            }
    } else {
        constexpr int HALF = (HL_FIX_ITERS + 1) / 2;
#pragma unroll
        for (int b0 = 0; b0 < HL_FIX_ITERS; b0 += HALF) {
            uint2 hi[HALF], lo[HALF];
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                const int it = b0 + k;
                if (it < HL_FIX_ITERS && ((m.vmask >> it) & 1u)) {
                    const float4 v = fix_value(*reinterpret_cast<const float4*>(pl + m.rd + it * 4096), sc, sh, has_aff, in_relu);
                    if constexpr (MODE == 2) hi[k] = hi_f16x4(v);
                    else split_f16x4(v, hi[k], lo[k]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                const int it = b0 + k;
                if (it < HL_FIX_ITERS && ((m.vmask >> it) & 1u)) {
                    *reinterpret_cast<uint2*>(pl + m.wr_hi + it * 4096) = hi[k];
                    if constexpr (MODE != 2) *reinterpret_cast<uint2*>(pl + m.wr_lo + it * 4096) = lo[k];
                }
            }
        }
        __syncwarp();
    }
}

// F16 = the single-launch fp16-split compensated variant (SS_MATH_F16X3, see common.cuh:split_f16x4): the workers rewrite
// every landed fp32 plane row in place as [hi | lo] fp16 halves and each (chunk, tap) issues six kind::f16 MMAs.
template <int BN, int MODE>
__global__ void __launch_bounds__(HL_THREADS, 1)
conv_halo_kernel(const HaloParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    using Cfg = HaloCfg<BN, MODE>;
    constexpr bool F16 = MODE != 0;
    constexpr int SB = Cfg::SB;
    constexpr int HL_NPL = Cfg::NPL;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment as an OFFSET from the extern __shared__ array: pointers derived this way keep the shared state space
    // (ld/st.shared); rounding a uintptr_t instead turns every access through them into a generic load / store
    unsigned char* base = smem_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    unsigned char* planes = base;                                            // HL_NPL plane slots
    unsigned char* bring = base + HL_NPL * HL_PLANE_BYTES;                   // SB weight tiles
    unsigned char* aux = bring + SB * Cfg::B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux + 2 * BN * sizeof(double));            // pa_full[3] pa_ready[3] pa_empty[3] pb_full[SB] pb_empty[SB] accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * HL_NPL + 2 * SB + 1);
    float* ssc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // scale[Cin], shift[Cin]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = uniform_warp_index();
    const int n0 = blockIdx.y * BN;
    // tile decode: blockIdx.x = ((b * D + d) * nTH + th) * nTW + tw
    const int tw_i = blockIdx.x % p.nTW, th_i = (blockIdx.x / p.nTW) % p.nTH;
    const int d = (blockIdx.x / (p.nTW * p.nTH)) % p.D, b = blockIdx.x / (p.nTW * p.nTH * p.D);
    const int h0 = th_i * HL_TH, w0 = tw_i * HL_TW;

    const uint32_t pa_full0 = h_smem_u32(bars), pa_ready0 = h_smem_u32(bars + HL_NPL), pa_empty0 = h_smem_u32(bars + 2 * HL_NPL),
                   pb_full0 = h_smem_u32(bars + 3 * HL_NPL), pb_empty0 = h_smem_u32(bars + 3 * HL_NPL + SB),
                   accum_bar = h_smem_u32(bars + 3 * HL_NPL + 2 * SB);
    const bool has_aff = (p.in_scale != nullptr);
    const bool in_relu = (p.in_act == SS_ACT_RELU);
    const bool fixup = F16 || has_aff || in_relu || p.a_lo;
    const int kchunks = p.Cin / 32;
    const int KD = p.KD;

    if (tid == 0) {
        for (int s = 0; s < HL_NPL; ++s) {
            h_mbar_init(pa_full0 + 8 * s, 1);
            h_mbar_init(pa_ready0 + 8 * s, HL_WORKERS);
            h_mbar_init(pa_empty0 + 8 * s, 1);
        }
        for (int s = 0; s < SB; ++s) {
            h_mbar_init(pb_full0 + 8 * s, 1);
            h_mbar_init(pb_empty0 + 8 * s, 1);
        }
        h_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (has_aff)
        for (int i = tid; i < p.Cin; i += HL_THREADS) {
            ssc[i] = __ldg(p.in_scale + (size_t)b * p.Cin + i);
            ssc[p.Cin + i] = __ldg(p.in_shift + (size_t)b * p.Cin + i);
        }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t planes_u32 = h_smem_u32(planes), bring_u32 = h_smem_u32(bring);

    if (warp == 8) {
        // ======================= A PRODUCER: 3 planes per 32-channel chunk (warp-uniform, elected issue) ====
        for (int L = 0; L < kchunks * KD; ++L) {
            const int slot = L % HL_NPL;
            const uint32_t use = (uint32_t)(L / HL_NPL);
            h_mbar_wait(pa_empty0 + 8 * slot, (use & 1u) ^ 1u);
            const uint32_t bar = pa_full0 + 8 * slot;
            mbar_expect_tx_elect(bar, HL_PLANE_ROWS * 128);
            tma_5d_elect(planes_u32 + slot * HL_PLANE_BYTES, &tmA, bar, (L / KD) * 32, w0 - 1, h0 - 1, d - KD / 2 + (L % KD), b);
            __syncwarp();
        }
    } else if (warp == 10) {
        // ======================= B PRODUCER: one weight tile per (chunk, tap) =======================
        const int ntaps = 9 * KD;
        for (int L = 0; L < kchunks * ntaps; ++L) {
            const int slot = L % SB;
            const uint32_t use = (uint32_t)(L / SB);
            h_mbar_wait(pb_empty0 + 8 * slot, (use & 1u) ^ 1u);
            const uint32_t bar = pb_full0 + 8 * slot;
            mbar_expect_tx_elect(bar, Cfg::B_BYTES);
            // weight tap in the tensor's (kd,kh,kw) order; with swapped axes the kernel's (kh,kw) are the tensor's (kw,kh)
            const int tl = L % ntaps, ce = tl % 9;
            const int wt = p.swap ? (tl - ce) + (ce % 3) * 3 + ce / 3 : tl;
            tma_2d_elect(bring_u32 + slot * Cfg::B_BYTES, &tmB, bar, (L / ntaps) * 32, wt * p.CoutP + n0);
            __syncwarp();
        }
    } else if (warp == 9) {
        // ======================= MMA ISSUER (warp-uniform; see common.cuh) ==========================
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t A_HI = umma_desc_hi(HL_HW * 128), B_HI = Cfg::B_HI;
        const uint32_t rdy0 = fixup ? pa_ready0 : pa_full0;
        int Lb = 0;
        for (int Lp = 0; Lp < kchunks * KD; ++Lp) {
            const int pslot = Lp % HL_NPL;
            h_mbar_wait(rdy0 + 8 * pslot, (uint32_t)(Lp / HL_NPL) & 1u);
            const uint32_t a_lo = umma_desc_lo(planes_u32 + pslot * HL_PLANE_BYTES);
#pragma unroll
            for (int ce = 0; ce < 9; ++ce, ++Lb) {
                const int bslot = Lb % SB;
                h_mbar_wait(pb_full0 + 8 * bslot, (uint32_t)(Lb / SB) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_lo = umma_desc_lo(bring_u32 + bslot * Cfg::B_BYTES);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t ao = (uint32_t)(((mt * 16 + ce / 3) * HL_HW + (ce % 3)) * 128) >> 4;
                    if constexpr (F16) {
                        constexpr uint32_t idesc16 = make_idesc_f16(128, BN);
#pragma unroll
                        for (int i = 0; i < (MODE == 2 ? 2 : 6); ++i)
                            umma_ss_f16<A_HI, B_HI>(tmem_base + (uint32_t)(mt * BN), a_lo + ao + kF16A[i], b_lo + kF16B[i], idesc16, (Lp | ce | i) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_ss_tf32<A_HI, B_HI>(tmem_base + (uint32_t)(mt * BN), a_lo + ao + 2 * k, b_lo + 2 * k, idesc, (Lp | ce | k) ? 1u : 0u);
                    }
                }
                umma_commit_elect(pb_empty0 + 8 * bslot);
                if (ce == 8) umma_commit_elect(pa_empty0 + 8 * pslot);
                __syncwarp();
            }
        }
        umma_commit_elect(accum_bar);
        __syncwarp();
    } else if (fixup) {
        // ======================= WORKERS: pending affine / ReLU / operand conversion, once per landed plane, in place =====
        const FixMap fm = make_fixmap(tid, h0, w0, p.H, p.W);
        const int chunk = tid & 7;
        for (int L = 0; L < kchunks * KD; ++L) {
            const int slot = L % HL_NPL;
            h_mbar_wait(pa_full0 + 8 * slot, (uint32_t)(L / HL_NPL) & 1u);
            const int dpl = d - KD / 2 + (L % KD), c0 = (L / KD) * 32;
            if ((unsigned)dpl < (unsigned)p.D) {
                unsigned char* pl = planes + slot * HL_PLANE_BYTES;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_aff) {
                    sc = *reinterpret_cast<const float4*>(ssc + c0 + chunk * 4);
                    sh = *reinterpret_cast<const float4*>(ssc + p.Cin + c0 + chunk * 4);
                }
                if constexpr (F16) {
                    fix_plane<MODE, false>(pl, fm, sc, sh, has_aff, in_relu);
                } else {
                    if (p.a_lo) fix_plane<0, true>(pl, fm, sc, sh, has_aff, in_relu);
                    else fix_plane<0, false>(pl, fm, sc, sh, has_aff, in_relu);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            h_mbar_arrive(pa_ready0 + 8 * slot);
        }
    }

    // ======================= EPILOGUE: 8 warps, warp = (M-tile, lane quarter) =======================
    if (warp < HL_WORKERS / 32) {
        h_mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, mt = warp >> 2;
        const int row = q * 32 + lane;
        const int oh = h0 + mt * 16 + row / HL_TW, ow = w0 + row % HL_TW;
        const bool valid = oh < p.H && ow < p.W;
        const size_t ov = ((size_t)b * p.D + d) * p.H * p.W + (size_t)oh * p.sH + (size_t)ow * p.sW;
        const bool vec_ok = ((p.out_ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        // column sums go through a per-warp 32x33 scratch tile (the plane ring is free once accum_bar fired): one
        // store + one load + two FP ops per value instead of the 5-round shuffle transpose
        // (stride 36: rows stay 16-byte aligned, so the same tile also transposes the chunk for the stores -- a lane holds 32
        // columns of ITS row, and storing them directly makes every STG.128 touch 32 different rows; through the tile 8 neighbouring
        // lanes write one row's 128 bytes and an instruction covers 4 full lines)
        float* scratch = reinterpret_cast<float*>(planes) + warp * (32 * 36);
        float* part = reinterpret_cast<float*>(planes) + 8 * 32 * 36 + warp * (2 * BN);      // per-warp column sums (no atomics: fixed summation order)
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const long long ov_ll = valid ? (long long)ov : -1;
        const int act = p.out_act;
        const bool has_bias = p.bias != nullptr;
        const bool want_stats = p.stats != nullptr && p.sr.has(d);
#pragma unroll 1
        for (int ci = 0; ci < BN / 32; ++ci) {
            uint32_t r[32];
            h_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + ci * 32), r);
            const int cbase = n0 + ci * 32;
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
            if constexpr (F16) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] *= p.acc_scale;
            }
            if (p.accumulate && valid) {                       // later pass of the compensated mode: add the partial result
                const float* src = p.y + ov * p.out_ldc + cbase;
                if (vec_ok && cbase + 32 <= p.Cout) {
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(src + k);
                        v[k] += t4.x; v[k + 1] += t4.y; v[k + 2] += t4.z; v[k + 3] += t4.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (cbase + k < p.Cout) v[k] += src[k];
                }
            }
            if (has_bias) {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (cbase + k < p.Cout) v[k] += __ldg(p.bias + cbase + k);
            }
            if (act == SS_ACT_RELU) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
            } else if (act == SS_ACT_GELU) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = gelu_erf(v[k]);
            } else if (act == SS_ACT_SWISH) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = swish_f(v[k]);
            }
            const bool full = vec_ok && cbase + 32 <= p.Cout;
            if (full || want_stats) {
#pragma unroll
                for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(scratch + lane * 36 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                __syncwarp();
            }
            if (full) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rr = 4 * j + (lane >> 3);
                    const long long ovr = __shfl_sync(0xffffffffu, ov_ll, rr);
                    const float4 t4 = *reinterpret_cast<const float4*>(scratch + rr * 36 + (lane & 7) * 4);
                    if (ovr >= 0) *reinterpret_cast<float4*>(p.y + (size_t)ovr * p.out_ldc + cbase + (lane & 7) * 4) = t4;
                }
            } else if (valid) {
                float* dst = p.y + ov * p.out_ldc + cbase;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (cbase + k < p.Cout) dst[k] = v[k];
            }
            if (want_stats) {
                float cs = 0.f, cq = 0.f;
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    const float x = ((vmask >> rr) & 1u) ? scratch[rr * 36 + lane] : 0.f;
                    cs += x;
                    cq = fmaf(x, x, cq);
                }
                part[2 * (ci * 32 + lane) + 0] = cs;
                part[2 * (ci * 32 + lane) + 1] = cq;
            }
            if (full || want_stats) __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (p.stats && p.sr.has(d)) {
        for (int i = tid; i < BN; i += HL_THREADS) {
            const int c = n0 + i;
            if (c < p.Cout) {
                const float* part0 = reinterpret_cast<const float*>(planes) + 8 * 32 * 36;
                double ts = 0.0, tq = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { ts += (double)part0[w * 2 * BN + 2 * i + 0]; tq += (double)part0[w * 2 * BN + 2 * i + 1]; }
                atomicAdd(p.stats + ((size_t)b * p.Cout + c) * 2 + 0, ts);
                atomicAdd(p.stats + ((size_t)b * p.Cout + c) * 2 + 1, tq);
            }
        }
    }
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

typedef CUresult (*HEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BN, int MODE>
static int launch_halo(const HaloParams& p, const CUtensorMap& tmA, const float* wk, HEncodeTiledFn encode, cudaStream_t st) {
    using Cfg = HaloCfg<BN, MODE>;
    constexpr int HL_NPL = Cfg::NPL;
    alignas(64) CUtensorMap tmB;
    cuuint64_t gdim[2] = {(cuuint64_t)p.Cin, (cuuint64_t)9 * p.KD * p.CoutP};
    cuuint64_t gstr[1] = {(cuuint64_t)p.Cin * 4};
    cuuint32_t box[2] = {Cfg::SINGLE ? 16u : 32u, (cuuint32_t)BN};       // single pass: the hi halves = first 64 bytes of every chunk row
    cuuint32_t estr[2] = {1, 1};
    if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wk), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               Cfg::SINGLE ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return set_arg_error("conv_halo: tensor map B");
    const size_t smem = 1024 + (size_t)HL_NPL * HL_PLANE_BYTES + (size_t)Cfg::SB * Cfg::B_BYTES + 2 * BN * sizeof(double) +
                        (3 * HL_NPL + 2 * Cfg::SB + 1) * sizeof(uint64_t) + 16 + 32 + 2 * (size_t)p.Cin * sizeof(float);
    static thread_local size_t configured = 0;
    if (smem > configured) {
        SS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)((long long)p.B * p.D * p.nTH * p.nTW), (unsigned)((p.CoutP + BN - 1) / BN), 1);
    conv_halo_kernel<BN, MODE><<<grid, HL_THREADS, smem, st>>>(p, tmA, tmB);
    return check_launch(MODE == 2 ? "conv_halo_f16_kernel" : MODE == 1 ? "conv_halo_f16x3_kernel" : "conv_halo_kernel");
}

// returns 1 if the layer was handled here
int try_conv_halo(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                  const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ConvPass& ps) {
    if (d->transposed || d->Cin % 32 != 0 || d->cout_packed < 64 || d->kh != 3 || d->kw != 3) return 0;
    if (!((d->kd == 3 && d->pd == 1) || (d->kd == 1 && d->pd == 0))) return 0;     // 3x3x3, or 2-D 3x3 (one plane per chunk)
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return 0;
    if (d->ph != 1 || d->pw != 1 || d->math != SS_MATH_TF32) return 0;
    if (d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win) return 0;
    // tile orientation: 32 x 8 voxels along (h, w) or, with the axes swapped, along (w, h) -- whichever wastes fewer rows
    auto tiles_of = [](int Hk, int Wk) { return (long long)((Hk + HL_TH - 1) / HL_TH) * ((Wk + HL_TW - 1) / HL_TW); };
    const long long t_norm = tiles_of(d->Hin, d->Win), t_swap = tiles_of(d->Win, d->Hin);
    const int swap = t_swap < t_norm ? 1 : 0;
    const int Hk = swap ? d->Win : d->Hin, Wk = swap ? d->Hin : d->Win;
    const int nTH = (Hk + HL_TH - 1) / HL_TH, nTW = (Wk + HL_TW - 1) / HL_TW;
    const double eff = (double)d->Hin * d->Win / ((double)nTH * HL_TH * nTW * HL_TW);
    if (eff < 0.7) return 0;                                       // too many wasted rows: the box kernel picks a better tile
    if ((long long)d->B * d->Din * nTH * nTW > 0x7fffffffLL) return 0;
    static HEncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) return 0;
        encode = reinterpret_cast<HEncodeTiledFn>(ptr);
    }
    HaloParams p;
    p.B = d->B; p.D = d->Din; p.H = Hk; p.W = Wk; p.Cin = d->Cin; p.Cout = d->Cout; p.CoutP = d->cout_packed;
    p.out_ldc = d->out_ldc; p.in_act = d->in_act; p.out_act = d->out_act; p.nTH = nTH; p.nTW = nTW; p.KD = d->kd;
    p.swap = swap; p.sH = swap ? 1 : d->Win; p.sW = swap ? d->Win : 1;
    p.in_scale = in_scale; p.in_shift = in_shift; p.bias = bias; p.y = y; p.stats = stats;
    p.a_lo = ps.a_lo; p.accumulate = ps.accumulate; p.acc_scale = ps.acc_scale; p.sr = stats_range_of(d); p.f16_n = ps.f16_n;
    const bool fixup = (in_scale != nullptr) || (d->in_act == SS_ACT_RELU) || ps.a_lo || ps.f16;
    alignas(64) CUtensorMap tmA;
    // tensor map dims in the kernel's order (C, w, h, D, B); the byte strides say which tensor axis each one walks
    const cuuint64_t str_w = (cuuint64_t)d->in_ldc * 4, str_h = (cuuint64_t)d->Win * d->in_ldc * 4;
    cuuint64_t gdim[5] = {(cuuint64_t)p.Cin, (cuuint64_t)Wk, (cuuint64_t)Hk, (cuuint64_t)p.D, (cuuint64_t)p.B};
    cuuint64_t gstr[4] = {swap ? str_h : str_w, swap ? str_w : str_h, (cuuint64_t)d->Hin * d->Win * d->in_ldc * 4,
                          (cuuint64_t)p.D * d->Hin * d->Win * d->in_ldc * 4};
    cuuint32_t box[5] = {32, HL_HW, HL_HH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (encode(&tmA, fixup ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, const_cast<float*>(x), gdim, gstr, box,
               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { *rc = set_arg_error("conv_halo: tensor map A"); return 1; }
    const int cp = d->cout_packed;
    const long long tiles = (long long)p.B * p.D * nTH * nTW;
    // column tile: the one that minimises (waves of CTAs on 148 SMs) x (work per CTA ~ BN)
    auto cost = [&](int bn) { const long long ctas = tiles * ((cp + bn - 1) / bn); return (double)((ctas + 147) / 148) * bn; };
    int best = 256;
    if (cp <= 64) best = 64;
    else if (cp <= 128) best = 128;
    else if (cp <= 192 && cp != 160) best = 192;
    else {
        double bc = cost(256) * 0.9;                                // 256-column tiles halve the A traffic per FLOP
        if (cp % 128 == 0 && cost(128) < bc) { best = 128; bc = cost(128); }
        if (cp % 160 == 0 && cost(160) < bc) { best = 160; bc = cost(160); }
    }
#define SS_HALO_LAUNCH(BN_)                                                                        \
    *rc = !ps.f16 ? launch_halo<BN_, 0>(p, tmA, w_kmajor, encode, st)                              \
                  : (ps.f16_n == 2 ? launch_halo<BN_, 2>(p, tmA, w_kmajor, encode, st) : launch_halo<BN_, 1>(p, tmA, w_kmajor, encode, st))
    if (best == 64) SS_HALO_LAUNCH(64);
    else if (best == 128) SS_HALO_LAUNCH(128);
    else if (best == 160) SS_HALO_LAUNCH(160);
    else if (best == 192) SS_HALO_LAUNCH(192);
    else SS_HALO_LAUNCH(256);
#undef SS_HALO_LAUNCH
    return 1;
}

}  // namespace ss
