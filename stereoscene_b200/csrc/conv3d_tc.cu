// tcgen05 + TMA implicit-GEMM convolution for the tensor-bound layers (Cin % 32 == 0, Cout % 4 == 0).
//
// Same contract as conv3d.cu (channels-last fp32 volumes, pending affine + activation of the
// producer layer applied on the way in, bias / activation / GroupNorm sums in the epilogue,
// ordinary / strided / dilated / transposed-by-parity-class gathers), built the Blackwell way:
//   * One CTA owns a 128-voxel OUTPUT BOX (TD x TH x TW voxels, chosen per layer) x BN channels.
//     For every K step (tap, 32-channel chunk) ONE thread issues two TMA tile loads:
//       A: cp.async.bulk.tensor.5d over the NDHWC input (box 32ch x TW x TH x TD x 1, traversal
//          stride = conv stride, start = box origin*stride + tap offset; out-of-bounds = zero fill,
//          which IS the conv zero padding; data type TFLOAT32 so the TMA unit rounds to TF32),
//       B: cp.async.bulk.tensor.2d over the K-major weights [tap*CoutP + n][Cin] (box 32 x BN),
//     both landing in the canonical K-major SWIZZLE_128B layout (one 128-byte row = 32 channels of
//     one voxel / one output channel) and signalling the stage's mbarrier with complete_tx bytes.
//   * If the input carries a pending affine / ReLU (the previous layer's GroupNorm etc.), 8 fix-up
//     warps read the landed A tile from shared memory once, apply the affine in registers (OOB rows
//     stay zero), round to TF32 and write it into a small ring of TENSOR-MEMORY columns
//     (tcgen05.st, lane = voxel row); the MMA then takes A from TMEM and B from shared memory, so
//     the fix-up adds no shared-memory write and removes the MMA's A read.  Plain inputs go
//     TMA -> shared memory -> tensor core with no thread touching the data.
//   * One elected lane issues 4 x tcgen05.mma (kind::tf32, M128 x N(BN) x K8) per stage into a TMEM
//     accumulator (128 lanes x BN fp32 columns) and tcgen05.commit's the stage back to the TMA
//     producer; after the last K step the 8 warps drain TMEM with tcgen05.ld (32 lanes x 32 columns),
//     add bias / activation, store 128-byte rows and reduce per-channel sums with a shuffle butterfly.
//   * STAGES-deep smem ring (A 16 KB + B BN*128 B per stage), full / ready / empty mbarriers.
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace ss {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                  // floats per K step = one 128-byte swizzle row
constexpr int TC_WORKERS = 256;            // 8 fix-up / epilogue warps
constexpr int TC_THREADS = TC_WORKERS + 64;   // + TMA warp + MMA warp
constexpr int TC_MAX_TAPS = 64;

struct TcParams {
    int B, Din, Hin, Win, Cin, Dout, Hout, Wout, Cout, CoutP;
    int kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw;
    int transposed, out_ldc, in_act, out_act;
    int cls_d, cls_h, cls_w;
    int TD, TH, TW;            // output box of one CTA (TD*TH*TW == 128)
    int nTD, nTH, nTW;         // boxes per class grid (sized for the largest class)
    const float* in_scale;
    const float* in_shift;
    const float* bias;
    float* y;
    double* stats;
    int a_lo, accumulate;      // ConvPass (common.cuh)
    float acc_scale;           // F16 variant: accumulator scale (power of two)
    int f16_n;                 // F16 variant: MMAs per K step (6 = compensated, 2 = fp16 single pass)
    int direct_store;          // A/B switch (STEREOSCENE_B200_TC_DIRECT_STORE=1): every lane stores its own row, as before round 2's transposed epilogue
    int ksplit;                // > 1: blockIdx.y = column tile * ksplit + K slice; raw partial tiles go to ws, splitk_reduce_kernel finishes
    float* ws;                 // float[ksplit][B*Dout*Hout*Wout][Cout]
    long long ws_slab;
    StatsRange sr;             // output planes that contribute to stats
};

__host__ __device__ inline int tc_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA tile loads (global -> shared, completion on an mbarrier via complete_tx::bytes)
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 in, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO>>4 in [16,30) (=1 for swizzled K-major), SBO>>4 in [32,46) (8 rows x 128 B = 1024 B),
// version=1 in [46,48), layout_type=2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// cute::UMMA::InstrDescriptor for kind::tf32: c_format F32 (1) at [4,6), a/b format TF32 (2) at
// [7,10)/[10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct TcCfg {
    // The stage count must be EVEN: two fix-up groups alternate K steps, and only with an even ring does the step that used a
    // slot one revolution earlier belong to the same group.  With 5 stages (160- / 192-column tiles until the end of round 2) it
    // belonged to the other group, nothing ordered a group's wait on `full[slot]` after that earlier step's data, and a parity
    // wait cannot tell "one phase behind" from "done": when TMA loads completed out of order under memory contention (a second
    // stream's kernels streaming beside this one), a group converted stale data, arrived on `ready` a phase early, and the
    // producer's next arrive.expect_tx trapped (Warp Illegal Instruction; found with cuda-gdb on tools/overlap_repro.py).
    static constexpr int STAGES = BN >= 160 ? 4 : BN >= 128 ? 6 : 4;
    static_assert(STAGES % 2 == 0, "see above");
    static constexpr int MIN_CTAS = BN <= 64 ? 2 : 1;
    static constexpr int A_BYTES = TC_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int ACC_COLS = (BN + 31) / 32 * 32;            // accumulator columns
    static constexpr int A_COL0 = ACC_COLS;                          // A-operand ring: STAGES x 32 columns
    static constexpr int TMEM_NEED = ACC_COLS + STAGES * TC_BK;
    static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;   // power of two >= need (2 CTAs/SM at 256)
};

// F16 = the single-launch fp16-split compensated variant (SS_MATH_F16X3, common.cuh:split_f16x4): the fix-up warps write the
// A operand into the TMEM ring as packed fp16 pairs [hi(32 ch) | lo(32 ch)] (32 columns, as in TF32 mode) and every K step
// issues six TS-mode kind::f16 MMAs against weight rows packed the same way.
template <int BN, bool F16>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<BN>::MIN_CTAS)
conv_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned stage ring (SWIZZLE_128B atoms are 1024 B)
    // 1024-byte alignment as an OFFSET from the extern __shared__ array: pointers derived this way keep the shared state space
    // (ld/st.shared); rounding a uintptr_t instead turns every access through them into a generic load / store
    unsigned char* ring = smem_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    unsigned char* aux = ring + STAGES * Cfg::STAGE_BYTES;
    int4* taps = reinterpret_cast<int4*>(aux);                             // [TC_MAX_TAPS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(taps + TC_MAX_TAPS) + 2 * BN);   // full[S], ready[S], empty[S], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
    int* s_ntaps = reinterpret_cast<int*>(tmem_slot + 1);
    float* ssc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [Cin] scale, [Cin] shift

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = uniform_warp_index();
    const int ncls = p.cls_d * p.cls_h * p.cls_w;
    const int b = blockIdx.z / ncls, cls = blockIdx.z % ncls;
    const int rd = cls / (p.cls_h * p.cls_w), rh = (cls / p.cls_w) % p.cls_h, rw = cls % p.cls_w;
    const int n0 = (int)(blockIdx.y / p.ksplit) * BN;
    const int Dc = (p.Dout - rd + p.cls_d - 1) / p.cls_d, Hc = (p.Hout - rh + p.cls_h - 1) / p.cls_h,
              Wc = (p.Wout - rw + p.cls_w - 1) / p.cls_w;
    const int isd = p.transposed ? 1 : p.sd, ish = p.transposed ? 1 : p.sh, isw = p.transposed ? 1 : p.sw;
    // output box of this CTA in class-grid coordinates
    const int tw_i = blockIdx.x % p.nTW, th_i = (blockIdx.x / p.nTW) % p.nTH, td_i = blockIdx.x / (p.nTW * p.nTH);
    const int q0d = td_i * p.TD, q0h = th_i * p.TH, q0w = tw_i * p.TW;
    if (q0d >= Dc || q0h >= Hc || q0w >= Wc) return;          // smaller parity classes have fewer boxes (uniform per CTA)

    const uint32_t full0 = smem_u32(bars), ready0 = smem_u32(bars + STAGES), empty0 = smem_u32(bars + 2 * STAGES),
                   accum_bar = smem_u32(bars + 3 * STAGES);
    const bool has_aff = (p.in_scale != nullptr);
    const bool in_relu = (p.in_act == SS_ACT_RELU);
    const bool fixup = F16 || has_aff || in_relu || p.a_lo;

    // ---- per-CTA setup ------------------------------------------------------------------------
    if (tid == 0) {
        int n = 0;
        for (int a = 0; a < p.kd; ++a)
            for (int c = 0; c < p.kh; ++c)
                for (int e = 0; e < p.kw; ++e) {
                    int od, oh, ow;
                    if (p.transposed) {
                        const int vd = rd + p.pd - a, vh = rh + p.ph - c, vw = rw + p.pw - e;
                        if (((vd % p.sd) + p.sd) % p.sd || ((vh % p.sh) + p.sh) % p.sh || ((vw % p.sw) + p.sw) % p.sw) continue;
                        od = tc_floordiv(vd, p.sd); oh = tc_floordiv(vh, p.sh); ow = tc_floordiv(vw, p.sw);
                    } else {
                        od = a * p.dd - p.pd; oh = c * p.dh - p.ph; ow = e * p.dw - p.pw;
                    }
                    taps[n++] = make_int4(od, oh, ow, (a * p.kh + c) * p.kw + e);
                }
        *s_ntaps = n;
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);              // one arrive.expect_tx by the TMA thread (+ tx bytes)
            mbar_init(ready0 + 8 * s, TC_WORKERS / 2);   // one fix-up group (4 warps) arrives per stage
            mbar_init(empty0 + 8 * s, 1);             // released by one tcgen05.commit
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (has_aff)
        for (int i = tid; i < p.Cin; i += TC_THREADS) {
            ssc[i] = __ldg(p.in_scale + (size_t)b * p.Cin + i);
            ssc[p.Cin + i] = __ldg(p.in_shift + (size_t)b * p.Cin + i);
        }
    if (warp == TC_WORKERS / 32 + 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ntaps = *s_ntaps;
    const int kchunks = p.Cin / TC_BK;
    // split-K: this CTA runs K steps [s0, s0 + nsteps) of the layer's ntaps * kchunks (the host makes every slice non-empty)
    const int ks = (int)(blockIdx.y % p.ksplit);
    const int per_slice = (ntaps * kchunks + p.ksplit - 1) / p.ksplit;
    const int s0 = ks * per_slice;
    const int nsteps = min(ntaps * kchunks, s0 + per_slice) - s0;
    const uint32_t ring_u32 = smem_u32(ring);

    if (warp == TC_WORKERS / 32) {
        // ======================= TMA PRODUCER (warp-uniform, elected issue; see common.cuh) =======
        for (int step = 0; step < nsteps; ++step) {
            const int slot = step % STAGES;
            const uint32_t use = (uint32_t)(step / STAGES);
            mbar_wait(empty0 + 8 * slot, (use & 1u) ^ 1u);
            const int tap = (s0 + step) % ntaps, c0 = ((s0 + step) / ntaps) * TC_BK;     // taps innermost: the box stays hot in L2
            const int4 tp = taps[tap];
            const uint32_t a_dst = ring_u32 + slot * Cfg::STAGE_BYTES;
            const uint32_t bar = full0 + 8 * slot;
            mbar_expect_tx_elect(bar, Cfg::STAGE_BYTES);
            tma_5d_elect(a_dst, &tmA, bar, c0, q0w * isw + tp.z, q0h * ish + tp.y, q0d * isd + tp.x, b);
            tma_2d_elect(a_dst + Cfg::A_BYTES, &tmB, bar, c0, tp.w * p.CoutP + n0);
            __syncwarp();
        }
    } else if (warp == TC_WORKERS / 32 + 1) {
        // ======================= MMA ISSUER (warp-uniform, elected issue) =========================
        constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
        constexpr uint32_t D_HI = umma_desc_hi(1024);
        const uint32_t wait0 = fixup ? ready0 : full0;
        for (int step = 0; step < nsteps; ++step) {
            const int slot = step % STAGES;
            const uint32_t use = (uint32_t)(step / STAGES);
            mbar_wait(wait0 + 8 * slot, use & 1u);
            tc_fence_after();
            const uint32_t a_addr = ring_u32 + slot * Cfg::STAGE_BYTES;
            const uint32_t a_lo = umma_desc_lo(a_addr), b_lo = umma_desc_lo(a_addr + Cfg::A_BYTES);
            if constexpr (F16) {
                constexpr uint32_t idesc16 = make_idesc_f16(TC_BM, BN);
                const uint32_t a_tmem = tmem_base + (uint32_t)(Cfg::A_COL0 + slot * TC_BK);
#pragma unroll
                for (int i = 0; i < 6; ++i)      // TMEM columns: 8 per K = 16 fp16 (two per 32-bit cell) -> 4 x the 16-byte descriptor units
                    if (i < p.f16_n)
                        umma_ts_f16<D_HI>(tmem_base, a_tmem + (uint32_t)(4 * kF16A[i]), b_lo + kF16B[i], idesc16, (step | i) ? 1u : 0u);
            } else if (fixup) {                       // A from the TMEM ring written by the fix-up warps
                const uint32_t a_tmem = tmem_base + (uint32_t)(Cfg::A_COL0 + slot * TC_BK);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)
                    umma_ts_tf32<D_HI>(tmem_base, a_tmem + (uint32_t)(8 * k), b_lo + 2 * k, idesc, (step | k) ? 1u : 0u);
            } else {
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)  // 8 TF32 = 32 bytes per MMA: advance start by 32 B
                    umma_ss_tf32<D_HI, D_HI>(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idesc, (step | k) ? 1u : 0u);
            }
            umma_commit_elect(empty0 + 8 * slot);            // frees the stage when these MMAs retire
            if (step == nsteps - 1) umma_commit_elect(accum_bar);
            __syncwarp();
        }
    } else if (fixup) {
        // ======================= FIX-UP WARPS: smem -> registers (affine, ReLU, TF32) -> TMEM ======
        // two groups of 4 warps alternate K steps, so each group has two stage times to hide the
        // barrier wake-up, shared-memory and tcgen05.st latencies
        const int q = warp & 3;                    // TMEM lane quarter of this warp
        const int grp = warp >> 2;                 // handles steps with step % 2 == grp
        const int r = q * 32 + lane;               // voxel row of this thread
        // validity of every tap for this row (bit t = tap t lands inside the input): the zero padding
        // written by the TMA unit must stay zero
        unsigned long long vmask = 0;
        {
            const int iw = (q0w + r % p.TW) * isw, ih = (q0h + (r / p.TW) % p.TH) * ish, id = (q0d + r / (p.TW * p.TH)) * isd;
            for (int t = 0; t < ntaps; ++t) {
                const int4 tp = taps[t];
                const bool ok = (unsigned)(id + tp.x) < (unsigned)p.Din && (unsigned)(ih + tp.y) < (unsigned)p.Hin &&
                                (unsigned)(iw + tp.z) < (unsigned)p.Win;
                vmask |= (unsigned long long)(ok ? 1 : 0) << t;
            }
        }
        uint32_t roff[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) roff[j] = (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)Cfg::A_COL0;
        if constexpr (F16) {
            for (int step = grp; step < nsteps; step += 2) {
                const int slot = step % STAGES;
                const uint32_t use = (uint32_t)(step / STAGES);
                const int tap = (s0 + step) % ntaps, c0 = ((s0 + step) / ntaps) * TC_BK;
                mbar_wait(full0 + 8 * slot, use & 1u);
                const unsigned char* a_src = ring + slot * Cfg::STAGE_BYTES;
                const bool ok = (vmask >> tap) & 1ull;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {       // 16 channels -> 8 packed hi words + 8 packed lo words
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(a_src + roff[hh * 4 + j]);
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 w = v[j];
                        if (has_aff) {
                            const float4 sc = *reinterpret_cast<const float4*>(ssc + c0 + hh * 16 + 4 * j);
                            const float4 sh = *reinterpret_cast<const float4*>(ssc + p.Cin + c0 + hh * 16 + 4 * j);
                            w.x = fmaf(w.x, sc.x, sh.x); w.y = fmaf(w.y, sc.y, sh.y); w.z = fmaf(w.z, sc.z, sh.z); w.w = fmaf(w.w, sc.w, sh.w);
                        }
                        if (in_relu) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
                        uint2 hi, lo;
                        split_f16x4(w, hi, lo);
                        oh[2 * j] = ok ? hi.x : 0u; oh[2 * j + 1] = ok ? hi.y : 0u;
                        ol[2 * j] = ok ? lo.x : 0u; ol[2 * j + 1] = ok ? lo.y : 0u;
                    }
                    tmem_st8_nowait(t_row + (uint32_t)(slot * TC_BK + hh * 8), oh);
                    tmem_st8_nowait(t_row + (uint32_t)(slot * TC_BK + 16 + hh * 8), ol);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(ready0 + 8 * slot);
            }
        } else {
        auto fix = [&](auto lo_tag) {
            constexpr bool LO = decltype(lo_tag)::value;
            for (int step = grp; step < nsteps; step += 2) {
                const int slot = step % STAGES;
                const uint32_t use = (uint32_t)(step / STAGES);
                const int tap = (s0 + step) % ntaps, c0 = ((s0 + step) / ntaps) * TC_BK;
                mbar_wait(full0 + 8 * slot, use & 1u);
                const unsigned char* a_src = ring + slot * Cfg::STAGE_BYTES;
                const bool ok = (vmask >> tap) & 1ull;
    #pragma unroll
                for (int hh = 0; hh < 2; ++hh) {       // two halves of 16 channels: 4 loads in flight, one tcgen05.st each
                    float4 v[4];
    #pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(a_src + roff[hh * 4 + j]);
                    uint32_t o[16];
    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 w = v[j];
                        if (has_aff) {
                            const float4 sc = *reinterpret_cast<const float4*>(ssc + c0 + hh * 16 + 4 * j);     // warp-uniform: broadcast
                            const float4 sh = *reinterpret_cast<const float4*>(ssc + p.Cin + c0 + hh * 16 + 4 * j);
                            w.x = fmaf(w.x, sc.x, sh.x); w.y = fmaf(w.y, sc.y, sh.y); w.z = fmaf(w.z, sc.z, sh.z); w.w = fmaf(w.w, sc.w, sh.w);
                        }
                        if (in_relu) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
                        o[4 * j + 0] = ok ? f2tf32_part<LO>(w.x) : 0u; o[4 * j + 1] = ok ? f2tf32_part<LO>(w.y) : 0u;
                        o[4 * j + 2] = ok ? f2tf32_part<LO>(w.z) : 0u; o[4 * j + 3] = ok ? f2tf32_part<LO>(w.w) : 0u;
                    }
                    tmem_st16_nowait(t_row + (uint32_t)(slot * TC_BK + hh * 16), o);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(ready0 + 8 * slot);
            }
        };
        SS_UNSWITCH_LO(p.a_lo, fix);
        }
    }

    // ======================= EPILOGUE: 8 warps drain TMEM ========================================
    if (warp < TC_WORKERS / 32) {
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = warp >> 2;                    // column half handled by this warp
        // bias of this warp's column chunks, one coalesced load per chunk issued BEFORE the accumulator wait (lane = column;
        // broadcast by shuffle below): 32 scalar loads per chunk after the wait were a serial L2 round trip each -- 1.3 us of
        // every chunk on the image encoder's 1x1 layers
        float bias_l[(BN / 32 + 1) / 2];
#pragma unroll
        for (int i = 0; i < (BN / 32 + 1) / 2; ++i) {
            const int c = n0 + (half * ((BN / 32 + 1) / 2) + i) * 32 + lane;
            bias_l[i] = (p.bias != nullptr && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
        }
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        const int qd = q0d + row / (p.TW * p.TH), qh = q0h + (row / p.TW) % p.TH, qw = q0w + row % p.TW;
        int ov = -1;
        bool st_row = false;                           // this row's output plane contributes to the GroupNorm sums
        if (qd < Dc && qh < Hc && qw < Wc) {
            const int od = qd * p.cls_d + rd, oh = qh * p.cls_h + rh, ow = qw * p.cls_w + rw;
            ov = ((b * p.Dout + od) * p.Hout + oh) * p.Wout + ow;
            st_row = p.sr.has(od);
        }
        constexpr int CHUNKS = BN / 32;
        constexpr int CPH = (CHUNKS + 1) / 2;          // chunks per half
        const bool vec_ok = ((p.out_ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        // column sums through a per-warp 32x33 scratch tile in the (now idle) operand ring
        float* scratch = reinterpret_cast<float*>(ring) + warp * (32 * 33);
        float* part = reinterpret_cast<float*>(ring) + 8 * 32 * 33 + q * (2 * BN);           // per-quarter column sums (no atomics: fixed summation order)
        float* trans = reinterpret_cast<float*>(ring) + 8 * 32 * 33 + 4 * 2 * 256 + warp * (32 * 36);   // store transposition tile (the ring is >= 80 KB)
        const int act = p.out_act;
        const bool has_bias = p.bias != nullptr;
        const bool want_stats = p.stats != nullptr;
#pragma unroll 1
        for (int ci = half * CPH; ci < min(CHUNKS, (half + 1) * CPH); ++ci) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ci * 32), r);
            const int cbase = n0 + ci * 32;
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
            if constexpr (F16) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] *= p.acc_scale;
            }
            if (p.ksplit > 1) {                                // split-K: the raw partial tile; bias / shortcut / activation in splitk_reduce_kernel
                if (ov >= 0) {
                    float* dst = p.ws + (size_t)ks * p.ws_slab + (size_t)ov * p.Cout + cbase;
                    if ((p.Cout & 3) == 0 && cbase + 32 <= p.Cout) {
#pragma unroll
                        for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            if (cbase + k < p.Cout) dst[k] = v[k];
                    }
                }
                continue;
            }
            if (p.accumulate && ov >= 0) {                     // later pass of the compensated mode: add the partial result
                const float* src = p.y + (size_t)ov * p.out_ldc + cbase;
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
                float bl = 0.f;
#pragma unroll
                for (int i = 0; i < CPH; ++i)
                    if (i == ci - half * CPH) bl = bias_l[i];
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] += __shfl_sync(0xffffffffu, bl, k);
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
            if (p.direct_store && vec_ok && cbase + 32 <= p.Cout) {
                if (ov >= 0) {
                    float* dst = p.y + (size_t)ov * p.out_ldc + cbase;
#pragma unroll
                    for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                }
            } else if (vec_ok && cbase + 32 <= p.Cout) {
                // A lane holds 32 columns of ITS row: storing them directly makes every STG.128 touch 32 different rows (32
                // LSU wavefronts per instruction; measured 6 us of a 9 us CTA on the 1x1 layers of the image encoder).  The
                // chunk is transposed through a per-warp 32 x 36 tile instead, so that 8 neighbouring lanes write one row's
                // 128 bytes and an instruction covers 4 full lines.
                float* tr = trans + lane * 36;
#pragma unroll
                for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(tr + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rr = 4 * j + (lane >> 3);
                    const int ovr = __shfl_sync(0xffffffffu, ov, rr);
                    const float4 t4 = *reinterpret_cast<const float4*>(trans + rr * 36 + (lane & 7) * 4);
                    if (ovr >= 0) *reinterpret_cast<float4*>(p.y + (size_t)ovr * p.out_ldc + cbase + (lane & 7) * 4) = t4;
                }
                __syncwarp();
            } else if (ov >= 0) {
                float* dst = p.y + (size_t)ov * p.out_ldc + cbase;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (cbase + k < p.Cout) dst[k] = v[k];
            }
            if (want_stats) {
#pragma unroll
                for (int k = 0; k < 32; ++k) scratch[lane * 33 + k] = st_row ? v[k] : 0.f;
                __syncwarp();
                float cs = 0.f, cq = 0.f;
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    const float x = scratch[rr * 33 + lane];
                    cs += x;
                    cq = fmaf(x, x, cq);
                }
                __syncwarp();
                part[2 * (ci * 32 + lane) + 0] = cs;
                part[2 * (ci * 32 + lane) + 1] = cq;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (p.stats) {
        for (int i = tid; i < BN; i += TC_THREADS) {
            const int c = n0 + i;
            if (c < p.Cout) {
                const float* part0 = reinterpret_cast<const float*>(ring) + 8 * 32 * 33;
                double ts = 0.0, tq = 0.0;
#pragma unroll
                for (int w = 0; w < 4; ++w) { ts += (double)part0[w * 2 * BN + 2 * i + 0]; tq += (double)part0[w * 2 * BN + 2 * i + 1]; }
                atomicAdd(p.stats + ((size_t)b * p.Cout + c) * 2 + 0, ts);
                atomicAdd(p.stats + ((size_t)b * p.Cout + c) * 2 + 1, tq);
            }
        }
    }
    if (warp == TC_WORKERS / 32 + 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

// y[r][c] = act(bias[c] + sum_s ws[s][r][c] (+ y[r][c] when accumulate)): the second half of a split-K layer (fixed summation order)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, long long slab, int S, const float* __restrict__ bias, float* __restrict__ y,
                     long long rows, int Cout, int out_ldc, int accumulate, int act) {
    const int cq = Cout >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cq) return;
    const long long r = i / cq;
    const int c = (int)(i % cq) * 4;
    float4 a = bias ? ldg_f4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
        const float4 t = *reinterpret_cast<const float4*>(ws + (size_t)s * slab + (size_t)r * Cout + c);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    float4* dst = reinterpret_cast<float4*>(y + (size_t)r * out_ldc + c);
    if (accumulate) { const float4 t = *dst; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    a.x = apply_act_sw(a.x, act); a.y = apply_act_sw(a.y, act); a.z = apply_act_sw(a.z, act); a.w = apply_act_sw(a.w, act);
    *dst = a;
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// choose the 128-voxel output box: powers of two, maximise the fraction of useful rows, prefer a long W
static void choose_box(int Dc, int Hc, int Wc, int sd, int sh, int sw, int& TD, int& TH, int& TW) {
    double best = -1.0;
    for (int tw = 128; tw >= 1; tw >>= 1)
        for (int th = 128 / tw; th >= 1; th >>= 1) {
            const int td = 128 / (tw * th);
            if (tw * sw > 256 || th * sh > 256 || td * sd > 256) continue;       // TMA box limit
            auto eff = [](int n, int t) { return (double)n / (double)(((n + t - 1) / t) * t); };
            const double u = eff(Dc, td) * eff(Hc, th) * eff(Wc, tw) + 1e-6 * tw;
            if (u > best) { best = u; TD = td; TH = th; TW = tw; }
        }
}

template <int BN, bool F16 = false>
static int launch_tc(TcParams& p, const float* x, int in_ldc, const float* wk, int ntaps_total, cudaStream_t st, float* ws, size_t ws_bytes) {
    using Cfg = TcCfg<BN>;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return set_arg_error("ss_conv3d_tc_fwd: cuTensorMapEncodeTiled is not available from the driver");
    const int ncls = p.cls_d * p.cls_h * p.cls_w;
    const int Dc = (p.Dout + p.cls_d - 1) / p.cls_d, Hc = (p.Hout + p.cls_h - 1) / p.cls_h, Wc = (p.Wout + p.cls_w - 1) / p.cls_w;
    const int isd = p.transposed ? 1 : p.sd, ish = p.transposed ? 1 : p.sh, isw = p.transposed ? 1 : p.sw;
    choose_box(Dc, Hc, Wc, isd, ish, isw, p.TD, p.TH, p.TW);
    p.nTD = (Dc + p.TD - 1) / p.TD; p.nTH = (Hc + p.TH - 1) / p.TH; p.nTW = (Wc + p.TW - 1) / p.TW;

    alignas(64) CUtensorMap tmA, tmB;
    {   // A: NDHWC input, 5-D (C, W, H, D, B)
        cuuint64_t gdim[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.Din, (cuuint64_t)p.B};
        cuuint64_t gstr[4] = {(cuuint64_t)in_ldc * 4, (cuuint64_t)p.Win * in_ldc * 4, (cuuint64_t)p.Hin * p.Win * in_ldc * 4,
                              (cuuint64_t)p.Din * p.Hin * p.Win * in_ldc * 4};
        cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)(p.TW * isw), (cuuint32_t)(p.TH * ish), (cuuint32_t)(p.TD * isd), 1};
        cuuint32_t estr[5] = {1, (cuuint32_t)isw, (cuuint32_t)ish, (cuuint32_t)isd, 1};
        // plain input: TFLOAT32 (the TMA unit rounds fp32 -> tf32); pending affine: raw FLOAT32, the fix-up
        // warps round once, after the affine
        const bool fixup = F16 || (p.in_scale != nullptr) || (p.in_act == SS_ACT_RELU) || p.a_lo;
        CUresult r = encode(&tmA, fixup ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, const_cast<float*>(x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_arg_error("ss_conv3d_tc_fwd: cuTensorMapEncodeTiled(A) failed");
    }
    {   // B: K-major weights [taps*CoutP][Cin]
        cuuint64_t gdim[2] = {(cuuint64_t)p.Cin, (cuuint64_t)ntaps_total * p.CoutP};
        cuuint64_t gstr[1] = {(cuuint64_t)p.Cin * 4};
        cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wk), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_arg_error("ss_conv3d_tc_fwd: cuTensorMapEncodeTiled(B) failed");
    }
    size_t smem = 1024 + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + TC_MAX_TAPS * sizeof(int4) + 2 * BN * sizeof(double) +
                  (3 * Cfg::STAGES + 1) * sizeof(uint64_t) + 16 + 2 * (size_t)p.Cin * sizeof(float) + 32;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        SS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // split-K (the caller lent a workspace): layers whose tiles fill less than half of the SMs while each walks a long K -- the
    // 1x1 projections of the image encoder's late stages (960 pixels x 2304..3840 channels -> 30..50 CTAs of 72..120 K steps)
    const int ntiles = (p.CoutP + BN - 1) / BN;
    const long long tiles = (long long)p.nTD * p.nTH * p.nTW * ntiles * p.B * ncls;
    const int nsteps_all = ntaps_total * (p.Cin / TC_BK);
    p.ksplit = 1;
    if (ws && ncls == 1 && !p.stats && !p.a_lo && tiles * 2 <= 148 && nsteps_all >= 16 && (p.Cout & 3) == 0 && (p.out_ldc & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 && (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
        const long long rows = (long long)p.B * p.Dout * p.Hout * p.Wout;
        int k = (int)std::min<long long>(std::min<long long>(148 / tiles, nsteps_all / 6), 8);
        while (k > 1 && (size_t)k * rows * p.Cout * sizeof(float) > ws_bytes) --k;
        if (k > 1) {
            const int per = (nsteps_all + k - 1) / k;
            p.ksplit = (nsteps_all + per - 1) / per;
            p.ws = ws;
            p.ws_slab = rows * p.Cout;
        }
    }
    dim3 grid((unsigned)(p.nTD * p.nTH * p.nTW), (unsigned)(ntiles * p.ksplit), (unsigned)(p.B * ncls));
    conv_tc_kernel<BN, F16><<<grid, TC_THREADS, smem, st>>>(p, tmA, tmB);
    int rc = check_launch(F16 ? "conv_tc_f16x3_kernel" : "conv_tc_kernel");
    if (rc != SS_OK || p.ksplit == 1) return rc;
    const long long rows = (long long)p.B * p.Dout * p.Hout * p.Wout, items = rows * (p.Cout / 4);
    splitk_reduce_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(p.ws, p.ws_slab, p.ksplit, p.bias, p.y, rows, p.Cout, p.out_ldc,
                                                                         p.accumulate, p.out_act);
    return check_launch("splitk_reduce_kernel");
}

}  // namespace ss

namespace ss {
int try_conv_march32(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                     const float* w_kmajor, const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ConvPass& ps);
int try_conv_tpose(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                   const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ss_conv3d_join* join, const ConvPass& ps);
int conv_tpose_join_supported(const ss_conv3d_desc* d);
int conv_march32_eligible(const ss_conv3d_desc* d);
int conv_pw_eligible(const ss_conv3d_desc* d);
int try_conv_pw(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ConvPass& ps);
int try_conv_halo(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                  const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ConvPass& ps);

// One TF32 launch of the convolution family: picks the kernel for the layer shape (d->math must be SS_MATH_TF32 here).
static int tc_dispatch(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                       const float* bias, float* y, double* stats, cudaStream_t st, const ConvPass& ps) {
    TcParams p;
    p.B = d->B; p.Din = d->Din; p.Hin = d->Hin; p.Win = d->Win; p.Cin = d->Cin;
    p.Dout = d->Dout; p.Hout = d->Hout; p.Wout = d->Wout; p.Cout = d->Cout; p.CoutP = d->cout_packed;
    p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.sd = d->sd; p.sh = d->sh; p.sw = d->sw;
    p.pd = d->pd; p.ph = d->ph; p.pw = d->pw; p.dd = d->dd; p.dh = d->dh; p.dw = d->dw;
    p.transposed = d->transposed; p.out_ldc = d->out_ldc; p.in_act = d->in_act; p.out_act = d->out_act;
    p.cls_d = d->transposed ? d->sd : 1; p.cls_h = d->transposed ? d->sh : 1; p.cls_w = d->transposed ? d->sw : 1;
    p.in_scale = in_scale; p.in_shift = in_shift; p.bias = bias; p.y = y; p.stats = stats;
    p.a_lo = ps.a_lo; p.accumulate = ps.accumulate; p.acc_scale = ps.acc_scale; p.sr = stats_range_of(d); p.f16_n = ps.f16_n;
    p.ksplit = 1; p.ws = nullptr; p.ws_slab = 0;
    static const int direct_store = [] { const char* e = getenv("STEREOSCENE_B200_TC_DIRECT_STORE"); return (e && e[0] == '1') ? 1 : 0; }();
    p.direct_store = direct_store;
    float* ws = reinterpret_cast<float*>(d->splitk_ws);
    const size_t ws_bytes = d->splitk_ws ? (size_t)d->splitk_ws_bytes : 0;
    SS_REQUIRE((long long)p.B * p.cls_d * p.cls_h * p.cls_w <= 65535, "ss_conv3d_tc_fwd: batch x parity classes > 65535");
    {
        int rcm = 0;
        // 32-channel 3x3x3 stride-1 layers: persistent marching kernel (halo planes + resident weights)
        if (try_conv_march32(d, x, in_scale, in_shift, w_kmajor, bias, y, stats, st, &rcm, ps)) return rcm;
        // pointwise layers with short K: persistent streaming GEMM with resident weights (no fp16 operand path)
        if (!ps.f16 && try_conv_pw(d, x, in_scale, in_shift, w_kmajor, bias, y, stats, st, &rcm, ps)) return rcm;
        // wide 3x3x3 stride-1 layers: halo-resident kernel (planes loaded once per chunk, two M tiles per weight tile)
        if (try_conv_halo(d, x, in_scale, in_shift, w_kmajor, bias, y, stats, st, &rcm, ps)) return rcm;
        // stride-2 transposed 3x3x3 layers: one CTA per input tile computes all 8 output parity classes
        if (try_conv_tpose(d, x, in_scale, in_shift, w_kmajor, bias, y, stats, st, &rcm, nullptr, ps)) return rcm;
    }
    const int cp = d->cout_packed;
    const int ntaps_total = d->kd * d->kh * d->kw;
    int best = 256;
    if (cp <= 32) best = 32;
    else if (cp <= 64) best = 64;
    else if (cp <= 128) best = 128;
    else if (cp <= 192 && cp != 160) best = 192;
    else {  // wide layers: the column tile that minimises (waves of CTAs on 148 SMs) x (work per CTA ~ BN); 256-column tiles
        // leave SMs idle on small grids, 160 columns turn the 300-CTA grid of the 640-channel 2-D layers into 240
        const int ncls = p.cls_d * p.cls_h * p.cls_w;
        const int Dc = (p.Dout + p.cls_d - 1) / p.cls_d, Hc = (p.Hout + p.cls_h - 1) / p.cls_h, Wc = (p.Wout + p.cls_w - 1) / p.cls_w;
        long long mt = (((long long)Dc * Hc * Wc + TC_BM - 1) / TC_BM) * p.B * ncls;
        if (p.Din == 1 && ntaps_total == 1) {            // 2-D pointwise layers: count the boxes the launch will really use (a 12 x 40
            int td, th, tw;                              // image is 3.75 tiles of 128 pixels but 5 boxes)
            choose_box(Dc, Hc, Wc, 1, 1, 1, td, th, tw);
            mt = (long long)((Dc + td - 1) / td) * ((Hc + th - 1) / th) * ((Wc + tw - 1) / tw) * p.B * ncls;
        }
        auto cost = [&](int bn) { const long long ctas = mt * ((cp + bn - 1) / bn); return (double)((ctas + 147) / 148) * bn; };
        double bc = cost(256) * 0.9;
        if (cp % 128 == 0 && cost(128) <= bc * 1.0001 / 0.9) { best = 128; bc = cost(128); }
        if (cp % 160 == 0 && cost(160) < bc) { best = 160; bc = cost(160); }
        if (p.Din == 1 && ntaps_total == 1 && cp % 192 == 0 && cost(192) < bc) { best = 192; bc = cost(192); }
        if (cp % 128 != 0 && cp % 160 != 0)              // 288- and 1344-wide layers of the image encoder: a partial last tile of 160
            for (int bn : {160, 192})                    // or 192 columns wastes less than 256-column tiles
                if (cost(bn) < bc) { best = bn; bc = cost(bn); }
    }
#define SS_TC_LAUNCH(BN_)                                                                                          \
    return ps.f16 ? launch_tc<BN_, true>(p, x, d->in_ldc, w_kmajor, ntaps_total, st, ws, ws_bytes) : launch_tc<BN_, false>(p, x, d->in_ldc, w_kmajor, ntaps_total, st, ws, ws_bytes)
    if (best == 32) SS_TC_LAUNCH(32);
    if (best == 64) SS_TC_LAUNCH(64);
    if (best == 128) SS_TC_LAUNCH(128);
    if (best == 160) SS_TC_LAUNCH(160);
    if (best == 192) SS_TC_LAUNCH(192);
    SS_TC_LAUNCH(256);
#undef SS_TC_LAUNCH
}

// The compensated mode (SS_MATH_TF32X3): three TF32 launches that accumulate into y, smallest terms first.
// w_kmajor holds the hi parts followed by the lo parts (2 x taps x cout_packed x Cin floats).
template <typename F>
static int tc_three_pass(const ss_conv3d_desc* d, const float* w_kmajor, F&& launch) {
    ss_conv3d_desc t = *d;
    t.math = SS_MATH_TF32;
    const size_t wn = (size_t)d->kd * d->kh * d->kw * d->cout_packed * d->Cin;
    const int acc = d->accumulate ? 1 : 0;           // in-place residual: the first (or only) pass adds what y already holds
    if (d->math == SS_MATH_TF32) return launch(&t, w_kmajor, ConvPass{0, acc}, true);
    if (d->math == SS_MATH_F16X3 || d->math == SS_MATH_F16) {      // single launch on fp16 operands (split inside the kernel)
        ConvPass ps{0, acc};
        ps.f16 = 1;
        ps.f16_n = d->math == SS_MATH_F16 ? 2 : 6;
        ps.acc_scale = d->acc_scale;
        return launch(&t, w_kmajor, ps, true);
    }
    ss_conv3d_desc part = t;
    part.out_act = SS_ACT_NONE;
    int rc = launch(&part, w_kmajor, ConvPass{1, acc}, false);               // lo(x) * hi(w)
    if (rc != SS_OK) return rc;
    rc = launch(&part, w_kmajor + wn, ConvPass{0, 1}, false);                // hi(x) * lo(w)
    if (rc != SS_OK) return rc;
    return launch(&t, w_kmajor, ConvPass{0, 1}, true);                       // hi(x) * hi(w) + bias, activation, statistics, join
}
}  // namespace ss

// 1 if the layer runs as ONE launch in the fp16-split compensated mode (halo-resident, marching, transposed or per-tap box
// kernel); layers that the pointwise streaming kernel serves keep the three-launch SS_MATH_TF32X3 path on that kernel.
extern "C" int ss_conv3d_tc_f16x3_supported(const ss_conv3d_desc* d) {
    if (!d || d->Cin % 32 != 0 || d->in_ldc % 4 != 0) return 0;
    ss_conv3d_desc t = *d;
    t.math = SS_MATH_TF32;
    // the pointwise streaming kernel has no fp16 operand path; for the 3-D layers it serves (hourglass redir, neck deblocks) three
    // streaming passes beat the box kernel, for the 2-D ones of the image encoder (depth-1 volumes of 10^5 pixels, N = 32) one
    // box-kernel launch beats three passes that re-read and re-write the output
    if (ss::conv_pw_eligible(&t) && d->Din > 1) return 0;
    return 1;
}

// wk: float[taps][cout_packed][Cin], K-major, values already rounded to TF32.
extern "C" int ss_conv3d_tc_fwd(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                                const float* w_kmajor, const float* bias, float* y, double* stats, void* stream) {
    using namespace ss;
    SS_REQUIRE(d && x && w_kmajor && y, "ss_conv3d_tc_fwd: null pointer");
    SS_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0, "ss_conv3d_tc_fwd: empty shape");
    SS_REQUIRE(d->Cin % TC_BK == 0 && d->in_ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
               "ss_conv3d_tc_fwd: needs Cin % 32 == 0 and 16-byte aligned channels-last input");
    SS_REQUIRE((reinterpret_cast<uintptr_t>(w_kmajor) & 15) == 0, "ss_conv3d_tc_fwd: weights must be 16-byte aligned");
    SS_REQUIRE(d->kd >= 1 && d->kd <= 4 && d->kh >= 1 && d->kh <= 4 && d->kw >= 1 && d->kw <= 4, "ss_conv3d_tc_fwd: kernel extent");
    SS_REQUIRE(d->sd >= 1 && d->sd <= 8 && d->sh >= 1 && d->sh <= 8 && d->sw >= 1 && d->sw <= 8, "ss_conv3d_tc_fwd: stride");
    SS_REQUIRE(d->cout_packed >= d->Cout && d->cout_packed % 8 == 0, "ss_conv3d_tc_fwd: cout_packed");
    SS_REQUIRE(d->in_ldc >= d->Cin && d->out_ldc >= d->Cout, "ss_conv3d_tc_fwd: ldc");
    SS_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "ss_conv3d_tc_fwd: scale/shift must come together");
    SS_REQUIRE(d->in_act == SS_ACT_NONE || d->in_act == SS_ACT_RELU, "ss_conv3d_tc_fwd: in_act");
    SS_REQUIRE(d->Cin <= 4096, "ss_conv3d_tc_fwd: Cin limited to 4096");
    SS_REQUIRE((long long)d->B * d->Dout * d->Hout * d->Wout < (1ll << 31), "ss_conv3d_tc_fwd: output too large");
    SS_REQUIRE(d->math == SS_MATH_TF32 || d->math == SS_MATH_TF32X3 || d->math == SS_MATH_F16X3 || d->math == SS_MATH_F16,
               "ss_conv3d_tc_fwd: TF32 / TF32X3 / F16X3 / F16 only (use ss_conv3d_fwd for 3xTF32 on mma.sync)");
    SS_REQUIRE(!(d->accumulate && stats), "ss_conv3d_tc_fwd: accumulate is not combined with statistics");
    if (d->math == SS_MATH_F16X3 || d->math == SS_MATH_F16) SS_REQUIRE(ss_conv3d_tc_f16x3_supported(d) == 1 && d->acc_scale > 0.f, "ss_conv3d_tc_fwd: F16X3 not offered for this layer (ask ss_conv3d_tc_f16x3_supported)");
    if (d->transposed) SS_REQUIRE(d->dd == 1 && d->dh == 1 && d->dw == 1, "ss_conv3d_tc_fwd: dilated transposed conv unsupported");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return tc_three_pass(d, w_kmajor, [&](const ss_conv3d_desc* dd, const float* w, const ConvPass& ps, bool last) {
        return tc_dispatch(dd, x, in_scale, in_shift, w, last ? bias : nullptr, y, last ? stats : nullptr, st, ps);
    });
}


// Convolution with the residual join fused into its epilogue (hourglass conv5 / conv6 + their joins,
// ViewTransformerLSSVoxel.py:92-95).  Served by the stride-2 transposed kernel; ss_conv3d_tc_join_supported tells the caller
// whether a layer qualifies (otherwise: ss_conv3d_tc_fwd followed by ss_affine_join_fwd).
extern "C" int ss_conv3d_tc_join_supported(const ss_conv3d_desc* d) {
    if (!d || d->Cin % 32 != 0 || d->in_ldc % 4 != 0) return 0;
    ss_conv3d_desc t = *d;
    t.math = SS_MATH_TF32;        // the compensated modes run the same kernel (three times, or once on fp16 halves)
    return ss::conv_tpose_join_supported(&t);
}

extern "C" int ss_conv3d_tc_join_fwd(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                                     const float* w_kmajor, const float* bias, const ss_conv3d_join* join, float* y, void* stream) {
    using namespace ss;
    SS_REQUIRE(d && x && w_kmajor && y && join, "ss_conv3d_tc_join_fwd: null pointer");
    SS_REQUIRE(d->Cin % TC_BK == 0 && d->in_ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(w_kmajor) & 15) == 0, "ss_conv3d_tc_join_fwd: alignment / Cin % 32");
    SS_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "ss_conv3d_tc_join_fwd: scale/shift must come together");
    SS_REQUIRE((join->out_scale == nullptr) == (join->out_shift == nullptr) && (join->res_scale == nullptr) == (join->res_shift == nullptr),
               "ss_conv3d_tc_join_fwd: join scale/shift must come together");
    SS_REQUIRE(!join->res || join->res_ldc >= d->Cout, "ss_conv3d_tc_join_fwd: res_ldc");
    SS_REQUIRE(join->res_act == SS_ACT_NONE || join->res_act == SS_ACT_RELU, "ss_conv3d_tc_join_fwd: res_act");
    SS_REQUIRE(d->in_act == SS_ACT_NONE || d->in_act == SS_ACT_RELU, "ss_conv3d_tc_join_fwd: in_act");
    SS_REQUIRE(d->out_ldc >= d->Cout && d->cout_packed >= d->Cout, "ss_conv3d_tc_join_fwd: ldc");
    SS_REQUIRE(d->math == SS_MATH_TF32 || d->math == SS_MATH_TF32X3 || d->math == SS_MATH_F16X3 || d->math == SS_MATH_F16,
               "ss_conv3d_tc_join_fwd: math mode");
    SS_REQUIRE(ss_conv3d_tc_join_supported(d), "ss_conv3d_tc_join_fwd: layer not supported by the fused kernel (query ss_conv3d_tc_join_supported)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return tc_three_pass(d, w_kmajor, [&](const ss_conv3d_desc* dd, const float* w, const ConvPass& ps, bool last) {
        int rc = 0;
        if (try_conv_tpose(dd, x, in_scale, in_shift, w, last ? bias : nullptr, y, nullptr, st, &rc, last ? join : nullptr, ps)) return rc;
        return set_arg_error("ss_conv3d_tc_join_fwd: layer not supported by the fused kernel");
    });
}
