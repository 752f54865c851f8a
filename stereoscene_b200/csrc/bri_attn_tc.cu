// BRI cross-volume attention on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM, TMA).
//
// Reference: projects/mmdet3d_plugin/occupancy/image2bev/attention.py:58-86 (see bri_attn.cu for the
// mma.sync / split-TF32 kernel that serves SS_MATH_3XTF32 and odd shapes).
//
// Algebra.  Q = wq*q+bq, K = wk*kv+bk, V = wv*kv+bv are scalar affines of the raw volumes, so with
//   S0[i,j] = sum_d q[d,i] kv[d,j]            (raw contraction, the only N^2 x D product)
// the energy is E[i,j] = wq*wk*S0[i,j] + bq*wk*sum_d kv[d,j] + (terms constant along j), and the
// softmax over keys j ignores the row constants.  Likewise
//   out[d,i] = gamma * (wv * sum_j kv[d,j] P'[i,j] + bv * sum_j P'[i,j]) / l[i] + kv[d,i],
//   P'[i,j] = exp(E[i,j]-m[i]) * conf[j],  l[i] = sum_j exp(E[i,j]-m[i]).
// Both contractions therefore run on the RAW q / kv tiles, which TMA drops into shared memory
// (rounded to TF32 by the TMA unit) without any transformation pass.
//
// Layout.  q, kv are [B][D][N] (depth-major, tokens contiguous).  A TMA box {32 tokens, DP depth rows}
// lands as DP rows of 128 B.  Loaded with SWIZZLE_128B_ATOM_32B it is the canonical MN-major tf32
// operand (layout type SWIZZLE_128B_BASE32B, 4 depth rows x 32 tokens per atom -- the only MN-major
// layout tf32 has) for S0 = q^T kv (A = q tile, B = kv tile, K = depth); loaded with SWIZZLE_128B it is
// the canonical K-major B operand (row = depth, 32 keys per 128-byte row) for O += P' kv^T (A = P' from
// tensor memory, N = depth, K = keys).  Pass 2 therefore brings each kv tile in twice (two tensor maps
// over the same global memory, the second copy is an L2 hit); nothing is transposed or re-packed.
//
// Two passes over the keys: pass 1 computes the exact row maximum (QK^T only), pass 2 recomputes
// QK^T, exponentiates against the final maximum and accumulates O in TMEM with NO running rescale
// (the extra QK^T pass costs ~1/3 more tensor work and removes the whole correction path).
// CTA = 128 queries x one key split; warps 0-3: softmax (thread = query row = TMEM lane),
// warp 4: TMA producer, warp 5: MMA issuer.  TMEM: 2 S/P buffers (P' overwrites S in place) + O.
#include <cuda.h>
#include "common.cuh"

namespace ss {

constexpr int BT_Q = 128;       // queries per CTA (UMMA M)
constexpr int BT_K = 64;        // keys per tile
constexpr int BT_NSB = 2;       // S/P buffers in tensor memory
constexpr int BT_SLOTS = 3;     // kv tile ring (a tile stays resident from its load until its PV product)
constexpr int BT_THREADS = 192;
constexpr int BT_TMEM_COLS = 256;
constexpr int BT_O_COL = BT_NSB * BT_K;

struct BriTcParams {
    const float* kv;
    const float* conf;       // [B][NP]  conf[j] (0 in the padding)
    const float* cb2;        // [B][NP]  log2(e)*bq*wk*sum_d kv[d,j]  (-inf in the padding: masks keys >= N)
    const float* params;     // device float[7] = wq,bq,wk,bk,wv,bv,gamma
    float* out;
    int out_ld;
    int B, D, DP, N, NP, keys_per_split;
    float* part_ml;
    float* part_o;
};

__device__ __forceinline__ uint32_t bt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bt_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "BTWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra BTWAIT_DONE;\n\t"
        "bra BTWAIT_LOOP;\n\t"
        "BTWAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bt_tma_3d_elect(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t"
        "}\n" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; descriptors assembled from runtime low / high words
__device__ __forceinline__ void bt_umma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void bt_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void bt_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void bt_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void bt_tmem_st16_nowait(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ float bt_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void bt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// conf[b,j] = max_d softmax_d(q)[d,j];  cb2[b,j] = log2(e)*bq*wk*sum_d kv[d,j];  padding j in [N,NP): 0 / -inf
__global__ void bri_prep_kernel(const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ params,
                                float* __restrict__ conf, float* __restrict__ cb2, int D, int N, int NP) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= NP) return;
    if (j >= N) {
        conf[(size_t)b * NP + j] = 0.f;
        cb2[(size_t)b * NP + j] = -INFINITY;
        return;
    }
    const float* qp = q + (size_t)b * D * N + j;
    const float* kp = kv + (size_t)b * D * N + j;
    float m = -INFINITY, sk = 0.f;
    for (int d = 0; d < D; ++d) {
        m = fmaxf(m, __ldg(qp + (size_t)d * N));
        sk += __ldg(kp + (size_t)d * N);
    }
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(__ldg(qp + (size_t)d * N) - m);
    conf[(size_t)b * NP + j] = 1.0f / s;
    cb2[(size_t)b * NP + j] = 1.4426950408889634f * __ldg(params + 1) * __ldg(params + 2) * sk;
}

__global__ void __launch_bounds__(BT_THREADS, 1)
bri_attn_tc_kernel(const BriTcParams p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKVmn,
                   const __grid_constant__ CUtensorMap tmKV) {
    extern __shared__ unsigned char bt_smem[];
    // 1024-byte alignment as an OFFSET from the extern __shared__ array: pointers derived this way keep the shared state space
    // (ld/st.shared); rounding a uintptr_t instead turns every access through them into a generic load / store
    unsigned char* base = bt_smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(bt_smem) & 1023u)) & 1023u);
    const int DP = p.DP;
    const uint32_t BOXB = (uint32_t)DP * 128u;                   // one [DP depth rows][32 tokens] box
    unsigned char* Qs = base;                                    // 4 boxes: 128 queries
    unsigned char* KVs = base + 4 * BOXB;                        // BT_SLOTS x (2 MN-major + 2 K-major boxes): 64 keys each
    const uint32_t SLOTB = 4 * BOXB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(KVs + (size_t)BT_SLOTS * SLOTB);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * BT_SLOTS + 4 * BT_NSB);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = uniform_warp_index();
    const int b = blockIdx.y, i0 = blockIdx.x * BT_Q;
    const int j_begin = blockIdx.z * p.keys_per_split;
    const int j_end = min(p.N, j_begin + p.keys_per_split);
    const int ntiles = (j_end - j_begin + BT_K - 1) / BT_K;

    const uint32_t q_full = bt_smem_u32(bars), kv_full0 = bt_smem_u32(bars + 1), kv_empty0 = bt_smem_u32(bars + 1 + BT_SLOTS),
                   s_full0 = bt_smem_u32(bars + 1 + 2 * BT_SLOTS), rd_done0 = s_full0 + 8 * BT_NSB, p_ready0 = rd_done0 + 8 * BT_NSB,
                   p_free0 = p_ready0 + 8 * BT_NSB, o_full = p_free0 + 8 * BT_NSB;
    if (tid == 0) {
        bt_mbar_init(q_full, 1);
        for (int s = 0; s < BT_SLOTS; ++s) { bt_mbar_init(kv_full0 + 8 * s, 1); bt_mbar_init(kv_empty0 + 8 * s, 1); }
        for (int s = 0; s < BT_NSB; ++s) {
            bt_mbar_init(s_full0 + 8 * s, 1);
            bt_mbar_init(rd_done0 + 8 * s, 128);
            bt_mbar_init(p_ready0 + 8 * s, 128);
            bt_mbar_init(p_free0 + 8 * s, 1);
        }
        bt_mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKVmn) : "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bt_smem_u32(tmem_slot)), "n"(BT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    bt_fence_before();
    __syncthreads();
    bt_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t Qs_u32 = bt_smem_u32(Qs), KVs_u32 = bt_smem_u32(KVs);

    // number of pass-1 uses of each S buffer (the barrier phase bookkeeping of pass 2 starts after them)
    int n1[BT_NSB];
#pragma unroll
    for (int s = 0; s < BT_NSB; ++s) n1[s] = (ntiles > s) ? (ntiles - s + BT_NSB - 1) / BT_NSB : 0;

    if (warp == 4) {
        // ======================= TMA PRODUCER =====================================================
        mbar_expect_tx_elect(q_full, 4 * BOXB);
#pragma unroll
        for (int m = 0; m < 4; ++m) bt_tma_3d_elect(Qs_u32 + m * BOXB, &tmQ, q_full, i0 + 32 * m, 0, b);
        __syncwarp();
        for (int L = 0; L < 2 * ntiles; ++L) {
            const int slot = L % BT_SLOTS;
            const uint32_t use = (uint32_t)(L / BT_SLOTS);
            bt_mbar_wait(kv_empty0 + 8 * slot, (use & 1u) ^ 1u);
            const uint32_t bar = kv_full0 + 8 * slot;
            const int j0 = j_begin + (L % ntiles) * BT_K;
            const uint32_t dst = KVs_u32 + (uint32_t)slot * SLOTB;
            mbar_expect_tx_elect(bar, L < ntiles ? 2 * BOXB : 4 * BOXB);
            bt_tma_3d_elect(dst, &tmKVmn, bar, j0, 0, b);                     // MN-major copy for S0 = q^T kv
            bt_tma_3d_elect(dst + BOXB, &tmKVmn, bar, j0 + 32, 0, b);
            if (L >= ntiles) {                                                // K-major copy for O += P' kv^T (pass 2 only)
                bt_tma_3d_elect(dst + 2 * BOXB, &tmKV, bar, j0, 0, b);
                bt_tma_3d_elect(dst + 3 * BOXB, &tmKV, bar, j0 + 32, 0, b);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ======================= MMA ISSUER (warp-uniform, elected issue) ==========================
        const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BT_K >> 3) << 17) |
                                ((uint32_t)(BT_Q >> 4) << 24);                     // A, B MN-major
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(DP >> 3) << 17) | ((uint32_t)(BT_Q >> 4) << 24);
        const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);           // K-major: SBO 1024 B, version 1, SWIZZLE_128B
        // MN-major tf32 operands exist only in the 128B-swizzle / 32B-atom layout (4 depth rows per K atom):
        // SBO = next 4 rows = 512 B, layout type SWIZZLE_128B_BASE32B (= TMA's SWIZZLE_128B_ATOM_32B)
        const uint32_t desc_hi_mn = (512u >> 4) | (1u << 14) | (1u << 29);
        const uint32_t lbo_mn = (BOXB >> 4) << 16;                                 // MN-major: next 32 tokens = next box
        const int ksteps = DP / 8;
        bt_mbar_wait(q_full, 0);
        auto issue_qk = [&](int g, int L) {
            const int buf = g % BT_NSB, u = g / BT_NSB, slot = L % BT_SLOTS;
            bt_mbar_wait(kv_full0 + 8 * slot, (uint32_t)(L / BT_SLOTS) & 1u);
            if (u >= 1) {
                if (g - BT_NSB < ntiles) bt_mbar_wait(rd_done0 + 8 * buf, (uint32_t)(u - 1) & 1u);
                else bt_mbar_wait(p_free0 + 8 * buf, (uint32_t)(u - 1 - n1[buf]) & 1u);
            }
            bt_fence_after();
            const uint32_t a0 = ((Qs_u32 >> 4) & 0x3FFFu) | lbo_mn;
            const uint32_t b0 = (((KVs_u32 + (uint32_t)slot * SLOTB) >> 4) & 0x3FFFu) | lbo_mn;
            const uint32_t dcol = tmem_base + (uint32_t)(buf * BT_K);
            for (int ks = 0; ks < ksteps; ++ks)
                bt_umma_ss(dcol, a0 + 64 * ks, desc_hi_mn, b0 + 64 * ks, desc_hi_mn, idesc1, ks ? 1u : 0u);
            umma_commit_elect(s_full0 + 8 * buf);
        };
        auto issue_pv = [&](int t) {
            const int g = ntiles + t, buf = g % BT_NSB, u = g / BT_NSB, slot = g % BT_SLOTS;
            bt_mbar_wait(p_ready0 + 8 * buf, (uint32_t)(u - n1[buf]) & 1u);
            bt_fence_after();
            const uint32_t kvb = KVs_u32 + (uint32_t)slot * SLOTB + 2 * BOXB;
#pragma unroll
            for (int kk = 0; kk < BT_K / 8; ++kk) {
                const uint32_t b_lo = (((kvb + (uint32_t)(kk >> 2) * BOXB) >> 4) & 0x3FFFu) | (1u << 16);
                bt_umma_ts(tmem_base + BT_O_COL, tmem_base + (uint32_t)(buf * BT_K + kk * 8), b_lo + 2 * (kk & 3), desc_hi, idesc2,
                           (t | kk) ? 1u : 0u);
            }
            umma_commit_elect(kv_empty0 + 8 * slot);
            umma_commit_elect(p_free0 + 8 * buf);
        };
        for (int t = 0; t < ntiles; ++t) {              // pass 1: row maxima
            issue_qk(t, t);
            umma_commit_elect(kv_empty0 + 8 * (t % BT_SLOTS));
            __syncwarp();
        }
        for (int t = 0; t < ntiles; ++t) {              // pass 2: S -> P' -> O
            issue_qk(ntiles + t, ntiles + t);
            if (t >= 1) issue_pv(t - 1);
            __syncwarp();
        }
        issue_pv(ntiles - 1);
        umma_commit_elect(o_full);
        __syncwarp();
    } else {
        // ======================= SOFTMAX / EPILOGUE: thread = query row ============================
        const int row = warp * 32 + lane;
        const int i = i0 + row;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float* cbb = p.cb2 + (size_t)b * p.NP;
        const float* cfb = p.conf + (size_t)b * p.NP;
        const float a2 = 1.4426950408889634f * __ldg(p.params + 0) * __ldg(p.params + 2);     // log2(e)*wq*wk
        const float wv = __ldg(p.params + 4), bv = __ldg(p.params + 5), gamma = __ldg(p.params + 6);
        float mx = -INFINITY;
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t % BT_NSB, u = t / BT_NSB;
            bt_mbar_wait(s_full0 + 8 * buf, (uint32_t)u & 1u);
            bt_fence_after();
            const int j0 = j_begin + t * BT_K;
#pragma unroll
            for (int c = 0; c < BT_K / 32; ++c) {
                uint32_t r[32];
                bt_tmem_ld32(lane_addr + (uint32_t)(buf * BT_K + c * 32), r);
                const float4* cb4 = reinterpret_cast<const float4*>(cbb + j0 + c * 32);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 cb = __ldg(cb4 + k4);
                    mx = fmaxf(mx, fmaf(a2, __uint_as_float(r[4 * k4 + 0]), cb.x));
                    mx = fmaxf(mx, fmaf(a2, __uint_as_float(r[4 * k4 + 1]), cb.y));
                    mx = fmaxf(mx, fmaf(a2, __uint_as_float(r[4 * k4 + 2]), cb.z));
                    mx = fmaxf(mx, fmaf(a2, __uint_as_float(r[4 * k4 + 3]), cb.w));
                }
            }
            bt_fence_before();
            bt_mbar_arrive(rd_done0 + 8 * buf);
        }
        float l = 0.f, rs = 0.f;
        for (int t = 0; t < ntiles; ++t) {
            const int g = ntiles + t, buf = g % BT_NSB, u = g / BT_NSB;
            bt_mbar_wait(s_full0 + 8 * buf, (uint32_t)u & 1u);
            bt_fence_after();
            const int j0 = j_begin + t * BT_K;
#pragma unroll
            for (int c = 0; c < BT_K / 32; ++c) {
                uint32_t r[32];
                const uint32_t col = lane_addr + (uint32_t)(buf * BT_K + c * 32);
                bt_tmem_ld32(col, r);
                const float4* cb4 = reinterpret_cast<const float4*>(cbb + j0 + c * 32);
                const float4* cf4 = reinterpret_cast<const float4*>(cfb + j0 + c * 32);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 cb = __ldg(cb4 + k4), cf = __ldg(cf4 + k4);
                    const float cbv[4] = {cb.x, cb.y, cb.z, cb.w}, cfv[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float pe = bt_ex2(fmaf(a2, __uint_as_float(r[4 * k4 + e]), cbv[e]) - mx);
                        l += pe;
                        const float pc = pe * cfv[e];
                        rs += pc;
                        r[4 * k4 + e] = f2tf32(pc);
                    }
                }
                bt_tmem_st16_nowait(col, r);
                bt_tmem_st16_nowait(col + 16, r + 16);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            bt_fence_before();
            bt_mbar_arrive(p_ready0 + 8 * buf);
        }
        // ---- epilogue: out = gamma * (wv*O + bv*rs) / l + kv, or the split's partial state
        bt_mbar_wait(o_full, 0);
        bt_fence_after();
        const bool split = gridDim.z > 1;
        const bool live = i < p.N;
        const size_t pbase = (size_t)blockIdx.z * gridDim.y + b;
        const float inv = 1.0f / l;
        if (split && live) {
            p.part_ml[(pbase * 2 + 0) * p.N + i] = mx * 0.6931471805599453f;       // back to natural-log units
            p.part_ml[(pbase * 2 + 1) * p.N + i] = l;
        }
        const float bvr = bv * rs;
        for (int c = 0; c < DP / 16; ++c) {
            uint32_t r[16];
            bt_tmem_ld16(lane_addr + (uint32_t)(BT_O_COL + c * 16), r);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int d = c * 16 + k;
                if (live && d < p.D) {
                    const float val = fmaf(wv, __uint_as_float(r[k]), bvr);
                    if (split) p.part_o[(pbase * p.D + d) * p.N + i] = val;
                    else {
                        const size_t e = ((size_t)b * p.D + d) * p.N + i;
                        p.out[e * p.out_ld] = fmaf(gamma, val * inv, __ldg(p.kv + e));
                    }
                }
            }
        }
        bt_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        bt_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BT_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*BtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// key splits that fill the 148 SMs best with one 128-query CTA per SM (each split >= 2 key tiles)
int bri_tc_key_splits(int B, int N) {
    const long long qt = (long long)((N + BT_Q - 1) / BT_Q) * B;
    int best = 1;
    double best_eff = 0.0;
    for (int ks = 1; ks <= 8; ks *= 2) {
        const int kps = ((N + ks - 1) / ks + BT_K - 1) / BT_K * BT_K;
        if (ks > 1 && (kps < 2 * BT_K || (long long)(ks - 1) * kps >= N)) break;
        const long long ctas = qt * ks;
        const double eff = (double)ctas / (double)(((ctas + 147) / 148) * 148);
        if (eff > best_eff + 0.05) { best_eff = eff; best = ks; }
    }
    return best;
}

size_t bri_tc_workspace_floats(int B, int D, int N) {
    const int NP = (N + BT_K - 1) / BT_K * BT_K;
    const int ks = bri_tc_key_splits(B, N);
    size_t fl = 2 * (size_t)B * NP;
    if (ks > 1) fl += (size_t)ks * B * N * (2 + D);
    return fl;
}

// returns 1 if the call was handled here (rc holds the result), 0 if the shape needs the mma.sync kernel
void bri_combine_launch(const float* part_ml, const float* part_o, const float* kv, const float* params, float* out, int out_ld,
                        int B, int D, int N, int KS, cudaStream_t st);

int try_bri_tc(const float* q, const float* kv, const float* params, float* ws, float* out, int out_ld, int B, int D, int N,
               cudaStream_t st, int* rc) {
    if (N % 4 != 0 || D > 128 || D < 1) return 0;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(kv)) & 15) return 0;
    static BtEncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) return 0;
        encode = reinterpret_cast<BtEncodeTiledFn>(ptr);
    }
    const int DP = (D + 15) / 16 * 16;
    const int NP = (N + BT_K - 1) / BT_K * BT_K;
    const int KS = bri_tc_key_splits(B, N);
    const int kps = ((N + KS - 1) / KS + BT_K - 1) / BT_K * BT_K;
    const size_t smem = 1024 + (size_t)(4 + 4 * BT_SLOTS) * DP * 128 + (2 + 2 * BT_SLOTS + 4 * BT_NSB) * sizeof(uint64_t) + 16;
    if (smem > 227 * 1024) return 0;                               // D > 112: tiles do not fit, mma.sync kernel
    alignas(64) CUtensorMap tmQ, tmKVmn, tmKV;
    cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)N * 4, (cuuint64_t)D * N * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)DP, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&tmQ, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(q), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        encode(&tmKVmn, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(kv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        encode(&tmKV, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(kv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        *rc = set_arg_error("bri_attn_tc: cuTensorMapEncodeTiled failed");
        return 1;
    }
    float* conf = ws;
    float* cb2 = ws + (size_t)B * NP;
    float* part_ml = cb2 + (size_t)B * NP;
    float* part_o = part_ml + (size_t)KS * B * N * 2;
    dim3 pgrid((NP + 127) / 128, B);
    bri_prep_kernel<<<pgrid, 128, 0, st>>>(q, kv, params, conf, cb2, D, N, NP);
    *rc = check_launch("bri_prep_kernel");
    if (*rc) return 1;
    BriTcParams p;
    p.kv = kv; p.conf = conf; p.cb2 = cb2; p.params = params;
    p.out = out; p.out_ld = out_ld; p.B = B; p.D = D; p.DP = DP; p.N = N; p.NP = NP; p.keys_per_split = kps;
    p.part_ml = part_ml; p.part_o = part_o;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(bri_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { *rc = set_cuda_error(e, "bri_attn_tc: smem attribute"); return 1; }
        configured = smem;
    }
    dim3 grid((N + BT_Q - 1) / BT_Q, B, KS);
    bri_attn_tc_kernel<<<grid, BT_THREADS, smem, st>>>(p, tmQ, tmKVmn, tmKV);
    *rc = check_launch("bri_attn_tc_kernel");
    if (*rc || KS == 1) return 1;
    bri_combine_launch(part_ml, part_o, kv, params, out, out_ld, B, D, N, KS, st);
    *rc = check_launch("bri_combine_kernel");
    return 1;
}

}  // namespace ss
