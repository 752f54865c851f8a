// Pointwise convolution as a persistent streaming GEMM on tcgen05: 1x1x1 stride-1 layers and "k = s" transposed
// convolutions (every input voxel produces s^3 output voxels, one weight matrix per output parity class) --
// the layers whose K is a few 32-channel chunks and whose cost is reading the input and writing the output once:
// hourglass redir2 (64 -> 64, ViewTransformerLSSVoxel.py:88), the encoder's input_proj (128 -> 128,
// resnet3d.py:143-148), the neck's deblocks (ConvTranspose3d k = s = 1 / 2 / 4, second_fpn_3d.py:53-59).
// The per-tap box kernel runs them as thousands of CTAs with 4-16 K steps each, i.e. it measures CTA set-up.
//
// Design (one CTA per SM): the weight matrix of the current (class, column half) GROUP stays RESIDENT in shared
// memory (K-major SWIZZLE_128B, one TMA box per 32-channel chunk, re-loaded only when the CTA's contiguous range of
// work items crosses into the next group); the CTA walks 128-voxel row tiles of the flat [V x Cin] input, whose
// 32-channel chunks ({32 ch, 128 rows} 2-D TMA boxes, 16 KB) stream through a chunk-granular ring -- a slot is
// released as soon as its 4 MMAs retire, so the ring always holds the next chunks whatever the tile size;
// accumulators live in a 4-slot TMEM ring and the epilogue of item t (tcgen05.ld, bias, activation, scatter to the
// class's output voxels, GroupNorm sums) overlaps the loads and MMAs of the following items.
// Pending affine / ReLU of the producer: 4 fix-up warps rewrite each landed chunk in place (TF32-rounded).
#include <cuda.h>
#include "common.cuh"

namespace ss {

constexpr int PW_ACC = 4;                            // TMEM accumulator ring
constexpr int PW_THREADS = 8 * 32 + 64;              // 4 epilogue warps, 4 fix-up / epilogue warps, producer, MMA
constexpr int PW_TILE = 128;                         // voxel rows per tile (UMMA M)
constexpr int PW_MAXRS = 12;                         // chunk ring slots (16 KB each)

struct PwParams {
    long long V;                                     // B * D * H * W input voxels
    long long T;                                     // row tiles = ceil(V / 128)
    long long total_items;                           // groups * T
    int vox_per_batch;                               // D*H*W (a multiple of 128)
    int D, H, W;                                     // input grid (class scatter)
    int s;                                           // 1, or the stride of a k = s transposed conv
    int NH;                                          // column halves per class (1 or 2)
    int Cin, KC, Cout, CoutP, NP, out_ldc, in_act, out_act, RS, scratch_floats;
    const float* in_scale;
    const float* in_shift;
    const float* bias;
    float* y;
    double* stats;
    int a_lo, accumulate;                            // ConvPass (common.cuh)
    StatsRange sr;                                   // output planes that contribute to stats
};

__device__ __forceinline__ uint32_t pw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pw_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void pw_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pw_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PWWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PWWAIT_DONE;\n\t"
        "bra PWWAIT_LOOP;\n\t"
        "PWWAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void pw_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

__global__ void __launch_bounds__(PW_THREADS, 1)
conv_pw_kernel(const PwParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW) {
    extern __shared__ unsigned char pw_smem[];
    // 1024-byte alignment as an OFFSET from the extern __shared__ array: pointers derived this way keep the shared state space
    unsigned char* base = pw_smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(pw_smem) & 1023u)) & 1023u);
    const int KC = p.KC, NP = p.NP, RS = p.RS;
    const uint32_t W_BYTES = (uint32_t)KC * NP * 128;
    constexpr uint32_t CH_BYTES = PW_TILE * 128;
    unsigned char* wres = base;                                   // KC chunks of [NP rows][128 B]
    unsigned char* ring = base + W_BYTES;                         // RS chunks of [128 rows][128 B]
    unsigned char* aux = ring + (size_t)RS * CH_BYTES;
    float* scratch = reinterpret_cast<float*>(aux);               // 8 warps x 32 x 32 floats, XOR-swizzled by float4: store transposition + GroupNorm column sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + p.scratch_floats);
    // barriers: w_full, w_empty, c_full[MAXRS], c_ready[MAXRS], c_empty[MAXRS], t_full[ACC], t_empty[ACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 3 * PW_MAXRS + 2 * PW_ACC);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = uniform_warp_index();
    const uint32_t w_full = pw_smem_u32(bars), w_empty = pw_smem_u32(bars + 1), c_full0 = pw_smem_u32(bars + 2),
                   c_ready0 = pw_smem_u32(bars + 2 + PW_MAXRS), c_empty0 = pw_smem_u32(bars + 2 + 2 * PW_MAXRS),
                   t_full0 = pw_smem_u32(bars + 2 + 3 * PW_MAXRS), t_empty0 = pw_smem_u32(bars + 2 + 3 * PW_MAXRS + PW_ACC);
    const bool has_aff = (p.in_scale != nullptr);
    const bool in_relu = (p.in_act == SS_ACT_RELU);
    const bool fixup = has_aff || in_relu || p.a_lo;
    const int G = NP / 32;                                        // 32-column groups of the accumulator
    const int wgroups = fixup ? 1 : (G >= 2 ? 2 : 1);             // epilogue warp groups (of 4 warps) that drain TMEM
    // this CTA's contiguous range of work items; item = group * T + tile, group = class * NH + column half
    const long long it_begin = p.total_items * blockIdx.x / gridDim.x;
    const long long it_end = p.total_items * (blockIdx.x + 1) / gridDim.x;

    if (tid == 0) {
        pw_mbar_init(w_full, 1);
        pw_mbar_init(w_empty, 1);
        for (int s = 0; s < PW_MAXRS; ++s) {
            pw_mbar_init(c_full0 + 8 * s, 1);
            pw_mbar_init(c_ready0 + 8 * s, 128);
            pw_mbar_init(c_empty0 + 8 * s, 1);
        }
        for (int a = 0; a < PW_ACC; ++a) {
            pw_mbar_init(t_full0 + 8 * a, 1);
            pw_mbar_init(t_empty0 + 8 * a, 128 * wgroups);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(pw_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ring_u32 = pw_smem_u32(ring), wres_u32 = pw_smem_u32(wres);

    if (warp == 8) {
        // ======================= TMA PRODUCER ======================================================
        uint32_t Lc = 0, nload = 0;
        long long cur_group = -1;
        for (long long it = it_begin; it < it_end; ++it) {
            const long long group = it / p.T, t = it - group * p.T;
            if (group != cur_group) {                                  // weights of the next (class, column half)
                cur_group = group;
                if (nload > 0) pw_mbar_wait(w_empty, (nload - 1) & 1u);     // every MMA of the previous group has retired
                const int wrow0 = (int)(group / p.NH) * p.CoutP + (int)(group % p.NH) * NP;
                mbar_expect_tx_elect(w_full, W_BYTES);
                for (int kc = 0; kc < KC; ++kc) tma_2d_elect(wres_u32 + (uint32_t)kc * NP * 128, &tmW, w_full, kc * 32, wrow0);
                ++nload;
                __syncwarp();
            }
            const int row0 = (int)(t * PW_TILE);
            for (int kc = 0; kc < KC; ++kc, ++Lc) {
                const uint32_t slot = Lc % (uint32_t)RS;
                pw_mbar_wait(c_empty0 + 8 * slot, ((Lc / (uint32_t)RS) & 1u) ^ 1u);
                const uint32_t bar = c_full0 + 8 * slot;
                mbar_expect_tx_elect(bar, CH_BYTES);
                tma_2d_elect(ring_u32 + slot * CH_BYTES, &tmA, bar, kc * 32, row0);
                __syncwarp();
            }
        }
    } else if (warp == 9) {
        // ======================= MMA ISSUER (warp-uniform, elected issue) ==========================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t D_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t rdy0 = fixup ? c_ready0 : c_full0;
        const uint32_t b0 = umma_desc_lo(wres_u32);
        uint32_t Lc = 0, L = 0, nload = 0;
        long long cur_group = -1;
        for (long long it = it_begin; it < it_end; ++it, ++L) {
            const long long group = it / p.T;
            if (group != cur_group) {
                if (cur_group >= 0) umma_commit_elect(w_empty);      // fires when the previous group's MMAs are done with the weights
                cur_group = group;
                pw_mbar_wait(w_full, nload & 1u);
                ++nload;
            }
            const uint32_t acc = L % PW_ACC;
            pw_mbar_wait(t_empty0 + 8 * acc, ((L / PW_ACC) & 1u) ^ 1u);
            const uint32_t dcol = tmem_base + acc * (uint32_t)NP;
            for (int kc = 0; kc < KC; ++kc, ++Lc) {
                const uint32_t slot = Lc % (uint32_t)RS;
                pw_mbar_wait(rdy0 + 8 * slot, (Lc / (uint32_t)RS) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ak = umma_desc_lo(ring_u32 + slot * CH_BYTES), bk = b0 + (uint32_t)kc * ((uint32_t)NP * 128 / 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss_tf32<D_HI, D_HI>(dcol, ak + 2 * k, bk + 2 * k, idesc, (kc | k) ? 1u : 0u);
                umma_commit_elect(c_empty0 + 8 * slot);
            }
            umma_commit_elect(t_full0 + 8 * acc);
            __syncwarp();
        }
    } else if (warp >= 4 && fixup) {
        // ======================= FIX-UP WARPS (4..7): pending affine / ReLU, once per landed chunk, in place ===
        const int ft = tid - 128;                      // 0..127
        const int chunk = ft & 7;                      // 16-byte chunk (4 channels) of a 128-byte row
        uint32_t Lc = 0;
        for (long long it = it_begin; it < it_end; ++it) {
            const long long t = it % p.T;
            const long long row0 = t * PW_TILE;
            for (int kc = 0; kc < KC; ++kc, ++Lc) {
                const uint32_t slot = Lc % (uint32_t)RS;
                pw_mbar_wait(c_full0 + 8 * slot, (Lc / (uint32_t)RS) & 1u);
                unsigned char* ch = ring + (size_t)slot * CH_BYTES;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_aff) {                                              // a tile never straddles two samples
                    const size_t bo = (size_t)(row0 / p.vox_per_batch) * p.Cin + kc * 32 + chunk * 4;
                    sc = ldg_f4(p.in_scale + bo);
                    sh = ldg_f4(p.in_shift + bo);
                }
                auto fix = [&](auto lo_tag) {
                    constexpr bool LO = decltype(lo_tag)::value;
                    for (int r = ft >> 3; r < PW_TILE; r += 16) {
                        if (row0 + r < p.V) {                                   // rows past the end stay zero
                            float4* ptr = reinterpret_cast<float4*>(ch + r * 128 + ((chunk ^ (r & 7)) << 4));
                            float4 x = *ptr;
                            if (has_aff) {
                                x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y); x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
                            }
                            if (in_relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                            uint4 o;
                            o.x = f2tf32_part<LO>(x.x); o.y = f2tf32_part<LO>(x.y); o.z = f2tf32_part<LO>(x.z); o.w = f2tf32_part<LO>(x.w);
                            *reinterpret_cast<uint4*>(ptr) = o;
                        }
                    }
                };
                SS_UNSWITCH_LO(p.a_lo, fix);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                pw_mbar_arrive(c_ready0 + 8 * slot);
            }
        }
    }
    if (warp < 8 && !(warp >= 4 && fixup) && (warp >> 2) < wgroups && it_begin < it_end) {
        // ======================= EPILOGUE WARPS: lane quarter q, column groups wg, wg + wgroups, ... ============
        const int q = warp & 3, wg = warp >> 2;
        const int row = q * 32 + lane;
        const bool vec_ok = ((p.out_ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        const bool want_stats = p.stats != nullptr;
        const int act = p.out_act;
        const int s = p.s, Ho = p.H * s, Wo = p.W * s, Do = p.D * s;
        float* sc = scratch + warp * (32 * 32);
        float run_s[4] = {0.f, 0.f, 0.f, 0.f}, run_q[4] = {0.f, 0.f, 0.f, 0.f};     // lane = column: running sums per owned column group
        int run_b = -1, run_n0 = 0;
        auto flush = [&]() {
            if (want_stats && run_b >= 0) {
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    const int c = run_n0 + (wg + gi * wgroups) * 32 + lane;
                    if (wg + gi * wgroups < G && c < p.Cout) {
                        atomicAdd(p.stats + ((size_t)run_b * p.Cout + c) * 2 + 0, (double)run_s[gi]);
                        atomicAdd(p.stats + ((size_t)run_b * p.Cout + c) * 2 + 1, (double)run_q[gi]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { run_s[i] = 0.f; run_q[i] = 0.f; }
        };
        uint32_t L = 0;
        for (long long it = it_begin; it < it_end; ++it, ++L) {
            const long long group = it / p.T, t = it - group * p.T;
            const int cls = (int)(group / p.NH), n0 = (int)(group % p.NH) * NP;          // output parity class, first output column
            const uint32_t acc = L % PW_ACC;
            const long long v = t * PW_TILE + row;
            const int tb = (int)((t * PW_TILE) / p.vox_per_batch);                      // tiles never straddle samples
            if (tb != run_b || n0 != run_n0) { flush(); run_b = tb; run_n0 = n0; }
            const bool valid = v < p.V;
            size_t ov = (size_t)v;                                                       // s == 1: output voxel = input voxel
            bool st_row = valid;                                                         // row contributes to the GroupNorm sums
            if (valid) {
                const int vv = (int)(v - (long long)tb * p.vox_per_batch);
                const int id = vv / (p.W * p.H);
                if (s > 1) {
                    const int iw = vv % p.W, ih = (vv / p.W) % p.H;
                    const int a = cls / (s * s), bb = (cls / s) % s, c = cls % s;
                    ov = (((size_t)tb * Do + (id * s + a)) * Ho + (ih * s + bb)) * Wo + (iw * s + c);
                    st_row = p.sr.has(id * s + a);
                } else {
                    st_row = p.sr.has(id);
                }
            }
            pw_mbar_wait(t_full0 + 8 * acc, (L / PW_ACC) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
                const int g = wg + gi * wgroups;
                if (g >= G) break;
                uint32_t r[32];
                pw_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)NP + (uint32_t)(g * 32), r);
                const int cbase = n0 + g * 32;
                float x[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(r[k]);
                if (p.accumulate && valid) {                   // later pass of the compensated mode: add the partial result
                    const float* src = p.y + ov * p.out_ldc + cbase;
                    if (vec_ok && cbase + 32 <= p.Cout) {
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(src + k);
                            x[k] += t4.x; x[k + 1] += t4.y; x[k + 2] += t4.z; x[k + 3] += t4.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            if (cbase + k < p.Cout) x[k] += src[k];
                    }
                }
                if (p.bias) {
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (cbase + k < p.Cout) x[k] += __ldg(p.bias + cbase + k);
                }
                if (act == SS_ACT_RELU) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k], 0.f);
                } else if (act == SS_ACT_GELU) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) x[k] = gelu_erf(x[k]);
                } else if (act == SS_ACT_SWISH) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) x[k] = swish_f(x[k]);
                }
                // the chunk goes through a 32 x 36 tile: 8 neighbouring lanes then store one row's 128 bytes (a lane storing its own
                // row makes every STG.128 touch 32 rows), and the GroupNorm column sums read the same tile
                const bool full = vec_ok && cbase + 32 <= p.Cout;
                if (full || want_stats) {
#pragma unroll
                    for (int k = 0; k < 32; k += 4)      // float4 k/4 of row `lane` sits at slot (k/4) ^ (lane & 7): every access below is conflict-free
                        *reinterpret_cast<float4*>(sc + lane * 32 + (((k >> 2) ^ (lane & 7)) << 2)) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
                    __syncwarp();
                }
                if (full) {
                    const long long ov_ll = valid ? (long long)ov : -1;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int rr = 4 * j + (lane >> 3);
                        const long long ovr = __shfl_sync(0xffffffffu, ov_ll, rr);
                        const float4 t4 = *reinterpret_cast<const float4*>(sc + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
                        if (ovr >= 0) *reinterpret_cast<float4*>(p.y + (size_t)ovr * p.out_ldc + cbase + (lane & 7) * 4) = t4;
                    }
                } else if (valid) {
                    float* dst = p.y + ov * p.out_ldc + cbase;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (cbase + k < p.Cout) dst[k] = x[k];
                }
                if (want_stats) {
                    const unsigned smask = __ballot_sync(0xffffffffu, st_row);
                    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};      // four independent chains
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        const float e = ((smask >> rr) & 1u) ? sc[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))] : 0.f;
                        cs[rr & 3] += e;
                        cq[rr & 3] = fmaf(e, e, cq[rr & 3]);
                    }
                    run_s[gi] += (cs[0] + cs[1]) + (cs[2] + cs[3]);
                    run_q[gi] += (cq[0] + cq[1]) + (cq[2] + cq[3]);
                }
                if (full || want_stats) __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            pw_mbar_arrive(t_empty0 + 8 * acc);                     // accumulator may be overwritten
        }
        flush();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

typedef CUresult (*PwEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// shape test of try_conv_pw (without the pending-input / shared-memory refinements): used to route the compensated modes
int conv_pw_eligible(const ss_conv3d_desc* d) {
    if (d->math != SS_MATH_TF32 || d->pd != 0 || d->ph != 0 || d->pw != 0 || d->dd != 1 || d->dh != 1 || d->dw != 1) return 0;
    int s = 1;
    if (d->kd == 1 && d->kh == 1 && d->kw == 1 && d->sd == 1 && d->sh == 1 && d->sw == 1) {
        if (d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win) return 0;
    } else if (d->transposed && d->kd == d->sd && d->kh == d->sh && d->kw == d->sw && d->sd == d->sh && d->sh == d->sw &&
               (d->sd == 2 || d->sd == 4)) {
        s = d->sd;
        if (d->Dout != d->Din * s || d->Hout != d->Hin * s || d->Wout != d->Win * s) return 0;
    } else {
        return 0;
    }
    if (d->Cin % 32 != 0 || d->Cin > 64 || d->cout_packed > 128) return 0;      // Cin > 64 with a forced fix-up goes to the box kernel
    const long long vpb = (long long)d->Din * d->Hin * d->Win, V = vpb * d->B;
    if (vpb % PW_TILE != 0 || V >= (1ll << 31) || V * s * s * s < 148LL * PW_TILE * 2) return 0;
    if (s > 1 && d->cout_packed % 32 != 0) return 0;
    return 1;
}

// returns 1 if the layer was handled here
int try_conv_pw(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift, const float* w_kmajor,
                const float* bias, float* y, double* stats, cudaStream_t st, int* rc, const ConvPass& ps) {
    if (d->math != SS_MATH_TF32 || d->pd != 0 || d->ph != 0 || d->pw != 0 || d->dd != 1 || d->dh != 1 || d->dw != 1) return 0;
    int s = 1;
    if (d->kd == 1 && d->kh == 1 && d->kw == 1 && d->sd == 1 && d->sh == 1 && d->sw == 1) {       // also ConvTranspose k = s = 1
        if (d->Dout != d->Din || d->Hout != d->Hin || d->Wout != d->Win) return 0;
    } else if (d->transposed && d->kd == d->sd && d->kh == d->sh && d->kw == d->sw && d->sd == d->sh && d->sh == d->sw &&
               (d->sd == 2 || d->sd == 4)) {                                                        // non-overlapping up-convolution
        s = d->sd;
        if (d->Dout != d->Din * s || d->Hout != d->Hin * s || d->Wout != d->Win * s) return 0;
    } else {
        return 0;
    }
    if (d->Cin % 32 != 0 || d->Cin > 512 || d->cout_packed > 128) return 0;
    const bool pending = (in_scale != nullptr) || (d->in_act == SS_ACT_RELU) || ps.a_lo;
    if (pending && d->Cin > 64) return 0;            // 4 fix-up warps cannot keep up with wide pending inputs: the box kernel's TMEM fix-up path wins
    const long long vpb = (long long)d->Din * d->Hin * d->Win, V = vpb * d->B;
    const int ncls = s * s * s;
    if (vpb % PW_TILE != 0 || V >= (1ll << 31) || V * ncls < 148LL * PW_TILE * 2) return 0;      // small volumes: the box kernel is fine
    if (s > 1 && d->cout_packed % 32 != 0) return 0;
    const int KC = d->Cin / 32;
    int NP = (d->cout_packed + 31) / 32 * 32, NH = 1;
    const int scratch_floats = 8 * 32 * 32;
    auto fixed_bytes = [&](int np) { return (size_t)1024 + (size_t)KC * np * 128 + (size_t)scratch_floats * sizeof(float) +
                                            (2 + 3 * PW_MAXRS + 2 * PW_ACC) * sizeof(uint64_t) + 64; };
    const size_t min_ring = (size_t)(KC < 4 ? 4 : (KC > 6 ? 4 : KC)) * PW_TILE * 128;              // at least 4 chunks in flight
    if (fixed_bytes(NP) + min_ring > 227 * 1024) {                                                  // weights too big: two column halves
        if (NP % 64 != 0) return 0;
        NP /= 2; NH = 2;
        if (fixed_bytes(NP) + min_ring > 227 * 1024) return 0;
    }
    int RS = (int)((227 * 1024 - fixed_bytes(NP)) / (PW_TILE * 128));
    if (RS > PW_MAXRS) RS = PW_MAXRS;
    if (RS < 2) return 0;
    static PwEncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) return 0;
        encode = reinterpret_cast<PwEncodeTiledFn>(ptr);
    }
    alignas(64) CUtensorMap tmA, tmW;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)d->Cin, (cuuint64_t)V};
        cuuint64_t gstr[1] = {(cuuint64_t)d->in_ldc * 4};
        cuuint32_t box[2] = {32, PW_TILE};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&tmA, pending ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(x), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { *rc = set_arg_error("conv_pw: tensor map A"); return 1; }
    }
    {   // K-major weights [taps * CoutP rows][Cin]: class c of a k = s transposed conv is tap c
        cuuint64_t gdim[2] = {(cuuint64_t)d->Cin, (cuuint64_t)ncls * d->cout_packed};
        cuuint64_t gstr[1] = {(cuuint64_t)d->Cin * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)NP};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w_kmajor), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { *rc = set_arg_error("conv_pw: tensor map W"); return 1; }
    }
    PwParams p;
    p.V = V; p.T = (V + PW_TILE - 1) / PW_TILE; p.total_items = p.T * ncls * NH; p.vox_per_batch = (int)vpb;
    p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.s = s; p.NH = NH;
    p.Cin = d->Cin; p.KC = KC; p.Cout = d->Cout; p.CoutP = d->cout_packed; p.NP = NP; p.out_ldc = d->out_ldc;
    p.in_act = d->in_act; p.out_act = d->out_act; p.RS = RS; p.scratch_floats = scratch_floats;
    p.in_scale = in_scale; p.in_shift = in_shift; p.bias = bias; p.y = y; p.stats = stats;
    p.a_lo = ps.a_lo; p.accumulate = ps.accumulate; p.sr = stats_range_of(d);
    const size_t smem = fixed_bytes(NP) + (size_t)RS * PW_TILE * 128;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { *rc = set_cuda_error(e, "conv_pw: smem attribute"); return 1; }
        configured = smem;
    }
    const unsigned grid = (unsigned)(p.total_items < 148 ? p.total_items : 148);
    conv_pw_kernel<<<grid, PW_THREADS, smem, st>>>(p, tmA, tmW);
    *rc = check_launch("conv_pw_kernel");
    return 1;
}

}  // namespace ss
