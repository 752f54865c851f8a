// Implicit-GEMM 3-D / 2-D convolution family on channels-last fp32 volumes.
//
//   y[b, o, :] = act_out( bias + sum_{tap, ci} A(x)[b, in(o, tap), ci] * w[tap, ci, :] )
//   A(x) = act_in(x * in_scale[b, ci] + in_shift[b, ci])        (the producer's pending GroupNorm /
//                                                                BatchNorm / SE gate, applied while
//                                                                the tile is staged: no extra pass)
//   stats[b, co] += (sum y, sum y^2)                              (feeds the next GroupNorm)
//
// One CTA computes a 128(voxels) x BN(output channels) tile; K runs over (tap, 32-channel chunk).
// Tiles are staged global -> registers (pending affine + activation + zero padding applied) ->
// padded shared memory, double buffered; the multiply is TF32 mma.sync (m16n8k8, fp32 accumulate)
// or the error-compensated 3xTF32 split.  Ordinary, strided, dilated and transposed convolutions
// share the kernel: a transposed convolution is decomposed into stride^3 output parity classes,
// each of which is an ordinary gather with its own (short) tap list, so no zero-inserted input
// is ever formed and no multiply is wasted.
//
// Replaces (reference, projects/mmdet3d_plugin/occupancy/): image2bev/ViewTransformerLSSVoxel.py
// :38-58, 66-88, 167-187, 239-241; image2bev/attention.py:93-112; backbones/resnet3d.py:18-32,
// 143-148, 196-198; necks/second_fpn_3d.py:53-59; dense_heads/occhead.py:100-107.
#include "common.cuh"

namespace ss {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int NT = 256;
constexpr int AS_LD = BK + 4;   // 36: a-fragment loads hit 32 distinct banks
constexpr int MAX_TAPS = 64;

struct ConvParams {
    int B, Din, Hin, Win, Cin, Dout, Hout, Wout, Cout, CoutP;
    int kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw;
    int transposed, in_ldc, out_ldc, in_act, out_act;
    int cls_d, cls_h, cls_w;     // parity classes per axis (stride for transposed, else 1)
    const float* x;
    const float* in_scale;
    const float* in_shift;
    const float* w;
    const float* bias;
    float* y;
    double* stats;
};

__host__ __device__ inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int BN>
struct SmemLayout {
    static constexpr int BS_LD = BN + 8;
    static constexpr size_t a_floats = 2 * BM * AS_LD;
    static constexpr size_t b_floats = 2 * BK * BS_LD;
};

template <int BN, int WM, int WN, bool PRECISE, bool VEC>
__global__ void __launch_bounds__(NT, (BN >= 128 || PRECISE) ? 1 : 2)
conv_igemm_kernel(const ConvParams p) {
    constexpr int BS_LD = BN + 8;
    constexpr int MF = BM / WM / 16;
    constexpr int NF = BN / WN / 8;
    static_assert(WM * WN == 8, "8 warps");
    static_assert(MF >= 1 && NF >= 1, "tile");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);                       // [2][BM][AS_LD]
    float* Bs = As + 2 * BM * AS_LD;                                      // [2][BK][BS_LD]
    int4* rowinfo = reinterpret_cast<int4*>(Bs + 2 * BK * BS_LD);         // [BM]
    int4* taps = rowinfo + BM;                                            // [MAX_TAPS]
    double* sstat = reinterpret_cast<double*>(taps + MAX_TAPS);           // [BN][2]
    float* ssc = reinterpret_cast<float*>(sstat + 2 * BN);                // [Cin] scale, [Cin] shift
    __shared__ int s_ntaps;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / WN, wn = warp % WN;

    const int ncls = p.cls_d * p.cls_h * p.cls_w;
    const int b = blockIdx.z / ncls;
    const int cls = blockIdx.z % ncls;
    const int rd = cls / (p.cls_h * p.cls_w), rh = (cls / p.cls_w) % p.cls_h, rw = cls % p.cls_w;
    const int n0 = blockIdx.y * BN;

    // extent of this parity class of the output and the input stride seen by its rows
    const int Dc = (p.Dout - rd + p.cls_d - 1) / p.cls_d;
    const int Hc = (p.Hout - rh + p.cls_h - 1) / p.cls_h;
    const int Wc = (p.Wout - rw + p.cls_w - 1) / p.cls_w;
    const int Mc = Dc * Hc * Wc;
    const int isd = p.transposed ? 1 : p.sd, ish = p.transposed ? 1 : p.sh, isw = p.transposed ? 1 : p.sw;
    const int m_base = blockIdx.x * BM;
    if (m_base >= Mc) return;   // (uniform per CTA)

    // ---- per-CTA tables -------------------------------------------------------------------
    if (tid == 0) {
        int n = 0;
        for (int a = 0; a < p.kd; ++a)
            for (int c = 0; c < p.kh; ++c)
                for (int e = 0; e < p.kw; ++e) {
                    int od, oh, ow;
                    if (p.transposed) {
                        int vd = rd + p.pd - a, vh = rh + p.ph - c, vw = rw + p.pw - e;
                        if (((vd % p.sd) + p.sd) % p.sd || ((vh % p.sh) + p.sh) % p.sh || ((vw % p.sw) + p.sw) % p.sw) continue;
                        od = floordiv(vd, p.sd); oh = floordiv(vh, p.sh); ow = floordiv(vw, p.sw);
                    } else {
                        od = a * p.dd - p.pd; oh = c * p.dh - p.ph; ow = e * p.dw - p.pw;
                    }
                    taps[n++] = make_int4(od, oh, ow, (a * p.kh + c) * p.kw + e);
                }
        s_ntaps = n;
    }
    if (tid < BM) {
        int m = m_base + tid;
        int4 ri = make_int4(0, 0, 0, -1);
        if (m < Mc) {
            int qd = m / (Hc * Wc), rem = m % (Hc * Wc);
            int qh = rem / Wc, qw = rem % Wc;
            int od = qd * p.cls_d + rd, oh = qh * p.cls_h + rh, ow = qw * p.cls_w + rw;
            ri = make_int4(qd * isd, qh * ish, qw * isw, ((b * p.Dout + od) * p.Hout + oh) * p.Wout + ow);
        }
        rowinfo[tid] = ri;
    }
    for (int i = tid; i < 2 * BN; i += NT) sstat[i] = 0.0;
    const bool has_aff = (p.in_scale != nullptr);
    if (has_aff) {
        for (int i = tid; i < p.Cin; i += NT) {
            ssc[i] = __ldg(p.in_scale + (size_t)b * p.Cin + i);
            ssc[p.Cin + i] = __ldg(p.in_shift + (size_t)b * p.Cin + i);
        }
    }
    __syncthreads();
    const int ntaps = s_ntaps;
    const int kchunks = VEC ? (p.Cin / BK) : 1;
    const int ktotal = ntaps * p.Cin;
    const int nsteps = VEC ? ntaps * kchunks : (ktotal + BK - 1) / BK;
    const bool in_relu = (p.in_act == SS_ACT_RELU);

    // ---- global -> register staging ---------------------------------------------------------
    constexpr int A_REGS = 16;
    constexpr int B_F4 = (BK * BN / 4 + NT - 1) / NT;     // float4 per thread for the weight tile
    float areg[A_REGS];
    float4 breg[B_F4];
    uint32_t avalid = 0;

    auto load_tiles = [&](int step) {
        if (VEC) {
            const int tap = step / kchunks, c0 = (step % kchunks) * BK;
            const int4 tp = taps[tap];
            const int col4 = (tid & 7) * 4;
            avalid = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = (tid >> 3) + 32 * j;
                const int4 ri = rowinfo[r];
                const int id = ri.x + tp.x, ih = ri.y + tp.y, iw = ri.z + tp.z;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ri.w >= 0 && (unsigned)id < (unsigned)p.Din && (unsigned)ih < (unsigned)p.Hin &&
                    (unsigned)iw < (unsigned)p.Win) {
                    const size_t vox = ((size_t)(b * p.Din + id) * p.Hin + ih) * p.Win + iw;
                    v = ldg_f4(p.x + vox * p.in_ldc + c0 + col4);
                    avalid |= 1u << j;           // affine / ReLU are applied at the smem store
                }
                areg[4 * j + 0] = v.x; areg[4 * j + 1] = v.y; areg[4 * j + 2] = v.z; areg[4 * j + 3] = v.w;
            }
            const int wrow0 = tp.w * p.Cin + c0;
#pragma unroll
            for (int j = 0; j < B_F4; ++j) {
                const int idx = tid + j * NT;
                const int kk = idx / (BN / 4), nn = (idx % (BN / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < BK * BN / 4 && n0 + nn < p.CoutP)
                    v = ldg_f4(p.w + (size_t)(wrow0 + kk) * p.CoutP + n0 + nn);
                breg[j] = v;
            }
        } else {
            // generic path (tiny Cin): K index = tap * Cin + ci, decoded per element
            const int k = step * BK + (tid & 31);
            const bool kvalid = k < ktotal;
            const int tap = kvalid ? k / p.Cin : 0, ci = kvalid ? k % p.Cin : 0;
            const int4 tp = taps[tap];
            float sc = 1.f, sh = 0.f;
            if (has_aff) { sc = ssc[ci]; sh = ssc[p.Cin + ci]; }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int r = (tid >> 5) + 8 * j;
                const int4 ri = rowinfo[r];
                const int id = ri.x + tp.x, ih = ri.y + tp.y, iw = ri.z + tp.z;
                float v = 0.f;
                if (kvalid && ri.w >= 0 && (unsigned)id < (unsigned)p.Din && (unsigned)ih < (unsigned)p.Hin &&
                    (unsigned)iw < (unsigned)p.Win) {
                    const size_t vox = ((size_t)(b * p.Din + id) * p.Hin + ih) * p.Win + iw;
                    v = __ldg(p.x + vox * p.in_ldc + ci);
                    if (has_aff) v = fmaf(v, sc, sh);
                    if (in_relu) v = fmaxf(v, 0.f);
                }
                areg[j] = v;
            }
#pragma unroll
            for (int j = 0; j < B_F4; ++j) {
                const int idx = tid + j * NT;
                const int kk = idx / (BN / 4), nn = (idx % (BN / 4)) * 4;
                const int kg = step * BK + kk;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < BK * BN / 4 && kg < ktotal && n0 + nn < p.CoutP) {
                    const int wrow = taps[kg / p.Cin].w * p.Cin + kg % p.Cin;
                    v = ldg_f4(p.w + (size_t)wrow * p.CoutP + n0 + nn);
                }
                breg[j] = v;
            }
        }
    };

    auto store_tiles = [&](int buf, int step) {
        float* A = As + buf * BM * AS_LD;
        float* Bt = Bs + buf * BK * BS_LD;
        if (VEC) {
            const int col4 = (tid & 7) * 4;
            const int c0 = (step % kchunks) * BK;
            float4 s = make_float4(1.f, 1.f, 1.f, 1.f), h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_aff) {
                s = *reinterpret_cast<const float4*>(ssc + c0 + col4);
                h = *reinterpret_cast<const float4*>(ssc + p.Cin + c0 + col4);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = (tid >> 3) + 32 * j;
                float4 v = make_float4(areg[4 * j], areg[4 * j + 1], areg[4 * j + 2], areg[4 * j + 3]);
                if (avalid & (1u << j)) {
                    if (has_aff) {
                        v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y);
                        v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
                    }
                    if (in_relu) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                }
                *reinterpret_cast<float4*>(A + r * AS_LD + col4) = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) A[((tid >> 5) + 8 * j) * AS_LD + (tid & 31)] = areg[j];
        }
#pragma unroll
        for (int j = 0; j < B_F4; ++j) {
            const int idx = tid + j * NT;
            const int kk = idx / (BN / 4), nn = (idx % (BN / 4)) * 4;
            if (idx < BK * BN / 4) *reinterpret_cast<float4*>(Bt + kk * BS_LD + nn) = breg[j];
        }
    };

    float acc[MF][NF][4];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    const int moff = wm * (BM / WM), noff = wn * (BN / WN);

    auto compute = [&](int buf) {
        const float* A = As + buf * BM * AS_LD;
        const float* Bt = Bs + buf * BK * BS_LD;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            uint32_t ah[MF][4], bh[NF][2];
            uint32_t al[PRECISE ? MF : 1][4], bl[PRECISE ? NF : 1][2];
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                const float* a0 = A + (moff + i * 16 + g) * AS_LD + ks * 8 + t;
                const float f0 = a0[0], f1 = a0[8 * AS_LD], f2 = a0[4], f3 = a0[8 * AS_LD + 4];
                ah[i][0] = f2tf32(f0); ah[i][1] = f2tf32(f1); ah[i][2] = f2tf32(f2); ah[i][3] = f2tf32(f3);
                if (PRECISE) {
                    al[i][0] = f2tf32(f0 - __uint_as_float(ah[i][0]));
                    al[i][1] = f2tf32(f1 - __uint_as_float(ah[i][1]));
                    al[i][2] = f2tf32(f2 - __uint_as_float(ah[i][2]));
                    al[i][3] = f2tf32(f3 - __uint_as_float(ah[i][3]));
                }
            }
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                const float* b0 = Bt + (ks * 8 + t) * BS_LD + noff + j * 8 + g;
                const float f0 = b0[0], f1 = b0[4 * BS_LD];
                bh[j][0] = f2tf32(f0); bh[j][1] = f2tf32(f1);
                if (PRECISE) {
                    bl[j][0] = f2tf32(f0 - __uint_as_float(bh[j][0]));
                    bl[j][1] = f2tf32(f1 - __uint_as_float(bh[j][1]));
                }
            }
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) {
                    if (PRECISE) {
                        mma_tf32(acc[i][j], al[i], bh[j]);
                        mma_tf32(acc[i][j], ah[i], bl[j]);
                    }
                    mma_tf32(acc[i][j], ah[i], bh[j]);
                }
        }
    };

    // ---- main loop: register-prefetch double buffering ---------------------------------------
    load_tiles(0);
    store_tiles(0, 0);
    __syncthreads();
    for (int step = 0; step < nsteps; ++step) {
        const int buf = step & 1;
        if (step + 1 < nsteps) load_tiles(step + 1);
        compute(buf);
        if (step + 1 < nsteps) store_tiles(buf ^ 1, step + 1);
        __syncthreads();
    }

    // ---- epilogue: bias, activation, store, per-channel sums ---------------------------------
    const bool vec2 = ((p.out_ldc & 1) == 0);
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        const int cl = noff + j * 8 + 2 * t;      // column inside the CTA tile
        const int c = n0 + cl;
        float bias0 = 0.f, bias1 = 0.f;
        if (p.bias) {
            if (c < p.Cout) bias0 = __ldg(p.bias + c);
            if (c + 1 < p.Cout) bias1 = __ldg(p.bias + c + 1);
        }
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < MF; ++i) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = moff + i * 16 + g + 8 * h;
                const int ov = rowinfo[r].w;
                if (ov < 0) continue;
                const float v0 = apply_act_sw(acc[i][j][2 * h] + bias0, p.out_act);
                const float v1 = apply_act_sw(acc[i][j][2 * h + 1] + bias1, p.out_act);
                float* dst = p.y + (size_t)ov * p.out_ldc + c;
                if (c + 1 < p.Cout && vec2) {
                    *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                } else {
                    if (c < p.Cout) dst[0] = v0;
                    if (c + 1 < p.Cout) dst[1] = v1;
                }
                s0 += v0; q0 += v0 * v0; s1 += v1; q1 += v1 * v1;
            }
        }
        if (p.stats) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                q0 += __shfl_xor_sync(0xffffffffu, q0, o);
                q1 += __shfl_xor_sync(0xffffffffu, q1, o);
            }
            if (g == 0) {
                atomicAdd(&sstat[2 * cl + 0], (double)s0);
                atomicAdd(&sstat[2 * cl + 1], (double)q0);
                atomicAdd(&sstat[2 * cl + 2], (double)s1);
                atomicAdd(&sstat[2 * cl + 3], (double)q1);
            }
        }
    }
    if (p.stats) {
        __syncthreads();
        for (int i = tid; i < BN; i += NT) {
            const int c = n0 + i;
            if (c < p.Cout) {
                atomicAdd(p.stats + ((size_t)b * p.Cout + c) * 2 + 0, sstat[2 * i + 0]);
                atomicAdd(p.stats + ((size_t)b * p.Cout + c) * 2 + 1, sstat[2 * i + 1]);
            }
        }
    }
}

template <int BN>
static size_t conv_smem_bytes(int Cin, bool has_aff) {
    size_t fl = 2 * BM * AS_LD + 2 * BK * (BN + 8);
    size_t bytes = fl * sizeof(float) + BM * sizeof(int4) + MAX_TAPS * sizeof(int4) + 2 * BN * sizeof(double);
    if (has_aff) bytes += 2 * (size_t)Cin * sizeof(float);
    return bytes;
}

template <int BN, int WM, int WN, bool PRECISE, bool VEC>
static int launch_conv(const ConvParams& p, cudaStream_t st) {
    const size_t smem = conv_smem_bytes<BN>(p.Cin, p.in_scale != nullptr);
    auto kern = conv_igemm_kernel<BN, WM, WN, PRECISE, VEC>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int ncls = p.cls_d * p.cls_h * p.cls_w;
    // largest parity class (class 0) bounds the tile count; smaller classes exit early
    const int Dc = (p.Dout + p.cls_d - 1) / p.cls_d, Hc = (p.Hout + p.cls_h - 1) / p.cls_h,
              Wc = (p.Wout + p.cls_w - 1) / p.cls_w;
    const long long Mc = (long long)Dc * Hc * Wc;
    dim3 grid((unsigned)((Mc + BM - 1) / BM), (unsigned)((p.CoutP + BN - 1) / BN), (unsigned)(p.B * ncls));
    kern<<<grid, NT, smem, st>>>(p);
    return check_launch("conv_igemm_kernel");
}

template <int BN, int WM, int WN>
static int dispatch_conv(const ConvParams& p, bool precise, bool vec, cudaStream_t st) {
    if (precise) return vec ? launch_conv<BN, WM, WN, true, true>(p, st) : launch_conv<BN, WM, WN, true, false>(p, st);
    return vec ? launch_conv<BN, WM, WN, false, true>(p, st) : launch_conv<BN, WM, WN, false, false>(p, st);
}

}  // namespace ss

namespace ss {
int try_conv_cout1(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                   const float* w_packed, const float* bias, float* y, double* stats, cudaStream_t st, int* rc);
int try_conv_cin_small(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                       const float* w_packed, const float* bias, float* y, double* stats, cudaStream_t st, int* rc);
}

extern "C" int ss_conv3d_fwd(const ss_conv3d_desc* d, const float* x, const float* in_scale, const float* in_shift,
                             const float* w_packed, const float* bias, float* y, double* stats, void* stream) {
    using namespace ss;
    SS_REQUIRE(d && x && w_packed && y, "ss_conv3d_fwd: null pointer");
    SS_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0, "ss_conv3d_fwd: empty shape");
    SS_REQUIRE(d->kd >= 1 && d->kd <= 4 && d->kh >= 1 && d->kh <= 4 && d->kw >= 1 && d->kw <= 4, "ss_conv3d_fwd: kernel extent");
    SS_REQUIRE(d->sd >= 1 && d->sh >= 1 && d->sw >= 1, "ss_conv3d_fwd: stride");
    SS_REQUIRE(d->cout_packed >= d->Cout && d->cout_packed % 8 == 0, "ss_conv3d_fwd: cout_packed must be Cout rounded up to 8");
    SS_REQUIRE(d->in_ldc >= d->Cin && d->out_ldc >= d->Cout, "ss_conv3d_fwd: ldc");
    SS_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "ss_conv3d_fwd: scale/shift must come together");
    SS_REQUIRE(d->in_act == SS_ACT_NONE || d->in_act == SS_ACT_RELU, "ss_conv3d_fwd: in_act");
    SS_REQUIRE(!d->accumulate, "ss_conv3d_fwd: accumulate is offered by ss_conv3d_tc_fwd only");
    SS_REQUIRE(!(in_scale && d->Cin > 4096), "ss_conv3d_fwd: pending affine limited to 4096 channels");
    SS_REQUIRE((long long)d->B * d->Dout * d->Hout * d->Wout < (1ll << 31), "ss_conv3d_fwd: output too large");
    SS_REQUIRE(!stats || d->stats_d1 <= d->stats_d0 || (d->stats_d0 <= 0 && d->stats_d1 >= d->Dout),
               "ss_conv3d_fwd: a restricted statistics plane range is only offered by ss_conv3d_tc_fwd");
    if (d->transposed) {
        SS_REQUIRE(d->dd == 1 && d->dh == 1 && d->dw == 1, "ss_conv3d_fwd: dilated transposed conv unsupported");
        SS_REQUIRE(d->Dout <= (d->Din - 1) * d->sd - 2 * d->pd + d->kd + d->sd - 1, "ss_conv3d_fwd: Dout");
    }
    ConvParams p;
    p.B = d->B; p.Din = d->Din; p.Hin = d->Hin; p.Win = d->Win; p.Cin = d->Cin;
    p.Dout = d->Dout; p.Hout = d->Hout; p.Wout = d->Wout; p.Cout = d->Cout; p.CoutP = d->cout_packed;
    p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.sd = d->sd; p.sh = d->sh; p.sw = d->sw;
    p.pd = d->pd; p.ph = d->ph; p.pw = d->pw; p.dd = d->dd; p.dh = d->dh; p.dw = d->dw;
    p.transposed = d->transposed; p.in_ldc = d->in_ldc; p.out_ldc = d->out_ldc;
    p.in_act = d->in_act; p.out_act = d->out_act;
    p.cls_d = d->transposed ? d->sd : 1; p.cls_h = d->transposed ? d->sh : 1; p.cls_w = d->transposed ? d->sw : 1;
    p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w_packed; p.bias = bias; p.y = y; p.stats = stats;
    SS_REQUIRE((long long)p.B * p.cls_d * p.cls_h * p.cls_w <= 65535, "ss_conv3d_fwd: batch x parity classes > 65535");

    const bool vec = (d->Cin % BK == 0) && (d->in_ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool precise = (d->math == SS_MATH_3XTF32);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {   // single-output-channel layers are bandwidth problems: FMA-pipe kernel (fp32 exact)
        int rc1 = 0;
        if (try_conv_cout1(d, x, in_scale, in_shift, w_packed, bias, y, stats, st, &rc1)) return rc1;
        // 1- or 2-channel inputs (MIE redir1): K = 27*Cin fits one warp-level GEMM with the weights in registers
        if (try_conv_cin_small(d, x, in_scale, in_shift, w_packed, bias, y, stats, st, &rc1)) return rc1;
    }
    const int cp = d->cout_packed;
    if (cp <= 8) return dispatch_conv<8, 8, 1>(p, precise, vec, st);
    if (cp <= 16) return dispatch_conv<16, 8, 1>(p, precise, vec, st);
    if (cp <= 32) return dispatch_conv<32, 8, 1>(p, precise, vec, st);
    if (cp <= 64 || (cp % 128 != 0 && cp % 64 == 0)) return dispatch_conv<64, 4, 2>(p, precise, vec, st);
    return dispatch_conv<128, 2, 4>(p, precise, vec, st);
}
