// BRI: confidence-weighted cross-volume attention, flash style.
//
// Reference: projects/mmdet3d_plugin/occupancy/image2bev/attention.py:58-86.  Tokens are the
// N = H*W pixels, the feature axis is the depth axis D (q, kv: [B,1,D,H,W]).  The reference
// materialises energy = Q^T K and attention = softmax(energy) as two [N,N] fp32 matrices (236 MB
// each at N = 7680) and multiplies with two cuBLAS bmm calls.  Here one CTA owns 64 queries and
// streams 64-key tiles through shared memory with an online softmax; nothing of size N^2 exists.
//   Q = wq*q+bq, K = wk*kv+bk, V = wv*kv+bv        (1x1x1 convs with one channel = scalar affine)
//   E[i,j] = sum_d Q[d,i] K[d,j]                    (no 1/sqrt(d))
//   A[i,j] = softmax_j(E[i,:])[j] * conf[j],  conf[j] = max_d softmax_d(q)[d,j]   (key-indexed)
//   out[d,i] = gamma * sum_j V[d,j] A[i,j] + kv[d,i]
// Contractions run on tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate); the energy product
// can use the 3xTF32 split because exp() amplifies its absolute error.
#include "common.cuh"

namespace ss {

constexpr int BQ = 64;          // queries per CTA
constexpr int BKEY = 64;        // keys per tile
constexpr int ATT_THREADS = 128;
constexpr int QK_LD = 72;       // [d][token] tiles used as (k=d, m/n=token) operands: bank = 8t+g
constexpr int V_LD = 68;        // [d][token] tile used as (n=d, k=token) operand:     bank = 4g+t
constexpr int P_LD = 68;

// conf[b,j] = max_d softmax_d(q)[d,j] = 1 / sum_d exp(q[d,j] - max_d q[:,j])
__global__ void bri_conf_kernel(const float* __restrict__ q, float* __restrict__ conf, int D, int N) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= N) return;
    const float* qp = q + (size_t)b * D * N + j;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(qp + (size_t)d * N));
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(__ldg(qp + (size_t)d * N) - m);
    conf[(size_t)b * N + j] = 1.0f / s;
}

template <int DP, bool PRECISE>   // DP = D rounded up to a multiple of 8
__global__ void __launch_bounds__(ATT_THREADS)
bri_attn_kernel(const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ params,
                const float* __restrict__ conf, float* __restrict__ out, int out_ld, int D, int N, int keys_per_split,
                float* __restrict__ part_ml, float* __restrict__ part_o) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                          // [DP][QK_LD]
    float* Ks = Qs + DP * QK_LD;             // [DP][QK_LD]
    float* Vs = Ks + DP * QK_LD;             // [DP][V_LD]
    float* Ps = Vs + DP * V_LD;              // [4][16][P_LD]
    float* cs = Ps + 4 * 16 * P_LD;          // [BKEY]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y, i0 = blockIdx.x * BQ;
    const float wq = __ldg(params + 0), bq = __ldg(params + 1), wk = __ldg(params + 2), bk = __ldg(params + 3),
                wv = __ldg(params + 4), bv = __ldg(params + 5), gamma = __ldg(params + 6);
    const float* qb = q + (size_t)b * D * N;
    const float* kvb = kv + (size_t)b * D * N;

    // stage Q tile (zero rows d >= D, zero columns i >= N)
    for (int idx = tid; idx < DP * BQ; idx += ATT_THREADS) {
        const int d = idx / BQ, i = idx % BQ;
        float v = 0.f;
        if (d < D && i0 + i < N) v = fmaf(wq, __ldg(qb + (size_t)d * N + i0 + i), bq);
        Qs[d * QK_LD + i] = v;
    }

    constexpr int NFO = DP / 8;              // output n-fragments (over d)
    float o_acc[NFO][4];
#pragma unroll
    for (int i = 0; i < NFO; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) o_acc[i][k] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float* Pw = Ps + warp * 16 * P_LD;
    const int m0 = warp * 16;

    // key split (flash-decoding): blockIdx.z owns keys [j_begin, j_end); partial (m, l, O) are merged by
    // bri_combine_kernel when gridDim.z > 1
    const int j_begin = blockIdx.z * keys_per_split;
    const int j_end = min(N, j_begin + keys_per_split);
    for (int j0 = j_begin; j0 < j_end; j0 += BKEY) {
        __syncthreads();                      // previous tile fully consumed (also covers Qs on entry)
        for (int idx = tid; idx < DP * BKEY; idx += ATT_THREADS) {
            const int d = idx / BKEY, j = idx % BKEY;
            float kvv = 0.f;
            const bool ok = (d < D && j0 + j < j_end);
            if (ok) kvv = __ldg(kvb + (size_t)d * N + j0 + j);
            Ks[d * QK_LD + j] = ok ? fmaf(wk, kvv, bk) : 0.f;
            Vs[d * V_LD + j] = ok ? fmaf(wv, kvv, bv) : 0.f;
        }
        if (tid < BKEY) cs[tid] = (j0 + tid < j_end) ? __ldg(conf + (size_t)b * N + j0 + tid) : 0.f;
        __syncthreads();

        // ---- S = Q^T K  (16 x 64 per warp)
        float s_acc[BKEY / 8][4];
#pragma unroll
        for (int i = 0; i < BKEY / 8; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) s_acc[i][k] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < DP / 8; ++ks) {
            const float* qa = Qs + (ks * 8 + t) * QK_LD + m0 + g;
            const float f0 = qa[0], f1 = qa[8], f2 = qa[4 * QK_LD], f3 = qa[4 * QK_LD + 8];
            uint32_t ah[4] = {f2tf32(f0), f2tf32(f1), f2tf32(f2), f2tf32(f3)};
            uint32_t al[4];
            if (PRECISE) {
                al[0] = f2tf32(f0 - __uint_as_float(ah[0])); al[1] = f2tf32(f1 - __uint_as_float(ah[1]));
                al[2] = f2tf32(f2 - __uint_as_float(ah[2])); al[3] = f2tf32(f3 - __uint_as_float(ah[3]));
            }
#pragma unroll
            for (int nf = 0; nf < BKEY / 8; ++nf) {
                const float* kb = Ks + (ks * 8 + t) * QK_LD + nf * 8 + g;
                const float e0 = kb[0], e1 = kb[4 * QK_LD];
                uint32_t bh[2] = {f2tf32(e0), f2tf32(e1)};
                if (PRECISE) {
                    uint32_t bl[2] = {f2tf32(e0 - __uint_as_float(bh[0])), f2tf32(e1 - __uint_as_float(bh[1]))};
                    mma_tf32(s_acc[nf], al, bh);
                    mma_tf32(s_acc[nf], ah, bl);
                }
                mma_tf32(s_acc[nf], ah, bh);
            }
        }
        // ---- online softmax over keys (rows g and g+8 of this warp)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float mx = -INFINITY;
#pragma unroll
            for (int nf = 0; nf < BKEY / 8; ++nf) {
                const int j = j0 + nf * 8 + 2 * t;
                if (j >= j_end) s_acc[nf][2 * h] = -INFINITY;
                if (j + 1 >= j_end) s_acc[nf][2 * h + 1] = -INFINITY;
                mx = fmaxf(mx, fmaxf(s_acc[nf][2 * h], s_acc[nf][2 * h + 1]));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[h], mx);
            const float corr = expf(m_run[h] - m_new);       // exp(-inf) = 0 on the first tile
            float ls = 0.f;
#pragma unroll
            for (int nf = 0; nf < BKEY / 8; ++nf) {
                const float p0 = expf(s_acc[nf][2 * h] - m_new), p1 = expf(s_acc[nf][2 * h + 1] - m_new);
                ls += p0 + p1;
                const int jl = nf * 8 + 2 * t;
                *reinterpret_cast<float2*>(Pw + (g + 8 * h) * P_LD + jl) = make_float2(p0 * cs[jl], p1 * cs[jl + 1]);
            }
            ls += __shfl_xor_sync(0xffffffffu, ls, 1);
            ls += __shfl_xor_sync(0xffffffffu, ls, 2);
            l_run[h] = l_run[h] * corr + ls;
            m_run[h] = m_new;
#pragma unroll
            for (int nf = 0; nf < NFO; ++nf) { o_acc[nf][2 * h] *= corr; o_acc[nf][2 * h + 1] *= corr; }
        }
        __syncwarp();
        // ---- O += P V^T  (16 x DP per warp, K = 64 keys)
#pragma unroll 2
        for (int ks = 0; ks < BKEY / 8; ++ks) {
            const float* pa = Pw + g * P_LD + ks * 8 + t;
            uint32_t a[4] = {f2tf32(pa[0]), f2tf32(pa[8 * P_LD]), f2tf32(pa[4]), f2tf32(pa[8 * P_LD + 4])};
#pragma unroll
            for (int nf = 0; nf < NFO; ++nf) {
                const float* vb = Vs + (nf * 8 + g) * V_LD + ks * 8 + t;
                uint32_t bb[2] = {f2tf32(vb[0]), f2tf32(vb[4])};
                mma_tf32(o_acc[nf], a, bb);
            }
        }
        __syncwarp();
    }

    // ---- epilogue: normalise, gamma-residual, store (or write the split's partial state)
    const bool split = gridDim.z > 1;
    const size_t pbase = ((size_t)blockIdx.z * gridDim.y + b);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = i0 + m0 + g + 8 * h;
        if (i >= N) continue;
        if (split) {
            if (t == 0) {
                part_ml[(pbase * 2 + 0) * N + i] = m_run[h];
                part_ml[(pbase * 2 + 1) * N + i] = l_run[h];
            }
#pragma unroll
            for (int nf = 0; nf < NFO; ++nf)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int d = nf * 8 + 2 * t + e;
                    if (d < D) part_o[(pbase * D + d) * N + i] = o_acc[nf][2 * h + e];
                }
        } else {
            const float inv = 1.0f / l_run[h];
#pragma unroll
            for (int nf = 0; nf < NFO; ++nf) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int d = nf * 8 + 2 * t + e;
                    if (d < D) {
                        const float res = __ldg(kvb + (size_t)d * N + i);
                        out[((size_t)b * D * N + (size_t)d * N + i) * out_ld] = fmaf(gamma, o_acc[nf][2 * h + e] * inv, res);
                    }
                }
            }
        }
    }
}

// merge the key splits: out = gamma * (sum_s O_s e^{m_s-m}) / (sum_s l_s e^{m_s-m}) + kv
__global__ void bri_combine_kernel(const float* __restrict__ part_ml, const float* __restrict__ part_o,
                                   const float* __restrict__ kv, const float* __restrict__ params,
                                   float* __restrict__ out, int out_ld, int B, int D, int N, int KS) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    if (i >= N) return;
    float m = -INFINITY;
    for (int s = 0; s < KS; ++s) m = fmaxf(m, part_ml[(((size_t)s * B + b) * 2 + 0) * N + i]);
    float w[8];
    float l = 0.f;
    for (int s = 0; s < KS; ++s) {
        w[s] = expf(part_ml[(((size_t)s * B + b) * 2 + 0) * N + i] - m);
        l += w[s] * part_ml[(((size_t)s * B + b) * 2 + 1) * N + i];
    }
    const float gamma = __ldg(params + 6), inv = 1.0f / l;
    for (int d = blockIdx.y; d < D; d += gridDim.y) {
        float o = 0.f;
        for (int s = 0; s < KS; ++s) o += w[s] * part_o[(((size_t)s * B + b) * D + d) * N + i];
        out[((size_t)b * D * N + (size_t)d * N + i) * out_ld] = fmaf(gamma, o * inv, __ldg(kv + ((size_t)b * D + d) * N + i));
    }
}

void bri_combine_launch(const float* part_ml, const float* part_o, const float* kv, const float* params, float* out, int out_ld,
                        int B, int D, int N, int KS, cudaStream_t st) {
    dim3 cgrid((N + 127) / 128, min(D, 8), B);
    bri_combine_kernel<<<cgrid, 128, 0, st>>>(part_ml, part_o, kv, params, out, out_ld, B, D, N, KS);
}

// tcgen05 kernel (bri_attn_tc.cu)
size_t bri_tc_workspace_floats(int B, int D, int N);
int try_bri_tc(const float* q, const float* kv, const float* params, float* ws, float* out, int out_ld, int B, int D, int N,
               cudaStream_t st, int* rc);

static int bri_key_splits(int B, int N) {
    // enough CTAs for ~2 waves of 2 CTAs/SM; at most 8 splits, each a multiple of the key tile
    const int qtiles = (N + BQ - 1) / BQ;
    int ks = 1;
    while (ks < 8 && (long long)qtiles * B * ks < 2 * 2 * 148 && (N / (ks * 2)) >= 4 * BKEY) ks *= 2;
    return ks;
}

template <int DP>
static int launch_bri(const float* q, const float* kv, const float* params, const float* conf, float* out, int out_ld,
                      int B, int D, int N, bool precise, int KS, float* part_ml, float* part_o, cudaStream_t st) {
    const size_t smem = (size_t)(2 * DP * QK_LD + DP * V_LD + 4 * 16 * P_LD + BKEY) * sizeof(float);
    const int kps = ((N + KS - 1) / KS + BKEY - 1) / BKEY * BKEY;
    dim3 grid((N + BQ - 1) / BQ, B, KS);
    static thread_local bool configured = false;
    if (!configured) {
        SS_CUDA(cudaFuncSetAttribute(bri_attn_kernel<DP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SS_CUDA(cudaFuncSetAttribute(bri_attn_kernel<DP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    if (precise)
        bri_attn_kernel<DP, true><<<grid, ATT_THREADS, smem, st>>>(q, kv, params, conf, out, out_ld, D, N, kps, part_ml, part_o);
    else
        bri_attn_kernel<DP, false><<<grid, ATT_THREADS, smem, st>>>(q, kv, params, conf, out, out_ld, D, N, kps, part_ml, part_o);
    int rc = check_launch("bri_attn_kernel");
    if (rc || KS == 1) return rc;
    dim3 cgrid((N + 127) / 128, min(D, 8), B);
    bri_combine_kernel<<<cgrid, 128, 0, st>>>(part_ml, part_o, kv, params, out, out_ld, B, D, N, KS);
    return check_launch("bri_combine_kernel");
}

}  // namespace ss

extern "C" size_t ss_bri_workspace_bytes(int B, int D, int N) {
    const int ks = ss::bri_key_splits(B, N);
    size_t fl = (size_t)B * N;                                   // conf
    if (ks > 1) fl += (size_t)ks * B * N * (2 + D);              // partial (m, l) and O
    const size_t fl_tc = ss::bri_tc_workspace_floats(B, D, N);   // the tcgen05 kernel lays the workspace out its own way
    return (fl > fl_tc ? fl : fl_tc) * sizeof(float);
}

extern "C" int ss_bri_attn_fwd(const float* q, const float* kv, const float* params, float* ws, size_t ws_bytes,
                               float* out, int out_ld, int B, int D, int N, int math, void* stream) {
    using namespace ss;
    SS_REQUIRE(q && kv && params && ws && out, "ss_bri_attn_fwd: null pointer");
    SS_REQUIRE(B > 0 && B <= 65535 && D > 0 && D <= 128 && N > 0 && out_ld >= 1, "ss_bri_attn_fwd: shape (D <= 128)");
    if (ws_bytes < ss_bri_workspace_bytes(B, D, N)) return SS_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (math == SS_MATH_TF32 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
        int rct = 0;
        if (try_bri_tc(q, kv, params, ws, out, out_ld, B, D, N, st, &rct)) return rct;
    }
    const int KS = bri_key_splits(B, N);
    float* conf_ws = ws;
    float* part_ml = ws + (size_t)B * N;
    float* part_o = part_ml + (size_t)KS * B * N * 2;
    dim3 cgrid((N + 127) / 128, B);
    bri_conf_kernel<<<cgrid, 128, 0, st>>>(q, conf_ws, D, N);
    int rc = check_launch("bri_conf_kernel");
    if (rc) return rc;
    const bool precise = (math == SS_MATH_3XTF32);
    const int dp = (D + 7) / 8 * 8;
#define SS_BRI(DPV) return launch_bri<DPV>(q, kv, params, conf_ws, out, out_ld, B, D, N, precise, KS, part_ml, part_o, st)
    if (dp <= 32) SS_BRI(32);
    if (dp <= 48) SS_BRI(48);
    if (dp <= 64) SS_BRI(64);
    if (dp <= 96) SS_BRI(96);
    if (dp <= 112) SS_BRI(112);
    SS_BRI(128);
#undef SS_BRI
}
