// Normalisation bookkeeping, residual joins, depth softmax and layout helpers.
//
// GroupNorm / BatchNorm(eval) / SE gates are "pending affines": the producer conv leaves
// per-(batch,channel) sums, ss_gn_finalize turns them into scale/shift vectors on [B,C], and
// the *consumer* applies act(x*scale+shift) while loading.  The only full-volume elementwise
// kernel left is the residual join, which has to materialise its result anyway.
#include "common.cuh"

namespace ss {

__global__ void gn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int B, int C, int groups, double count,
                                   float eps, float* __restrict__ scale, float* __restrict__ shift, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i % C;
    const int cpg = C / groups, g0 = (c / cpg) * cpg;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < cpg; ++k) {
        s += stats[((size_t)b * C + g0 + k) * 2 + 0];
        q += stats[((size_t)b * C + g0 + k) * 2 + 1];
    }
    const double n = count * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = __ldg(gamma + c) * rstd;
    scale[(size_t)b * ld + c] = sc;
    shift[(size_t)b * ld + c] = __ldg(beta + c) - (float)mean * sc;
}

// one CTA per batch sample; C <= 1024, Cmid <= 128
__global__ void ca3d_gate_kernel(const double* __restrict__ stats, double count, float* __restrict__ scale,
                                 float* __restrict__ shift, const float* __restrict__ w1,
                                 const float* __restrict__ b1, const float* __restrict__ w2,
                                 const float* __restrict__ b2, int C, int Cmid) {
    extern __shared__ float sm[];
    float* pool = sm;          // [C]
    float* hid = sm + C;       // [Cmid]
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mean_raw = (float)(stats[((size_t)b * C + c) * 2] / count);
        pool[c] = fmaf(scale[(size_t)b * C + c], mean_raw, shift[(size_t)b * C + c]);
    }
    __syncthreads();
    for (int m = threadIdx.x; m < Cmid; m += blockDim.x) {
        float a = __ldg(b1 + m);
        for (int c = 0; c < C; ++c) a = fmaf(__ldg(w1 + (size_t)m * C + c), pool[c], a);
        hid[m] = gelu_erf(a);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = __ldg(b2 + c);
        for (int m = 0; m < Cmid; ++m) a = fmaf(__ldg(w2 + (size_t)c * Cmid + m), hid[m], a);
        const float gate = 1.0f / (1.0f + expf(-gelu_erf(a)));
        scale[(size_t)b * C + c] *= gate;
        shift[(size_t)b * C + c] *= gate;
    }
}

// out = act( alpha * A(x) + A(r) ); 4 channels per thread when VEC
template <bool VEC>
__global__ void affine_join_kernel(const float* __restrict__ x, const float* __restrict__ xs,
                                   const float* __restrict__ xh, int x_act, const float* __restrict__ r,
                                   const float* __restrict__ rs, const float* __restrict__ rh, int r_act,
                                   const float* __restrict__ alpha_p, int out_act, long long V, int C,
                                   int x_ldc, int r_ldc, int out_ldc, float* __restrict__ out) {
    const int b = blockIdx.y;
    const float alpha = alpha_p ? __ldg(alpha_p) : 1.0f;
    constexpr int W = VEC ? 4 : 1;
    const int cw = C / W;
    const long long total = V * cw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long v = i / cw;
        const int c = (int)(i % cw) * W;
        const size_t vox = (size_t)b * V + v;
        float xv[W], rv[W];
        if (VEC) {
            float4 t4 = ldg_f4(x + vox * x_ldc + c);
            xv[0] = t4.x; xv[1] = t4.y; xv[2] = t4.z; xv[3] = t4.w;
            if (r) {
                float4 u4 = ldg_f4(r + vox * r_ldc + c);
                rv[0] = u4.x; rv[1] = u4.y; rv[2] = u4.z; rv[3] = u4.w;
            }
        } else {
            xv[0] = __ldg(x + vox * x_ldc + c);
            if (r) rv[0] = __ldg(r + vox * r_ldc + c);
        }
        float o[W];
#pragma unroll
        for (int k = 0; k < W; ++k) {
            float a = xv[k];
            if (xs) a = fmaf(a, __ldg(xs + (size_t)b * C + c + k), __ldg(xh + (size_t)b * C + c + k));
            a = apply_act(a, x_act) * alpha;
            if (r) {
                float e = rv[k];
                if (rs) e = fmaf(e, __ldg(rs + (size_t)b * C + c + k), __ldg(rh + (size_t)b * C + c + k));
                a += apply_act(e, r_act);
            }
            o[k] = apply_act(a, out_act);
        }
        if (VEC) *reinterpret_cast<float4*>(out + vox * out_ldc + c) = make_float4(o[0], o[1], o[2], o[3]);
        else out[vox * out_ldc + c] = o[0];
    }
}

// per-(batch,channel) sum and sum of squares of a pending volume act(x*scale+shift): a CTA owns a slab of
// voxels, a thread owns channels tid, tid+blockDim, ... (consecutive threads read consecutive channels)
__global__ void channel_sums_kernel(const float* __restrict__ x, const float* __restrict__ xs,
                                    const float* __restrict__ xh, int act, long long V, int C, int ldc,
                                    double* __restrict__ stats) {
    const int b = blockIdx.y;
    const long long per = (V + gridDim.x - 1) / gridDim.x;
    const long long v0 = (long long)blockIdx.x * per;
    const long long v1 = (v0 + per < V) ? v0 + per : V;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = xs ? __ldg(xs + (size_t)b * C + c) : 1.0f;
        const float sh = xs ? __ldg(xh + (size_t)b * C + c) : 0.0f;
        float s = 0.f, q = 0.f;
        for (long long v = v0; v < v1; ++v) {
            const float a = apply_act(fmaf(__ldg(x + ((size_t)b * V + v) * ldc + c), sc, sh), act);
            s += a;
            q = fmaf(a, a, q);
        }
        if (v1 > v0) {
            atomicAdd(stats + ((size_t)b * C + c) * 2 + 0, (double)s);
            atomicAdd(stats + ((size_t)b * C + c) * 2 + 1, (double)q);
        }
    }
}

// softmax over D of x[b][d][p]; one thread per pixel column, coalesced across p.
__global__ void softmax_d_kernel(const float* __restrict__ x, long long xbs, float* __restrict__ y, long long ybs,
                                 int D, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= P) return;
    const float* xp = x + (size_t)b * xbs + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(xp + (size_t)d * P));
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(__ldg(xp + (size_t)d * P) - m);
    const float inv = 1.0f / s;
    float* yp = y + (size_t)b * ybs + p;
    for (int d = 0; d < D; ++d) yp[(size_t)d * P] = expf(__ldg(xp + (size_t)d * P) - m) * inv;
}

// [B][C][V] -> [B][V][ldc] via a 32x32 shared-memory transpose
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long long V, int ldc) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long v0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long v = v0 + threadIdx.x;
        if (c < C && v < V) tile[j][threadIdx.x] = __ldg(x + ((size_t)b * C + c) * V + v);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long v = v0 + j;
        const int c = c0 + threadIdx.x;
        if (c < C && v < V) y[((size_t)b * V + v) * ldc + c] = tile[threadIdx.x][j];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long long V, int ldc) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long v0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long v = v0 + j;
        const int c = c0 + threadIdx.x;
        if (c < C && v < V) tile[j][threadIdx.x] = __ldg(x + ((size_t)b * V + v) * ldc + c);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long v = v0 + threadIdx.x;
        if (c < C && v < V) y[((size_t)b * C + c) * V + v] = tile[threadIdx.x][j];
    }
}

}  // namespace ss

using namespace ss;

// GroupNorm finalize whose result is also multiplied by up to two per-(batch,channel) gates (> 0, e.g. SE gates:
// relu(gn(y)) * g == relu(gn(y) * g)), one (scale, shift) pair per gate: out_k = (scale * gate_k, shift * gate_k).
__global__ void gn_finalize_gated_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, int B, int C, int groups, double count, float eps,
                                         const float* __restrict__ gate1, float* __restrict__ scale1, float* __restrict__ shift1,
                                         const float* __restrict__ gate2, float* __restrict__ scale2, float* __restrict__ shift2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i % C;
    const int cpg = C / groups, g0 = (c / cpg) * cpg;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < cpg; ++k) {
        s += stats[((size_t)b * C + g0 + k) * 2 + 0];
        q += stats[((size_t)b * C + g0 + k) * 2 + 1];
    }
    const double n = count * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = __ldg(gamma + c) * rstd;
    const float sh = __ldg(beta + c) - (float)mean * sc;
    const float g1 = __ldg(gate1 + i);
    scale1[i] = sc * g1;
    shift1[i] = sh * g1;
    if (gate2) {
        const float g2 = __ldg(gate2 + i);
        scale2[i] = sc * g2;
        shift2[i] = sh * g2;
    }
}

// ASPP image-pooling branch folded into the pending shift of the fusing 1x1 conv's BatchNorm (ViewTransformerLSSBEVDepth.py:
// 373-379, 394-406): the branch is constant over the map, so its share of conv1 is a per-(batch, channel) vector.
//   pooled[c] = sum_x[b][c] / count;  t = W1 pooled;  g = relu(GroupNorm_groups(t));  u = Wp g;
//   shift_out[b][j] = bn_shift[b][j] + bn_scale[b][j] * u[j]
// Two launches of (B x mid/32) CTAs, 8 warps each, a warp computes 4 outputs at a time with float4 loads (the two 640 x 640
// matrix-vector products are bound by streaming 3.3 MB of weights: one CTA per sample took 0.11 ms, 20 per sample take ~5 us).
constexpr int AP_ROWS = 32;      // outputs per CTA

__device__ __forceinline__ void ap_matvec_rows(const float* __restrict__ w, const float* __restrict__ x_s, int K, int row0, int nrows,
                                               float* out4 /* [4] per warp-iteration */, int warp, int lane, float* dst_s) {
    // rows row0 + warp*4 .. +3 of w (row-major [.][K]) times x_s[K]
    for (int r = 0; r < 4; ++r) {
        const int o = row0 + warp * 4 + r;
        float a = 0.f;
        if (warp * 4 + r < nrows) {
            const float* wr = w + (size_t)o * K;
            if ((K & 3) == 0) {
                for (int c = lane * 4; c < K; c += 128) {
                    const float4 wv = ldg_f4(wr + c);
                    a = fmaf(wv.x, x_s[c], a); a = fmaf(wv.y, x_s[c + 1], a); a = fmaf(wv.z, x_s[c + 2], a); a = fmaf(wv.w, x_s[c + 3], a);
                }
            } else {
                for (int c = lane; c < K; c += 32) a = fmaf(__ldg(wr + c), x_s[c], a);
            }
        }
        out4[r] = warp_sum(a);
    }
    if (lane == 0)
        for (int r = 0; r < 4; ++r)
            if (warp * 4 + r < nrows) dst_s[warp * 4 + r] = out4[r];
}

__global__ void __launch_bounds__(256)
aspp_pool_t_kernel(const double* __restrict__ stats, double count, const float* __restrict__ w1, float* __restrict__ t, int C, int mid) {
    extern __shared__ float sm[];
    float* pooled = sm;                 // [C]
    float* res = sm + C;                // [AP_ROWS]
    const int b = blockIdx.y, row0 = blockIdx.x * AP_ROWS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < C; c += blockDim.x) pooled[c] = (float)(stats[((size_t)b * C + c) * 2] / count);
    __syncthreads();
    float o4[4];
    ap_matvec_rows(w1, pooled, C, row0, min(AP_ROWS, mid - row0), o4, warp, lane, res);
    __syncthreads();
    if (tid < AP_ROWS && row0 + tid < mid) t[(size_t)b * mid + row0 + tid] = res[tid];
}

__global__ void __launch_bounds__(256)
aspp_pool_shift_kernel(const float* __restrict__ t, const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps,
                       const float* __restrict__ wp, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                       float* __restrict__ shift_out, int mid) {
    extern __shared__ float sm[];
    float* g = sm;                      // [mid]   relu(GroupNorm(t))
    float* red = sm + mid;              // [2 * groups]
    float* res = red + 2 * groups;      // [AP_ROWS]
    const int b = blockIdx.y, row0 = blockIdx.x * AP_ROWS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int o = tid; o < mid; o += blockDim.x) g[o] = t[(size_t)b * mid + o];
    __syncthreads();
    const int cpg = mid / groups;
    for (int gi = warp; gi < groups; gi += blockDim.x >> 5) {      // one warp per group: mean and variance over its channels
        float s = 0.f, q = 0.f;
        for (int k = lane; k < cpg; k += 32) { const float v = g[gi * cpg + k]; s += v; q = fmaf(v, v, q); }
        s = warp_sum(s); q = warp_sum(q);
        if (lane == 0) {
            const float mean = s / cpg;
            float var = q / cpg - mean * mean;
            if (var < 0.f) var = 0.f;
            red[2 * gi] = mean;
            red[2 * gi + 1] = rsqrtf(var + eps);
        }
    }
    __syncthreads();
    for (int o = tid; o < mid; o += blockDim.x) {
        const int gi = o / cpg;
        g[o] = fmaxf((g[o] - red[2 * gi]) * red[2 * gi + 1] * __ldg(gamma + o) + __ldg(beta + o), 0.f);
    }
    __syncthreads();
    float o4[4];
    ap_matvec_rows(wp, g, mid, row0, min(AP_ROWS, mid - row0), o4, warp, lane, res);
    __syncthreads();
    if (tid < AP_ROWS && row0 + tid < mid) {
        const size_t j = (size_t)b * mid + row0 + tid;
        shift_out[j] = fmaf(__ldg(bn_scale + j), res[tid], __ldg(bn_shift + j));
    }
}

extern "C" int ss_gn_finalize_gated(const double* stats, const float* gamma, const float* beta, int B, int C, int groups,
                                    double count, float eps, const float* gate1, float* scale1, float* shift1,
                                    const float* gate2, float* scale2, float* shift2, void* stream) {
    SS_REQUIRE(stats && gamma && beta && gate1 && scale1 && shift1, "ss_gn_finalize_gated: null pointer");
    SS_REQUIRE(!gate2 || (scale2 && shift2), "ss_gn_finalize_gated: second gate needs its outputs");
    SS_REQUIRE(B > 0 && C > 0 && groups > 0 && C % groups == 0 && count > 0, "ss_gn_finalize_gated: shape");
    const int n = B * C;
    gn_finalize_gated_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, gamma, beta, B, C, groups, count, eps, gate1,
                                                                                scale1, shift1, gate2, scale2, shift2);
    return check_launch("gn_finalize_gated_kernel");
}

extern "C" int ss_aspp_pool_shift(const double* stats, double count, const float* w1, const float* gamma, const float* beta, int groups,
                                  float eps, const float* w_pool, const float* bn_scale, const float* bn_shift, float* shift_out,
                                  float* t_ws, int B, int C, int mid, void* stream) {
    SS_REQUIRE(stats && w1 && gamma && beta && w_pool && bn_scale && bn_shift && shift_out && t_ws, "ss_aspp_pool_shift: null pointer");
    SS_REQUIRE(B > 0 && C > 0 && mid > 0 && groups > 0 && groups <= 32 && mid % groups == 0 && count > 0, "ss_aspp_pool_shift: shape");
    SS_REQUIRE(((size_t)C + AP_ROWS) * 4 <= 48 * 1024 && ((size_t)mid + 2 * groups + AP_ROWS) * 4 <= 48 * 1024,
               "ss_aspp_pool_shift: channel counts too large");
    dim3 grid((mid + AP_ROWS - 1) / AP_ROWS, B);
    cudaStream_t st = (cudaStream_t)stream;
    aspp_pool_t_kernel<<<grid, 256, ((size_t)C + AP_ROWS) * sizeof(float), st>>>(stats, count, w1, t_ws, C, mid);
    int rc = check_launch("aspp_pool_t_kernel");
    if (rc != SS_OK) return rc;
    aspp_pool_shift_kernel<<<grid, 256, ((size_t)mid + 2 * groups + AP_ROWS) * sizeof(float), st>>>(t_ws, gamma, beta, groups, eps, w_pool,
                                                                                                  bn_scale, bn_shift, shift_out, mid);
    return check_launch("aspp_pool_shift_kernel");
}

extern "C" int ss_gn_finalize(const double* stats, const float* gamma, const float* beta, int B, int C, int groups,
                              double count, float eps, float* scale, float* shift, int ld_out, void* stream) {
    SS_REQUIRE(stats && gamma && beta && scale && shift, "ss_gn_finalize: null pointer");
    SS_REQUIRE(B > 0 && C > 0 && groups > 0 && C % groups == 0 && ld_out >= C && count > 0, "ss_gn_finalize: shape");
    const int n = B * C;
    gn_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, gamma, beta, B, C, groups, count, eps,
                                                                          scale, shift, ld_out);
    return check_launch("gn_finalize_kernel");
}

extern "C" int ss_ca3d_gate(const double* stats, double count, float* scale, float* shift, const float* w1,
                            const float* b1, const float* w2, const float* b2, int B, int C, int Cmid, void* stream) {
    SS_REQUIRE(stats && scale && shift && w1 && b1 && w2 && b2, "ss_ca3d_gate: null pointer");
    SS_REQUIRE(B > 0 && C > 0 && C <= 1024 && Cmid > 0 && Cmid <= 128 && count > 0, "ss_ca3d_gate: shape");
    ca3d_gate_kernel<<<B, 128, (C + Cmid) * sizeof(float), (cudaStream_t)stream>>>(stats, count, scale, shift, w1, b1,
                                                                                   w2, b2, C, Cmid);
    return check_launch("ca3d_gate_kernel");
}

extern "C" int ss_affine_join_fwd(const float* x, const float* x_scale, const float* x_shift, int x_act,
                                  const float* r, const float* r_scale, const float* r_shift, int r_act,
                                  const float* alpha, int out_act, int B, long long V, int C, int x_ldc, int r_ldc,
                                  int out_ldc, float* out, void* stream) {
    SS_REQUIRE(x && out, "ss_affine_join_fwd: null pointer");
    SS_REQUIRE(x_act <= SS_ACT_GELU && r_act <= SS_ACT_GELU && out_act <= SS_ACT_GELU, "ss_affine_join_fwd: activations NONE | RELU | GELU");
    SS_REQUIRE(B > 0 && B <= 65535 && V > 0 && C > 0, "ss_affine_join_fwd: shape");
    SS_REQUIRE((x_scale == nullptr) == (x_shift == nullptr) && (r_scale == nullptr) == (r_shift == nullptr),
               "ss_affine_join_fwd: scale/shift must come together");
    const bool vec = (C % 4 == 0) && (x_ldc % 4 == 0) && (out_ldc % 4 == 0) && (!r || r_ldc % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(r)) & 15) == 0;
    const long long total = V * (vec ? C / 4 : C);
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = 148LL * 16;     // grid-stride: 16 CTAs per SM
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)B);
    if (vec)
        affine_join_kernel<true><<<grid, threads, 0, (cudaStream_t)stream>>>(x, x_scale, x_shift, x_act, r, r_scale, r_shift,
                                                                            r_act, alpha, out_act, V, C, x_ldc, r_ldc,
                                                                            out_ldc, out);
    else
        affine_join_kernel<false><<<grid, threads, 0, (cudaStream_t)stream>>>(x, x_scale, x_shift, x_act, r, r_scale, r_shift,
                                                                             r_act, alpha, out_act, V, C, x_ldc, r_ldc,
                                                                             out_ldc, out);
    return check_launch("affine_join_kernel");
}

extern "C" int ss_channel_sums_fwd(const float* x, const float* x_scale, const float* x_shift, int x_act, int B,
                                   long long V, int C, int x_ldc, double* stats, void* stream) {
    SS_REQUIRE(x && stats, "ss_channel_sums_fwd: null pointer");
    SS_REQUIRE(B > 0 && B <= 65535 && V > 0 && C > 0 && x_ldc >= C, "ss_channel_sums_fwd: shape");
    SS_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), "ss_channel_sums_fwd: scale/shift must come together");
    long long blocks = (V + 15) / 16;                      // >= 16 voxels per CTA keeps the atomics sparse
    const long long cap = 148LL * 4;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)B);
    channel_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_scale, x_shift, x_act, V, C, x_ldc, stats);
    return check_launch("channel_sums_kernel");
}

extern "C" int ss_softmax_d_fwd(const float* x, long long x_batch_stride, float* y, long long y_batch_stride, int B,
                                int D, int P, void* stream) {
    SS_REQUIRE(x && y, "ss_softmax_d_fwd: null pointer");
    SS_REQUIRE(B > 0 && B <= 65535 && D > 0 && P > 0, "ss_softmax_d_fwd: shape");
    dim3 grid((P + 127) / 128, B);
    softmax_d_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, x_batch_stride, y, y_batch_stride, D, P);
    return check_launch("softmax_d_kernel");
}

extern "C" int ss_nchw_to_nhwc(const float* x, float* y, int B, int C, long long V, int out_ldc, void* stream) {
    SS_REQUIRE(x && y && B > 0 && B <= 65535 && C > 0 && V > 0 && out_ldc >= C, "ss_nchw_to_nhwc: arguments");
    dim3 grid((unsigned)((V + 31) / 32), (C + 31) / 32, B), block(32, 8);
    SS_REQUIRE((C + 31) / 32 <= 65535, "ss_nchw_to_nhwc: too many channels");
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, C, V, out_ldc);
    return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int ss_nhwc_to_nchw(const float* x, float* y, int B, int C, long long V, int in_ldc, void* stream) {
    SS_REQUIRE(x && y && B > 0 && B <= 65535 && C > 0 && V > 0 && in_ldc >= C, "ss_nhwc_to_nchw: arguments");
    dim3 grid((unsigned)((V + 31) / 32), (C + 31) / 32, B), block(32, 8);
    SS_REQUIRE((C + 31) / 32 <= 65535, "ss_nhwc_to_nchw: too many channels");
    nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, C, V, in_ldc);
    return check_launch("nhwc_to_nchw_kernel");
}
