// LSS lift (x) splat: voxel-index quantisation, a sorted point->voxel index, the fused
// depth-probability x image-feature outer product with per-voxel accumulation, and the exact
// operator-level replacement of the external mmdet3d.ops.bev_pool.
//
// Reference (projects/mmdet3d_plugin/occupancy/image2bev/ViewTransformerLSSVoxel.py):
//   lift :517-519 materialises depth_prob (x) img_feat as a [B,N,D,H,W,C] tensor (440 MB),
//   voxel_pooling :432-476 quantises geom to voxel indices (:441, trunc toward zero, int64),
//   filters (:447-451, boolean gather of an [N,128] matrix) and calls bev_pool (:473).
// Here the index (sorted point ids + CSR offsets per voxel) depends only on the calibration, the
// lifted volume is never formed, every output voxel is written exactly once (no memset, no
// atomics) and the accumulation order is ascending point id, i.e. that of a sequential
// index_add_, so the sums are reproducible bit for bit.
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace ss {

// rank of a point in OUR output layout [b][x][y][z]; dropped points get the sentinel `nvox_total`
__global__ void splat_rank_kernel(const float* __restrict__ geom, float3 dx, float3 bx, int nx, int ny, int nz,
                                  long long P, long long total, int32_t nvox_total, int32_t* __restrict__ coords,
                                  int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / P);
    // exactly ((geom - (bx - dx/2.)) / dx).long(): fp32 subtract, IEEE fp32 divide, truncate
    const float ox = __fsub_rn(bx.x, __fdiv_rn(dx.x, 2.0f));
    const float oy = __fsub_rn(bx.y, __fdiv_rn(dx.y, 2.0f));
    const float oz = __fsub_rn(bx.z, __fdiv_rn(dx.z, 2.0f));
    const float gx = __ldg(geom + 3 * i + 0), gy = __ldg(geom + 3 * i + 1), gz = __ldg(geom + 3 * i + 2);
    const float fx = __fdiv_rn(__fsub_rn(gx, ox), dx.x);
    const float fy = __fdiv_rn(__fsub_rn(gy, oy), dx.y);
    const float fz = __fdiv_rn(__fsub_rn(gz, oz), dx.z);
    // float -> int64 truncation; values far outside the grid only need to stay outside it
    const long long lx = (long long)fx, ly = (long long)fy, lz = (long long)fz;
    const bool kept = lx >= 0 && lx < nx && ly >= 0 && ly < ny && lz >= 0 && lz < nz;
    const int ix = (int)max(min(lx, (long long)INT32_MAX), (long long)INT32_MIN);
    const int iy = (int)max(min(ly, (long long)INT32_MAX), (long long)INT32_MIN);
    const int iz = (int)max(min(lz, (long long)INT32_MAX), (long long)INT32_MIN);
    if (coords) reinterpret_cast<int4*>(coords)[i] = make_int4(ix, iy, iz, kept ? 1 : 0);
    keys[i] = kept ? (((b * nx + ix) * ny + iy) * nz + iz) : nvox_total;
    vals[i] = (int32_t)i;
}

// CSR offsets by binary search over the sorted keys: start[v] = lower_bound(keys, v)
__global__ void csr_offsets_kernel(const int32_t* __restrict__ sorted_keys, long long n, int32_t nvox_total,
                                   int32_t* __restrict__ start) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nvox_total) return;
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(sorted_keys + mid) < v) lo = mid + 1; else hi = mid;
    }
    start[v] = (int32_t)lo;
}

// one warp per output voxel; each lane owns 4 consecutive channels of every 128-channel slab
__global__ void __launch_bounds__(256)
lift_splat_kernel(const float* __restrict__ depth_prob, const float* __restrict__ img_feat,
                  const int32_t* __restrict__ order, const int32_t* __restrict__ start, float* __restrict__ out,
                  int D, int HW, int C, long long nvox_total, int nvox_per_batch) {
    const long long v = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v >= nvox_total) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(v / nvox_per_batch);
    const int s = __ldg(start + v), e = __ldg(start + v + 1);
    const long long P = (long long)D * HW;
    if (C <= 128) {
        // The point ids and their depth probabilities are fetched lane-parallel, 32 points at a time, and handed round by shuffle:
        // the loop body is then left with ONE dependent load (the feature row) instead of a chain of three, and consecutive
        // iterations overlap.  Same ascending-point summation order as before.
        const int c0 = lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = s; base < e; base += 32) {
            const int cnt = min(32, e - base);
            int n_l = 0;
            float dp_l = 0.f;
            if (lane < cnt) {
                n_l = __ldg(order + base + lane);           // global point id = b*P + d*HW + pix
                dp_l = __ldg(depth_prob + n_l);
            }
            const int pix_l = (int)((n_l - b * P) % HW);
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const int pix = __shfl_sync(0xffffffffu, pix_l, i);
                const float dp = __shfl_sync(0xffffffffu, dp_l, i);
                if (c0 < C) {
                    const float4 f = ldg_f4(img_feat + ((size_t)b * HW + pix) * C + c0);
                    // separate multiply and add (no fma): reproduces "materialise the product, then sum"
                    acc.x = __fadd_rn(acc.x, __fmul_rn(dp, f.x));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(dp, f.y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(dp, f.z));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(dp, f.w));
                }
            }
        }
        if (c0 < C) st_cs_f4(out + (size_t)v * C + c0, acc);
        return;
    }
    for (int c0 = lane * 4; c0 < C; c0 += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = s; i < e; ++i) {
            const int n = __ldg(order + i);                 // global point id = b*P + d*HW + pix
            const int pix = (int)((n - b * P) % HW);
            const float dp = __ldg(depth_prob + n);
            const float4 f = ldg_f4(img_feat + ((size_t)b * HW + pix) * C + c0);
            // separate multiply and add (no fma): reproduces "materialise the product, then sum"
            acc.x = __fadd_rn(acc.x, __fmul_rn(dp, f.x));
            acc.y = __fadd_rn(acc.y, __fmul_rn(dp, f.y));
            acc.z = __fadd_rn(acc.z, __fmul_rn(dp, f.z));
            acc.w = __fadd_rn(acc.w, __fmul_rn(dp, f.w));
        }
        st_cs_f4(out + (size_t)v * C + c0, acc);
    }
}

// ---- exact bev_pool drop-in ------------------------------------------------------------------
__global__ void bevpool_rank_kernel(const int64_t* __restrict__ coords, long long N, int B, int D, int H, int W,
                                    int32_t nvox_total, int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const long long x = coords[4 * i + 0], y = coords[4 * i + 1], z = coords[4 * i + 2], b = coords[4 * i + 3];
    const bool ok = x >= 0 && x < H && y >= 0 && y < W && z >= 0 && z < D && b >= 0 && b < B;
    keys[i] = ok ? (int32_t)((((b * D + z) * H + x) * W) + y) : nvox_total;
    vals[i] = (int32_t)i;
}

// thread per (voxel, channel); consecutive threads = consecutive voxels (y fastest) -> coalesced
// NCDHW stores
__global__ void bevpool_sum_kernel(const float* __restrict__ feats, const int32_t* __restrict__ order,
                                   const int32_t* __restrict__ start, float* __restrict__ out, int C,
                                   long long nvox_per_batch, long long nvox_total) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (v >= nvox_total) return;
    const int s = __ldg(start + v), e = __ldg(start + v + 1);
    float acc = 0.f;
    for (int i = s; i < e; ++i) acc = __fadd_rn(acc, __ldg(feats + (size_t)__ldg(order + i) * C + c));
    const long long b = v / nvox_per_batch, r = v % nvox_per_batch;
    out[((size_t)b * C + c) * nvox_per_batch + r] = acc;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sort_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)n);
    return bytes;
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_splat_index_workspace_bytes(long long n_points) {
    // keys_in, keys_out, vals_in + cub temp
    return 3 * align256((size_t)n_points * sizeof(int32_t)) + align256(sort_temp_bytes(n_points));
}

static int end_bit_for(long long nvox_total) {
    int bits = 1;
    while ((1ll << bits) <= nvox_total) ++bits;
    return bits;
}

extern "C" int ss_splat_build_index(const float* geom, const float* dx3, const float* bx3, int nx, int ny, int nz,
                                    int B, long long P, int32_t* coords, int32_t* order, int32_t* voxel_start,
                                    void* ws, size_t ws_bytes, void* stream) {
    SS_REQUIRE(geom && dx3 && bx3 && order && voxel_start && ws, "ss_splat_build_index: null pointer");
    SS_REQUIRE(nx > 0 && ny > 0 && nz > 0 && B > 0 && P > 0, "ss_splat_build_index: shape");
    const long long total = (long long)B * P;
    const long long nvox = (long long)B * nx * ny * nz;
    SS_REQUIRE(total < (1ll << 31) && nvox < (1ll << 31) - 1, "ss_splat_build_index: too many points / voxels");
    if (ws_bytes < ss_splat_index_workspace_bytes(total)) return SS_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const float* h = dx3;          // HOST float[3]
    const float* hb = bx3;         // HOST float[3]
    char* w = static_cast<char*>(ws);
    const size_t seg = align256((size_t)total * sizeof(int32_t));
    int32_t* keys_in = reinterpret_cast<int32_t*>(w);
    int32_t* keys_out = reinterpret_cast<int32_t*>(w + seg);
    int32_t* vals_in = reinterpret_cast<int32_t*>(w + 2 * seg);
    void* temp = w + 3 * seg;
    size_t temp_bytes = sort_temp_bytes(total);
    const int threads = 256;
    splat_rank_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(
        geom, make_float3(h[0], h[1], h[2]), make_float3(hb[0], hb[1], hb[2]), nx, ny, nz, P, total, (int32_t)nvox, coords,
        keys_in, vals_in);
    int rc = check_launch("splat_rank_kernel");
    if (rc) return rc;
    SS_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, order, (int)total, 0,
                                            end_bit_for(nvox), st));
    g_launches.fetch_add(1);
    csr_offsets_kernel<<<(unsigned)((nvox + 1 + threads - 1) / threads), threads, 0, st>>>(keys_out, total, (int32_t)nvox,
                                                                                          voxel_start);
    return check_launch("csr_offsets_kernel");
}

extern "C" int ss_lift_splat_fwd(const float* depth_prob, const float* img_feat, const int32_t* order,
                                 const int32_t* voxel_start, float* out, int B, int D, int H, int W, int C, int nx,
                                 int ny, int nz, void* stream) {
    SS_REQUIRE(depth_prob && img_feat && order && voxel_start && out, "ss_lift_splat_fwd: null pointer");
    SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && nx > 0 && ny > 0 && nz > 0,
               "ss_lift_splat_fwd: shape (C must be a multiple of 4)");
    const long long nvox_b = (long long)nx * ny * nz, nvox = nvox_b * B;
    SS_REQUIRE(nvox < (1ll << 31) && (long long)B * D * H * W < (1ll << 31), "ss_lift_splat_fwd: too large");
    const int warps = 8;
    lift_splat_kernel<<<(unsigned)((nvox + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        depth_prob, img_feat, order, voxel_start, out, D, H * W, C, nvox, (int)nvox_b);
    return check_launch("lift_splat_kernel");
}

extern "C" size_t ss_bev_pool_workspace_bytes(long long n_points, long long n_voxels) {
    return 4 * align256((size_t)n_points * sizeof(int32_t)) + align256((size_t)(n_voxels + 1) * sizeof(int32_t)) +
           align256(sort_temp_bytes(n_points));
}

extern "C" int ss_bev_pool_fwd(const float* feats, const int64_t* coords, long long N, int C, int B, int D, int H,
                               int W, float* out, void* ws, size_t ws_bytes, void* stream) {
    SS_REQUIRE(out && ws, "ss_bev_pool_fwd: null pointer");
    SS_REQUIRE(N >= 0 && C > 0 && C <= 65535 && B > 0 && D > 0 && H > 0 && W > 0, "ss_bev_pool_fwd: shape");
    const long long nvox_b = (long long)D * H * W, nvox = nvox_b * B;
    SS_REQUIRE(N < (1ll << 31) && nvox < (1ll << 31) - 1, "ss_bev_pool_fwd: too many points / voxels");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {   // empty input: all voxels are zero
        SS_CUDA(cudaMemsetAsync(out, 0, (size_t)nvox * C * sizeof(float), st));
        return SS_OK;
    }
    SS_REQUIRE(feats && coords, "ss_bev_pool_fwd: null pointer");
    if (ws_bytes < ss_bev_pool_workspace_bytes(N, nvox)) return SS_ERR_WORKSPACE;
    char* w = static_cast<char*>(ws);
    const size_t seg = align256((size_t)N * sizeof(int32_t));
    int32_t* keys_in = reinterpret_cast<int32_t*>(w);
    int32_t* keys_out = reinterpret_cast<int32_t*>(w + seg);
    int32_t* vals_in = reinterpret_cast<int32_t*>(w + 2 * seg);
    int32_t* order = reinterpret_cast<int32_t*>(w + 3 * seg);
    int32_t* start = reinterpret_cast<int32_t*>(w + 4 * seg);
    void* temp = w + 4 * seg + align256((size_t)(nvox + 1) * sizeof(int32_t));
    size_t temp_bytes = sort_temp_bytes(N);
    const int threads = 256;
    bevpool_rank_kernel<<<(unsigned)((N + threads - 1) / threads), threads, 0, st>>>(coords, N, B, D, H, W, (int32_t)nvox,
                                                                                    keys_in, vals_in);
    int rc = check_launch("bevpool_rank_kernel");
    if (rc) return rc;
    SS_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, order, (int)N, 0,
                                            end_bit_for(nvox), st));
    g_launches.fetch_add(1);
    csr_offsets_kernel<<<(unsigned)((nvox + 1 + threads - 1) / threads), threads, 0, st>>>(keys_out, N, (int32_t)nvox, start);
    rc = check_launch("csr_offsets_kernel");
    if (rc) return rc;
    dim3 grid((unsigned)((nvox + threads - 1) / threads), C);
    bevpool_sum_kernel<<<grid, threads, 0, st>>>(feats, order, start, out, C, nvox_b, nvox);
    return check_launch("bevpool_sum_kernel");
}
