// NVLink peer-memory collectives of the X-slab sharded mode (stereoscene_b200/xshard.py): the halo
// exchange between neighbouring slabs and the all-reduce of the GroupNorm sums, written as plain kernels that store straight
// into the peers' memory and synchronise through flag words -- no NCCL launch, no host round trip, capturable in a CUDA graph.
//
// Memory model.  Every rank owns one pool (cudaMalloc) that it exports through CUDA IPC; all ranks open all pools, so a rank
// holds a device pointer to every peer's pool.  Tensors that are exchanged live at the SAME offset in every pool (the ranks
// run the same allocation sequence).  A collective call gets a private flag word (and, for the all-reduce, a private slot
// area) at a fixed offset; flags are never reset: a call of forward number `epoch` (a device counter, bumped once per
// forward) waits until the flag holds a value >= epoch, and forward numbers only grow.
//
// Ordering.  Data stores to the peer are followed by __threadfence_system() and a release of the flag by the last CTA
// (ticket counter); the waiting thread polls the flag with a volatile load and fences before the kernel ends, so every later
// kernel of the stream sees the neighbour's planes.  A rank can only run ahead of its neighbour by one collective (each call
// waits for the neighbour's call of the same number), and buffers are not reused inside a forward, so a push never lands in
// memory the peer is still reading.
#include <cuda_runtime.h>
#include <cstring>
#include "common.cuh"

namespace ss {

__device__ __forceinline__ void flag_release(volatile int* flag, int value) {
    __threadfence_system();
    *flag = value;
    __threadfence_system();
}
__device__ __forceinline__ void flag_wait(const volatile int* flag, int value) {
    while (*flag < value) { __nanosleep(64); }
    __threadfence_system();
}

struct HaloPushParams {
    const float4* src_lo;      // my plane 1   -> lower neighbour's plane n+1      (nullptr: no lower neighbour)
    const float4* src_hi;      // my plane n   -> upper neighbour's plane 0        (nullptr: no upper neighbour)
    float4* dst_lo;            // lower neighbour's plane n+1 (peer memory)
    float4* dst_hi;            // upper neighbour's plane 0   (peer memory)
    float4* edge_lo;           // my plane 0   (filled locally when there is no lower neighbour; nullptr: leave)
    float4* edge_hi;           // my plane n+1 (filled locally when there is no upper neighbour)
    long long plane_vec;       // float4 per plane
    int edge_replicate;        // 1: copy the edge plane into the outer halo, 0: zero it
    // flag words of one call site, int[4] at the same pool offset on every rank:
    //   [0] my lower neighbour is READY (its buffer is quiescent)   [1] my upper neighbour is READY
    //   [2] my lower neighbour's plane has ARRIVED in my plane 0     [3] my upper neighbour's plane has ARRIVED in my plane n+1
    volatile int* mine;
    volatile int* lo;          // lower neighbour's flag words (peer memory) or nullptr
    volatile int* hi;          // upper neighbour's
    unsigned int* ticket;      // CTA ticket counter (zero on entry, reset by the last CTA)
    const int* epoch;
};

// Two handshakes per exchange.  READY: a rank's producer kernel also writes (garbage) into the halo planes of its own output, so a
// neighbour may only push once this rank's stream has reached its own exchange call -- i.e. once the producer has finished.
// ARRIVED: the consumer kernel of this rank may only start once both neighbours' planes have landed.
__global__ void halo_push_kernel(const HaloPushParams p) {
    const int e = *p.epoch;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {                              // my buffer is quiescent: tell both neighbours
            if (p.lo) flag_release(p.lo + 1, e);            //   I am the lower neighbour's UPPER neighbour
            if (p.hi) flag_release(p.hi + 0, e);
        }
        if (p.lo) flag_wait(p.mine + 0, e);
        if (p.hi) flag_wait(p.mine + 1, e);
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p.src_lo)
        for (long long i = i0; i < p.plane_vec; i += stride) p.dst_lo[i] = __ldg(p.src_lo + i);
    else if (p.edge_lo)
        for (long long i = i0; i < p.plane_vec; i += stride) p.edge_lo[i] = p.edge_replicate ? p.edge_lo[i + p.plane_vec] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.src_hi)
        for (long long i = i0; i < p.plane_vec; i += stride) p.dst_hi[i] = __ldg(p.src_hi + i);
    else if (p.edge_hi)
        for (long long i = i0; i < p.plane_vec; i += stride) p.edge_hi[i] = p.edge_replicate ? p.edge_hi[i - p.plane_vec] : make_float4(0.f, 0.f, 0.f, 0.f);
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        if (p.lo) flag_release(p.lo + 3, e);                // my plane 1 sits in the lower neighbour's plane n+1
        if (p.hi) flag_release(p.hi + 2, e);
        if (p.lo) flag_wait(p.mine + 2, e);
        if (p.hi) flag_wait(p.mine + 3, e);
        *p.ticket = 0u;
    }
}

constexpr int kMaxPeers = 16;
struct StatsReduceParams {
    double* stats;                      // my sums, n doubles: in = local, out = sum over ranks
    int n, world, rank;
    double* slots[kMaxPeers];           // slots[r] = rank r's slot area (world x n doubles), peer memory for r != rank
    volatile int* flags[kMaxPeers];     // flags[r] = rank r's flag array (world ints): flags[r][me] = "rank me has written its sums"
    const int* epoch;
};

// one CTA: write my sums into slot[me] of every rank, raise my flag at every rank, wait for everyone's flag here, add up in
// rank order (the same order on every rank, so all ranks finalise bit-identical statistics)
__global__ void stats_allreduce_kernel(const StatsReduceParams p) {
    const int e = *p.epoch;
    for (int r = 0; r < p.world; ++r)
        for (int i = threadIdx.x; i < p.n; i += blockDim.x) p.slots[r][(size_t)p.rank * p.n + i] = p.stats[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < p.world) flag_release(p.flags[threadIdx.x] + p.rank, e);
    if (threadIdx.x < p.world) flag_wait(p.flags[p.rank] + threadIdx.x, e);
    __syncthreads();
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < p.world; ++r) s += ((volatile double*)p.slots[p.rank])[(size_t)r * p.n + i];
        p.stats[i] = s;
    }
}

__global__ void epoch_bump_kernel(int* epoch) { *epoch += 1; }

}  // namespace ss

using namespace ss;

// ---- pool management (host) ---------------------------------------------------------------------------------------------
extern "C" int ss_peer_pool_alloc(size_t bytes, void** ptr) {
    SS_REQUIRE(ptr && bytes > 0, "ss_peer_pool_alloc: bad arguments");
    SS_CUDA(cudaMalloc(ptr, bytes));
    SS_CUDA(cudaMemset(*ptr, 0, bytes));
    return SS_OK;
}
extern "C" int ss_peer_pool_free(void* ptr) {
    SS_CUDA(cudaFree(ptr));
    return SS_OK;
}
extern "C" int ss_peer_ipc_export(void* ptr, void* handle64) {
    SS_REQUIRE(ptr && handle64, "ss_peer_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SS_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
    return SS_OK;
}
extern "C" int ss_peer_ipc_open(const void* handle64, int peer_device, void** ptr) {
    SS_REQUIRE(handle64 && ptr, "ss_peer_ipc_open: null pointer");
    int dev = 0;
    SS_CUDA(cudaGetDevice(&dev));
    if (peer_device != dev) {
        int can = 0;
        SS_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
        SS_REQUIRE(can, "ss_peer_ipc_open: no peer access between the two devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_cuda_error(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    SS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SS_OK;
}
extern "C" int ss_peer_ipc_close(void* ptr) {
    SS_CUDA(cudaIpcCloseMemHandle(ptr));
    return SS_OK;
}

// ---- collectives --------------------------------------------------------------------------------------------------------
extern "C" int ss_peer_epoch_bump(int* epoch, void* stream) {
    SS_REQUIRE(epoch, "ss_peer_epoch_bump: null pointer");
    epoch_bump_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(epoch);
    return check_launch("epoch_bump_kernel");
}

extern "C" int ss_peer_halo_push(float* buf, float* peer_lo_buf, float* peer_hi_buf, long long plane_floats, int n, int edge_replicate,
                                 int* my_flags, int* peer_lo_flags, int* peer_hi_flags, unsigned int* ticket, const int* epoch,
                                 void* stream) {
    SS_REQUIRE(buf && my_flags && ticket && epoch && n >= 1, "ss_peer_halo_push: bad arguments");
    SS_REQUIRE(plane_floats % 4 == 0 && (reinterpret_cast<uintptr_t>(buf) & 15) == 0, "ss_peer_halo_push: planes must be 16-byte multiples");
    SS_REQUIRE((peer_lo_buf == nullptr) == (peer_lo_flags == nullptr) && (peer_hi_buf == nullptr) == (peer_hi_flags == nullptr),
               "ss_peer_halo_push: a neighbour needs both its buffer and its flags");
    HaloPushParams p;
    const long long pv = plane_floats / 4;
    float4* b4 = reinterpret_cast<float4*>(buf);
    p.plane_vec = pv;
    p.src_lo = peer_lo_buf ? b4 + pv : nullptr;
    p.src_hi = peer_hi_buf ? b4 + (long long)n * pv : nullptr;
    p.dst_lo = peer_lo_buf ? reinterpret_cast<float4*>(peer_lo_buf) + (long long)(n + 1) * pv : nullptr;
    p.dst_hi = peer_hi_buf ? reinterpret_cast<float4*>(peer_hi_buf) : nullptr;
    p.edge_lo = peer_lo_buf ? nullptr : b4;
    p.edge_hi = peer_hi_buf ? nullptr : b4 + (long long)(n + 1) * pv;
    p.edge_replicate = edge_replicate;
    p.mine = my_flags;
    p.lo = peer_lo_flags;
    p.hi = peer_hi_flags;
    p.ticket = ticket;
    p.epoch = epoch;
    const int threads = 256;
    long long want = (pv + threads - 1) / threads;
    const int blocks = (int)(want < 1 ? 1 : (want > 64 ? 64 : want));   // a few dozen CTAs saturate one NVLink direction for ~1 MB planes
    halo_push_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("halo_push_kernel");
}

extern "C" int ss_peer_stats_allreduce(double* stats, int n, int world, int rank, double* const* slots, int* const* flags,
                                       const int* epoch, void* stream) {
    SS_REQUIRE(stats && slots && flags && epoch && n > 0, "ss_peer_stats_allreduce: bad arguments");
    SS_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "ss_peer_stats_allreduce: world size");
    StatsReduceParams p;
    p.stats = stats; p.n = n; p.world = world; p.rank = rank; p.epoch = epoch;
    for (int r = 0; r < world; ++r) { p.slots[r] = slots[r]; p.flags[r] = flags[r]; }
    stats_allreduce_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("stats_allreduce_kernel");
}
