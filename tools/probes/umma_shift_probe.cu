// Hardware probe: does a K-major SWIZZLE_128B UMMA descriptor accept (a) a start address shifted by
// whole 128-byte rows and (b) a stride-byte-offset that is not a multiple of 1024 B?  If the swizzle is a
// function of the absolute shared-memory address, one halo tile in smem can serve all 27 taps of a
// 3x3x3 stencil through shifted descriptors.  Prints the fraction of correct outputs per variant.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(float* out, int shift_rows, int sbo_bytes, int base_off_mode) {
    extern __shared__ unsigned char dyn[];
    unsigned char* buf = (unsigned char*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    float* A = (float*)buf;                       // 256 rows x 128 B, absolute-address swizzle
    float* Bm = (float*)(buf + 256 * 128);        // 16 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 * 32; i += blockDim.x) {
        int r = i / 32, k = i % 32, c = k / 4, e = k % 4;
        A[r * 32 + ((c ^ (r & 7)) * 4) + e] = (float)((r % 64) * 32 + k);
    }
    for (int i = tid; i < 16 * 32; i += blockDim.x) {
        int n = i / 32, k = i % 32, c = k / 4, e = k % 4;
        Bm[n * 32 + ((c ^ (n & 7)) * 4) + e] = (k == n) ? 1.0f : 0.0f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(A) + shift_rows * 128;
        const uint32_t b_addr = smem_u32(Bm);
        uint32_t bo = base_off_mode ? ((a_addr >> 7) & 7) : 0;
        auto desc = [](uint32_t addr, uint32_t sbo, uint32_t baseoff) {
            uint32_t lo = ((addr >> 4) & 0x3FFF) | (1u << 16);
            uint32_t hi = (sbo >> 4) | (1u << 14) | (baseoff << 17) | (2u << 29);
            return ((uint64_t)hi << 32) | lo;
        };
        const uint64_t ad = desc(a_addr, (uint32_t)sbo_bytes, bo), bd = desc(b_addr, 1024, 0);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        for (int k = 0; k < 4; ++k) {
            uint32_t acc = k ? 1u : 0u;
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}"
                         ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid < 128) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)((tid / 32) * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < 16; ++n) out[tid * 16 + n] = __uint_as_float(r[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    float* d; cudaMalloc(&d, 128 * 16 * 4);
    const size_t smem = 1024 + 256 * 128 + 16 * 128;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct V { int shift, sbo, bo; } vs[] = {{0, 1024, 0}, {3, 1024, 0}, {3, 1024, 1}, {0, 1280, 0}, {3, 1280, 0}, {3, 1280, 1}, {11, 1280, 0}, {11, 1280, 1}, {5, 2304, 0}};
    for (auto v : vs) {
        cudaMemset(d, 0, 128 * 16 * 4);
        probe<<<1, 128, smem>>>(d, v.shift, v.sbo, v.bo);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> h(128 * 16);
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int good = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                int r = v.shift + (m / 8) * (v.sbo / 128) + (m % 8);
                float want = (float)((r % 64) * 32 + n);
                good += (h[m * 16 + n] == want);
            }
        printf("shift=%2d sbo=%4d base_off_mode=%d : %4d/2048 correct (%s)  sample d[9][3]=%g want %g\n", v.shift, v.sbo, v.bo, good,
               cudaGetErrorString(e), h[9 * 16 + 3], (float)(((v.shift + 1 * (v.sbo / 128) + 1) % 64) * 32 + 3));
    }
    return 0;
}
