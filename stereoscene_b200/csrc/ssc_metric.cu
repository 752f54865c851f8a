// Semantic-scene-completion scores of a predicted label volume against the ground truth: the step after
// the hot path (trilinear x2 + argmax -> labels; reference: projects/mmdet3d_plugin/utils/ssc_metric.py
// :62-85 `update`, :109-168 the two score functions; call site occupancy/apis/test.py:113-115).
//
// The reference loops over the 20 classes with three full-volume boolean reductions each (plus host-side
// masking copies).  Here one pass builds the C x C confusion matrix of the remapped labels and the three
// completion counts; tp/fp/fn per class are row/column sums of that matrix (host side, 400 numbers).
//   remap (ssc_metric.py:113-114, 147-148): where target == ignore both prediction and target become 0
//   (so ignored voxels land in cell [0][0] of the semantic matrix, exactly like the reference, whose
//   in-place edit of y_true makes its second `y_true != 255` mask all-true);
//   completion (ssc_metric.py:109-141): over voxels with target != ignore (and the optional masks),
//   occupied = label > 0:  tp = true & pred, fp = !true & pred, fn = true & !pred;
//   semantic  (ssc_metric.py:143-168): over voxels selected by `nonempty` only.
#include "common.cuh"

namespace ss {

constexpr int SSC_MAX_C = 32;
constexpr int SSC_THREADS = 256;

template <typename TT>
__global__ void __launch_bounds__(SSC_THREADS)
ssc_confusion_kernel(const uint8_t* __restrict__ pred, const TT* __restrict__ target, const uint8_t* __restrict__ nonempty,
                     const uint8_t* __restrict__ nonsurface, long long n, int C, int ignore,
                     unsigned long long* __restrict__ counts) {
    __shared__ unsigned int hist[SSC_MAX_C * SSC_MAX_C + 3 + SSC_MAX_C];
    for (int i = threadIdx.x; i < C * C + 3 + C; i += SSC_THREADS) hist[i] = 0u;
    __syncthreads();
    unsigned int ctp = 0, cfp = 0, cfn = 0;
    for (long long i = (long long)blockIdx.x * SSC_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * SSC_THREADS) {
        const int t0 = (int)target[i];
        int pr = (int)pred[i];
        const bool ign = (t0 == ignore);
        const bool ne = nonempty ? nonempty[i] != 0 : true;
        const bool ns = nonsurface ? nonsurface[i] != 0 : true;
        const int t = ign ? 0 : t0;
        if (ign) pr = 0;
        if (!ign && ne && ns) {
            const bool bt = t > 0, bp = pr > 0;
            ctp += (bt && bp); cfp += (!bt && bp); cfn += (bt && !bp);
        }
        if (ne && (unsigned)t < (unsigned)C) {
            if ((unsigned)pr < (unsigned)C) atomicAdd(&hist[t * C + pr], 1u);
            else atomicAdd(&hist[C * C + 3 + t], 1u);       // prediction outside the class range: a miss (fn) of the target class,
        }                                                   // a false positive of none (ssc_metric.py:157-163 counts it the same way)
    }
    // completion counts: warp reduce, then one shared atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ctp += __shfl_xor_sync(0xffffffffu, ctp, o);
        cfp += __shfl_xor_sync(0xffffffffu, cfp, o);
        cfn += __shfl_xor_sync(0xffffffffu, cfn, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&hist[C * C + 0], ctp); atomicAdd(&hist[C * C + 1], cfp); atomicAdd(&hist[C * C + 2], cfn);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C + 3 + C; i += SSC_THREADS)
        if (hist[i]) atomicAdd(counts + i, (unsigned long long)hist[i]);
}

}  // namespace ss

extern "C" int ss_ssc_confusion_fwd(const uint8_t* pred, const void* target, int target_elem_bytes, const uint8_t* nonempty,
                                    const uint8_t* nonsurface, long long n, int C, int ignore_label, long long* counts,
                                    void* stream) {
    using namespace ss;
    SS_REQUIRE(n >= 0 && C > 0 && C <= SSC_MAX_C, "ss_ssc_confusion_fwd: shape (C <= 32)");
    SS_REQUIRE(target_elem_bytes == 1 || target_elem_bytes == 8, "ss_ssc_confusion_fwd: target must be uint8 or int64");
    SS_REQUIRE(counts, "ss_ssc_confusion_fwd: null pointer");
    if (n == 0) return SS_OK;                               // empty volume: nothing to count
    SS_REQUIRE(pred && target, "ss_ssc_confusion_fwd: null pointer");
    long long blocks = (n + (long long)SSC_THREADS * 16 - 1) / ((long long)SSC_THREADS * 16);   // >= 16 voxels per thread
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
    if (target_elem_bytes == 1)
        ssc_confusion_kernel<uint8_t><<<(unsigned)blocks, SSC_THREADS, 0, st>>>(pred, static_cast<const uint8_t*>(target), nonempty,
                                                                               nonsurface, n, C, ignore_label, cnt);
    else
        ssc_confusion_kernel<long long><<<(unsigned)blocks, SSC_THREADS, 0, st>>>(pred, static_cast<const long long*>(target), nonempty,
                                                                                 nonsurface, n, C, ignore_label, cnt);
    return check_launch("ssc_confusion_kernel");
}
