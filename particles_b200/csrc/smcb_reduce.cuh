// smcb_reduce.cuh -- grid-wide deterministic reduction ("last block done").
#pragma once
#include "smcb_common.cuh"

namespace smcb {

// Every block deposits K partial triples; the block that draws the last ticket merges
// all partials in a FIXED order (so the result does not depend on block scheduling)
// and returns true with the totals valid in thread 0.  The ticket counter wraps back
// to 0 (atomicInc modulo), so it never needs a memset between launches.
template <int BLOCK, int K>
__device__ __forceinline__ bool grid_merge_lse3(Lse3 (&mine)[K], double *partials /* grid x 4K */,
                                                unsigned int *ticket, Lse3 *smem,
                                                Lse3 (&total)[K]) {
    __shared__ bool s_last;
#pragma unroll
    for (int j = 0; j < K; j++) mine[j] = lse3_block_reduce<BLOCK>(mine[j], smem);
    if (threadIdx.x == 0) {
        double *p = partials + (size_t)blockIdx.x * 4 * K;
#pragma unroll
        for (int j = 0; j < K; j++) {
            p[4 * j + 0] = mine[j].m; p[4 * j + 1] = mine[j].s; p[4 * j + 2] = mine[j].q;
        }
        __threadfence();
        unsigned int tk = atomicInc(ticket, gridDim.x - 1);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
#pragma unroll
    for (int j = 0; j < K; j++) {
        Lse3 acc = lse3_empty();
        for (int k = threadIdx.x; k < (int)gridDim.x; k += BLOCK) {
            const volatile double *p = partials + (size_t)k * 4 * K + 4 * j;
            acc = lse3_merge(acc, Lse3{p[0], p[1], p[2]});
        }
        total[j] = lse3_block_reduce<BLOCK>(acc, smem);
    }
    return true;
}

// Weights.__init__ scalars from the merged triple (resampling.py:217-226):
//   log_mean = m + log(s / N);  ESS = 1 / sum (w/s)^2 = s^2 / q
// All -inf (m == -inf) or any +inf (m == +inf) give NaN everywhere, as NumPy does.
__device__ __forceinline__ void weights_scalars(const Lse3 &a, double n, double &log_mean,
                                                double &ess) {
    if (a.m == -CUDART_INF || a.m == CUDART_INF || a.m != a.m) {
        log_mean = CUDART_NAN;
        ess = CUDART_NAN;
        return;
    }
    log_mean = a.m + log(a.s / n);
    ess = (a.s * a.s) / a.q;
}

}  // namespace smcb
