// k_stat_depth.cuh — Stats::statDepth (stats.cpp:56-83) for a batch of reads: the bases of every mapped read added to the
// genome-scale depth bins of `step` bases (Stats::mGenomeDepth, stats.cpp:40-46: 1 + target_len / step bins per contig), one
// THREAD per read, 64-bit atomics on the bins (the depth half of Stats).
#pragma once

#include "device_common.cuh"

namespace gcb {

constexpr int DEPTH_THREADS = 256;

__global__ void __launch_bounds__(DEPTH_THREADS) stat_depth_kernel(const int32_t *tid, const int32_t *pos, const int32_t *len, int64_t n, int32_t step,
                                                                    const int64_t *bin_off, int32_t n_targets, unsigned long long *depth) {
    GCB_GRID_DEP();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = tid[i];
    if (t >= n_targets || t < 0) return;  // stats.cpp:60-61
    const int start = pos[i], l = len[i];
    const int end = start + l;
    const int leftPos = start / step, rightPos = end / step;  // (C division: a start in (-step, 0) lands in bin 0, as in the reference)
    const int64_t bins = bin_off[t + 1] - bin_off[t];
    if ((int64_t)rightPos >= bins || leftPos < 0) return;  // stats.cpp:68-69
    unsigned long long *d = depth + bin_off[t];
    if (leftPos == rightPos) {
        atomicAdd(d + leftPos, (unsigned long long)(long long)l);
    } else {
        const int leftLen = (leftPos + 1) * step - start, rightLen = end - rightPos * step;
        atomicAdd(d + leftPos, (unsigned long long)(long long)leftLen);
        atomicAdd(d + rightPos, (unsigned long long)(long long)rightLen);
        for (int p = leftPos + 1; p < rightPos; p++) atomicAdd(d + p, (unsigned long long)step);
    }
}

}  // namespace gcb
