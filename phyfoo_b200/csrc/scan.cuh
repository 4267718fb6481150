// scan.cuh -- genome-scan / classification front-end on the device: consumers of the
// per-offset motif scores that must not ship them to the host first (config 5: 2*10^9 window scores = 16 GB).
//
// Reference semantics (classification/ClassificationUtil.java):
//   determineThresholdsForSpecifitiesPerPos :244-264   threshold = sorted negative scores[(int)((n - 1) * specificity)]
//   determineThresholdForSpecifity          :161-180   threshold = sorted scores[(int)(n * specificity)]
//   determineSensitivityPerSeq              :266-278   fraction of alignments with at least one score > threshold
//   determinePositionSet                    :287-298   every (alignment, offset) with score > threshold
//   getScores                               :191-219   = phyfoo_score_windows
#pragma once
#include <cub/device/device_radix_sort.cuh>

// one block per alignment: number of offsets with score > threshold, the first such offset (-1 if none), the maximal
// score and its first offset (util/Util.java:528-538: strict '>' => first maximum)
__global__ void __launch_bounds__(256) scan_hits_kernel(const double* __restrict__ scores, const int64_t* __restrict__ win_off, double thr,
                                                        int32_t* __restrict__ n_hits, int32_t* __restrict__ first_hit,
                                                        double* __restrict__ max_score, int32_t* __restrict__ arg_max) {
    __shared__ int s_cnt[8], s_first[8], s_arg[8];
    __shared__ double s_max[8];
    const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int64_t w0 = win_off[a];
    const int n = (int)(win_off[a + 1] - w0);
    int cnt = 0, first = 0x7fffffff, arg = 0x7fffffff;
    double best = -INFINITY;
    for (int j = tid; j < n; j += T) {
        const double v = scores[w0 + j];
        if (v > thr) { ++cnt; first = min(first, j); }
        if (v > best) { best = v; arg = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        first = min(first, __shfl_down_sync(0xffffffffu, first, o));
        const double ov = __shfl_down_sync(0xffffffffu, best, o);
        const int oa = __shfl_down_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
    }
    if ((tid & 31) == 0) { s_cnt[tid >> 5] = cnt; s_first[tid >> 5] = first; s_max[tid >> 5] = best; s_arg[tid >> 5] = arg; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < (T + 31) / 32; ++i) {
            cnt += s_cnt[i]; first = min(first, s_first[i]);
            if (s_max[i] > best || (s_max[i] == best && s_arg[i] < arg)) { best = s_max[i]; arg = s_arg[i]; }
        }
        if (n_hits) n_hits[a] = cnt;
        if (first_hit) first_hit[a] = first == 0x7fffffff ? -1 : first;
        if (max_score) max_score[a] = best;
        if (arg_max) arg_max[a] = arg == 0x7fffffff ? (n > 0 ? 0 : -1) : arg;
    }
}

__global__ void scan_pick_kernel(const double* __restrict__ sorted, const int64_t* __restrict__ idx, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sorted[idx[i]];
}
