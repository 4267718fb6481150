// score_common.cuh -- what the scoring kernels share and what the run-time compiled kernels (jit_o0.hpp) need as well: the
// packed-column loader, the alignment search, the parameter block of the F81-structured order-0 kernel and the named barrier
// of its window groups.  Device code without host includes (compiles under NVRTC; included by phyfoo.cu after tree_pass.cuh).
#pragma once

__device__ __forceinline__ phyfoo::ColWords load_column(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ gaps,
                                                int64_t n_cols, int nb, int64_t c) {
    phyfoo::ColWords cw;
    cw.b = bases[c];
    if (nb > 1) cw.b |= (uint64_t)bases[n_cols + c] << 32;
    cw.g = gaps[c];
    return cw;
}

// largest a with off[a] <= w  (off has n+1 entries, off[0] = 0, w < off[n])
__device__ __forceinline__ int find_alignment(const int64_t* __restrict__ off, int n, int64_t w) {
    int lo = 0, hi = n;  // invariant: off[lo] <= w < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= w) lo = mid; else hi = mid;
    }
    return lo;
}

// the same with a first guess from the average alignment length (equal-length alignments: no search at all)
__device__ __forceinline__ int find_alignment_hint(const int64_t* __restrict__ off, int n, int64_t w, int64_t n_windows) {
    int a = (int)(((double)w * (double)n) / (double)n_windows);
    a = min(max(a, 0), n - 1);
    if (off[a] > w) {                              // guess too far right: gallop left, then bisect
        int step = 1, hi = a;
        while (a > 0 && off[a] > w) { hi = a; a = max(0, a - step); step <<= 1; }
        while (hi - a > 1) { const int mid = (a + hi) >> 1; if (off[mid] <= w) a = mid; else hi = mid; }
    } else if (off[a + 1] <= w) {                  // too far left
        int step = 1, lo = a + 1;
        a = lo;
        while (a < n - 1 && off[a + 1] <= w) { lo = a + 1; a = min(n - 1, a + step); step <<= 1; }
        int hi = a + 1;                            // off[lo] <= w < off[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= w) lo = mid; else hi = mid; }
        a = lo;
    }
    return a;
}

struct ScoreFParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, I, S, Ef, prog_len;
    const double* tab;       // compact tables [Ef][W]
    const int* prog;
    double* out; int32_t* sweeps;
    unsigned long long* counter_next; unsigned long long* counter_sweeps;
    int max_sweeps, min_sweeps; double tol;
    double* gstate;          // nullable: per-thread state in global memory instead of shared
    int n_slots;             // internal nodes with internal children (StateF)
    int group_lanes;         // lanes of a group (a multiple of 32); windows per group = group_lanes / W
    int n_groups;            // groups per block
};

__device__ __forceinline__ int named_bar_or(int id, int n_threads, bool pred) {
    int r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %3, 0;\n\tbar.red.or.pred q, %1, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"(id), "r"(n_threads), "r"((int)pred) : "memory");
    return r;
}

