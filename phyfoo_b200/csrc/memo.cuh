// memo.cuh -- "pattern table" path of the order-0 scoring kernels for few species (3 S <= 18 bits).
//
// For an order-0 model the trees of a window do not interact: the free energy of tree t after k sweeps depends only
// on (the column's pattern, t, k).  The window-level stop rule (MeanFieldForBayesNet.java:102-107,204-209) couples
// them only through the SUM over t.  With S = 5 species there are at most 5^5 = 3125 distinct columns, so instead of
// running ~10 sweeps for each of the n_windows * W trees, this path
//   1. marks the column patterns that occur in the data set (once per data set),
//   2. runs the mean field for every (occurring pattern, position) for k = 0..max_sweeps and stores the whole
//      free-energy trajectory F_k           (memo_traj_kernel: the same meanfield_pass as the direct kernel),
//   3. scores every window by walking k: F_k(window) = sum_t traj[t][pattern(c0 + t)][k] in tree order, applying the
//      reference's stop rule to that sum  (memo_window_kernel: pure gathers, no transcendental);
//   4. trajectories are stored up to a sweep cap (default 64 of max_sweeps = 100: the trajectory kernel is a chain of
//      dependent passes, so its time is proportional to the cap); the rare windows whose stop rule has not fired by then
//      are listed on the device and scored by the direct kernel (score_o0_kernel, ITEMS variant) -- same bits either way.
// The per-tree values and the order of the sum over t are those of score_o0_kernel, so scores, sweep counts and
// arg-max calls are bit-identical to the direct path (tests/test_gpu_memo.py compares the two paths bit for bit).
// Exact mode is the same with a one-entry "trajectory" (-log P(column | t)).
#pragma once

struct MemoSet {                 // per (data set): which patterns occur (forward and reverse-complement use)
    int code_bits = 0;           // 3 S
    int n_codes = 0;             // 1 << code_bits
    int cap = 0;                 // upper bound of the number of occurring patterns (host-side, for allocations)
    int32_t* d_dense_of = nullptr;   // [n_codes] dense id or -1
    int32_t* d_code_of = nullptr;    // [cap]
    int32_t* d_count = nullptr;      // [1] number of occurring patterns D (device-resident; kernels read it)
};

// canonical pattern code of a packed column: base bits with 00 under gaps | gap bits << 2S
__device__ __forceinline__ uint32_t memo_code(ColWords cw, int S) {
    uint32_t b = (uint32_t)cw.b & ((1u << (2 * S)) - 1u);
    const uint32_t g = cw.g & ((1u << S) - 1u);
    for (int s = 0; s < S; ++s)
        if ((g >> s) & 1u) b &= ~(3u << (2 * s));
    return b | (g << (2 * S));
}
__device__ __forceinline__ ColWords memo_words(uint32_t code, int S) {
    ColWords cw; cw.b = code & ((1u << (2 * S)) - 1u); cw.g = code >> (2 * S); return cw;
}

__global__ void memo_mark_kernel(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ gaps, int64_t n_cols, int nb, int S,
                                 int32_t* __restrict__ flags) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const ColWords cw = load_column(bases, gaps, n_cols, nb, c);
    flags[memo_code(cw, S)] = 1;
    flags[memo_code(complement(cw), S)] = 1;      // reverse-complement windows read the complemented column
}

// flags -> dense ids (ascending code order), one block; flags array is overwritten with dense_of
__global__ void __launch_bounds__(1024) memo_compact_kernel(int32_t* __restrict__ flags_dense, int n_codes, int32_t* __restrict__ code_of,
                                                            int cap, int32_t* __restrict__ count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_codes; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const int f = (c < n_codes && flags_dense[c]) ? 1 : 0;
        int v = f;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_warp[w];
        const int id = before + v - f;
        if (c < n_codes) flags_dense[c] = f ? id : -1;
        if (f && id < cap) code_of[id] = c;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_base = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = s_base;
}

struct MemoTrajParams {
    const int32_t* code_of; const int32_t* count;
    int W, I, S, E, prog_len, cap, K1, RS;  // K1 = trajectory length (sweep cap + 1, or 1 in exact mode); RS = row stride
    int exact;
    const double* tab;                      // [E][W] (logs for the mean field, probabilities for exact mode)
    const int* prog;
    double* traj;                           // [(t * cap + d) * RS + k], RS a multiple of 4 (32-byte rows)
};
struct MemoTrajJobs { MemoTrajParams job[2]; };   // blockIdx.y selects the job (background and motif in one launch)

// four lanes per (pattern d, position t) -- lane a owns symbol a, meanfield_pass_quad -- : the mean-field trajectory
// F_0..F_cap of that tree.  The grid is small (patterns x positions trees) and every tree is a chain of cap + 1
// dependent passes, so the kernel is latency-bound: four lanes per tree shorten the chain per lane (one exp instead of
// four) and quadruple the warps in flight.  Same bits as the one-thread-per-tree pass of the direct kernel.
__global__ void __launch_bounds__(128) memo_traj_kernel(MemoTrajJobs jobs) {
    extern __shared__ double smem[];
    const MemoTrajParams& p = jobs.job[blockIdx.y];
    const int T = blockDim.x, tid = threadIdx.x;
    const int TPB = T >> 2;                                 // trees per block
    double* s_tab = smem;                                   // [E][W]
    double* s_state = s_tab + (size_t)p.E * p.W;            // [I*8][TPB]
    int* s_prog = (int*)(s_state + (size_t)p.I * 8 * TPB);
    for (int i = tid; i < p.E * p.W; i += T) s_tab[i] = p.tab[i];
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    __syncthreads();
    TreeProg tp;
    tp.n_internal = p.I; tp.n_leaves = p.S; tp.n_nodes = 0;
    tp.par_internal = s_prog;
    tp.node_of = s_prog + p.I;
    tp.leaf_begin = s_prog + 2 * p.I;
    tp.leaf_node = s_prog + 3 * p.I + 1;
    tp.leaf_species = s_prog + 3 * p.I + 1 + p.S;
    const int slot = tid >> 2, a = tid & 3;
    const int qb = (tid & 31) & ~3;
    const unsigned mask = 0xFu << qb;
    const State st{s_state + slot, TPB};
    const int D = *p.count;
    const long long n = (long long)D * p.W;
    // warp-uniform trip count: quads past the end recompute the last tree and do not store
    const long long first = (long long)blockIdx.x * TPB + ((tid >> 5) << 3);     // first tree of this warp
    for (long long it0 = first; it0 < n; it0 += (long long)gridDim.x * TPB) {
        const long long it_raw = it0 + ((tid & 31) >> 2);
        const bool live = it_raw < n;
        const long long it = live ? it_raw : n - 1;
        const int d = (int)(it / p.W), t = (int)(it - (long long)d * p.W);
        const ColWords cw = memo_words((uint32_t)p.code_of[d], p.S);
        const Tab tb{s_tab + t, p.W};
        double* out = p.traj + ((size_t)t * p.cap + d) * p.RS;
        if (p.exact) {
            if (a == 0 && live) out[0] = -log(pruning_likelihood(tp, tb, st, cw));
            continue;
        }
        for (int k = 0; k < p.K1; ++k) {
            const double F = meanfield_pass_quad<true>(tp, tb, st, cw, k > 0, a, qb, mask);
            if (a == 0 && live) out[k] = F;
        }
    }
}

struct __align__(32) MemoRow4 { double v[4]; };
// one 32-byte sector in one request (LDG.E.256 on sm_100a), through the read-only path
__device__ __forceinline__ MemoRow4 memo_load4(const MemoRow4* p) {
    MemoRow4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}

struct MemoWindowParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, S, cap, K1, RS;
    const int32_t* dense_of; const double* traj;
    double* out; int32_t* sweeps; unsigned long long* counter_sweeps;
    int max_sweeps, min_sweeps; double tol; int exact;
    // windows whose stop rule has not fired within the stored trajectory go to the direct kernel
    int* pend_count; int64_t* pend_col0; uint8_t* pend_strand; int64_t* pend_dst;
};

// one thread per window: walk k in steps of four (one 32-byte sector per tree and step), summing the W trajectories in
// tree order, until the reference's stop rule fires
__global__ void __launch_bounds__(256) memo_window_kernel(MemoWindowParams p) {
    const long long n_items = (long long)p.n_windows * p.n_strands;
    unsigned long long ksum = 0;
    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (long long)gridDim.x * blockDim.x) {
        const int si = (int)(item / p.n_windows);
        const int64_t w = item - (long long)si * p.n_windows;
        const int strand = p.n_strands == 2 ? si : p.strand0;
        const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
        const int64_t c0 = p.col_off[a] + (w - p.win_off[a]);
        const MemoRow4* row[16];
        // W <= 16 here (the launcher falls back to the direct kernel otherwise): trajectory rows of the W trees
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            if (t < p.W) {
                const int64_t c = strand ? c0 + (p.W - 1 - t) : c0 + t;       // jstacs StrandModel.java:636-639
                ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                if (strand) cw = complement(cw);
                const int d = p.dense_of[memo_code(cw, p.S)];
                row[t] = (const MemoRow4*)(p.traj + ((size_t)t * p.cap + d) * p.RS);
            }
        }
        double Fsum = 0.0, old = 0.0;
        int k = 0;                     // sweeps done
        bool conv = false, done = false, pending = false;
        for (int kb = 0; !done; kb += 4) {
            if (kb >= p.K1) { pending = true; break; }                         // trajectory exhausted
            double f[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if (t < p.W) {
                    const MemoRow4 r = memo_load4(&row[t][kb >> 2]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) f[j] += r.v[j];                // tree-major, like calcFreeEnergy(0, W-1)
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (done || pending) continue;
                const int kk = kb + j;                                         // index into the trajectory
                if (kk >= p.K1) { pending = true; continue; }
                if (kk == 0) {
                    Fsum = old = f[0];                                         // F_0, MeanFieldForBayesNet.java:105
                    if (p.exact) done = true;
                } else {
                    Fsum = f[j]; ++k;
                    if (fabs(old - Fsum) < p.tol) conv = true;                 // :204-207 (sticky)
                    old = Fsum;
                }
                if (!((k < p.max_sweeps && !conv) || k < p.min_sweeps)) done = true;   // :107
            }
            if (pending) break;
        }
        if (pending) {
            const int idx = atomicAdd(p.pend_count, 1);
            p.pend_col0[idx] = c0; p.pend_strand[idx] = (uint8_t)strand; p.pend_dst[idx] = item;
            continue;
        }
        p.out[item] = -Fsum;                                                   // PhyloBayesModel.java:259
        if (p.sweeps) p.sweeps[item] = k;
        ksum += (unsigned long long)k;
    }
    // sweep statistics: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) ksum += __shfl_down_sync(0xffffffffu, ksum, o);
    if ((threadIdx.x & 31) == 0 && ksum) atomicAdd(p.counter_sweeps, ksum);
}
