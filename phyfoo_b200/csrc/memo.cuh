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
    int32_t* d_flags = nullptr;      // [n_codes] 1 where a column of the data set has this code (forward strand only)
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
    flags[memo_code(load_column(bases, gaps, n_cols, nb, c), S)] = 1;     // the complemented column: memo_compact_kernel
}

// code of the complemented column (reverse-complement windows read it): complement is an involution on the canonical
// codes (base bits 00 under gaps)
__device__ __forceinline__ uint32_t memo_comp_code(uint32_t code, int S) {
    return memo_code(complement(memo_words(code, S)), S);
}
__device__ __forceinline__ bool memo_occurs(const int32_t* __restrict__ flags, uint32_t c, int S) {
    const uint32_t mask = (1u << (2 * S)) - 1u, b = c & mask, g = c >> (2 * S);
    uint32_t under = 0;                                                   // base bits that sit under a gap
    for (int s = 0; s < S; ++s) under |= ((g >> s) & 1u) ? 3u << (2 * s) : 0u;
    if (b & under) return false;                                          // not a canonical code: no column has it
    const uint32_t comp = ((~b) & mask & ~under) | (g << (2 * S));        // = memo_comp_code(c, S)
    return (flags[c] | flags[comp]) != 0;
}

// flags -> dense ids in ascending code order; a code occurs if it or its complement was marked.  One block: every warp
// owns a segment of the codes (coalesced reads, ballots instead of block-wide scans); two block barriers in total.
__global__ void __launch_bounds__(1024) memo_compact_kernel(const int32_t* __restrict__ flags, int n_codes, int S,
                                                            int32_t* __restrict__ dense_of, int32_t* __restrict__ code_of,
                                                            int cap, int32_t* __restrict__ count) {
    __shared__ int s_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, n_warps = blockDim.x >> 5;
    const int seg = (((n_codes + n_warps - 1) / n_warps) + 31) & ~31;
    const int lo = min(w * seg, n_codes), hi = min(lo + seg, n_codes);
    // (a segment of up to 1024 codes is 32 rounds: the lane keeps its 32 answers in a register for the second pass)
    const bool keep = hi - lo <= 1024;
    unsigned mine_bits = 0;
    int total = 0, round = 0;
    for (int base = lo; base < hi; base += 32, ++round) {
        const int c = base + lane;
        const bool f = c < hi && memo_occurs(flags, (uint32_t)c, S);
        if (keep && f) mine_bits |= 1u << round;
        total += __popc(__ballot_sync(0xffffffffu, f));
    }
    if (lane == 0) s_tot[w] = total;
    __syncthreads();
    if (w == 0) {
        const int mine = lane < n_warps ? s_tot[lane] : 0;
        int v = mine;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        s_tot[lane] = v - mine;                 // exclusive
        if (lane == 31) *count = v;
    }
    __syncthreads();
    int running = s_tot[w];
    round = 0;
    for (int base = lo; base < hi; base += 32, ++round) {
        const int c = base + lane;
        const bool f = keep ? ((mine_bits >> round) & 1u) != 0 : (c < hi && memo_occurs(flags, (uint32_t)c, S));
        const unsigned m = __ballot_sync(0xffffffffu, f);
        const int id = running + __popc(m & ((1u << lane) - 1u));
        if (c < hi) dense_of[c] = f ? id : -1;
        if (f && id < cap) code_of[id] = c;
        running += __popc(m);
    }
}

struct MemoTrajParams {
    const int32_t* code_of; const int32_t* count;
    int W, I, S, E, prog_len, cap, K1, RS;  // K1 = trajectory length (sweep cap + 1, or 1 in exact mode); RS = row stride
    int exact;
    const double* tab;                      // [E][W] (logs for the mean field, probabilities for exact mode)
    const int* prog;
    double* traj;                           // [(t * cap + d) * RS + k], RS a multiple of 4 (32-byte rows)
    // passes k0 .. k1-1 of the trajectory in this launch; a launch with k0 > 0 resumes from the state a previous launch
    // saved (state: [I * 8][D * W] doubles), a launch with k1 < K1 saves it
    int k0, k1;
    double* state;
    // wavefront variant (memo_traj_wave_kernel): node i of sweep k runs at step 2 k + depth(i); sched[parity][slot] = the
    // internal node a node slot works on at steps of that parity (-1: none)
    int dmax, ML, MCH;                      // deepest internal node, most leaf / internal children of one node
    signed char sched[2][4];
    signed char depth[8];
};
struct MemoTrajJobs { MemoTrajParams job[2]; };   // blockIdx.y selects the job (background and motif in one launch)

// four lanes per (pattern d, position t) -- lane a owns symbol a, meanfield_pass_quad -- : the mean-field trajectory
// F_0..F_cap of that tree.  The grid is small (patterns x positions trees) and every tree is a chain of cap + 1
// dependent passes, so the kernel is latency-bound: four lanes per tree shorten the chain per lane (one exp instead of
// four) and quadruple the warps in flight.  Same bits as the one-thread-per-tree pass of the direct kernel.
__global__ void __launch_bounds__(128) memo_traj_kernel(MemoTrajJobs jobs) {
    extern __shared__ double smem[];
    const MemoTrajParams& p = jobs.job[blockIdx.y];
    if (p.k0 >= p.k1) return;                               // this job has no passes in this phase
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
        if (p.k0 > 0) {
            for (int e = a; e < p.I * 8; e += 4) s_state[(size_t)e * TPB + slot] = p.state[(size_t)e * n + it];
            __syncwarp();
        }
        for (int k = p.k0; k < p.k1; ++k) {
            const double F = meanfield_pass_quad<true>(tp, tb, st, cw, k > 0, a, qb, mask);
            if (a == 0 && live) out[k] = F;
        }
        if (p.k1 < p.K1 && live)
            for (int e = a; e < p.I * 8; e += 4) p.state[(size_t)e * n + it] = s_state[(size_t)e * TPB + slot];
    }
}

// The same trajectories as a WAVEFRONT over (node, sweep): in the reference's sweep order a node needs its parent of the
// same sweep and its children of the previous sweep, so node i of sweep k can run at step tau = 2 k + depth(i) and nodes
// of alternating depth -- belonging to consecutive sweeps -- run side by side.  LPT lanes per tree = node slots x four
// symbols; the chain of dependent node updates shrinks from I per sweep to 2 per sweep.  Per node the arithmetic is that
// of meanfield_pass_quad (CANONICAL ARITHMETIC in tree_pass.cuh): children leave their message in a slot of their own
// and the parent adds the slots in child order to its base, which is the order `cs(p) += m` produces in the sequential
// pass; the per-node free-energy terms of a sweep are kept until its deepest node is done and then summed in node order.
// Same bits as memo_traj_kernel and as the direct kernel (tests/test_gpu_memo.py).
// EARLY STOP (launches that cover the whole trajectory): the sweep map is deterministic, so once the q of every node
// equals, bit for bit, the q of the previous sweep (fixed point) or of the sweep before that (two-cycle in the last
// bits), the rest of the trajectory repeats the last two values; a warp stops when its trees all got there and fills
// the rows.  (C2: 31 of 65 passes per warp on average.)
// Per-tree shared-memory block: q[I][4] base[I][4] msg[I][4] q2[I][4] ring[4 sweeps][I][2] flags, padded to 4 (mod 16)
// doubles so that the quads of the trees of a warp fall into different banks.
// IT: the number of internal nodes when it is known at compile time (0: read from the parameters) -- with it the strides
// of the per-tree block are immediates instead of multiplications that the register cap makes the compiler redo per step
template <int LPT, int IT>
__global__ void __launch_bounds__(128, 6) memo_traj_wave_kernel(MemoTrajJobs jobs) {
    extern __shared__ double smem[];
    const MemoTrajParams& p = jobs.job[blockIdx.y];
    if (p.k0 >= p.k1) return;
    const int T = blockDim.x, tid = threadIdx.x, I = IT ? IT : p.I;
    constexpr int TPW = 32 / LPT;
    const int TPB = T / LPT;
    int ST = 26 * I + 5; ST += (4 - (ST & 15) + 16) & 15;
    double* s_tab = smem;                                   // [W][E]: entry offsets become immediates of the loads
    double* s_tree = s_tab + (size_t)p.E * p.W;             // [TPB][ST]
    int* s_prog = (int*)(s_tree + (size_t)TPB * ST);
    for (int i = tid; i < p.E * p.W; i += T) s_tab[(i % p.W) * p.E + i / p.W] = p.tab[i];
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    for (int i = tid; i < TPB; i += T) s_tree[(size_t)i * ST + 26 * I] = 0.0;     // the message of an absent child
    __syncthreads();
    const int* par_internal = s_prog;
    const int* node_of = s_prog + I;
    const int* leaf_begin = s_prog + 2 * I;
    const int* leaf_node = s_prog + 3 * I + 1;
    const int* leaf_species = leaf_node + p.S;
    const int* ichild_begin = leaf_species + p.S;
    const int* ichild = ichild_begin + I + 1;
    const int lane = tid & 31, sub = lane % LPT, slot = sub >> 2, a = sub & 3;
    const int qb = lane & ~3;
    const unsigned mask = 0xFu << qb;
    double* tr = s_tree + (size_t)(tid / LPT) * ST;
    // per-tree block: q at 0, base at 4 I, msg at 8 I, the q of the sweep before at 12 I (reached through nd_own below)
    double* Q = tr; double* MSG = tr + 8 * I; double* RING = tr + 16 * I;
    int* FLAG = (int*)(tr + 24 * I);                        // [4 sweeps][I]
    const double* ZERO = tr + 26 * I;
    double* FLAST = tr + 26 * I + 1;                        // [0] F of the sweep before the last completed one, [1] of the last,
                                                            // [2] (as int) the last sweep's periodicity flags
    const bool early = p.k0 == 0 && p.k1 == p.K1;
    constexpr unsigned LANE0 = LPT == 8 ? 0x01010101u : 0x00010001u;
    const int D = *p.count;
    const long long n = (long long)D * p.W;
    const long long first = (long long)blockIdx.x * TPB + (long long)(tid >> 5) * TPW;
    for (long long it0 = first; it0 < n; it0 += (long long)gridDim.x * TPB) {
        const long long it_raw = it0 + lane / LPT;
        const bool live = it_raw < n;
        const long long it = live ? it_raw : n - 1;
        const int d = (int)(it / p.W), t = (int)(it - (long long)d * p.W);
        const ColWords cw = memo_words((uint32_t)p.code_of[d], p.S);
        const Tab tb{s_tab + t * p.E, 1};
        double* out = p.traj + ((size_t)t * p.cap + d) * p.RS;
        unsigned gm = 0;                                    // internal nodes with an unobserved leaf child
        for (int i = 0; i < I; ++i)
            for (int k = leaf_begin[i]; k < leaf_begin[i + 1]; ++k)
                if (cw.obs(leaf_species[k]) == 4) gm |= 1u << i;
        __syncwarp();                                       // the previous tree of this lane group is completely done
        if (p.k0 > 0) {
            for (int e = sub; e < 12 * I; e += LPT) tr[e] = p.state[(size_t)it * (12 * I) + e];
            __syncwarp();
        }
        const double lpia = tb(a);
        // this lane's node at even and at odd steps, with everything that does not change from sweep to sweep: table
        // offset, parent, children, and c[a] = sum over the observed leaf children of log T(a, x) (the sequential pass
        // re-reads it every sweep; the value is the same)
        int nd_i[2];                                               // node (or -1) | depth << 8
        int nd_to[2], nd_tm[2], nd_qp[2], nd_m1[2], nd_m2[2];      // offsets into smem (32-bit, shared address space)
        int nd_own[2];                                             // this lane's q entry; base / msg / q2 follow at 4 I steps
        const int I4 = 4 * I, tro = (int)(tr - smem);
        double nd_c[2];
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const int i_s = p.sched[par][slot];
            const int i = i_s < 0 ? 0 : i_s;
            nd_i[par] = (i_s & 0xff) | ((int)p.depth[i] << 8);
            const int e0 = i ? tab_of(node_of[i]) : 0;
            nd_own[par] = tro + i * 4 + a;
            nd_to[par] = t * p.E + e0 + a;                  // own rows: + 4 l
            nd_tm[par] = t * p.E + e0 + a * 4;              // message row: + h
            nd_qp[par] = (int)(Q - smem) + (i ? par_internal[i] * 4 : 0);
            const int cb = ichild_begin[i], ce = ichild_begin[i + 1];
            nd_m1[par] = cb < ce ? (int)(MSG - smem) + ichild[cb] * 4 + a : (int)(ZERO - smem);
            nd_m2[par] = cb + 1 < ce ? (int)(MSG - smem) + ichild[cb + 1] * 4 + a : (int)(ZERO - smem);

            double c = 0.0;
            for (int k = leaf_begin[i]; k < leaf_begin[i + 1]; ++k) {
                const int x = cw.obs(leaf_species[k]);
                if (x < 4) c += tb(tab_of(leaf_node[k]) + a * 4 + x);
            }
            nd_c[par] = c;
        }
        // (the fields of p sit behind a run-time index into the kernel parameters: copies for the loop)
        const int k0 = p.k0, k1 = p.k1, dmax = p.dmax;
        const bool many_children = p.MCH > 2;
        const int tau_end = 2 * (k1 - 1) + dmax;
        bool tree_done = false;
        int kc_stop = -1;                                   // >= 0: the warp stopped after completing this sweep
        for (int tau0 = 2 * k0; tau0 <= tau_end && kc_stop < 0; tau0 += 2) {
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const int tau = tau0 + par;
                if (tau > tau_end || kc_stop >= 0) break;
                const int i_s = (int)(signed char)(nd_i[par] & 0xff);
                const int i = i_s < 0 ? 0 : i_s;
                const int k = (tau - (nd_i[par] >> 8)) >> 1;
                const bool act = i_s >= 0 && k >= k0 && k < k1;
                const bool update = act && k > 0;
                const double* to = smem + nd_to[par];
                const double* tm = smem + nd_tm[par];
                double own;
                {
                    // (the root runs the chain on its own q and drops the result; so does its message below)
                    const double2 p01 = *(const double2*)(smem + nd_qp[par]), p23 = *(const double2*)(smem + nd_qp[par] + 2);
                    double acc = fma(p01.x, to[0], 0.0);
                    acc = fma(p01.y, to[4], acc);
                    acc = fma(p23.x, to[8], acc);
                    acc = fma(p23.y, to[12], acc);
                    own = i == 0 ? lpia : acc;
                }
                // base + the children's messages in child order; an absent child adds +0.0, which leaves every value but
                // -0.0 unchanged, and the base (a sum starting from +0.0) is never -0.0
                double* const mine = smem + nd_own[par];               // q; + I4 base, + 2 I4 msg, + 3 I4 q of the sweep before
                double cs = (mine[I4] + smem[nd_m1[par]]) + smem[nd_m2[par]];
                if (many_children)                                     // further internal children (multifurcating trees)
                    for (int cc = ichild_begin[i] + 2; cc < ichild_begin[i + 1]; ++cc) cs += MSG[ichild[cc] * 4 + a];
                const long long q_old1 = __double_as_longlong(mine[0]), q_old2 = __double_as_longlong(mine[3 * I4]);
                const double s = own + cs;
                const double ex = update ? s : 0.0;
                const double e = phy_exp(ex);
                const double ev0 = __shfl_sync(0xffffffffu, e, qb), ev1 = __shfl_sync(0xffffffffu, e, qb + 1);
                const double ev2 = __shfl_sync(0xffffffffu, e, qb + 2), ev3 = __shfl_sync(0xffffffffu, e, qb + 3);
                const double z = ((ev0 + ev1) + ev2) + ev3;
                const double rz = phy_rcp(z);
                const double lz = phy_log(z);
                const double q0 = phy_mul(ev0, rz), q1 = phy_mul(ev1, rz), q2 = phy_mul(ev2, rz), q3 = phy_mul(ev3, rz);
                const double qa = phy_mul(e, rz);
                const double lq = ex - lz;
                double m = fma(q0, tm[0], 0.0);
                m = fma(q1, tm[1], m);
                m = fma(q2, tm[2], m);
                m = fma(q3, tm[3], m);
                const double c = nd_c[par];
                const double ft = quad_sum4(0xffffffffu, qb, phy_mul(qa, (lq - own) - c));
                // bit 0: this node's q equals that of the previous sweep, bit 1: that of the sweep before (all four symbols)
                int fl = (__double_as_longlong(qa) == q_old1 ? 1 : 0) | (__double_as_longlong(qa) == q_old2 ? 2 : 0);
                fl &= __shfl_xor_sync(0xffffffffu, fl, 1);
                fl &= __shfl_xor_sync(0xffffffffu, fl, 2);
                double tsum = 0.0, Fgap = 0.0;
                if ((gm >> i) & 1u) {
                    const double qs = ((q0 + q1) + q2) + q3;
                    const double lp0 = tb(0), lp1 = tb(1), lp2 = tb(2), lp3 = tb(3);
                    for (int k2 = leaf_begin[i]; k2 < leaf_begin[i + 1]; ++k2) {
                        if (cw.obs(leaf_species[k2]) < 4) continue;
                        const double xl = phy_mul(qs, lpia);
                        const double exl = update ? xl : 0.0;
                        const double el = phy_exp(exl);
                        const double l0 = __shfl_sync(mask, el, qb), l1 = __shfl_sync(mask, el, qb + 1);
                        const double l2 = __shfl_sync(mask, el, qb + 2), l3 = __shfl_sync(mask, el, qb + 3);
                        const double zl = ((l0 + l1) + l2) + l3;
                        const double rzl = phy_rcp(zl);
                        const double lzl = phy_log(zl);
                        double tt = fma(phy_mul(l0, rzl), lp0, 0.0);
                        tt = fma(phy_mul(l1, rzl), lp1, tt);
                        tt = fma(phy_mul(l2, rzl), lp2, tt);
                        tt = fma(phy_mul(l3, rzl), lp3, tt);
                        const double ql = phy_mul(el, rzl);
                        Fgap += quad_sum4(mask, qb, phy_mul(ql, (exl - lzl) - xl));
                        tsum += tt;
                    }
                }
                if (act) {
                    mine[3 * I4] = __longlong_as_double(q_old1);
                    mine[0] = qa;
                    mine[2 * I4] = m;
                    mine[I4] = c + tsum;
                    if (a == 0) { RING[((k & 3) * I + i) * 2] = ft; RING[((k & 3) * I + i) * 2 + 1] = Fgap; FLAG[(k & 3) * I + i] = fl; }
                }
                __syncwarp();
                // the sweep whose deepest node ran in this step is complete: its free energy, terms in node order
                const int kc2 = tau - dmax;
                if (kc2 >= 2 * k0 && !(kc2 & 1)) {
                    const int kc = kc2 >> 1;
                    if (sub == 0) {
                        double F = 0.0;
                        int fa = kc >= 2 ? 3 : kc;                             // sweep 1 has one predecessor, sweep 0 none
                        for (int j = 0; j < I; ++j) {
                            F += RING[((kc & 3) * I + j) * 2];
                            if ((gm >> j) & 1u) F += RING[((kc & 3) * I + j) * 2 + 1];
                            fa &= FLAG[(kc & 3) * I + j];
                        }
                        if (live) out[kc] = F;
                        FLAST[0] = FLAST[1]; FLAST[1] = F; *(int*)(FLAST + 2) = fa;
                        tree_done = tree_done || fa != 0;
                    }
                    if (early && __ballot_sync(0xffffffffu, sub == 0 && tree_done) == LANE0) kc_stop = kc;
                }
            }
        }
        if (kc_stop >= 0) {
            // every tree of the warp is periodic from here on.  F of a sweep depends on the q of that sweep and of the one
            // before: at a fixed point (bit 0 at the last sweep) every later value equals the last one, in a two-cycle the
            // last two values alternate
            __syncwarp();
            const double Fc = FLAST[1];
            const double Fp = (*(const int*)(FLAST + 2) & 1) ? Fc : FLAST[0];
            if (live)
                for (int j = kc_stop + 1 + sub; j < p.K1; j += LPT) out[j] = ((j - kc_stop) & 1) ? Fp : Fc;
        }
        if (p.k1 < p.K1 && live)       // (the step loop ended with a __syncwarp: every lane's stores are visible)
            for (int e = sub; e < 12 * I; e += LPT) p.state[(size_t)it * (12 * I) + e] = tr[e];
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
    // windows whose stop rule has not fired within the K1 stored passes are listed: for the second gather pass (when the
    // trajectories are computed in two phases) or for the direct kernel
    int* pend_count; int64_t* pend_col0; uint8_t* pend_strand; int64_t* pend_dst;
    // LIST variant: the windows to score (the list a previous pass left)
    const int* in_count; const int64_t* in_col0; const uint8_t* in_strand; const int64_t* in_dst;
};

// one thread per window: walk k in steps of four (one 32-byte sector per tree and step), summing the W trajectories in
// tree order, until the reference's stop rule fires
template <bool LIST>
__global__ void __launch_bounds__(256, 4) memo_window_kernel(MemoWindowParams p) {
    const long long n_items = LIST ? (long long)*p.in_count : (long long)p.n_windows * p.n_strands;
    unsigned long long ksum = 0;
    for (long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx0 < n_items; idx0 += (long long)gridDim.x * blockDim.x) {
        long long item; int strand; int64_t c0;
        if (LIST) {
            item = p.in_dst[idx0]; strand = p.in_strand[idx0]; c0 = p.in_col0[idx0];
        } else {
            item = idx0;
            const int si = (int)(item / p.n_windows);
            const int64_t w = item - (long long)si * p.n_windows;
            strand = p.n_strands == 2 ? si : p.strand0;
            const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
            c0 = p.col_off[a] + (w - p.win_off[a]);
        }
        uint32_t row[16];              // in 32-byte units: a 32-bit offset per tree keeps the kernel at 64 registers
        const MemoRow4* const base = (const MemoRow4*)p.traj;
        const uint32_t RS4 = (uint32_t)p.RS >> 2;
        // W <= 16 here (the launcher falls back to the direct kernel otherwise): trajectory rows of the W trees
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            if (t < p.W) {
                const int64_t c = strand ? c0 + (p.W - 1 - t) : c0 + t;       // jstacs StrandModel.java:636-639
                ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                if (strand) cw = complement(cw);
                const int d = p.dense_of[memo_code(cw, p.S)];
                row[t] = ((uint32_t)t * (uint32_t)p.cap + (uint32_t)d) * RS4;
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
                    const MemoRow4 r = memo_load4(base + (row[t] + (uint32_t)(kb >> 2)));
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
