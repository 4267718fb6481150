// tree_pass_f81.cuh -- order-0 mean field, the pass of tree_pass.cuh restructured for substitution tables with F81
// structure (evolution/FS81alpha.java:76-92, FS81beta.java:81-103: one off-diagonal value per target symbol) and the kernel
// around it (score_o0f_kernel).  Same fixed-point iteration as MeanFieldForBayesNet.java:108-202 / :232-365, sums
// re-associated; what changes against tree_pass.cuh::meanfield_pass:
//  * a table row is (Lo[b], D[b] = Ld[b] - Lo[b]) -- log off-diagonal and log diagonal minus it -- so the 16-term contractions
//    collapse:   own(a) = sum_l q_par(l) logT[l][a] = q_par(a) D(a) + tot_par Lo(a)
//                msg(l) = sum_h q(h) logT[l][h]     = q(l) D(l) + sum_h q(h) Lo(h)           (message to the tree parent)
//                observed leaf x: c(a) += Lo(x) + (a == x ? D(x) : 0)
//  * the free energy needs no logarithm per node: the node's terms sum_a q(a) (log q(a) - own(a) - c(a)) equal
//    sum_a q(a) m(a) - log z with m = the children's messages and gap-leaf sums (pass at the uniform start: -(own + c) and
//    z = 4), and the log z of a tree's nodes are accumulated as ONE logarithm of the product of their mantissas plus ln 2
//    times the sum of their exponents (net1n.cuh has the same scheme);
//  * the four exponentials of a node are straight-line code (n1n_exp4), the reciprocal skips the library's special cases.
// Kernel: one thread per tree as before, but the trees of a window exchange their free energies inside a GROUP of a few
// warps (named barrier, lcm(W, 32) lanes where that is small) instead of the whole block, once per pass instead of twice
// (the next window of a slot is fetched ahead of time), with reduction of the "still active" vote in the same barrier.
//
// The N1N_HD / PHY_HD parts compile for the host too (tests/emul).
#pragma once
#include "net1n.cuh"
#include "tree_pass.cuh"

namespace phyfoo {

// compact table accessor: entries [0..3] log pi, then per tree node n >= 1 eight entries at 4 + 8 (n - 1): Lo[4] | D[4]
PHY_HD int tabf_of(int node) { return 4 + 8 * (node - 1); }

// Per-thread state of the pass, strided so that consecutive threads touch consecutive doubles.  Only what a later node
// update reads is kept: q(s, .) and the message sums cs(s, .) of the internal nodes that HAVE internal children (slot s; the
// q of the others is read by nobody during scoring), and one scalar per node for the sums of its gap leaves (touched only
// when the column has an unobserved symbol).  A 12-species binary tree: 6 slots + 11 scalars = 472 bytes instead of 704.
struct StateF {
    double* base;
    int stride;
    int n_slots;
    PHY_HD double& q(int s, int a) const { return base[(s * 8 + a) * stride]; }
    PHY_HD double& cs(int s, int a) const { return base[(s * 8 + 4 + a) * stride]; }
    PHY_HD double& ts(int i) const { return base[(n_slots * 8 + i) * stride]; }
};
PHY_HD int statef_doubles(int n_slots, int n_internal) { return n_slots * 8 + n_internal; }

// One pass over the tree; acc collects the tree's free energy (terms and the log z accumulator).
// slot[i]: state slot of internal node i, -1 when it has no internal child.
PHY_HD void meanfield_pass_f81(const TreeProg& tp, const int* slot, const Tab& tb, const StateF& st, const ColWords& cw, bool update,
                               const double* exptab, N1nAcc& acc) {
    double lpi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) lpi[a] = tb(a);
    const bool col_gap = cw.g != 0;
    for (int i = 0; i < tp.n_internal; ++i) {
        const int p = tp.par_internal[i];
        const int sl = slot[i], sp = i ? slot[p] : -1;
        double Lo[4], D[4], oc[4];
        if (i == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) oc[a] = lpi[a];
        } else {
            const int e0 = tabf_of(tp.node_of[i]);
            double qp[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { Lo[a] = tb(e0 + a); D[a] = tb(e0 + 4 + a); qp[a] = st.q(sp, a); }
            const double tot = ((qp[0] + qp[1]) + qp[2]) + qp[3];
#pragma unroll
            for (int a = 0; a < 4; ++a) oc[a] = fma(qp[a], D[a], tot * Lo[a]);      // own rows, MeanField :117-135
        }
        // observed leaf children (:147-148,181-182); gap leaves are handled below
        bool any_gap = false;
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = cw.obs(tp.leaf_species[k]);
            if (x < 4) {
                const int e0 = tabf_of(tp.leaf_node[k]);
                const double lo = tb(e0 + x), d = tb(e0 + 4 + x);
#pragma unroll
                for (int a = 0; a < 4; ++a) oc[a] += a == x ? lo + d : lo;
            } else any_gap = true;
        }
        // children's part from the previous pass (:139-191): messages of the internal children, sums of the gap leaves
        double m[4] = {0.0, 0.0, 0.0, 0.0}, ex[4], g[4], e[4], qi[4];
        if (sl >= 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) m[a] = st.cs(sl, a);
        }
        if (col_gap) {
            const double tsv = st.ts(i);
#pragma unroll
            for (int a = 0; a < 4; ++a) m[a] += tsv;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            ex[a] = update ? oc[a] + m[a] : 0.0;
            g[a] = update ? m[a] : -oc[a];
        }
        n1n_exp4(ex, exptab, e);                                  // :192, no max shift
        const double z = ((e[0] + e[1]) + e[2]) + e[3];           // :195-198
        double fu = e[0] * g[0];
        fu = fma(e[1], g[1], fu); fu = fma(e[2], g[2], fu); fu = fma(e[3], g[3], fu);
        const int zhi = n1n_hi(z);
        const unsigned zef = (unsigned)zhi >> 20;
        const bool znormal = zef - 24u < 2000u;
        double rz = n1n_rcp_normal(z);
        if (znormal) { acc.E += (int)zef - 1023; acc.M *= n1n_with_hi(z, (zhi & 0x000fffff) | 0x3ff00000); }
        else {
            const N1nZR zr = n1n_z_rare(z);
            rz = zr.rz;
            acc.Fslow += zr.lz;
        }
        acc.Fq = fma(fu, rz, acc.Fq);
#pragma unroll
        for (int a = 0; a < 4; ++a) qi[a] = e[a] * rz;                              // :199-201
        if (sl >= 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) { st.q(sl, a) = qi[a]; st.cs(sl, a) = 0.0; }   // the children add their messages below
        }
        if (i != 0) {
            // message to the parent: expectation of the log table over this node's new q
            double sw = qi[0] * Lo[0];
            sw = fma(qi[1], Lo[1], sw); sw = fma(qi[2], Lo[2], sw); sw = fma(qi[3], Lo[3], sw);
#pragma unroll
            for (int l = 0; l < 4; ++l) st.cs(sp, l) += fma(qi[l], D[l], sw);
        }
        if (col_gap) {
            double tsum = 0.0;
            if (any_gap) {
                // unobserved leaves: every table row is log pi ("independent" table, bayesNet/CPF.java:146-152); their own
                // update only needs q of the parent, which is final for this sweep.  Free-energy terms: entropy (:250-267) and
                // the hidden non-root term (:327-358) sum to -log z_l (pass at the uniform start: -sum_a q_l(a) x_l(a) - log 4)
                const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
                for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
                    if (cw.obs(tp.leaf_species[k]) < 4) continue;
                    double xl[4], xe[4], el[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { xl[a] = qs * lpi[a]; xe[a] = update ? xl[a] : 0.0; }
                    n1n_exp4(xe, exptab, el);
                    const double zl = ((el[0] + el[1]) + el[2]) + el[3];
                    const N1nZR zr = n1n_z_rare(zl);                                // (rare path: gaps) library reciprocal, log z directly
                    double t = 0.0, fl = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const double ql = el[a] * zr.rz;
                        t = fma(ql, lpi[a], t);
                        fl = fma(ql, update ? 0.0 : -xl[a], fl);
                    }
                    acc.Fq += fl;
                    acc.Fslow += zr.lz;
                    tsum += t;
                }
            }
            st.ts(i) = tsum;
        }
    }
}

// slot[i] for every internal node from the CSR of internal children; returns the number of slots
PHY_HD int f81_slots(int n_internal, const int* ichild_begin, int* slot) {
    int n = 0;
    for (int i = 0; i < n_internal; ++i) slot[i] = ichild_begin[i + 1] > ichild_begin[i] ? n++ : -1;
    return n;
}

}  // namespace phyfoo

#if defined(__CUDACC__)
__global__ void __launch_bounds__(512, 1) score_o0f_kernel(ScoreFParams p) {
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int W = p.W, GL = p.group_lanes, wpg = GL / W;           // windows per group
    double* s_tab = smem;                                            // [Ef][W]
    double* s_exp = s_tab + (size_t)p.Ef * W;                        // [64]
    double* s_F = s_exp + 64;                                        // [2][T]
    long long* s_next = (long long*)(s_F + 2 * T);                   // [n_groups][wpg] item fetched ahead for the window slot
    double* s_state = (double*)(s_next + (size_t)p.n_groups * wpg);
    const int state_doubles = phyfoo::statef_doubles(p.n_slots, p.I);
    int* s_prog = (int*)(s_state + (p.gstate ? 0 : (size_t)state_doubles * T));
    int* s_slot = s_prog + p.prog_len;                               // [I]
    for (int i = tid; i < p.Ef * W; i += T) s_tab[i] = p.tab[i];
    for (int i = tid; i < 64; i += T) s_exp[i] = phyfoo::d_phy_exp_tab[i];
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];

    phyfoo::TreeProg tp;
    tp.n_internal = p.I; tp.n_leaves = p.S; tp.n_nodes = 0;
    tp.par_internal = s_prog;
    tp.node_of = s_prog + p.I;
    tp.leaf_begin = s_prog + 2 * p.I;
    tp.leaf_node = s_prog + 3 * p.I + 1;
    tp.leaf_species = s_prog + 3 * p.I + 1 + p.S;
    __syncthreads();
    if (tid == 0) phyfoo::f81_slots(p.I, s_prog + 3 * p.I + 1 + 2 * p.S, s_slot);

    const int grp = tid / GL, gl = tid - grp * GL;                   // group and lane within it
    const int ws = gl / W, t = gl - ws * W;                          // window slot within the group, tree (motif position)
    const bool lane_ok = ws < wpg;
    const int bar_id = 1 + grp;
    long long* my_next = s_next + (size_t)grp * wpg + ws;
    const phyfoo::Tab tb{s_tab + t, W};
    phyfoo::StateF st;
    st.n_slots = p.n_slots;
    if (p.gstate) { st.base = p.gstate + ((size_t)blockIdx.x * T + tid); st.stride = gridDim.x * T; }
    else { st.base = s_state + tid; st.stride = T; }
    const long long n_items = (long long)p.n_windows * p.n_strands;

    // Every window slot holds one item fetched ahead.  Protocol: when a window takes its item (a register copy, `pre`, that
    // all its lanes loaded after an earlier barrier), its first lane fetches the next one into shared memory; the lanes load
    // it after the barrier of the same pass.  The slot is not written again before the window's next take, at least three
    // passes later (the pass at the uniform start and min_sweeps >= 1 sweeps ... the stop rule needs at least one sweep).
    if (lane_ok && t == 0) *my_next = (long long)atomicAdd(p.counter_next, 1ULL);
    __syncthreads();

    bool need = lane_ok, active = lane_ok, upd = false, conv = false, pre_valid = lane_ok;
    int k = 0, buf = 0;
    long long item = -1, pre = lane_ok ? *my_next : 0;
    double old = 0.0;
    __syncthreads();                                                 // everybody has its copy before a first lane fetches again
    phyfoo::ColWords cw; cw.b = 0; cw.g = 0;
    double* const fwin = s_F + grp * GL + ws * W;                    // this window's slots of buffer 0

    for (;;) {
        if (need && active) {
            item = pre;
            pre_valid = false;
            if (t == 0) *my_next = (long long)atomicAdd(p.counter_next, 1ULL);
            if (item >= n_items) active = false;
            else {
                const int si = (int)(item / p.n_windows);
                const int64_t w = item - (long long)si * p.n_windows;
                const int strand = p.n_strands == 2 ? si : p.strand0;
                const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
                const int64_t c0 = p.col_off[a] + (w - p.win_off[a]);
                // jstacs StrandModel.java:636-639 / Sequence.java:1321-1333: reverse the columns, complement the symbols
                const int64_t c = strand ? c0 + (W - 1 - t) : c0 + t;
                cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                if (strand) cw = complement(cw);
                k = 0; conv = false; upd = false; need = false;
            }
        }
        double F = 0.0;
        if (active) {
            phyfoo::N1nAcc acc; acc.Fq = 0.0; acc.M = 1.0; acc.E = 0; acc.Fslow = 0.0;
            phyfoo::meanfield_pass_f81(tp, s_slot, tb, st, cw, upd, s_exp, acc);
            F = phyfoo::n1n_free_energy(acc);
        }
        s_F[buf * T + tid] = F;
        // one barrier per pass, among the warps of this group only: publishes the trees' free energies, the items fetched
        // ahead during the previous pass, and votes on the exit
        if (!named_bar_or(bar_id, GL, active)) break;
        if (!pre_valid && lane_ok) { pre = *my_next; pre_valid = true; }
        if (active) {
            double Fsum = 0.0;
            const double* fp = fwin + buf * T;
            for (int m = 0; m < W; ++m) Fsum += fp[m];               // tree-major, like calcFreeEnergy(0, W-1)
            bool done;
            if (!upd) { old = Fsum; upd = true; done = false; }                   // F_0, MeanFieldForBayesNet.java:105
            else {
                if (fabs(old - Fsum) < p.tol) conv = true;                        // :204-207 (sticky)
                old = Fsum; ++k;
                done = !((k < p.max_sweeps && !conv) || k < p.min_sweeps);        // :107
            }
            if (done) {
                if (t == 0) {
                    p.out[item] = -Fsum;                                          // PhyloBayesModel.java:259
                    if (p.sweeps) p.sweeps[item] = k;
                    atomicAdd(p.counter_sweeps, (unsigned long long)k);
                }
                need = true;
            }
        }
        buf ^= 1;
    }
}
#endif  // __CUDACC__
