// netg.cuh -- the mean field of the reference on an ARBITRARY network of trees, one thread per item.
//
// Used for background models of order >= 2 (models/PhyloBackground.java:381-398: tree t has the trees t-1 .. t-order as
// horizontal parents, every node the same node of those trees), whose conditional probability functions have up to 4^(order+1)
// rows per node.  The specialised kernels (tree_pass*.cuh: order 0; net1*.cuh: one horizontal parent) rest on contractions
// written out for their table shapes; this one walks parent and child lists like the reference does and is the statement of
// algorithm/MeanFieldForBayesNet.java on the device:
//   init              :52-69     every hidden node uniform
//   update            :108-202   Gauss-Seidel over the nodes in net order; own rows with the OBSERVED parents not filtered
//                                (:117-135), children with the observed co-parents filtered (:139-191), exp without a max shift
//   free energy       :232-365   entropy part and expectation part over a range of trees
//   stop rule         :102-107,204-209 on the free energy of the first `len` trees (len = columns of the scored sub-sequence),
//                                sticky convergence flag, 2 <= sweeps <= 100
// An item names the column each tree observes explicitly, so that the ranges shorter than order + 1 columns -- whose trailing
// trees keep the observation of an earlier call in the reference (MeanFieldForBayesNet.java:84) -- are
// items like any other.  Unobserved leaves take the independent table (bayesNet/CPF.java:146-152,218-224).
// The per-item function compiles for the host as well (tests/emul).
#pragma once
#include <cstdint>
#include "fp64_math.cuh"
#include "tree_pass.cuh"

#if defined(__CUDACC__)
#define NETG_HD __host__ __device__ __forceinline__
#else
#define NETG_HD inline
#endif

namespace phyfoo {

// node record: NETG_REC ints
//   0 n_par   1 par_off (into the int list)   2 n_chi   3 chi_off (pairs: child id, index of this node among the child's parents)
//   4 tab_off (doubles: log CPF [rows][4])    5 ind_off (log of the independent table; = tab_off for inner nodes)
//   6 rows = 4^n_par                           7 species of a leaf, -1 for an inner node
//   8 tree
constexpr int NETG_REC = 9;
constexpr int NETG_MAX_TREES = 8;

struct NetgView {
    int n_nodes, n_trees, nodes_per_tree;
    const int* rec;        // [n_nodes][NETG_REC]
    const int* lists;      // parent ids and (child, index) pairs
    const double* tab;     // log tables
};

// q of this item: q[(g * 4 + a) * stride]
struct NetgState {
    double* q; size_t stride;
    NETG_HD double& at(int g, int a) const { return q[(size_t)(g * 4 + a) * stride]; }
};

struct NetgObs {
    ColWords cw[NETG_MAX_TREES];
    // 0..3 observed symbol, 4 hidden
    NETG_HD int of(const int* r) const { return r[7] < 0 ? 4 : cw[r[8]].obs(r[7]); }
};

NETG_HD int netg_digit(int l, int p) { return (l >> (2 * p)) & 3; }

// prod times the factors of the parents of row l of node `r`, skipping parent `skip` (-1: none), in the order of the parent
// list: hidden parents contribute their q; agrees = every observed parent has the symbol the row names
NETG_HD double netg_parent_product(const NetgView& v, const NetgObs& ob, const NetgState& st, const int* r, int l, int skip, double prod,
                                   bool& agrees) {
    agrees = true;
    const int* par = v.lists + r[1];
    for (int p = 0; p < r[0]; ++p) {
        const int* pr = v.rec + (size_t)par[p] * NETG_REC;
        const int hp = netg_digit(l, p);
        const int po = ob.of(pr);
        if (po == 4) { if (p != skip) prod *= st.at(par[p], hp); }
        else if (hp != po) agrees = false;
    }
    return prod;
}

// free energy over the trees [t0, t1]  (MeanFieldForBayesNet.java:232-365)
NETG_HD double netg_free_energy(const NetgView& v, const NetgObs& ob, const NetgState& st, int t0, int t1) {
    double part1 = 0.0, part2 = 0.0;
    const int g0 = t0 * v.nodes_per_tree, g1 = (t1 + 1) * v.nodes_per_tree;
    for (int g = g0; g < g1; ++g) {
        const int* r = v.rec + (size_t)g * NETG_REC;
        double tmp = 0.0;
        if (ob.of(r) == 4)
            for (int h = 0; h < 4; ++h) { const double qh = st.at(g, h); if (qh > 0) tmp += qh * phy_log(qh); }
        part1 += tmp;
    }
    for (int g = g0; g < g1; ++g) {
        const int* r = v.rec + (size_t)g * NETG_REC;
        const int o = ob.of(r);
        const double* lt = v.tab + (o == 4 ? r[5] : r[4]);      // an unobserved leaf: the independent table
        double tmp = 0.0;
        if (r[0] == 0 && o == 4) {                                 // :282-291
            for (int h = 0; h < 4; ++h) tmp += st.at(g, h) * lt[h];
        } else if (o != 4) {                                       // :292-326
            for (int l = 0; l < r[6]; ++l) {
                bool agrees;
                double prod = netg_parent_product(v, ob, st, r, l, -1, 1.0, agrees);
                if (!agrees) prod = 0.0;
                prod *= lt[l * 4 + o];
                tmp += prod;
            }
        } else {                                                   // :327-358
            for (int h2 = 0; h2 < 4; ++h2) {
                const double qh = st.at(g, h2);
                for (int l = 0; l < r[6]; ++l) {
                    bool agrees;
                    double prod = netg_parent_product(v, ob, st, r, l, -1, 1.0, agrees);
                    if (!agrees) prod = 0.0;
                    prod *= qh * lt[l * 4 + h2];
                    tmp += prod;
                }
            }
        }
        part2 += tmp;
    }
    return part1 - part2;
}

// one Gauss-Seidel sweep over all nodes of the net  (MeanFieldForBayesNet.java:108-202)
NETG_HD void netg_sweep(const NetgView& v, const NetgObs& ob, const NetgState& st) {
    for (int g = 0; g < v.n_nodes; ++g) {
        const int* r = v.rec + (size_t)g * NETG_REC;
        if (ob.of(r) != 4) continue;
        const double* lt = v.tab + r[5];                           // g is hidden: a leaf reads its independent table
        const int* par = v.lists + r[1];
        const int* chi = v.lists + r[3];
        double nq[4];
        for (int a = 0; a < 4; ++a) {
            double acc = 0.0;
            for (int l = 0; l < r[6]; ++l) {                       // own rows, :117-135 (observed parents NOT filtered)
                double prod = 1.0;
                for (int p = 0; p < r[0]; ++p)
                    if (ob.of(v.rec + (size_t)par[p] * NETG_REC) == 4) prod *= st.at(par[p], netg_digit(l, p));
                acc += prod * lt[l * 4 + a];
            }
            for (int c = 0; c < r[2]; ++c) {                       // children, :139-191
                const int ci = chi[2 * c], idx = chi[2 * c + 1];
                const int* cr = v.rec + (size_t)ci * NETG_REC;
                const int co = ob.of(cr);
                const double* ct = v.tab + (co == 4 ? cr[5] : cr[4]);
                for (int ac = 0; ac < 4; ++ac) {
                    if (co != 4 && co != ac) continue;
                    const double qc = co != 4 ? 1.0 : st.at(ci, ac);
                    for (int l = 0; l < cr[6]; ++l) {
                        if (netg_digit(l, idx) != a) continue;
                        bool agrees;
                        const double prod = netg_parent_product(v, ob, st, cr, l, idx, qc, agrees);
                        if (agrees) acc += prod * ct[l * 4 + ac];
                    }
                }
            }
            nq[a] = phy_exp(acc);                                  // :192 (no max shift)
        }
        const double sum = ((nq[0] + nq[1]) + nq[2]) + nq[3];      // :195-198
        for (int a = 0; a < 4; ++a) st.at(g, a) = nq[a] / sum;     // :199-201
    }
}

// One item: mean field to the reference's stop rule on the first `len` trees; out_first = sum of the free energies of the
// single trees 0 .. min(len, n_trees - 1) - 1 (PhyloBackground.java:273-283), out_last = free energy of the last tree alone
// (:284-297, meaningful for len = n_trees).  Returns the number of sweeps.
NETG_HD int netg_item(const NetgView& v, const NetgObs& ob, const NetgState& st, int len, int max_sweeps, int min_sweeps, double tol,
                      double* out_first, double* out_last) {
    for (int g = 0; g < v.n_nodes; ++g)
        for (int a = 0; a < 4; ++a) st.at(g, a) = 0.25;
    double old = netg_free_energy(v, ob, st, 0, len - 1);
    bool conv = false;
    int k = 0;
    while ((k < max_sweeps && !conv) || k < min_sweeps) {
        netg_sweep(v, ob, st);
        const double cur = netg_free_energy(v, ob, st, 0, len - 1);
        if (fabs(old - cur) < tol) conv = true;
        old = cur;
        ++k;
    }
    double first = 0.0;
    const int nf = len < v.n_trees - 1 ? len : v.n_trees - 1;
    for (int t = 0; t < nf; ++t) first += netg_free_energy(v, ob, st, t, t);
    *out_first = first;
    *out_last = netg_free_energy(v, ob, st, v.n_trees - 1, v.n_trees - 1);
    return k;
}

// Expected counts of the table entries under the fixed q of one item: the coefficients netg_free_energy multiplies the log table
// entries with (over ALL trees), times w, handed to add(entry index into the table array, value); returns w * sum q log q (the
// entropy part).  With them the fixed-q objective of optimizing/AbstractFreeEnergyOptimizer.java:95-114 is
// f(theta) = sum_e C[e] log table_theta[e] - ENT for any parameter set.
template <class Add>
NETG_HD double netg_counts(const NetgView& v, const NetgObs& ob, const NetgState& st, double w, Add&& add) {
    double part1 = 0.0;
    for (int g = 0; g < v.n_nodes; ++g) {
        const int* r = v.rec + (size_t)g * NETG_REC;
        const int o = ob.of(r);
        const int base = o == 4 ? r[5] : r[4];
        if (o == 4)
            for (int h = 0; h < 4; ++h) { const double qh = st.at(g, h); if (qh > 0) part1 += qh * phy_log(qh); }
        if (r[0] == 0 && o == 4) {
            for (int h = 0; h < 4; ++h) add(base + h, w * st.at(g, h));
        } else if (o != 4) {
            for (int l = 0; l < r[6]; ++l) {
                bool agrees;
                const double prod = netg_parent_product(v, ob, st, r, l, -1, 1.0, agrees);
                if (agrees && prod != 0.0) add(base + l * 4 + o, w * prod);
            }
        } else {
            for (int l = 0; l < r[6]; ++l) {
                bool agrees;
                const double prod = netg_parent_product(v, ob, st, r, l, -1, 1.0, agrees);
                if (!agrees || prod == 0.0) continue;
                for (int h2 = 0; h2 < 4; ++h2) add(base + l * 4 + h2, w * (prod * st.at(g, h2)));
            }
        }
    }
    return w * part1;
}

}  // namespace phyfoo

#if defined(__CUDACC__)
struct NetgParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    phyfoo::NetgView view;
    const int64_t* item_cols;   // [n_trees][n_items]: the column tree t of item u observes
    const int32_t* item_len;    // NULL: every item spans all trees
    int64_t n_items;
    double* q; size_t q_stride; // [n_nodes * 4][q_stride], one slot per thread of the grid
    double* out_first; double* out_last; int32_t* sweeps;
    unsigned long long* sweep_sum;
    int max_sweeps, min_sweeps; double tol;
};

__global__ void __launch_bounds__(128) score_netg_kernel(NetgParams p) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    phyfoo::NetgState st{p.q + slot, p.q_stride};
    unsigned long long sw = 0;
    for (int64_t u = (int64_t)slot; u < p.n_items; u += (int64_t)gridDim.x * blockDim.x) {
        phyfoo::NetgObs ob;
        for (int t = 0; t < p.view.n_trees; ++t) ob.cw[t] = load_column(p.bases, p.gaps, p.n_cols, p.nb, p.item_cols[(size_t)t * p.n_items + u]);
        double f = 0.0, l = 0.0;
        const int k = phyfoo::netg_item(p.view, ob, st, p.item_len ? p.item_len[u] : p.view.n_trees, p.max_sweeps, p.min_sweeps, p.tol, &f, &l);
        p.out_first[u] = f;
        p.out_last[u] = l;
        if (p.sweeps) p.sweeps[u] = k;
        sw += (unsigned long long)k;
    }
    if (p.sweep_sum && sw) atomicAdd(p.sweep_sum, sw);
}
#endif

#if defined(__CUDACC__)
// M-step of a background of order >= 1: mean field of every window, then its expected counts into counts[] (one slot per table
// entry, global atomics) and its entropy part into *entropy.  item u = window u; its weight is weight_of_column[u] when given --
// the reference indexes the per-COLUMN weight vector with the WINDOW number (models/PhyloBackground.java:114,178-186 against
// optimizing/AbstractFreeEnergyOptimizer.java:104-108), kept.
struct NetgCountParams {
    NetgParams s;
    const double* aln_w; const int64_t* col_off; int n_aln;   // aln_w NULL: every weight 1
    double* counts; double* entropy;
};
__global__ void __launch_bounds__(128) netg_count_kernel(NetgCountParams p) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    phyfoo::NetgState st{p.s.q + slot, p.s.q_stride};
    double ent = 0.0;
    for (int64_t u = (int64_t)slot; u < p.s.n_items; u += (int64_t)gridDim.x * blockDim.x) {
        phyfoo::NetgObs ob;
        for (int t = 0; t < p.s.view.n_trees; ++t) ob.cw[t] = load_column(p.s.bases, p.s.gaps, p.s.n_cols, p.s.nb, p.s.item_cols[(size_t)t * p.s.n_items + u]);
        double f = 0.0, l = 0.0;
        phyfoo::netg_item(p.s.view, ob, st, p.s.view.n_trees, p.s.max_sweeps, p.s.min_sweeps, p.s.tol, &f, &l);
        const double w = p.aln_w ? p.aln_w[find_alignment(p.col_off, p.n_aln, u)] : 1.0;
        ent += phyfoo::netg_counts(p.s.view, ob, st, w, [&](int e, double val) { atomicAdd(p.counts + e, val); });
    }
    if (ent != 0.0) atomicAdd(p.entropy, ent);
}
#endif
