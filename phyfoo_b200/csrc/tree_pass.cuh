// tree_pass.cuh -- per-thread work of the order-0 kernels: one thread owns one phylogenetic tree
// (= one aligned column under the parameters of one motif position) and runs the mean-field
// coordinate ascent / free energy (reference: algorithm/MeanFieldForBayesNet.java:92-211,232-365)
// or Felsenstein pruning (what algorithm/SimpleNodeElimination.java:63-69 equals for order 0).
//
// Everything here is __host__ __device__ so tests/emul can run the identical arithmetic on the CPU
// when no GPU is around; the shipped library only ever calls it from kernels.
//
// Restructuring relative to the Java loops (same fixed point iteration, sums re-associated):
//  * logs of the CPF tables are hoisted into per-parameter-set tables (the Java takes Math.log
//    inside the innermost loops);
//  * a sweep and the free energy after it are fused: in DFS pre-order the parent of a node is
//    already final for this sweep when the node is updated, so every free-energy term is final
//    at the moment its last participant has been updated;
//  * the child->parent message  msg_c[a] = sum_h q_c(h) log T_c(a,h)  is computed once, used for
//    the free energy of this sweep and accumulated into the parent's "child sum" for the next one;
//  * observed leaves never get state: their contribution log T_c(a, x_c) is re-read per sweep;
//  * an unobserved (gap) leaf is handled inline at its parent (its table is the "independent"
//    pi row, bayesNet/CPF.java:146-152, so it only needs the parent's q of the same sweep).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PHY_HD __host__ __device__ __forceinline__
#else
#define PHY_HD inline
#endif

namespace phyfoo {

// Flattened tree "program" shared by all trees of a model (host builds it, kernels read it from
// shared memory).  Internal nodes are numbered 0..I-1 in DFS pre-order (root = 0).
struct TreeProg {
    int n_internal;          // I
    int n_leaves;            // S
    int n_nodes;             // N
    const int* par_internal; // [I]   internal index of the parent (-1 for the root)
    const int* node_of;      // [I]   tree node id (pre-order) of internal node i
    const int* leaf_begin;   // [I+1] CSR into leaf_node / leaf_species: leaf children of internal i
    const int* leaf_node;    // [S]   tree node id of the leaf
    const int* leaf_species; // [S]   species (row of the alignment) observed at that leaf
};

// Per-position log tables: [0..3] = log pi, then for tree node id n >= 1: 16 doubles at
// 4 + 16*(n-1), row-major [parent value a][own value b] = log P(b | a).
PHY_HD const double* lt_of(const double* tb, int node) { return tb + 4 + 16 * (node - 1); }

// Observation of species s in a packed column (include/phyfoo.h "packed columns"): 0..3, or 4 = gap.
PHY_HD int column_obs(const uint32_t* bases, const uint32_t* gaps, int64_t plane_stride, int s, bool revcomp) {
    uint32_t g = (gaps[(int64_t)(s >> 5) * plane_stride] >> (s & 31)) & 1u;
    uint32_t b = (bases[(int64_t)(s >> 4) * plane_stride] >> (2 * (s & 15))) & 3u;
    if (revcomp) b = 3u - b;  // UnobservableDNAAlphabet.java:88-95: complement, gap stays gap
    return g ? 4 : (int)b;
}

// State accessors: q[i][a] and child-sum cs[i][a], strided so that consecutive threads touch
// consecutive doubles (conflict-free shared memory / coalesced global memory).
struct State {
    double* base;
    int stride;  // number of threads sharing the buffer
    PHY_HD double& q(int i, int a) const { return base[(size_t)(i * 8 + a) * stride]; }
    PHY_HD double& cs(int i, int a) const { return base[(size_t)(i * 8 + 4 + a) * stride]; }
};

// Optional sink for the converged q (M-step "prepare"): internal nodes then leaves.
struct QSink {
    double* q_internal;  // [I][4] or nullptr
    double* q_leaf;      // [S][4] or nullptr (rows of observed leaves are left untouched)
};

// One pass over the tree.  update == false: q stays as it is (pass 0 evaluates the free energy of the
// uniform start and primes the child sums).  Returns this tree's free energy F = part1 - part2.
PHY_HD double meanfield_pass(const TreeProg& tp, const double* tb, const State& st, const uint32_t* bases,
                             const uint32_t* gaps, int64_t plane_stride, bool revcomp, bool update,
                             const QSink* sink) {
    double F = 0.0;
    const double* lpi = tb;
    for (int i = 0; i < tp.n_internal; ++i) {
        const int p = tp.par_internal[i];
        const double* lt = (i == 0) ? nullptr : lt_of(tb, tp.node_of[i]);
        double qi[4], lq[4];
        if (update) {
            double s[4];
            if (i == 0) {
#pragma unroll
                for (int a = 0; a < 4; ++a) s[a] = lpi[a] + st.cs(0, a);
            } else {
                double qp[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) qp[l] = st.q(p, l);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < 4; ++l) acc += qp[l] * lt[l * 4 + a];  // own rows, MeanField :117-135
                    s[a] = acc + st.cs(i, a);                                 // + children, :139-191
                }
            }
            double e[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) e[a] = exp(s[a]);                     // :192, no max shift
            const double z = (e[0] + e[1]) + (e[2] + e[3]);
            const double lz = log(z);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                qi[a] = e[a] / z;                                             // :199-201
                lq[a] = s[a] - lz;
                st.q(i, a) = qi[a];
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                qi[a] = st.q(i, a);
                lq[a] = (qi[a] > 0.0) ? log(qi[a]) : 0.0;
            }
        }
        // entropy term, :250-267 (q == 0 rows are skipped there; here 0 * finite = 0)
#pragma unroll
        for (int a = 0; a < 4; ++a) F += (qi[a] > 0.0) ? qi[a] * lq[a] : 0.0;

        // own expectation, :282-291 (root) / :327-358 (hidden non-root), and message to the parent
        if (i == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) F -= qi[a] * lpi[a];
        } else {
            double cross = 0.0;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double m = 0.0;
#pragma unroll
                for (int h = 0; h < 4; ++h) m += qi[h] * lt[l * 4 + h];
                cross += st.q(p, l) * m;
                st.cs(p, l) += m;
            }
            F -= cross;
        }

        // leaf children: restart this node's child sum with their contributions
        double c[4] = {0.0, 0.0, 0.0, 0.0};   // observed leaves: sum_c log T_c(a, x_c)   (:147-148,181-182)
        double tsum = 0.0;                     // gap leaves: same value for every a
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int sp = tp.leaf_species[k];
            const int x = column_obs(bases, gaps, plane_stride, sp, revcomp);
            if (x < 4) {
                const double* ltc = lt_of(tb, tp.leaf_node[k]);
#pragma unroll
                for (int a = 0; a < 4; ++a) c[a] += ltc[a * 4 + x];
            } else {
                // unobserved leaf: every table row is log pi ("independent" table); its own update only
                // needs q of the parent, which is final for this sweep
                const double qs = (qi[0] + qi[1]) + (qi[2] + qi[3]);
                double ql[4], lql[4];
                if (update) {
                    double e[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) e[a] = exp(qs * lpi[a]);
                    const double z = (e[0] + e[1]) + (e[2] + e[3]);
                    const double lz = log(z);
#pragma unroll
                    for (int a = 0; a < 4; ++a) { ql[a] = e[a] / z; lql[a] = qs * lpi[a] - lz; }
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) { ql[a] = 0.25; lql[a] = -1.3862943611198906; }
                }
                if (sink && sink->q_leaf) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) sink->q_leaf[k * 4 + a] = ql[a];
                }
                double t = 0.0, ent = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) { t += ql[a] * lpi[a]; ent += (ql[a] > 0.0) ? ql[a] * lql[a] : 0.0; }
                F += ent - qs * t;             // entropy (:250-267) and hidden non-root term (:327-358)
                tsum += t;
            }
        }
        {   // observed-children expectation, :292-326
            double d = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a) d += qi[a] * c[a];
            F -= d;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) c[a] += tsum;
#pragma unroll
        for (int a = 0; a < 4; ++a) st.cs(i, a) = c[a];
        if (sink && sink->q_internal) {
#pragma unroll
            for (int a = 0; a < 4; ++a) sink->q_internal[i * 4 + a] = qi[a];
        }
    }
    return F;
}

PHY_HD void meanfield_init(const TreeProg& tp, const State& st) {
    for (int i = 0; i < tp.n_internal; ++i)
#pragma unroll
        for (int a = 0; a < 4; ++a) { st.q(i, a) = 0.25; st.cs(i, a) = 0.0; }
}

// Free energy of one tree for FIXED q (M-step objective, optimizing/AbstractFreeEnergyOptimizer.java:95-114):
// q_internal [I][4], q_leaf [S][4] (rows of gap leaves only), tables tb of the parameter set under test.
PHY_HD double free_energy_fixed_q(const TreeProg& tp, const double* tb, const double* q_internal,
                                  const double* q_leaf, const uint32_t* bases, const uint32_t* gaps,
                                  int64_t plane_stride, bool revcomp) {
    double F = 0.0;
    const double* lpi = tb;
    for (int i = 0; i < tp.n_internal; ++i) {
        const double* qi = q_internal + i * 4;
#pragma unroll
        for (int a = 0; a < 4; ++a) F += (qi[a] > 0.0) ? qi[a] * log(qi[a]) : 0.0;
        if (i == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) F -= qi[a] * lpi[a];
        } else {
            const double* qp = q_internal + tp.par_internal[i] * 4;
            const double* lt = lt_of(tb, tp.node_of[i]);
            double cross = 0.0;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double m = 0.0;
#pragma unroll
                for (int h = 0; h < 4; ++h) m += qi[h] * lt[l * 4 + h];
                cross += qp[l] * m;
            }
            F -= cross;
        }
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = column_obs(bases, gaps, plane_stride, tp.leaf_species[k], revcomp);
            if (x < 4) {
                const double* ltc = lt_of(tb, tp.leaf_node[k]);
                double d = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) d += qi[a] * ltc[a * 4 + x];
                F -= d;
            } else {
                const double* ql = q_leaf + k * 4;
                double t = 0.0, ent = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) { t += ql[a] * lpi[a]; ent += (ql[a] > 0.0) ? ql[a] * log(ql[a]) : 0.0; }
                F += ent - ((qi[0] + qi[1]) + (qi[2] + qi[3])) * t;
            }
        }
    }
    return F;
}

// Felsenstein pruning: P(column) for one tree, gaps as missing data.  Tables here are the
// PROBABILITIES (not logs): pt[0..3] = pi, node n >= 1 at 4 + 16*(n-1), [a][b] = P(b | a).
// Scratch: st.q(i, a) holds the partial likelihood of internal node i.
PHY_HD double pruning_likelihood(const TreeProg& tp, const double* pt, const State& st, const uint32_t* bases,
                                 const uint32_t* gaps, int64_t plane_stride, bool revcomp) {
    for (int i = tp.n_internal - 1; i >= 0; --i) {
        double part[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) part[a] = 1.0;
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = column_obs(bases, gaps, plane_stride, tp.leaf_species[k], revcomp);
            const double* t = lt_of(pt, tp.leaf_node[k]);
#pragma unroll
            for (int a = 0; a < 4; ++a)
                part[a] *= (x < 4) ? t[a * 4 + x] : ((t[a * 4] + t[a * 4 + 1]) + (t[a * 4 + 2] + t[a * 4 + 3]));
        }
        // internal children were processed earlier (they have larger pre-order indices) and left
        // their message in cs(i, .) as a running product
#pragma unroll
        for (int a = 0; a < 4; ++a) part[a] *= st.cs(i, a);
        if (i == 0) {
            double p = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a) p += pt[a] * part[a];
            return p;
        }
        const double* t = lt_of(pt, tp.node_of[i]);
        const int p = tp.par_internal[i];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double m = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) m += t[a * 4 + b] * part[b];
            st.cs(p, a) *= m;
        }
    }
    return 0.0;
}

PHY_HD void pruning_init(const TreeProg& tp, const State& st) {
    for (int i = 0; i < tp.n_internal; ++i)
#pragma unroll
        for (int a = 0; a < 4; ++a) st.cs(i, a) = 1.0;
}

}  // namespace phyfoo
