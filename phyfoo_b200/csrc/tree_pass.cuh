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
//    pi row, bayesNet/CPF.java:146-152, so it only needs the parent's q of the same sweep);
//  * "pass 0" (free energy of the uniform start, MeanFieldForBayesNet.java:105) is the same code as
//    an update pass with the exponent forced to 0 (exp(0)/4 = 0.25 exactly), so threads of one warp
//    that are in different sweeps of different windows never diverge.
#pragma once
#include "fp64_math.cuh"      // brings <math.h> / <stdint.h> (or their stand-ins under NVRTC)

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

// Table accessor: entry e of this thread's position.  Entries: [0..3] = log pi (or pi), then for
// tree node id n >= 1 sixteen entries at 4 + 16*(n-1), row-major [parent value a][own value b].
// Device layout is [entry][position] (stride = W) so that the W threads of a window group read
// W consecutive doubles; the host emulation uses stride 1.
struct Tab {
    const double* base;
    int stride;
    PHY_HD double operator()(int e) const { return base[e * stride]; }
};
PHY_HD int tab_of(int node) { return 4 + 16 * (node - 1); }

// One packed column held in registers (include/phyfoo.h "packed columns", S <= 32): 2 bits per
// species in b, 1 gap bit per species in g.
struct ColWords {
    uint64_t b;
    uint32_t g;
    // 0..3, or 4 = unobserved
    PHY_HD int obs(int s) const { return ((g >> s) & 1u) ? 4 : (int)((b >> (2 * s)) & 3u); }
};
// jstacs UnobservableDNAAlphabet.java:88-95: complement = 3 - code, gap stays gap.  Gap species
// store base bits 00, which complement to 11; obs() looks at the gap bit first, so that is harmless.
PHY_HD ColWords complement(ColWords c) { c.b = ~c.b; return c; }

// State accessors: q[i][a] and child-sum cs[i][a], strided so that consecutive threads touch
// consecutive doubles (conflict-free shared memory / coalesced global memory).
struct State {
    double* base;
    int stride;  // number of threads sharing the buffer
    // 32-bit index arithmetic (the buffers are far below 2^31 doubles): one IMAD per access instead of a 64-bit chain
    PHY_HD double& q(int i, int a) const { return base[(i * 8 + a) * stride]; }
    PHY_HD double& cs(int i, int a) const { return base[(i * 8 + 4 + a) * stride]; }
};

// Sink for the converged q of one tree (M-step "prepare"): entry (i, a) at q_internal[(i*4+a)*stride],
// gap-leaf entry (CSR leaf slot k, a) at q_leaf[(k*4+a)*stride]; rows of observed leaves are left untouched.
struct QSink {
    double* q_internal;
    double* q_leaf;
    size_t stride;
};

// a product that must stay a product: never contracted into an FMA with a following add, so that the one-thread and the
// four-lane mapping of the pass round identically
PHY_HD double phy_mul(double x, double y) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(x, y);
#else
    return x * y;
#endif
}

// correctly rounded reciprocal (= 1.0 / z bit for bit, without the general division's slow paths)
PHY_HD double phy_rcp(double z) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(z);
#else
    return 1.0 / z;
#endif
}

// One pass over the tree.  update == false: the exponents are forced to 0, i.e. q is (re)set to
// uniform; this pass evaluates the free energy of the uniform start and primes the child sums.
// Returns this tree's free energy F = part1 - part2.
//
// CANONICAL ARITHMETIC (shared bit for bit by meanfield_pass below -- one thread per tree -- and by
// meanfield_pass_quad -- four lanes per tree, lane a = symbol a; tests/test_gpu_memo.py compares the two):
//   own[a]  = sum_l q_par[l] * logT[l][a]        (fma chain l = 0..3 from 0; root: own[a] = log pi[a])
//   s[a]    = own[a] + cs[a]                     (cs = child sums left by the previous pass)
//   e[a]    = exp(update ? s[a] : 0);  z = ((e0 + e1) + e2) + e3;  rz = 1 / z;  lz = log z
//   q[a]    = e[a] * rz;  lq[a] = (update ? s[a] : 0) - lz
//   c[a]    = sum over observed leaf children of logT_leaf[a][x]   (leaf order, from 0)
//   f[a]    = q[a] * ((lq[a] - own[a]) - c[a])    entropy, own expectation (sum_a q[a] own[a] is the term of
//             MeanFieldForBayesNet.java:282-291 / :327-358) and the observed leaves' terms (:292-326)
//   F      += ((f0 + f1) + f2) + f3
//   m[l]    = sum_h q[h] * logT[l][h]            (fma chain h = 0..3 from 0);  cs_par[l] += m[l]
//   gap leaf (table = pi in every row): qs = ((q0 + q1) + q2) + q3; x[a] = qs * log pi[a]; el[a] = exp(update ? x[a] : 0);
//             zl, rzl, lzl as above; ql[a] = el[a] * rzl; lql[a] = (update ? x[a] : 0) - lzl;
//             tl = sum_a ql[a] * log pi[a] (fma chain from 0); fl[a] = ql[a] * (lql[a] - x[a]); F += ((fl0 + fl1) + fl2) + fl3
//   cs[a]   = c[a] + (sum of tl over the gap leaves, leaf order, from 0)
PHY_HD double meanfield_pass(const TreeProg& tp, const Tab& tb, const State& st, const ColWords& cw,
                             bool update) {
    double F = 0.0;
    double lpi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) lpi[a] = tb(a);
    for (int i = 0; i < tp.n_internal; ++i) {
        const int p = tp.par_internal[i];
        double lt[16];
        double own[4], s[4], qp[4];
        if (i == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) own[a] = lpi[a];
        } else {
            const int e0 = tab_of(tp.node_of[i]);
#pragma unroll
            for (int k = 0; k < 16; ++k) lt[k] = tb(e0 + k);
#pragma unroll
            for (int l = 0; l < 4; ++l) qp[l] = st.q(p, l);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < 4; ++l) acc = fma(qp[l], lt[l * 4 + a], acc);   // own rows, MeanField :117-135
                own[a] = acc;
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) s[a] = own[a] + st.cs(i, a);                    // + children, :139-191
        double e[4], qi[4], lq[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) e[a] = phy_exp(update ? s[a] : 0.0);            // :192, no max shift
        const double z = ((e[0] + e[1]) + e[2]) + e[3];                             // :195-198
        const double rz = phy_rcp(z);
        const double lz = phy_log(z);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            qi[a] = phy_mul(e[a], rz);                                              // :199-201
            lq[a] = (update ? s[a] : 0.0) - lz;                                     // = log q without a second log
            st.q(i, a) = qi[a];
        }
        // message to the parent: expectation of the log table over this node's new q
        if (i != 0) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double m = 0.0;
#pragma unroll
                for (int h = 0; h < 4; ++h) m = fma(qi[h], lt[l * 4 + h], m);
                st.cs(p, l) += m;
            }
        }
        // leaf children: restart this node's child sum with their contributions
        double c[4] = {0.0, 0.0, 0.0, 0.0};   // observed leaves: sum_c log T_c(a, x_c)   (:147-148,181-182)
        double tsum = 0.0;                     // gap leaves: same value for every a
        double Fgap = 0.0;
        bool any_gap = false;
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = cw.obs(tp.leaf_species[k]);
            if (x < 4) {
                const int e0 = tab_of(tp.leaf_node[k]);
#pragma unroll
                for (int a = 0; a < 4; ++a) c[a] += tb(e0 + a * 4 + x);
            } else any_gap = true;
        }
        {
            double f[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) f[a] = phy_mul(qi[a], (lq[a] - own[a]) - c[a]);
            F += ((f[0] + f[1]) + f[2]) + f[3];
        }
        if (any_gap) {
            // unobserved leaves: every table row is log pi ("independent" table, bayesNet/CPF.java:146-152); their own
            // update only needs q of the parent, which is final for this sweep
            const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
            for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
                if (cw.obs(tp.leaf_species[k]) < 4) continue;
                double xl[4], el[4], ql[4], fl[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) { xl[a] = phy_mul(qs, lpi[a]); el[a] = phy_exp(update ? xl[a] : 0.0); }
                const double zl = ((el[0] + el[1]) + el[2]) + el[3];
                const double rzl = phy_rcp(zl);
                const double lzl = phy_log(zl);
                double t = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    ql[a] = phy_mul(el[a], rzl);
                    const double lql = (update ? xl[a] : 0.0) - lzl;
                    t = fma(ql[a], lpi[a], t);
                    fl[a] = phy_mul(ql[a], lql - xl[a]);      // entropy (:250-267) and hidden non-root term (:327-358)
                }
                Fgap += ((fl[0] + fl[1]) + fl[2]) + fl[3];
                tsum += t;
            }
            F += Fgap;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) st.cs(i, a) = c[a] + tsum;
    }
    return F;
}

#if defined(__CUDACC__)
// The same pass with FOUR LANES PER TREE (lane a owns symbol a): the canonical arithmetic above, with the four exp of a
// node on four lanes and two exchanges per node (e, f).  State layout: Q(i, a) / CS(i, a) at base[(i * 8 + a) * stride]
// / base[(i * 8 + 4 + a) * stride] exactly like State, so both mappings can run on the same buffers.  qb = first lane of
// the quad, mask = its four lanes; every lane of the quad returns the tree's F.
__device__ __forceinline__ double quad_sum4(unsigned mask, int qb, double v) {
    const double v0 = __shfl_sync(mask, v, qb), v1 = __shfl_sync(mask, v, qb + 1);
    const double v2 = __shfl_sync(mask, v, qb + 2), v3 = __shfl_sync(mask, v, qb + 3);
    return ((v0 + v1) + v2) + v3;
}
// WARP_UNIFORM: every lane of the warp runs the pass together (the caller guarantees it), so the node-level exchanges
// use the full mask (a runtime partial mask costs a convergence sequence per collective); gap leaves always use `mask`.
template <bool WARP_UNIFORM>
__device__ __forceinline__ double meanfield_pass_quad(const TreeProg& tp, const Tab& tb, const State& st, const ColWords& cw,
                                                      bool update, int a, int qb, unsigned mask) {
    const unsigned nm = WARP_UNIFORM ? 0xffffffffu : mask;
    double F = 0.0;
    const double lpia = tb(a);
    for (int i = 0; i < tp.n_internal; ++i) {
        const int p = tp.par_internal[i];
        const int e0 = i ? tab_of(tp.node_of[i]) : 0;
        double own;
        if (i == 0) own = lpia;
        else {
            own = 0.0;
#pragma unroll
            for (int l = 0; l < 4; ++l) own = fma(st.q(p, l), tb(e0 + l * 4 + a), own);
        }
        const double s = own + st.cs(i, a);
        const double ex = update ? s : 0.0;
        const double e = phy_exp(ex);
        const double ev0 = __shfl_sync(nm, e, qb), ev1 = __shfl_sync(nm, e, qb + 1);
        const double ev2 = __shfl_sync(nm, e, qb + 2), ev3 = __shfl_sync(nm, e, qb + 3);
        const double z = ((ev0 + ev1) + ev2) + ev3;
        const double rz = phy_rcp(z);
        const double lz = phy_log(z);
        const double q0 = phy_mul(ev0, rz), q1 = phy_mul(ev1, rz), q2 = phy_mul(ev2, rz), q3 = phy_mul(ev3, rz);
        const double qa = phy_mul(e, rz);
        const double lq = ex - lz;
        __syncwarp(nm);                        // the quad has read the old q of this node's neighbours
        st.q(i, a) = qa;
        if (i != 0) {
            // lane l = a computes m[l] and adds it to its own slot of the parent's child sums
            double m = fma(q0, tb(e0 + a * 4 + 0), 0.0);
            m = fma(q1, tb(e0 + a * 4 + 1), m);
            m = fma(q2, tb(e0 + a * 4 + 2), m);
            m = fma(q3, tb(e0 + a * 4 + 3), m);
            st.cs(p, a) += m;
        }
        double c = 0.0;
        bool any_gap = false;
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = cw.obs(tp.leaf_species[k]);
            if (x < 4) c += tb(tab_of(tp.leaf_node[k]) + a * 4 + x);
            else any_gap = true;
        }
        F += quad_sum4(nm, qb, phy_mul(qa, (lq - own) - c));
        double tsum = 0.0;
        if (any_gap) {
            const double qs = ((q0 + q1) + q2) + q3;
            const double lp0 = tb(0), lp1 = tb(1), lp2 = tb(2), lp3 = tb(3);
            double Fgap = 0.0;
            for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
                if (cw.obs(tp.leaf_species[k]) < 4) continue;
                const double xl = phy_mul(qs, lpia);
                const double exl = update ? xl : 0.0;
                const double el = phy_exp(exl);
                const double l0 = __shfl_sync(mask, el, qb), l1 = __shfl_sync(mask, el, qb + 1);
                const double l2 = __shfl_sync(mask, el, qb + 2), l3 = __shfl_sync(mask, el, qb + 3);
                const double zl = ((l0 + l1) + l2) + l3;
                const double rzl = phy_rcp(zl);
                const double lzl = phy_log(zl);
                double t = fma(phy_mul(l0, rzl), lp0, 0.0);
                t = fma(phy_mul(l1, rzl), lp1, t);
                t = fma(phy_mul(l2, rzl), lp2, t);
                t = fma(phy_mul(l3, rzl), lp3, t);
                const double ql = phy_mul(el, rzl);
                Fgap += quad_sum4(mask, qb, phy_mul(ql, (exl - lzl) - xl));
                tsum += t;
            }
            F += Fgap;
        }
        st.cs(i, a) = c + tsum;
        __syncwarp(nm);                        // q(i, .) and cs are visible to the quad before the next node reads them
    }
    return F;
}
#endif

// Before pass 0 of a new window only the child sums need a defined value (q is overwritten by pass 0;
// cs(i, .) is restarted when node i is visited, before any child adds to it), so nothing to do.
// The state slots can hold anything finite-or-not: pass 0 never reads them into a result
// (update == false selects 0.0 instead of s[a]); q(p, .) of the parent is rewritten by pass 0 before
// the child reads it.

// After the last sweep of a window: write the converged q of this tree (internal nodes from the state, gap
// leaves recomputed from their parent's final q with the arithmetic of the pass, so the bits are the same)
// and the tree's entropy term  sum_hidden sum_h q log q  (MeanFieldForBayesNet.java:250-267), which does not
// depend on the parameters and is therefore computed once for the whole M-step.
PHY_HD double store_q(const TreeProg& tp, const Tab& tb, const State& st, const ColWords& cw, const QSink& sink) {
    double ent = 0.0;
    double lpi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) lpi[a] = tb(a);
    for (int i = 0; i < tp.n_internal; ++i) {
        double qi[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            qi[a] = st.q(i, a);
            sink.q_internal[(size_t)(i * 4 + a) * sink.stride] = qi[a];
            ent += (qi[a] > 0.0) ? qi[a] * phy_log(qi[a]) : 0.0;
        }
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            if (cw.obs(tp.leaf_species[k]) < 4) continue;
            const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
            double el[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) el[a] = phy_exp(qs * lpi[a]);
            const double rzl = 1.0 / (((el[0] + el[1]) + el[2]) + el[3]);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double ql = el[a] * rzl;
                sink.q_leaf[(size_t)(k * 4 + a) * sink.stride] = ql;
                ent += (ql > 0.0) ? ql * phy_log(ql) : 0.0;
            }
        }
    }
    return ent;
}

// E_q[ sum_nodes log CPF ] of one tree for FIXED q ("part2" of calcFreeEnergy, MeanFieldForBayesNet.java:269-364)
// under the tables tb of the parameter set under test; free energy = entropy - this.
// M-step objective: optimizing/AbstractFreeEnergyOptimizer.java:95-114.
PHY_HD double expected_log_cpf(const TreeProg& tp, const Tab& tb, const double* q_internal, const double* q_leaf,
                               size_t qstride, const ColWords& cw) {
    double acc = 0.0;
    double lpi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) lpi[a] = tb(a);
    for (int i = 0; i < tp.n_internal; ++i) {
        double qi[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) qi[a] = q_internal[(size_t)(i * 4 + a) * qstride];
        if (i == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) acc += qi[a] * lpi[a];                       // :282-291
        } else {
            const int pi_ = tp.par_internal[i];
            const int e0 = tab_of(tp.node_of[i]);
            double cross = 0.0;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double m = 0.0;
#pragma unroll
                for (int h = 0; h < 4; ++h) m += qi[h] * tb(e0 + l * 4 + h);
                cross += q_internal[(size_t)(pi_ * 4 + l) * qstride] * m;           // :327-358
            }
            acc += cross;
        }
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = cw.obs(tp.leaf_species[k]);
            if (x < 4) {
                const int e0 = tab_of(tp.leaf_node[k]);
                double d = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) d += qi[a] * tb(e0 + a * 4 + x);         // :292-326
                acc += d;
            } else {
                double t = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) t += q_leaf[(size_t)(k * 4 + a) * qstride] * lpi[a];
                acc += (((qi[0] + qi[1]) + qi[2]) + qi[3]) * t;                     // independent table: every row is log pi
            }
        }
    }
    return acc;
}

// The same expectation under K parameter sets at once (forward-difference gradient: K = dim + 1): the q loads are
// shared, the per-set arithmetic and its order are those of expected_log_cpf, so acc[k] carries the same bits.
// tabs: K tables of E entries each, entry e of set k at tabs[k * E + e].
template <int K>
PHY_HD void expected_log_cpf_multi(const TreeProg& tp, const double* tabs, int E, const double* q_internal,
                                   const double* q_leaf, size_t qstride, const ColWords& cw, double* acc) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int i = 0; i < tp.n_internal; ++i) {
        double qi[4], qp[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) qi[a] = q_internal[(size_t)(i * 4 + a) * qstride];
        if (i == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double* tb = tabs + (size_t)k * E;
                double r = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) r += qi[a] * tb[a];
                acc[k] += r;
            }
        } else {
            const int pi_ = tp.par_internal[i];
            const int e0 = tab_of(tp.node_of[i]);
#pragma unroll
            for (int l = 0; l < 4; ++l) qp[l] = q_internal[(size_t)(pi_ * 4 + l) * qstride];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double* tb = tabs + (size_t)k * E + e0;
                double cross = 0.0;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    double m = 0.0;
#pragma unroll
                    for (int h = 0; h < 4; ++h) m += qi[h] * tb[l * 4 + h];
                    cross += qp[l] * m;
                }
                acc[k] += cross;
            }
        }
        for (int c = tp.leaf_begin[i]; c < tp.leaf_begin[i + 1]; ++c) {
            const int x = cw.obs(tp.leaf_species[c]);
            if (x < 4) {
                const int e0 = tab_of(tp.leaf_node[c]);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double* tb = tabs + (size_t)k * E + e0;
                    double d = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) d += qi[a] * tb[a * 4 + x];
                    acc[k] += d;
                }
            } else {
                double ql[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) ql[a] = q_leaf[(size_t)(c * 4 + a) * qstride];
                const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double* tb = tabs + (size_t)k * E;
                    double t = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) t += ql[a] * tb[a];
                    acc[k] += qs * t;
                }
            }
        }
    }
}

// Felsenstein pruning: P(column) for one tree, gaps as missing data.  Tables here are the
// PROBABILITIES (not logs).  Scratch: st.cs(i, a) holds the running product of the messages of the
// internal children of internal node i.
PHY_HD double pruning_likelihood(const TreeProg& tp, const Tab& pt, const State& st, const ColWords& cw) {
    for (int i = 0; i < tp.n_internal; ++i)
#pragma unroll
        for (int a = 0; a < 4; ++a) st.cs(i, a) = 1.0;
    for (int i = tp.n_internal - 1; i >= 0; --i) {
        double part[4];
        // internal children were processed earlier (they have larger pre-order indices) and left
        // their messages in cs(i, .) as a running product
#pragma unroll
        for (int a = 0; a < 4; ++a) part[a] = st.cs(i, a);
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            const int x = cw.obs(tp.leaf_species[k]);
            const int e0 = tab_of(tp.leaf_node[k]);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double v = (x < 4) ? pt(e0 + a * 4 + x)
                                         : (((pt(e0 + a * 4) + pt(e0 + a * 4 + 1)) + pt(e0 + a * 4 + 2)) + pt(e0 + a * 4 + 3));
                part[a] *= v;
            }
        }
        if (i == 0) {
            double p = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a) p += pt(a) * part[a];
            return p;
        }
        const int e0 = tab_of(tp.node_of[i]);
        const int p = tp.par_internal[i];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double m = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) m += pt(e0 + a * 4 + b) * part[b];
            st.cs(p, a) *= m;
        }
    }
    return 0.0;
}

}  // namespace phyfoo
