// model_host.hpp -- host-side model state of libphyfoo: flattens the species tree into the
// "tree program" the kernels walk and turns (evolution model, branch lengths, pi) into the
// per-position CPF tables and their logs.
//
// Reference semantics restated here (it runs once per parameter set, on the host, like the Java):
//   bayesNet/VirtualTree.java:82-96      CPFs from the evolution model, one tree per position
//   evolution/FS81alpha.java:76-92       P(b|a) = (a==b)(1-alpha) + alpha*pi_b, alpha = branch "length"
//   evolution/FS81beta.java:81-103       alpha' = exp(-v/(1-sum pi^2)) clamped to {0,1} within 1e-4
//   evolution/HKY.java:77-106 (+ :32)    linear HKY whose transversion rate is always 0 as implemented
//   bayesNet/CPF.java:56-59,238-256      row index l = sum_p 4^p * value(parent p); parent 0 = tree parent
//   util/MatrixLinearisation.java:15-19  pi = softmax over consecutive blocks of 4 (lambda space)
#pragma once
#include <cmath>
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/phyfoo.h"

namespace phyfoo {

struct ModelHost {
    int kind = PHYFOO_MODEL_FG, W = 0, order = 0, N = 0, S = 0, I = 0;
    int evol = PHYFOO_EVOL_F81ALPHA;
    double hky_ratio = 0.0;                  // PHYFOO_EVOL_HKY_RATIO only: beta = alpha * hky_ratio
    int mode = PHYFOO_MODE_MEANFIELD;
    int max_sweeps = 100, min_sweeps = 2;
    double tol = 1e-5;
    std::vector<int> parent, leaf_species;
    std::vector<double> branch;              // [W][N]
    std::vector<std::vector<double>> pi;     // per position [ctx][4], ctx = 4^{#horizontal parents of the root}
    // tree program (see tree_pass.cuh)
    std::vector<int> par_internal, node_of, leaf_begin, leaf_node, leaf_sp, internal_of;
    std::vector<int> ichild_begin, ichild;   // CSR of the internal tree children of internal node i (order-1 kernel)
    // order-0 tables: E entries per position, host layout [W][E]
    int E = 0;
    std::vector<double> logtab, probtab;
    // order >= 1 tables: per position, per node, [4^P][4] (see ctx_rows); host layout described in tables1()
    uint64_t version = 0;                    // bumped on every parameter change (device copy is lazy)

    int ctx_rows(int pos) const {            // PhyloBayesModel.java:338-339 / PhyloBackground.java:367
        if (kind == PHYFOO_MODEL_FG) return (order >= 1 && pos > 0) ? 4 : 1;
        int r = 1;
        for (int j = 0; j < order && j < pos; ++j) r *= 4;
        return r;
    }

    void init(const phyfoo_model_desc& d) {
        kind = d.kind; order = d.order; N = d.n_nodes; evol = d.evol; mode = d.mode;
        W = (kind == PHYFOO_MODEL_FG) ? d.width : d.order + 1;
        if (d.max_sweeps > 0) max_sweeps = d.max_sweeps;
        if (d.min_sweeps > 0) min_sweeps = d.min_sweeps;
        if (d.tol > 0) tol = d.tol;
        if (N < 2 || W < 1) throw std::invalid_argument("model: need >= 2 tree nodes and width >= 1");
        if (order < 0) throw std::invalid_argument("model: order must be >= 0");
        if (evol == PHYFOO_EVOL_HKY_RATIO) {
            hky_ratio = d.hky_ratio;
            if (order != 0) throw std::invalid_argument("HKY with its ratio in effect is offered for order-0 models only (the order-1 kernels assume F81 tables)");
            if (!(hky_ratio > 0) || !std::isfinite(hky_ratio)) throw std::invalid_argument("HKY: the transversion / transition ratio must be positive");
        }
        if (kind != PHYFOO_MODEL_FG && kind != PHYFOO_MODEL_BG) throw std::invalid_argument("model: unknown kind");
        if (kind == PHYFOO_MODEL_FG && order != 0 && order != 1)
            throw std::invalid_argument("motif order must be 0 or 1 (models/PhyloBayesModel.java:363)");
        if (mode != PHYFOO_MODE_MEANFIELD && mode != PHYFOO_MODE_EXACT) throw std::invalid_argument("model: unknown mode");
        if (d.allow_independent_leaf == 0)
            throw std::invalid_argument("CPF.ALLOW_RETURN_INDEPENDENT_TRANSITION=false is not supported");
        if (!d.parent || !d.leaf_species || !d.branch || !d.pi) throw std::invalid_argument("model: null array in descriptor");
        parent.assign(d.parent, d.parent + N);
        leaf_species.assign(d.leaf_species, d.leaf_species + N);
        if (parent[0] != -1) throw std::invalid_argument("model: node 0 must be the root (DFS pre-order)");
        std::vector<int> nchild(N, 0);
        for (int k = 1; k < N; ++k) {
            if (parent[k] < 0 || parent[k] >= k) throw std::invalid_argument("model: nodes must be in DFS pre-order");
            nchild[parent[k]]++;
        }
        S = 0;
        for (int k = 0; k < N; ++k) {
            bool leaf = nchild[k] == 0;
            if (leaf != (leaf_species[k] >= 0)) throw std::invalid_argument("model: leaf_species must be >= 0 exactly at leaves");
            if (leaf) S++;
        }
        std::vector<char> seen(S, 0);
        for (int k = 0; k < N; ++k)
            if (leaf_species[k] >= 0) {
                if (leaf_species[k] >= S || seen[leaf_species[k]]) throw std::invalid_argument("model: leaf_species must be a permutation of 0..S-1");
                seen[leaf_species[k]] = 1;
            }
        branch.assign(d.branch, d.branch + (size_t)W * N);
        pi.resize(W);
        size_t off = 0;
        for (int t = 0; t < W; ++t) {
            int n = ctx_rows(t) * 4;
            pi[t].assign(d.pi + off, d.pi + off + n);
            off += n;
        }
        build_prog();
        E = 4 + 16 * (N - 1);
        if (order == 0) {
            logtab.assign((size_t)W * E, 0.0);
            probtab.assign((size_t)W * E, 0.0);
            for (int t = 0; t < W; ++t) build_tables(t);
        }
        version++;
    }

    void build_prog() {
        internal_of.assign(N, -1);
        node_of.clear(); par_internal.clear();
        std::vector<int> nchild(N, 0);
        for (int k = 1; k < N; ++k) nchild[parent[k]]++;
        for (int k = 0; k < N; ++k)
            if (nchild[k] > 0) {
                internal_of[k] = (int)node_of.size();
                node_of.push_back(k);
                par_internal.push_back(k == 0 ? -1 : internal_of[parent[k]]);
            }
        I = (int)node_of.size();
        leaf_begin.assign(I + 1, 0);
        leaf_node.clear(); leaf_sp.clear();
        for (int i = 0; i < I; ++i) {
            leaf_begin[i] = (int)leaf_node.size();
            for (int k = 1; k < N; ++k)
                if (parent[k] == node_of[i] && nchild[k] == 0) { leaf_node.push_back(k); leaf_sp.push_back(leaf_species[k]); }
        }
        leaf_begin[I] = (int)leaf_node.size();
        ichild_begin.assign(I + 1, 0);
        ichild.clear();
        for (int i = 0; i < I; ++i) {
            ichild_begin[i] = (int)ichild.size();
            for (int k = 1; k < N; ++k)
                if (parent[k] == node_of[i] && nchild[k] > 0) ichild.push_back(internal_of[k]);
        }
        ichild_begin[I] = (int)ichild.size();
    }

    // the flattened program as one int array: par_internal[I] node_of[I] leaf_begin[I+1] leaf_node[S] leaf_species[S]
    // ichild_begin[I+1] ichild[I-1]
    std::vector<int> prog_flat() const {
        std::vector<int> v;
        v.insert(v.end(), par_internal.begin(), par_internal.end());
        v.insert(v.end(), node_of.begin(), node_of.end());
        v.insert(v.end(), leaf_begin.begin(), leaf_begin.end());
        v.insert(v.end(), leaf_node.begin(), leaf_node.end());
        v.insert(v.end(), leaf_sp.begin(), leaf_sp.end());
        v.insert(v.end(), ichild_begin.begin(), ichild_begin.end());
        v.insert(v.end(), ichild.begin(), ichild.end());
        return v;
    }

    // 4x4 transition for one context row of pi and one branch; out[a*4+b] = P(b | a)
    void transition(const double* p, double len, double* out) const {
        if (evol == PHYFOO_EVOL_F81ALPHA) {
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) out[a * 4 + b] = (a == b) ? (1 - len + p[b] * len) : (p[b] * len);
        } else if (evol == PHYFOO_EVOL_F81BETA) {
            double beta = 1. / (1. - p[0] * p[0] - p[1] * p[1] - p[2] * p[2] - p[3] * p[3]);
            double ta = std::exp(-len * beta);
            if (ta < 1e-4) ta = 0; else if (ta > 1 - 1e-4) ta = 1;
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) out[a * 4 + b] = (a == b) ? (ta + p[b] * (1. - ta)) : (p[b] * (1. - ta));
        } else if (evol == PHYFOO_EVOL_HKY_AS_IMPLEMENTED) {
            const double al = len, be = len * 0.0;
            const double aA = al * p[0], aC = al * p[1], aG = al * p[2], aT = al * p[3];
            const double bA = be * p[0], bC = be * p[1], bG = be * p[2], bT = be * p[3];
            const double v[16] = {1 - (aG + bT + bC), bC, aG, bT, bA, 1 - (aT + bA + bG), bG, aT,
                                  aA, bC, 1 - (aA + bT + bC), bT, bA, aC, bG, 1 - (aC + bA + bG)};
            for (int i = 0; i < 16; ++i) out[i] = v[i];
        } else if (evol == PHYFOO_EVOL_HKY_RATIO) {      // HKY.java:77-106 as written, with the ratio the constructor was meant to store
            const double al = len, be = len * hky_ratio;
            const double aA = al * p[0], aC = al * p[1], aG = al * p[2], aT = al * p[3];
            const double bA = be * p[0], bC = be * p[1], bG = be * p[2], bT = be * p[3];
            const double v[16] = {1 - (aG + bT + bC), bC, aG, bT, bA, 1 - (aT + bA + bG), bG, aT,
                                  aA, bC, 1 - (aA + bT + bC), bT, bA, aC, bG, 1 - (aC + bA + bG)};
            for (int i = 0; i < 16; ++i) out[i] = v[i];
        } else throw std::invalid_argument("model: unknown evolution model id");
    }

    // order-0 tables of position t: [log pi (4)] [node 1 (16)] ... [node N-1 (16)]
    void build_tables(int t) { build_tables_into(t, pi[t].data(), &logtab[(size_t)t * E], &probtab[(size_t)t * E]); }
    // same, for an arbitrary stationary distribution p[4] (M-step: parameter sets under test); pt may be null
    void build_tables_into(int t, const double* p, double* lt, double* pt) const {
        for (int a = 0; a < 4; ++a) { if (pt) pt[a] = p[a]; lt[a] = std::log(p[a]); }
        for (int k = 1; k < N; ++k) {
            double tr[16];
            transition(p, branch[(size_t)t * N + k], tr);
            for (int i = 0; i < 16; ++i) { if (pt) pt[tab_off(k) + i] = tr[i]; lt[tab_off(k) + i] = std::log(tr[i]); }
        }
    }
    static int tab_off(int node) { return 4 + 16 * (node - 1); }

    // compact order-0 log tables for substitution tables with F81 structure (tree_pass_f81.cuh): per position
    // Ef = 4 + 8 (N - 1) entries: log pi, then per node Lo[b] (log off-diagonal, the same for every a != b) | D[b] = log
    // diagonal - Lo[b].  Returns false when a table does not have that structure bit for bit or an entry is not finite.
    int Ef() const { return 4 + 8 * (N - 1); }
    bool tables0f(std::vector<double>& out) const {
        const int ef = Ef();
        out.assign((size_t)W * ef, 0.0);
        for (int t = 0; t < W; ++t) {
            const double* pt = &probtab[(size_t)t * E];
            const double* lt = &logtab[(size_t)t * E];
            double* o = &out[(size_t)t * ef];
            for (int a = 0; a < 4; ++a) { o[a] = lt[a]; if (!std::isfinite(o[a])) return false; }
            for (int k = 1; k < N; ++k)
                for (int b = 0; b < 4; ++b) {
                    const int off_row = (b + 1) & 3;
                    for (int a = 0; a < 4; ++a)
                        if (a != b && pt[tab_off(k) + a * 4 + b] != pt[tab_off(k) + off_row * 4 + b]) return false;
                    const double lo = lt[tab_off(k) + off_row * 4 + b], ld = lt[tab_off(k) + b * 4 + b];
                    if (!std::isfinite(lo) || !std::isfinite(ld)) return false;
                    o[4 + 8 * (k - 1) + b] = lo;
                    o[4 + 8 * (k - 1) + 4 + b] = ld - lo;
                }
        }
        return true;
    }

    // CPF table of (position t, node k) for ANY order, reference layout [4^P][4], row l = a_treeparent + 4*ctx
    // (root: rows = ctx).  VirtualTree.java:82-96.
    std::vector<double> cpf(int t, int k) const {
        const int ctx = ctx_rows(t);
        std::vector<double> out;
        if (k == 0) { out = pi[t]; return out; }
        out.resize((size_t)ctx * 16);
        for (int c = 0; c < ctx; ++c) transition(&pi[t][(size_t)c * 4], branch[(size_t)t * N + k], &out[(size_t)c * 16]);
        return out;
    }

    // ---- generic net (netg.cuh): any order, the reference's parent / child lists and full CPF tables ----------------------
    // horizontal parents of tree t, nearest first (PhyloBackground.java:389-393; PhyloBayesModel.java:363-365)
    std::vector<int> horizontal_parents(int t) const {
        std::vector<int> v;
        if (kind == PHYFOO_MODEL_FG) { if (order == 1 && t > 0) v.push_back(t - 1); }
        else for (int j = 0; j < order; ++j) if (t - j > 0) v.push_back(t - j - 1);
        return v;
    }
    // rec [W * N][9], lists (parent ids; (child id, index among the child's parents) pairs), tab (log CPF, log independent table)
    void netg_build(std::vector<int>& rec, std::vector<int>& lists, std::vector<double>& tab) const {
        const int G = W * N;
        rec.assign((size_t)G * 9, 0); lists.clear(); tab.clear();
        std::vector<std::vector<int>> par(G), chi(G);
        for (int t = 0; t < W; ++t)
            for (int k = 0; k < N; ++k) {
                if (k > 0) par[t * N + k].push_back(t * N + parent[k]);
                for (int k2 = k + 1; k2 < N; ++k2) if (parent[k2] == k) chi[t * N + k].push_back(t * N + k2);
            }
        for (int t = 0; t < W; ++t)                                   // BayesNetHandler.connectVirtualTrees :230-256, in call order
            for (int hp : horizontal_parents(t))
                for (int k = 0; k < N; ++k) { chi[hp * N + k].push_back(t * N + k); par[t * N + k].push_back(hp * N + k); }
        for (int t = 0; t < W; ++t) {
            const int ctx = ctx_rows(t);
            for (int k = 0; k < N; ++k) {
                const int g = t * N + k;
                int* r = &rec[(size_t)g * 9];
                r[0] = (int)par[g].size(); r[1] = (int)lists.size();
                lists.insert(lists.end(), par[g].begin(), par[g].end());
                r[2] = (int)chi[g].size(); r[3] = (int)lists.size();
                for (int c : chi[g]) {
                    int idx = -1;
                    for (size_t q = 0; q < par[c].size(); ++q) if (par[c][q] == g) idx = (int)q;
                    lists.push_back(c); lists.push_back(idx);
                }
                int rows = 1;
                for (int q = 0; q < r[0]; ++q) rows *= 4;
                r[6] = rows; r[7] = leaf_species[k]; r[8] = t;
                const size_t before = tab.size();
                netg_node_logtab(t, k, tab);
                r[4] = (int)before;
                r[5] = leaf_species[k] >= 0 ? (int)before + rows * 4 : (int)before;
                if ((int)(tab.size() - before) != rows * 4 * (leaf_species[k] >= 0 ? 2 : 1)) throw std::invalid_argument("netg: table shape");
            }
        }
    }
    // log CPF of node (t, k) [rows][4], followed for a leaf by the log of its independent table (pi of the row's context,
    // FS81alpha.java:52-64) -- appended to out
    void netg_node_logtab(int t, int k, std::vector<double>& out) const {
        for (double x : cpf(t, k)) out.push_back(std::log(x));
        if (leaf_species[k] >= 0)
            for (int c = 0; c < ctx_rows(t); ++c)
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) out.push_back(std::log(pi[t][(size_t)c * 4 + b]));
    }
    // the tables of tree t alone, in netg_build's order (its slice of the table array)
    void netg_tree_logtab(int t, std::vector<double>& out) const {
        out.clear();
        for (int k = 0; k < N; ++k) netg_node_logtab(t, k, out);
    }

    bool tables_finite() const {
        if (order >= 1) {
            std::vector<double> t1;
            tables1(t1);
            for (double v : t1) if (!std::isfinite(v)) return false;
            return true;
        }
        for (double v : logtab) if (!std::isfinite(v)) return false;
        return true;
    }

    // order-1 motif tables (net1.cuh): per position E1 = 16 + 32 (N-1) logs; root [ctx][4] at 0; node k >= 1: the
    // reference's table [l = a_treeparent + 4 ctx][b] has F81 structure (one value for every a != b), stored as
    // off-diagonal Lo[ctx][b] at 16 + 32 (k-1) and diagonal Ld[ctx][b] 16 further.  Unused rows (t = 0 has a single
    // context) are 0.  Throws if a table does not have that structure bit for bit.
    int E1() const { return 16 + 32 * (N - 1); }
    // tables of position t for an arbitrary parameter set pi_t ([ctx_rows(t)][4]); out[E1]
    void build_tables1_into(int t, const double* pi_t, double* o) const {
        const int ctx = ctx_rows(t);
        for (int i = 0; i < E1(); ++i) o[i] = 0.0;
        for (int i = 0; i < ctx * 4; ++i) o[i] = std::log(pi_t[i]);
        for (int k = 1; k < N; ++k) {
            double* lo = o + 16 + 32 * (k - 1);
            double* ld = lo + 16;
            for (int c = 0; c < ctx; ++c) {
                double tr[16];                                   // [a][b] for context c
                transition(pi_t + (size_t)c * 4, branch[(size_t)t * N + k], tr);
                for (int b = 0; b < 4; ++b) {
                    const double off = tr[((b + 1) & 3) * 4 + b];
                    for (int a = 0; a < 4; ++a)
                        if (a != b && tr[a * 4 + b] != off)
                            throw std::invalid_argument("order-1 kernel needs F81-structured substitution tables");
                    lo[c * 4 + b] = std::log(off);
                    ld[c * 4 + b] = std::log(tr[b * 4 + b]);
                }
            }
        }
    }
    // scoring-kernel layout (net1.cuh): per position E1v = 32 + 64 (N-1): root [Lpi | LpiT], node k >= 1
    // [Lo | Ld | LoT | LdT] (T = transposed, [b][ctx]); W + 1 positions, the last one all zero
    int E1v() const { return 32 + 64 * (N - 1); }
    void tables1v(std::vector<double>& out) const {
        const int e1 = E1(), ev = E1v();
        std::vector<double> t1;
        tables1(t1);
        out.assign((size_t)(W + 1) * ev, 0.0);
        for (int t = 0; t < W; ++t) {
            const double* src = &t1[(size_t)t * e1];
            double* o = &out[(size_t)t * ev];
            for (int c = 0; c < 4; ++c)
                for (int b = 0; b < 4; ++b) { o[c * 4 + b] = src[c * 4 + b]; o[16 + b * 4 + c] = src[c * 4 + b]; }
            for (int k = 1; k < N; ++k) {
                const double* lo = src + 16 + 32 * (k - 1);
                double* d = o + 32 + 64 * (k - 1);
                for (int c = 0; c < 4; ++c)
                    for (int b = 0; b < 4; ++b) {
                        d[c * 4 + b] = lo[c * 4 + b]; d[16 + c * 4 + b] = lo[16 + c * 4 + b];
                        d[32 + b * 4 + c] = lo[c * 4 + b]; d[48 + b * 4 + c] = lo[16 + c * 4 + b];
                    }
            }
        }
    }
    void tables1(std::vector<double>& out) const {
        const int e1 = E1();
        out.assign((size_t)W * e1, 0.0);
        for (int t = 0; t < W; ++t) build_tables1_into(t, pi[t].data(), &out[(size_t)t * e1]);
    }

    // ---- wavefront kernel (net1w.cuh) ----
    // tables: per position 64 N logs, node k block [Lo | Ld | LoT | LdT] at 64 k, root block [Lpi | Lpi | LpiT | LpiT];
    // W + 1 positions (the last all zero) followed by one zero block of 64 for null list entries
    int E1w() const { return 64 * (N + 1); }   // block N of every position is all zero (null leaf entries)
    void tables1w(std::vector<double>& out) const {
        const int e1 = E1(), ew = E1w();
        std::vector<double> t1;
        tables1(t1);
        out.assign((size_t)(W + 1) * ew, 0.0);
        for (int t = 0; t < W; ++t) {
            const double* src = &t1[(size_t)t * e1];
            double* o = &out[(size_t)t * ew];
            for (int c = 0; c < 4; ++c)
                for (int b = 0; b < 4; ++b) {
                    o[c * 4 + b] = o[16 + c * 4 + b] = src[c * 4 + b];
                    o[32 + b * 4 + c] = o[48 + b * 4 + c] = src[c * 4 + b];
                }
            for (int k = 1; k < N; ++k) {
                const double* lo = src + 16 + 32 * (k - 1);
                double* d = o + 64 * k;
                for (int c = 0; c < 4; ++c)
                    for (int b = 0; b < 4; ++b) {
                        d[c * 4 + b] = lo[c * 4 + b]; d[16 + c * 4 + b] = lo[16 + c * 4 + b];
                        d[32 + b * 4 + c] = lo[c * 4 + b]; d[48 + b * 4 + c] = lo[16 + c * 4 + b];
                    }
            }
        }
    }
    // program: par_internal[I] node_of[I] leaf_node[S] leaf_species[S] ich[I][MC] lf[I][ML] sched[n_steps][4]
    // (padded child / leaf lists, -1 = null) and the list schedule of the W*I node updates into steps of <= 4:
    // node (t, i) may run once (t, parent), (t-1, i) and (t-1, child) for every internal child are in EARLIER steps.
    void wave_program(std::vector<int>& prog, int& MC, int& ML, int& n_steps) const {
        MC = 1; ML = 1;
        for (int i = 0; i < I; ++i) {
            MC = std::max(MC, ichild_begin[i + 1] - ichild_begin[i]);
            ML = std::max(ML, leaf_begin[i + 1] - leaf_begin[i]);
        }
        std::vector<int> depth(I, 0);
        for (int i = 1; i < I; ++i) depth[i] = depth[par_internal[i]] + 1;
        const int n = W * I;
        std::vector<int> done(n, -1);                       // step in which the node runs
        std::vector<int> sched;
        int left = n, step = 0;
        while (left > 0) {
            std::vector<std::pair<int, int>> ready;         // (priority, node id)
            for (int id = 0; id < n; ++id) {
                if (done[id] >= 0) continue;
                const int t = id / I, i = id % I;
                auto ok = [&](int tt, int ii) { const int d = done[tt * I + ii]; return d >= 0 && d < step; };
                bool r = true;
                if (i > 0 && !ok(t, par_internal[i])) r = false;
                if (r && t > 0) {
                    if (!ok(t - 1, i)) r = false;
                    for (int j = ichild_begin[i]; r && j < ichild_begin[i + 1]; ++j) if (!ok(t - 1, ichild[j])) r = false;
                }
                if (r) ready.push_back({depth[i] + 2 * t, id});
            }
            std::sort(ready.begin(), ready.end());
            for (int sl = 0; sl < 4; ++sl) {
                if (sl < (int)ready.size()) {
                    const int id = ready[sl].second;
                    done[id] = step; --left;
                    sched.push_back(((id / I) << 16) | (id % I));
                } else sched.push_back(-1);
            }
            ++step;
        }
        n_steps = step;
        prog.clear();
        prog.insert(prog.end(), par_internal.begin(), par_internal.end());
        prog.insert(prog.end(), node_of.begin(), node_of.end());
        prog.insert(prog.end(), leaf_node.begin(), leaf_node.end());
        prog.insert(prog.end(), leaf_sp.begin(), leaf_sp.end());
        for (int i = 0; i < I; ++i)
            for (int j = 0; j < MC; ++j) prog.push_back(ichild_begin[i] + j < ichild_begin[i + 1] ? ichild[ichild_begin[i] + j] : -1);
        for (int i = 0; i < I; ++i)
            for (int j = 0; j < ML; ++j) prog.push_back(leaf_begin[i] + j < leaf_begin[i + 1] ? leaf_begin[i] + j : -1);
        prog.insert(prog.end(), sched.begin(), sched.end());
        // node records [I][2 + 3 MC + 2 ML]: par, 64 node | per child: internal index (self if null), 64 node(child), live flag
        // | per leaf: CSR slot (S if null), 64 node(leaf) (64 N = the zero block if null)
        for (int i = 0; i < I; ++i) {
            prog.push_back(par_internal[i]);
            prog.push_back(64 * node_of[i]);
            for (int j = 0; j < MC; ++j) {
                const bool has = ichild_begin[i] + j < ichild_begin[i + 1];
                const int ci = has ? ichild[ichild_begin[i] + j] : i;
                prog.push_back(ci); prog.push_back(64 * node_of[ci]); prog.push_back(has ? 1 : 0);
            }
            for (int j = 0; j < ML; ++j) {
                const bool has = leaf_begin[i] + j < leaf_begin[i + 1];
                prog.push_back(has ? leaf_begin[i] + j : S);
                prog.push_back(has ? 64 * leaf_node[leaf_begin[i] + j] : 64 * N);
            }
        }
    }

    // ---- node-per-thread kernel (net1n.cuh) ----
    // One thread updates one hidden node (all four symbols); a window is worked on by 4 node slots, a warp holds 8 windows
    // (lane = 8 slot + window).  Everything a node update touches is addressed through a precomputed record of offsets, in
    // doubles, into the warp's region of shared memory (rows of 32 doubles = four values of a node for 8 windows):
    //   region: zero row | AB rows [I] (64 doubles each) | message rows [I] || one-hot row | q rows [W*I]
    // (the part behind || may live in global memory instead -- behind L1 -- which leaves room for twice the warps per SM;
    // offsets are the same in both placements)
    // Record layout (n1n_rec_len(MC) ints): qself, qpar, qnself, qnpar, ab, msg, tabnext, flags (1 live, 2 last position,
    // 4 message from the node's own new q: W = 1), then one message row per child (the zero row for a null entry).
    // AB row of node i: the context sums the update of (t, i) needs of the SAME node at t-1, B(a) = sum_c q_prev(c) Lo[c][a]
    // and (first half of the row) D(a) = A(a) - B(a) = sum_c q_prev(c) (Ld - Lo)[c][a]; produced by the update of (t-1, i) -- in place, the row's only reader is (t, i)
    // itself -- and by the update of the last position for position 0 of the next sweep (virtual one-hot predecessor).
    // Message row of node j: what j contributes to the exponents of its tree parent at the NEXT position,
    // msg(a) = q_j(a) A_j(a) + sum_{a' != a} q_j(a') B_j(a') with the old q of (t+1, j) and the new context sums; produced by
    // the update of (t, j), read by (t+1, parent(j)), which runs before (t+1, j) overwrites it.
    static int n1n_rec_len(int MC) { return (8 + MC + 3) & ~3; }
    // a node's 32 table entries are padded to 34 doubles: the four node slots of a warp read the same entry of four
    // different nodes at once, and with a row of 256 bytes these would all sit in the same banks
    static const int N1N_TAB_STRIDE = 34;
    struct N1nLayout { int q0, zero_row, onehot_row, ab0, msg0, sregion, region; };
    N1nLayout n1n_layout() const {
        N1nLayout L;
        L.zero_row = 0; L.ab0 = 32; L.msg0 = L.ab0 + I * 64; L.sregion = L.msg0 + I * 32;
        L.onehot_row = L.sregion; L.q0 = L.onehot_row + 32; L.region = L.q0 + W * I * 32;
        return L;
    }
    // tables of the internal nodes for the kernel's shared memory: [W][I][34] = Lo[ctx][b] | D[ctx][b] = Ld - Lo | 2 pad
    // (root: log pi and zeros): every use of the diagonal entries is linear, q_par(a) Ld + (tot - q_par(a)) Lo =
    // q_par(a) D + tot Lo, which saves a third of the multiplications of a node update
    void tables1n(std::vector<double>& out) const {
        const int e1 = E1();
        std::vector<double> t1;
        tables1(t1);
        out.assign((size_t)W * I * N1N_TAB_STRIDE, 0.0);
        for (int t = 0; t < W; ++t)
            for (int i = 0; i < I; ++i) {
                const int k = node_of[i];
                double* o = &out[((size_t)t * I + i) * N1N_TAB_STRIDE];
                const double* src = &t1[(size_t)t * e1];
                if (k == 0) for (int j = 0; j < 16; ++j) { o[j] = src[j]; o[16 + j] = 0.0; }
                else for (int j = 0; j < 16; ++j) {
                    const double lo = src[16 + 32 * (k - 1) + j], ld = src[16 + 32 * (k - 1) + 16 + j];
                    o[j] = lo; o[16 + j] = ld - lo;
                }
            }
    }
    // Free energy of the uniform start (MeanFieldForBayesNet.java:105: F_0) without a pass over the net: at q = 1/4 every
    // node contributes -1/4 sum_a (own(a) + cobs(a)) - log 4, and own does not depend on the window, so
    // F_0 = n1n_f0_const() - 1/4 sum over nodes and symbols of the window's observed-leaf sums.
    double n1n_f0_const() const {
        std::vector<double> tab;
        tables1n(tab);
        double C = 0.0;
        for (int t = 0; t < W; ++t)
            for (int i = 0; i < I; ++i) {
                const double* T = &tab[((size_t)t * I + i) * N1N_TAB_STRIDE];
                double own = 0.0;
                for (int a = 0; a < 4; ++a) {
                    double B = 0.0, D = 0.0;                       // context sums: one-hot predecessor at t = 0, uniform otherwise
                    if (t == 0) { B = T[a]; D = T[16 + a]; }
                    else for (int c = 0; c < 4; ++c) { B += 0.25 * T[c * 4 + a]; D += 0.25 * T[16 + c * 4 + a]; }
                    const double qp = i == 0 ? (a == 0 ? 1.0 : 0.0) : 0.25;      // the root's virtual parent is one-hot (D = 0 there)
                    own += qp * D + 1.0 * B;
                }
                C -= 0.25 * own + std::log(4.0);
            }
        return C;
    }
    // tables of the leaves (window set-up only, read from global memory): [W][S (CSR leaf slot)][32] = Lo | Ld
    void leaf_tables1n(std::vector<double>& out) const {
        const int e1 = E1();
        std::vector<double> t1;
        tables1(t1);
        out.assign((size_t)W * S * 32, 0.0);
        for (int t = 0; t < W; ++t)
            for (int kk = 0; kk < S; ++kk)
                for (int j = 0; j < 32; ++j) out[((size_t)t * S + kk) * 32 + j] = t1[(size_t)t * e1 + 16 + 32 * (leaf_node[kk] - 1) + j];
    }
    // program: header [n_steps, MC, rec_len, records offset] | leaf_begin[I+1] | leaf_species[S] | pad to a multiple of 4 |
    // records [n_steps][4][rec_len] (16-byte aligned: the kernel loads them as int4).
    // List schedule of the W*I node updates into steps of <= 4, longest remaining dependency chain first: node (t, i) may run
    // once (t, parent), (t-1, i) and (t-1, child) for every internal child are in EARLIER steps (see net1w.cuh for why any
    // such schedule reproduces the reference's sweep).
    void node_program(std::vector<int>& prog, int& MC, int& n_steps) const {
        MC = 1;
        for (int i = 0; i < I; ++i) MC = std::max(MC, ichild_begin[i + 1] - ichild_begin[i]);
        const int n = W * I, RL = n1n_rec_len(MC);
        const N1nLayout L = n1n_layout();
        auto deps = [&](int id, std::vector<int>& d) {
            d.clear();
            const int t = id / I, i = id % I;
            if (i > 0) d.push_back(t * I + par_internal[i]);
            if (t > 0) {
                d.push_back((t - 1) * I + i);
                for (int j = ichild_begin[i]; j < ichild_begin[i + 1]; ++j) d.push_back((t - 1) * I + ichild[j]);
            }
        };
        std::vector<int> height(n, 1), d;
        for (int id = n - 1; id >= 0; --id) {               // ids are in a topological order
            deps(id, d);
            for (int p : d) height[p] = std::max(height[p], height[id] + 1);
        }
        // list scheduling with the priority "longest remaining chain first", ties and near-ties broken by a fixed
        // pseudo-random perturbation; the shortest of a few hundred trials is kept (C3: 36 steps instead of 38 for 132
        // nodes; the steps are what every sweep of every window costs)
        std::vector<std::vector<int>> dep(n);
        for (int id = 0; id < n; ++id) { deps(id, d); dep[id] = d; }
        std::vector<int> sched;
        int step = 0;
        {
            uint64_t rng = 0x9E3779B97F4A7C15ULL;
            auto next01 = [&]() { rng = rng * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(rng >> 11) * (1.0 / 9007199254740992.0); };
            int best_steps = -1;
            std::vector<double> key(n);
            std::vector<int> done(n), trial;
            for (int tr = 0; tr < 300; ++tr) {
                const double w = tr == 0 ? 0.0 : 2.0 * next01();
                for (int id = 0; id < n; ++id) key[id] = -(double)height[id] + (tr == 0 ? 0.0 : w * next01());
                std::fill(done.begin(), done.end(), -1);
                trial.clear();
                int left = n, st = 0;
                while (left > 0) {
                    std::vector<std::pair<double, int>> ready;
                    for (int id = 0; id < n; ++id) {
                        if (done[id] >= 0) continue;
                        bool r = true;
                        for (int p : dep[id]) if (done[p] < 0 || done[p] >= st) { r = false; break; }
                        if (r) ready.push_back({key[id], id});
                    }
                    std::sort(ready.begin(), ready.end());
                    for (int sl = 0; sl < 4; ++sl) {
                        if (sl < (int)ready.size()) { done[ready[sl].second] = st; --left; trial.push_back(ready[sl].second); }
                        else trial.push_back(-1);
                    }
                    ++st;
                    if (best_steps >= 0 && st >= best_steps) break;       // cannot beat the best one
                }
                if (left == 0 && (best_steps < 0 || st < best_steps)) { best_steps = st; sched = trial; }
                if (best_steps == (n + 3) / 4) break;                     // four nodes in every step: optimal
            }
            step = best_steps;
        }
        n_steps = step;
        prog.clear();
        prog.push_back(n_steps); prog.push_back(MC); prog.push_back(RL); prog.push_back(0);
        prog.insert(prog.end(), leaf_begin.begin(), leaf_begin.end());
        prog.insert(prog.end(), leaf_sp.begin(), leaf_sp.end());
        while (prog.size() & 3) prog.push_back(0);
        prog[3] = (int)prog.size();
        auto qrow = [&](int t, int i) { return L.q0 + (t * I + i) * 32; };
        for (int s = 0; s < n_steps; ++s)
            for (int sl = 0; sl < 4; ++sl) {
                std::vector<int> r(RL, 0);
                const int id = sched[s * 4 + sl];
                if (id >= 0) {
                    const int t = id / I, i = id % I, p = par_internal[i];
                    const bool last = t + 1 == W;
                    r[0] = qrow(t, i);
                    r[1] = p >= 0 ? qrow(t, p) : L.onehot_row;
                    r[2] = qrow(last ? 0 : t + 1, i);          // the last position: the message for position 0 needs q(0, i)
                    r[3] = (last || p < 0) ? r[1] : qrow(t + 1, p);
                    r[4] = L.ab0 + i * 64;
                    r[5] = L.msg0 + i * 32;
                    r[6] = (((last ? 0 : t + 1) * I) + i) * N1N_TAB_STRIDE;
                    r[7] = 1 | (last ? 2 : 0) | (W == 1 ? 4 : 0);
                    for (int j = 0; j < MC; ++j) {
                        const bool has = ichild_begin[i] + j < ichild_begin[i + 1];
                        r[8 + j] = has ? L.msg0 + ichild[ichild_begin[i] + j] * 32 : L.zero_row;
                    }
                }
                prog.insert(prog.end(), r.begin(), r.end());
            }
    }

    void set_pi(int t, const double* v, int n) {
        if (t < 0 || t >= W || n != (int)pi[t].size()) throw std::invalid_argument("set_pi: bad position or size");
        pi[t].assign(v, v + n);
        if (order == 0) build_tables(t);
        version++;
    }
    void set_branch(int t, int node, double len) {
        if (t < 0 || t >= W || node <= 0 || node >= N) throw std::invalid_argument("set_branch: bad position or node");
        branch[(size_t)t * N + node] = len;
        if (order == 0) build_tables(t);
        version++;
    }
};

// util/MatrixLinearisation.java:15-19 through jstacs Normalisation.logSumNormalisation (:352-375):
// per block of 4: shift by the max, exp, sum, divide (skipped when the sum is exactly 1).
inline void lambda_to_pi(const double* lambda, double* pi, int n) {
    for (int i = 0; i < n; i += 4) {
        double off = lambda[i];
        for (int k = 1; k < 4; ++k) off = lambda[i + k] > off ? lambda[i + k] : off;
        double sum = 0;
        for (int k = 0; k < 4; ++k) { pi[i + k] = std::exp(lambda[i + k] - off); sum += pi[i + k]; }
        if (sum != 1.0) for (int k = 0; k < 4; ++k) pi[i + k] /= sum;
    }
}

}  // namespace phyfoo
