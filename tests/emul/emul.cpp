// Host emulation of the kernels' per-thread arithmetic (phyfoo_b200/csrc/tree_pass.cuh) so that the
// numerics can be checked against the oracle on a machine without a GPU.  TEST INFRASTRUCTURE:
// built and used by tests/test_emul_host.py only; the shipped library never runs this code path.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../phyfoo_b200/csrc/model_host.hpp"
#include "../../phyfoo_b200/csrc/tree_pass.cuh"

using namespace phyfoo;

static TreeProg prog_of(const ModelHost& m) {
    TreeProg tp;
    tp.n_internal = m.I; tp.n_leaves = m.S; tp.n_nodes = m.N;
    tp.par_internal = m.par_internal.data(); tp.node_of = m.node_of.data(); tp.leaf_begin = m.leaf_begin.data();
    tp.leaf_node = m.leaf_node.data(); tp.leaf_species = m.leaf_sp.data();
    return tp;
}

static ColWords words_of(const uint8_t* col, int S, bool rc) {
    ColWords cw; cw.b = 0; cw.g = 0;
    for (int s = 0; s < S; ++s) {
        if (col[s] >= 4) cw.g |= 1u << s;
        else cw.b |= (uint64_t)col[s] << (2 * s);
    }
    return rc ? complement(cw) : cw;
}

extern "C" int emul_score_windows(const phyfoo_model_desc* d, const uint8_t* codes, int L, int strand,
                                  double* out, int* sweeps) {
    try {
        ModelHost m; m.init(*d);
        const int S = m.S, W = m.W;
        TreeProg tp = prog_of(m);
        std::vector<double> state((size_t)W * m.I * 8, 1e300);   // poisoned: pass 0 must not depend on it
        for (int j = 0; j + W <= L; ++j) {
            auto pass = [&](bool update) {
                double sum = 0;
                for (int t = 0; t < W; ++t) {
                    State st{state.data() + (size_t)t * m.I * 8, 1};
                    int col = strand ? j + W - 1 - t : j + t;
                    ColWords cw = words_of(codes + (size_t)col * S, S, strand != 0);
                    if (m.mode == PHYFOO_MODE_EXACT) {
                        Tab pt{&m.probtab[(size_t)t * m.E], 1};
                        sum += -std::log(pruning_likelihood(tp, pt, st, cw));
                    } else {
                        Tab lt{&m.logtab[(size_t)t * m.E], 1};
                        sum += meanfield_pass(tp, lt, st, cw, update);
                    }
                }
                return sum;
            };
            if (m.mode == PHYFOO_MODE_EXACT) { out[j] = -pass(false); if (sweeps) sweeps[j] = 0; continue; }
            double old = pass(false), cur = old;
            int k = 0; bool conv = false;
            while ((k < m.max_sweeps && !conv) || k < m.min_sweeps) {
                cur = pass(true);
                if (std::fabs(old - cur) < m.tol) conv = true;
                old = cur; ++k;
            }
            out[j] = -cur;
            if (sweeps) sweeps[j] = k;
        }
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// Fixed-q objective the way the M-step kernels compute it: mean field to convergence per window (store_q), then
// values[k] = sum_t G_t - ENT for the parameter sets k = 0 (lambda) and 1..4 (lambda + eps e_i) of position `pos`.
extern "C" int emul_fe_values(const phyfoo_model_desc* d, const uint8_t* cols /*[n][W][S]*/, int n, const double* weights,
                              int pos, const double* lambda, double eps, double* values /*[5]*/) {
    try {
        ModelHost m; m.init(*d);
        const int S = m.S, W = m.W, I = m.I, E = m.E;
        TreeProg tp = prog_of(m);
        const size_t n_trees = (size_t)n * W;
        std::vector<double> qint((size_t)4 * I * n_trees), qleaf((size_t)4 * S * n_trees), ent(n_trees);
        std::vector<double> state((size_t)W * I * 8, 1e300);
        for (int u = 0; u < n; ++u) {
            auto words = [&](int t) { return words_of(cols + ((size_t)u * W + t) * S, S, false); };
            auto pass = [&](bool update) {
                double sum = 0;
                for (int t = 0; t < W; ++t) {
                    State st{state.data() + (size_t)t * I * 8, 1};
                    Tab lt{&m.logtab[(size_t)t * E], 1};
                    sum += meanfield_pass(tp, lt, st, words(t), update);
                }
                return sum;
            };
            double old = pass(false);
            int k = 0; bool conv = false;
            while ((k < m.max_sweeps && !conv) || k < m.min_sweeps) {
                double cur = pass(true);
                if (std::fabs(old - cur) < m.tol) conv = true;
                old = cur; ++k;
            }
            for (int t = 0; t < W; ++t) {
                State st{state.data() + (size_t)t * I * 8, 1};
                Tab lt{&m.logtab[(size_t)t * E], 1};
                const size_t tree = (size_t)t * n + u;
                ent[tree] = store_q(tp, lt, st, words(t), QSink{qint.data() + tree, qleaf.data() + tree, n_trees});
            }
        }
        double ENT = 0;
        for (int u = 0; u < n; ++u) { double e = 0; for (int t = 0; t < W; ++t) e += ent[(size_t)t * n + u]; ENT += weights[u] * e; }
        std::vector<double> G(W, 0.0);
        for (int t = 0; t < W; ++t)
            for (int u = 0; u < n; ++u) {
                const size_t tree = (size_t)t * n + u;
                double e1[1];
                expected_log_cpf_multi<1>(tp, &m.logtab[(size_t)t * E], E, qint.data() + tree, qleaf.data() + tree, n_trees,
                                          words_of(cols + ((size_t)u * W + t) * S, S, false), e1);
                G[t] += weights[u] * e1[0];
                Tab lt{&m.logtab[(size_t)t * E], 1};
                double single = expected_log_cpf(tp, lt, qint.data() + tree, qleaf.data() + tree, n_trees, words_of(cols + ((size_t)u * W + t) * S, S, false));
                if (single != e1[0]) throw std::runtime_error("expected_log_cpf_multi<1> differs from expected_log_cpf");
            }
        std::vector<double> tabs((size_t)5 * E), x(lambda, lambda + 4);
        double pi[4];
        lambda_to_pi(x.data(), pi, 4);
        m.build_tables_into(pos, pi, &tabs[0], nullptr);
        for (int i = 0; i < 4; ++i) {
            double keep = x[i]; x[i] += eps;
            lambda_to_pi(x.data(), pi, 4);
            m.build_tables_into(pos, pi, &tabs[(size_t)(1 + i) * E], nullptr);
            x[i] = keep;
        }
        double gk[5] = {0, 0, 0, 0, 0};
        for (int u = 0; u < n; ++u) {
            const size_t tree = (size_t)pos * n + u;
            double e5[5];
            expected_log_cpf_multi<5>(tp, tabs.data(), E, qint.data() + tree, qleaf.data() + tree, n_trees,
                                      words_of(cols + ((size_t)u * W + pos) * S, S, false), e5);
            for (int k = 0; k < 5; ++k) gk[k] += weights[u] * e5[k];
        }
        double rest = 0;
        for (int t = 0; t < W; ++t) if (t != pos) rest += G[t];
        for (int k = 0; k < 5; ++k) values[k] = (rest + gk[k]) - ENT;
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// phy_exp / phy_log (csrc/fp64_math.cuh) on the host, for the accuracy test
#include "../../phyfoo_b200/csrc/fp64_math.cuh"
extern "C" void emul_exp(const double* x, int n, double* out) { for (int i = 0; i < n; ++i) out[i] = phy_exp(x[i]); }
extern "C" void emul_log(const double* x, int n, double* out) { for (int i = 0; i < n; ++i) out[i] = phy_log(x[i]); }

// For every (column, position) tree: the first pass k >= 1 whose state (q and child sums) equals the state after pass k-1
// bit for bit -- from there on the trajectory is constant -- or -1 when that does not happen within max_passes.
extern "C" int emul_fixed_point(const phyfoo_model_desc* d, const uint8_t* cols /*[n][S]*/, int n, int max_passes, int* out /*[n][W]*/) {
    try {
        ModelHost m; m.init(*d);
        TreeProg tp = prog_of(m);
        std::vector<double> state((size_t)m.I * 8), prev, prev2;
        for (int c = 0; c < n; ++c)
            for (int t = 0; t < m.W; ++t) {
                State st{state.data(), 1};
                ColWords cw = words_of(cols + (size_t)c * m.S, m.S, false);
                Tab lt{&m.logtab[(size_t)t * m.E], 1};
                std::fill(state.begin(), state.end(), 0.0);
                int found = -1;
                for (int k = 0; k < (max_passes < 0 ? -max_passes : max_passes); ++k) {
                    prev2 = prev; prev = state;
                    meanfield_pass(tp, lt, st, cw, k > 0);
                    if (k > 0 && std::memcmp(prev.data(), state.data(), state.size() * sizeof(double)) == 0) { found = k; break; }
                    if (max_passes < 0 && k > 1 && std::memcmp(prev2.data(), state.data(), state.size() * sizeof(double)) == 0) { found = k; break; }
                }
                out[(size_t)c * m.W + t] = found;
            }
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}
