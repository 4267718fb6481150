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

// ---------------------------------------------------------------------------------------------------------------------------
// Host model of the wavefront trajectory kernel (csrc/memo.cuh::memo_traj_wave_kernel): node i of sweep k at step
// 2 k + depth(i), children's messages in slots of their own added in child order, per-node free-energy terms summed in node
// order when the sweep's deepest node is done, early stop at a bitwise fixed point or two-cycle of q with the periodic
// fill.  Not the CUDA code, but the same algorithm step for step, so the schedule and the fill logic can be checked against
// the sequential pass (meanfield_pass) without a GPU: traj[c][t][k] for k < K1 and the pass at which each tree stopped.
// ---------------------------------------------------------------------------------------------------------------------------
namespace {
struct WaveTree {
    int I; std::vector<int> depth, par; std::vector<std::vector<int>> kids;
    std::vector<double> Q, Q2, BASE, MSG;       // [I][4]
};

// one node update with the canonical arithmetic of tree_pass.cuh; returns the node's free-energy term(s)
void wave_node(const TreeProg& tp, const Tab& tb, const ColWords& cw, WaveTree& w, int i, bool update, double& ft, double& fgap,
               bool& any_gap, int& flags) {
    double lpi[4], own[4], s[4], e[4], q[4], lq[4];
    for (int a = 0; a < 4; ++a) lpi[a] = tb(a);
    const int e0 = i ? tab_of(tp.node_of[i]) : 0;
    for (int a = 0; a < 4; ++a) {
        if (i == 0) own[a] = lpi[a];
        else {
            double acc = 0.0;
            for (int l = 0; l < 4; ++l) acc = fma(w.Q[w.par[i] * 4 + l], tb(e0 + l * 4 + a), acc);
            own[a] = acc;
        }
        double cs = w.BASE[i * 4 + a];
        for (int c : w.kids[i]) cs += w.MSG[c * 4 + a];
        s[a] = own[a] + cs;
        e[a] = phy_exp(update ? s[a] : 0.0);
    }
    const double z = ((e[0] + e[1]) + e[2]) + e[3];
    const double rz = phy_rcp(z), lz = phy_log(z);
    flags = 3;
    for (int a = 0; a < 4; ++a) {
        q[a] = phy_mul(e[a], rz);
        lq[a] = (update ? s[a] : 0.0) - lz;
        if (std::memcmp(&q[a], &w.Q[i * 4 + a], 8) != 0) flags &= ~1;
        if (std::memcmp(&q[a], &w.Q2[i * 4 + a], 8) != 0) flags &= ~2;
    }
    double c[4] = {0, 0, 0, 0};
    any_gap = false;
    for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
        const int x = cw.obs(tp.leaf_species[k]);
        if (x < 4) { for (int a = 0; a < 4; ++a) c[a] += tb(tab_of(tp.leaf_node[k]) + a * 4 + x); }
        else any_gap = true;
    }
    double f[4];
    for (int a = 0; a < 4; ++a) f[a] = phy_mul(q[a], (lq[a] - own[a]) - c[a]);
    ft = ((f[0] + f[1]) + f[2]) + f[3];
    double tsum = 0.0; fgap = 0.0;
    if (any_gap) {
        const double qs = ((q[0] + q[1]) + q[2]) + q[3];
        for (int k = tp.leaf_begin[i]; k < tp.leaf_begin[i + 1]; ++k) {
            if (cw.obs(tp.leaf_species[k]) < 4) continue;
            double xl[4], el[4], fl[4];
            for (int a = 0; a < 4; ++a) { xl[a] = phy_mul(qs, lpi[a]); el[a] = phy_exp(update ? xl[a] : 0.0); }
            const double zl = ((el[0] + el[1]) + el[2]) + el[3];
            const double rzl = phy_rcp(zl), lzl = phy_log(zl);
            double t = 0.0;
            for (int a = 0; a < 4; ++a) {
                const double ql = phy_mul(el[a], rzl);
                t = fma(ql, lpi[a], t);
                fl[a] = phy_mul(ql, ((update ? xl[a] : 0.0) - lzl) - xl[a]);
            }
            fgap += ((fl[0] + fl[1]) + fl[2]) + fl[3];
            tsum += t;
        }
    }
    double m[4] = {0, 0, 0, 0};
    if (i != 0)
        for (int l = 0; l < 4; ++l) { double acc = 0.0; for (int h = 0; h < 4; ++h) acc = fma(q[h], tb(e0 + l * 4 + h), acc); m[l] = acc; }
    for (int a = 0; a < 4; ++a) {
        w.Q2[i * 4 + a] = w.Q[i * 4 + a];
        w.Q[i * 4 + a] = q[a];
        w.MSG[i * 4 + a] = m[a];
        w.BASE[i * 4 + a] = c[a] + tsum;
    }
}
}  // namespace

extern "C" int emul_wave_traj(const phyfoo_model_desc* d, const uint8_t* cols /*[n][S]*/, int n, int K1, int early,
                              double* traj /*[n][W][K1]*/, int* stopped /*[n][W]: last computed sweep*/) {
    try {
        ModelHost m; m.init(*d);
        TreeProg tp = prog_of(m);
        WaveTree w; w.I = m.I; w.depth.assign(m.I, 0); w.par = m.par_internal; w.kids.assign(m.I, {});
        int dmax = 0;
        for (int i = 1; i < m.I; ++i) { w.depth[i] = w.depth[w.par[i]] + 1; dmax = std::max(dmax, w.depth[i]); w.kids[w.par[i]].push_back(i); }
        for (int c = 0; c < n; ++c)
            for (int t = 0; t < m.W; ++t) {
                const ColWords cw = words_of(cols + (size_t)c * m.S, m.S, false);
                Tab lt{&m.logtab[(size_t)t * m.E], 1};
                double* out = traj + ((size_t)c * m.W + t) * K1;
                w.Q.assign(m.I * 4, 1e300); w.Q2 = w.BASE = w.MSG = w.Q;          // poisoned: nothing may depend on it
                std::vector<double> ring_ft(4 * m.I), ring_fg(4 * m.I);
                std::vector<int> ring_fl(4 * m.I), ring_gap(4 * m.I);
                double Fcur = 0, Fprev = 0; int fa_last = 0, kc_stop = -1;
                const int tau_end = 2 * (K1 - 1) + dmax;
                for (int tau = 0; tau <= tau_end && kc_stop < 0; ++tau) {
                    // all nodes of this step read the state the previous steps left (no node of a step reads another one's output)
                    WaveTree before = w;
                    for (int i = 0; i < m.I; ++i) {
                        if ((tau - w.depth[i]) & 1 || tau < w.depth[i]) continue;
                        const int k = (tau - w.depth[i]) >> 1;
                        if (k >= K1) continue;
                        WaveTree scratch = before;
                        double ft, fg; bool gap; int fl;
                        wave_node(tp, lt, cw, scratch, i, k > 0, ft, fg, gap, fl);
                        for (int a = 0; a < 4; ++a) {
                            w.Q[i * 4 + a] = scratch.Q[i * 4 + a]; w.Q2[i * 4 + a] = scratch.Q2[i * 4 + a];
                            w.MSG[i * 4 + a] = scratch.MSG[i * 4 + a]; w.BASE[i * 4 + a] = scratch.BASE[i * 4 + a];
                        }
                        const int r = (k & 3) * m.I + i;
                        ring_ft[r] = ft; ring_fg[r] = fg; ring_fl[r] = fl; ring_gap[r] = gap;
                    }
                    const int kc2 = tau - dmax;
                    if (kc2 >= 0 && !(kc2 & 1)) {
                        const int kc = kc2 >> 1;
                        double F = 0.0; int fa = kc >= 2 ? 3 : kc;
                        for (int j = 0; j < m.I; ++j) {
                            const int r = (kc & 3) * m.I + j;
                            F += ring_ft[r];
                            if (ring_gap[r]) F += ring_fg[r];
                            fa &= ring_fl[r];
                        }
                        out[kc] = F; Fprev = Fcur; Fcur = F; fa_last = fa;
                        if (early && fa != 0) kc_stop = kc;
                    }
                }
                stopped[(size_t)c * m.W + t] = kc_stop < 0 ? K1 - 1 : kc_stop;
                if (kc_stop >= 0)
                    for (int j = kc_stop + 1; j < K1; ++j) out[j] = ((j - kc_stop) & 1) ? ((fa_last & 1) ? Fcur : Fprev) : Fcur;
            }
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// the same trajectories from the sequential pass (what the direct kernel runs per tree)
extern "C" int emul_seq_traj(const phyfoo_model_desc* d, const uint8_t* cols, int n, int K1, double* traj) {
    try {
        ModelHost m; m.init(*d);
        TreeProg tp = prog_of(m);
        std::vector<double> state((size_t)m.I * 8);
        for (int c = 0; c < n; ++c)
            for (int t = 0; t < m.W; ++t) {
                State st{state.data(), 1};
                const ColWords cw = words_of(cols + (size_t)c * m.S, m.S, false);
                Tab lt{&m.logtab[(size_t)t * m.E], 1};
                std::fill(state.begin(), state.end(), 1e300);
                for (int k = 0; k < K1; ++k) traj[((size_t)c * m.W + t) * K1 + k] = meanfield_pass(tp, lt, st, cw, k > 0);
            }
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// the list schedule of the order-1 wavefront kernel (model_host.hpp::wave_program): sched[n_steps][4] entries (t << 16 | i, -1 = idle)
extern "C" int emul_wave_schedule(const phyfoo_model_desc* d, int* sched_out, int max_entries, int* n_steps, int* n_internal,
                                  int* par_internal_out /*[I]*/) {
    try {
        ModelHost m; m.init(*d);
        std::vector<int> prog; int MC, ML, ns;
        m.wave_program(prog, MC, ML, ns);
        const int off = 2 * m.I + 2 * m.S + m.I * MC + m.I * ML;
        if (4 * ns > max_entries) return -2;
        for (int k = 0; k < 4 * ns; ++k) sched_out[k] = prog[off + k];
        for (int i = 0; i < m.I; ++i) par_internal_out[i] = m.par_internal[i];
        *n_steps = ns; *n_internal = m.I;
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Host run of the node-per-thread order-1 kernel's arithmetic (csrc/net1n.cuh::n1n_update, the very code the kernel inlines)
// under the kernel's program (model_host.hpp::node_program): one window at a time in window slot `wslot` of a warp region,
// the four node slots of a step run one after the other -- in both orders, which must agree bit for bit because no node of
// a step may read what another node of the same step writes.  Windows must be free of gaps (the kernel sends the others to
// the net1w.cuh kernel).
// ---------------------------------------------------------------------------------------------------------------------------
#include "../../phyfoo_b200/csrc/net1n.cuh"

namespace {
template <int MC>
int run_o1n(const ModelHost& m, const std::vector<int>& prog, const uint8_t* codes, int L, int strand, int wslot, double* out, int* sweeps) {
    const int W = m.W, I = m.I, S = m.S;
    const int n_steps = prog[0], RL = prog[2];
    const int* leaf_begin = prog.data() + 4;
    const int* leaf_species = leaf_begin + I + 1;
    const int* rec = prog.data() + prog[3];
    const ModelHost::N1nLayout Lt = m.n1n_layout();
    std::vector<double> tab, leaf_tab;
    m.tables1n(tab);
    m.leaf_tables1n(leaf_tab);
    std::vector<double> region(Lt.region, 0.0), gc(Lt.region, 0.0);
    double exp16[16];                                            // 2^(j/16): the table of n1n_exp4_t16
    for (int j = 0; j < 16; ++j) exp16[j] = h_phy_exp_tab[4 * j];
    for (int w = 0; w < 8; ++w) region[Lt.onehot_row + 2 * w] = 1.0;
    for (int j = 0; j + W <= L; ++j) {
        // set-up: observed-leaf sums, uniform start, context sums of position 0
        for (int t = 0; t < W; ++t) {
            const int col = strand ? j + W - 1 - t : j + t;
            ColWords cw = words_of(codes + (size_t)col * S, S, strand != 0);
            ColWords cwp; cwp.b = 0; cwp.g = 0;
            if (t > 0) cwp = words_of(codes + (size_t)(strand ? col + 1 : col - 1) * S, S, strand != 0);
            if (cw.g) return -3;
            for (int i = 0; i < I; ++i)
                for (int a = 0; a < 4; ++a) {
                    double c = 0.0;
                    for (int kk = leaf_begin[i]; kk < leaf_begin[i + 1]; ++kk) {
                        const int sp = leaf_species[kk];
                        const int x = cw.obs(sp), y = t > 0 ? cwp.obs(sp) : 0;
                        c += leaf_tab[(size_t)(t * S + kk) * 32 + (a == x ? 16 : 0) + y * 4 + x];
                    }
                    const int o = Lt.q0 + (t * I + i) * 32 + (a >> 1) * 16 + 2 * wslot + (a & 1);
                    gc[o] = c; region[o] = 0.25;
                }
        }
        for (int i = 0; i < I; ++i)
            n1n_setup_node(region.data() + 2 * wslot, tab.data(), Lt.ab0 + i * 64, Lt.msg0 + i * 32, i * ModelHost::N1N_TAB_STRIDE);
        // F_0 in closed form (the kernel's set-up does the same), the first pass is the first sweep
        double csum = 0.0;
        for (int n = 0; n < W * I; ++n) {
            const double* g = gc.data() + Lt.q0 + n * 32 + 2 * wslot;
            csum += (g[0] + g[1]) + (g[16] + g[17]);
        }
        bool upd = true, conv = false; int k = 0; double old = fma(-0.25, csum, m.n1n_f0_const());
        {   // the closed form against an actual pass at the uniform start (on a copy of the state)
            std::vector<double> tmp = region;
            N1nAcc a0[4];
            for (auto& a : a0) { a.Fq = 0.0; a.M = 1.0; a.E = 0; a.Fslow = 0.0; }
            for (int step = 0; step < n_steps; ++step)
                for (int slot = 0; slot < 4; ++slot) {
                    const int* r = rec + (size_t)(step * 4 + slot) * RL;
                    double* reg = tmp.data() + 2 * wslot;
                    const double* g = gc.data() + 2 * wslot + r[0];
                    const double cobs[4] = {g[0], g[1], g[16], g[17]};
                    double qx[4], qpn[4];
                    n1n_ld4(reg + r[2], qx);
                    n1n_ld4(reg + r[3], qpn);
                    n1n_update<MC>(reg, reg, tab.data(), exp16, r, cobs, qx, qpn, false, a0[slot]);
                }
            const double F0 = (n1n_free_energy(a0[0]) + n1n_free_energy(a0[1])) + (n1n_free_energy(a0[2]) + n1n_free_energy(a0[3]));
            if (std::fabs(F0 - old) > 1e-12 * std::fabs(F0)) return -6;
            for (int o = 0; o < Lt.region; ++o) if (std::fabs(tmp[o] - region[o]) > 1e-13 * std::fabs(region[o])) return -7;   // ... and leaves the state alone
        }
        for (;;) {
            N1nAcc acc[4];
            for (auto& a : acc) { a.Fq = 0.0; a.M = 1.0; a.E = 0; a.Fslow = 0.0; }
            for (int step = 0; step < n_steps; ++step) {
                std::vector<double> alt = region;
                N1nAcc acc_alt[4] = {acc[0], acc[1], acc[2], acc[3]};
                for (int order = 0; order < 2; ++order)
                    for (int q = 0; q < 4; ++q) {
                        const int slot = order ? 3 - q : q;
                        const int* r = rec + (size_t)(step * 4 + slot) * RL;
                        double* reg = (order ? alt.data() : region.data()) + 2 * wslot;
                        const double* g = gc.data() + 2 * wslot + r[0];
                        const double cobs[4] = {g[0], g[1], g[16], g[17]};
                        double qx[4], qpn[4];
                        n1n_ld4(reg + r[2], qx);
                        n1n_ld4(reg + r[3], qpn);
                        n1n_update<MC>(reg, reg, tab.data(), exp16, r, cobs, qx, qpn, upd, order ? acc_alt[slot] : acc[slot]);
                    }
                for (int o = 0; o < Lt.region; ++o)
                    if (std::memcmp(&alt[o], &region[o], 8) != 0) return -4;
                if ((step & 63) == 63) for (auto& a : acc) n1n_fold(a);
            }
            const double f0 = n1n_free_energy(acc[0]), f1 = n1n_free_energy(acc[1]), f2 = n1n_free_energy(acc[2]), f3 = n1n_free_energy(acc[3]);
            const double F = (f0 + f1) + (f2 + f3);            // the kernel's butterfly over lanes w, w+8, w+16, w+24
            if (std::fabs(old - F) < m.tol) conv = true;
            old = F; ++k;
            if (!((k < m.max_sweeps && !conv) || k < m.min_sweeps)) { out[j] = -F; if (sweeps) sweeps[j] = k; break; }
        }
    }
    return 0;
}
}  // namespace

extern "C" int emul_score_o1n(const phyfoo_model_desc* d, const uint8_t* codes, int L, int strand, int wslot, double* out, int* sweeps,
                              int* n_steps_out) {
    try {
        ModelHost m; m.init(*d);
        if (m.order != 1) return -2;
        std::vector<int> prog; int MC, ns;
        m.node_program(prog, MC, ns);
        if (n_steps_out) *n_steps_out = ns;
        if (MC == 1) return run_o1n<1>(m, prog, codes, L, strand, wslot, out, sweeps);
        if (MC == 2) return run_o1n<2>(m, prog, codes, L, strand, wslot, out, sweeps);
        if (MC == 3) return run_o1n<3>(m, prog, codes, L, strand, wslot, out, sweeps);
        return -5;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Host run of the F81-structured order-0 pass (csrc/tree_pass_f81.cuh::meanfield_pass_f81) with the window-level stop rule of
// score_o0f_kernel: per pass every tree contributes (terms, log z accumulator) -> n1n_free_energy, the window's free energy
// is their tree-major sum.
// ---------------------------------------------------------------------------------------------------------------------------
#include "../../phyfoo_b200/csrc/tree_pass_f81.cuh"

extern "C" int emul_score_o0f(const phyfoo_model_desc* d, const uint8_t* codes, int L, int strand, double* out, int* sweeps) {
    try {
        ModelHost m; m.init(*d);
        if (m.order != 0) return -2;
        std::vector<double> tabf;
        if (!m.tables0f(tabf)) return -3;
        const int S = m.S, W = m.W, ef = m.Ef();
        TreeProg tp = prog_of(m);
        std::vector<int> slot(m.I);
        const int n_slots = f81_slots(m.I, m.ichild_begin.data(), slot.data());
        const int sd = statef_doubles(n_slots, m.I);
        std::vector<double> state((size_t)W * sd, std::nan(""));          // poisoned: the pass at the uniform start must not use it
        for (int j = 0; j + W <= L; ++j) {
            auto pass = [&](bool update) {
                double sum = 0;
                for (int t = 0; t < W; ++t) {
                    StateF st{state.data() + (size_t)t * sd, 1, n_slots};
                    const int col = strand ? j + W - 1 - t : j + t;
                    const ColWords cw = words_of(codes + (size_t)col * S, S, strand != 0);
                    Tab lt{&tabf[(size_t)t * ef], 1};
                    N1nAcc acc; acc.Fq = 0.0; acc.M = 1.0; acc.E = 0; acc.Fslow = 0.0;
                    meanfield_pass_f81(tp, slot.data(), lt, st, cw, update, h_phy_exp_tab, acc);
                    sum += n1n_free_energy(acc);
                }
                return sum;
            };
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

// ---- alignment-based models of order >= 1: the kernels' term (ab.cuh::ab_term) on the host, through the shared expansion ----
#include "../../phyfoo_b200/csrc/ab_expand.hpp"
// row[L]: codes of one species row; variant 0 = motif (context inside [l - min(order, pos_mot), l]), 1 = background (pos_mot = l)
extern "C" double emul_ab_term(const double* cond /* table of this position */, int order, const uint8_t* row, int l, int pos_mot, int variant) {
    const int kmax = pos_mot < order ? pos_mot : order;
    int sym[AB_MAX_ORDER + 1];
    bool any_gap = false, any_obs = false;
    int h = 0, pw = 1;
    for (int k = 0; k <= kmax; ++k, pw *= 4) {
        sym[k] = row[l - k];
        if (sym[k] == 4) any_gap = true; else { any_obs = true; h += sym[k] * pw; }
    }
    if (!any_obs) return std::log(0.25);
    if (!any_gap) return std::log(cond[h] / 1);
    int idx[AB_MAX_EXPAND], pre, n;
    ab_expand(sym, kmax, variant, idx, pre, n);
    const int size = n - pre;
    double sum = 0.0;
    if (variant) { for (int i = pre; i < n; ++i) sum += cond[idx[i]] / size; }
    else { for (int i = pre; i < n; ++i) sum += cond[idx[i]]; sum /= size; }
    return std::log(sum);
}
// the training loops' expansion: hashes and their common divisor
extern "C" int emul_ab_expand(const uint8_t* row, int l, int kmax, int variant, int* idx_out) {
    int sym[AB_MAX_ORDER + 1];
    for (int k = 0; k <= kmax; ++k) sym[k] = row[l - k];
    int idx[AB_MAX_EXPAND], pre, n;
    ab_expand(sym, kmax, variant, idx, pre, n);
    for (int i = pre; i < n; ++i) idx_out[i - pre] = idx[i];
    return n - pre;
}

// ---- generic net (netg.cuh): the per-item mean field on explicit (tree -> column) items ---------------------------------
#include "../../phyfoo_b200/csrc/netg.cuh"
extern "C" int emul_netg_items(const phyfoo_model_desc* d, const uint8_t* codes /*[L][S]*/, int n_items, const int* cols /*[n_items][T]*/,
                               const int* len, double* out_first, double* out_last, int* sweeps) {
    try {
        ModelHost m; m.init(*d);
        std::vector<int> rec, lists; std::vector<double> tab;
        m.netg_build(rec, lists, tab);
        NetgView v{m.W * m.N, m.W, m.N, rec.data(), lists.data(), tab.data()};
        std::vector<double> q((size_t)v.n_nodes * 4, 1e300);
        NetgState st{q.data(), 1};
        for (int u = 0; u < n_items; ++u) {
            NetgObs ob;
            for (int t = 0; t < m.W; ++t) ob.cw[t] = words_of(codes + (size_t)cols[(size_t)u * m.W + t] * m.S, m.S, false);
            sweeps[u] = netg_item(v, ob, st, len[u], m.max_sweeps, m.min_sweeps, m.tol, &out_first[u], &out_last[u]);
        }
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}

// expected counts of the generic net's table entries over a list of full windows (netg_counts), and the table array itself
extern "C" int emul_netg_tab(const phyfoo_model_desc* d, double* tab_out, int cap) {
    try {
        ModelHost m; m.init(*d);
        std::vector<int> rec, lists; std::vector<double> tab;
        m.netg_build(rec, lists, tab);
        if (tab_out) for (int i = 0; i < (int)tab.size() && i < cap; ++i) tab_out[i] = tab[i];
        return (int)tab.size();
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}
extern "C" int emul_netg_counts(const phyfoo_model_desc* d, const uint8_t* codes, int n_items, const int* cols, const double* weights,
                                double* counts_out, double* entropy_out) {
    try {
        ModelHost m; m.init(*d);
        std::vector<int> rec, lists; std::vector<double> tab;
        m.netg_build(rec, lists, tab);
        NetgView v{m.W * m.N, m.W, m.N, rec.data(), lists.data(), tab.data()};
        std::vector<double> q((size_t)v.n_nodes * 4, 1e300);
        NetgState st{q.data(), 1};
        for (size_t i = 0; i < tab.size(); ++i) counts_out[i] = 0.0;
        double ent = 0.0;
        for (int u = 0; u < n_items; ++u) {
            NetgObs ob;
            for (int t = 0; t < m.W; ++t) ob.cw[t] = words_of(codes + (size_t)cols[(size_t)u * m.W + t] * m.S, m.S, false);
            double f, l;
            netg_item(v, ob, st, m.W, m.max_sweeps, m.min_sweeps, m.tol, &f, &l);
            ent += netg_counts(v, ob, st, weights[u], [&](int e, double val) { counts_out[e] += val; });
        }
        *entropy_out = ent;
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emul: %s\n", e.what()); return -1; }
}
