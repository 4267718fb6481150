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
                        sum += meanfield_pass(tp, lt, st, cw, update, nullptr);
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
