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

extern "C" int emul_score_windows(const phyfoo_model_desc* d, const uint8_t* codes, int L, int strand,
                                  double* out, int* sweeps) {
    try {
        ModelHost m; m.init(*d);
        const int S = m.S, W = m.W, nb = (S + 15) / 16, ng = (S + 31) / 32;
        std::vector<uint32_t> bases((size_t)nb * L, 0), gaps((size_t)ng * L, 0);
        for (int c = 0; c < L; ++c)
            for (int s = 0; s < S; ++s) {
                uint8_t v = codes[(size_t)c * S + s];
                if (v >= 4) gaps[(size_t)(s / 32) * L + c] |= 1u << (s % 32);
                else bases[(size_t)(s / 16) * L + c] |= (uint32_t)v << (2 * (s % 16));
            }
        TreeProg tp = prog_of(m);
        std::vector<double> state((size_t)W * m.I * 8);
        for (int j = 0; j + W <= L; ++j) {
            std::vector<double> F(W);
            auto pass = [&](bool update) {
                double sum = 0;
                for (int t = 0; t < W; ++t) {
                    State st{state.data() + (size_t)t * m.I * 8, 1};
                    int col = strand ? j + W - 1 - t : j + t;
                    if (m.mode == PHYFOO_MODE_EXACT) {
                        pruning_init(tp, st);
                        sum += std::log(pruning_likelihood(tp, &m.probtab[(size_t)t * m.tbl_stride], st, &bases[col], &gaps[col], L, strand != 0));
                    } else
                        sum += meanfield_pass(tp, &m.logtab[(size_t)t * m.tbl_stride], st, &bases[col], &gaps[col], L, strand != 0, update, nullptr);
                }
                return sum;
            };
            if (m.mode == PHYFOO_MODE_EXACT) { out[j] = pass(false); if (sweeps) sweeps[j] = 0; continue; }
            for (int t = 0; t < W; ++t) { State st{state.data() + (size_t)t * m.I * 8, 1}; meanfield_init(tp, st); }
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
