// phyfoo_oracle.cpp -- CPU restatement of PhyFoo's site-scoring / optimiser hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under phyfoo_b200/ may import, link or execute this
// file; it is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  The product path is CUDA (phyfoo_b200/csrc) and fails loudly
// when the extension is missing.
//
// PARITY: no JVM exists in this image and the reference (single-threaded Java on a patched
// Jstacs 1.5) ships no tests, golden vectors or fixtures for this path, so there is no JVM run to
// compare with.  The arithmetic is pinned by the reference's OWN STATEMENTS instead: tests/refexec.py
// translates MeanFieldForBayesNet.init / optimizeByNormalisation / calcFreeEnergy, the evolution
// models' reinit and the Jstacs E-step (getNewWeights / modify / logSumNormalisation) from the
// reference's files at test time and executes them; this oracle reproduces their outputs to the
// last bit (tests/test_reference_statements.py, golden vectors in tests/golden/refexec_meanfield.npz).
// What those methods do not cover (selector, objective drivers, strand model, optimiser) is
// restated only.  Further pins (tests/test_oracle_*.py): closed forms (star tree under F81alpha),
// brute-force enumeration of the joint for exact mode, the invariant -F <= log P, and an
// independently written numpy restatement (oracle/np_restatement.py) agreeing to 1e-12.
// The only outputs of reference code in the image -- its bundled synthetic data, written by its own
// sampler -- pin the MODEL statistically (tests/test_generator_pins.py: column-pattern histograms
// against the exact likelihood under the described trees, chi-square over 1024 patterns), not the
// mean-field numbers bit for bit.
//
// Every function cites the reference lines it follows.  "X.java:N" is
// /root/reference/main/java/src/X.java; "jstacs!/p:N" is a source file inside
// /root/reference/libs/jstacs.jar.
//
// Build: make -C oracle   (g++ -O2 -pthread -shared -fPIC) -> oracle/libphyfoo_oracle.so

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <thread>

namespace orc {

static const int ALPHA = 4;  // io/Alphabet.java:5-7 (size = 4)

// ------------------------------------------------------------------------------------------
// Newick -> tree template.  io/NewickToBayesNet.java:38-208: the root exists before parsing;
// '(' creates a child of the current node and descends, ',' closes the current node and
// creates a sibling, ')' closes and ascends, ':' announces the distance of the node being
// closed, ';' ends.  Nodes are therefore numbered in DFS pre-order.  Leaves are the nodes
// without children, left to right (:204-206).  Whitespace <= '\n' separates tokens (:301).
// NHX brackets / quoting / escapes (:152-193) are not restated (unused by PhyFoo's data).
// ------------------------------------------------------------------------------------------
struct TreeTpl {
    int N = 0;
    std::vector<int> parent;
    std::vector<double> dist;
    std::vector<std::vector<int>> children;
    std::vector<std::string> label;
    std::vector<int> leaves;
};

static TreeTpl parse_newick(const std::string& s) {
    TreeTpl t;
    auto add = [&](int par) {
        t.parent.push_back(par);
        t.dist.push_back(0.0);  // Java field default, BayesNetNode.java:46
        t.children.emplace_back();
        t.label.emplace_back();
        if (par >= 0) t.children[par].push_back(t.N);
        return t.N++;
    };
    int n = add(-1);
    std::string label;
    double d = -1.0;
    bool distance = false, finished = false;
    auto close_node = [&](int node) {  // setLabel + setDistanceToParent, :123-127,:133-137
        if (!label.empty()) t.label[node] = label;
        if (d != -1.0) { t.dist[node] = d; d = -1.0; }
        label.clear();
    };
    size_t i = 0;
    while (i < s.size() && !finished) {
        unsigned char c = (unsigned char)s[i];
        if (c <= '\n') { ++i; continue; }
        if (c == '(') { n = add(n); ++i; }
        else if (c == ')') {
            close_node(n);
            if (t.parent[n] < 0) throw std::runtime_error("newick: unbalanced ')'");
            n = t.parent[n]; ++i;
        } else if (c == ',') {
            close_node(n);
            if (t.parent[n] < 0) throw std::runtime_error("newick: ',' at root level");
            n = add(t.parent[n]); ++i;
        } else if (c == ':') { distance = true; ++i; }
        else if (c == ';') { finished = true; ++i; }
        else {
            size_t j = i;
            while (j < s.size()) {
                unsigned char cj = (unsigned char)s[j];
                if (cj <= '\n' || cj == '(' || cj == ')' || cj == ',' || cj == ':' || cj == ';') break;
                ++j;
            }
            std::string w = s.substr(i, j - i);
            if (distance) { d = std::stod(w); distance = false; }  // :87-89
            else label += w;                                       // :91
            i = j;
        }
    }
    for (int k = 0; k < t.N; ++k)
        if (t.children[k].empty()) t.leaves.push_back(k);
    return t;
}

// ------------------------------------------------------------------------------------------
// Bayes net substrate (bayesNet/BayesNetNode.java:108-117, CPF.java:47-64,238-256,
// BayesNetHandler.java:47-55,230-256).
// ------------------------------------------------------------------------------------------
struct Node {
    std::vector<int> par, chi;  // global ids in the reference's insertion order
    bool phylo_leaf = false, phylo_root = false;
    int tree = 0, local = 0;
    double dist = 0.0;
    bool observed = false;
    int obs = 0;
    int size = 1;                        // 4^{#parents}, CPF.java:48
    std::vector<double> cpf, indep;      // [size][4] row-major
    std::vector<double> lcpf, lindep;    // log of the above (hoisted; identical bits to Math.log in place)
    int parent_index(int p) const {      // BayesNetNode.java:154-156
        for (size_t k = 0; k < par.size(); ++k) if (par[k] == p) return (int)k;
        return -1;
    }
    bool is_root() const { return par.empty(); }  // BayesNetNode.java:144-146
    // CPF.get(l,obs) / getCondProb(): independent table for an unobserved phylo leaf, CPF.java:146-152,218-224
    const double* table(bool allow_indep) const {
        return (phylo_leaf && !observed && allow_indep) ? indep.data() : cpf.data();
    }
    const double* ltable(bool allow_indep) const {
        return (phylo_leaf && !observed && allow_indep) ? lindep.data() : lcpf.data();
    }
};

enum Evol { F81ALPHA = 0, F81BETA = 1, HKY_AS_IMPLEMENTED = 2, HKY_RATIO = 3 };
enum Mode { MEANFIELD = 0, EXACT_ELIMINATION = 1, EXACT_PRUNING = 2 };

struct Model {
    int kind = 0;   // 0 = motif (PhyloBayesModel), 1 = background (PhyloBackground)
    int W = 0;      // number of trees in the net (motif length, or order+1 for the background)
    int order = 0;
    int N = 0, S = 0;
    TreeTpl tpl;
    std::vector<Node> nodes;
    std::vector<int> leaf_local;             // local node index of leaf s
    int evol = F81ALPHA;
    double hky_ratio = 0.0;                  // HKY_AS_IMPLEMENTED: never reaches the matrices (HKY.java:32 self-assigns);
                                             // HKY_RATIO (opt-in): the ratio the constructor was meant to store
    std::vector<std::vector<double>> pi;     // per tree [dim][4], dim = 4^{#parents of that tree's root}
    int mode = MEANFIELD;
    bool allow_indep = true;                 // CPF.ALLOW_RETURN_INDEPENDENT_TRANSITION, CPF.java:14
    bool log_in_loop = false;                // cost model of the Java (Math.log inside the loops); same bits
    // statistics
    long long sweeps_total = 0, calls_total = 0;
    // instrumented term counter (SURVEY.md section 8d counting rule): every prod*log CPF accumulation = 2 flop,
    // every q factor beyond the first = 1 flop, normalisation = 7 flop per hidden node, q log q = 2 flop;
    // exp/log calls are counted separately as transcendentals (logs of CPF entries are per-parameter constants)
    bool count_terms = false;
    long long c_upd_flops = 0, c_upd_transc = 0, c_fe_flops = 0, c_fe_transc = 0, c_sweeps = 0, c_fe_calls = 0;

    int gid(int t, int k) const { return t * N + k; }
};

// BayesNetHandler.connectVirtualTrees (:230-256): node l of t2 becomes a child of node l of t1.
static void connect_trees(Model& m, int t1, int t2) {
    for (int k = 0; k < m.N; ++k) {
        m.nodes[m.gid(t1, k)].chi.push_back(m.gid(t2, k));
        m.nodes[m.gid(t2, k)].par.push_back(m.gid(t1, k));
    }
}

static void init_tree_params(Model& m, int t);

// models/PhyloBayesModel.java:356-370 (motif; order 0 or 1) and models/PhyloBackground.java:381-398.
static Model* build_model(int kind, int W, int order, const std::string& newick, int evol, double hky_ratio) {
    Model* m = new Model();
    m->kind = kind; m->order = order; m->evol = evol; m->hky_ratio = hky_ratio;
    m->W = (kind == 0) ? W : order + 1;
    m->tpl = parse_newick(newick);
    m->N = m->tpl.N; m->S = (int)m->tpl.leaves.size();
    m->leaf_local = m->tpl.leaves;
    m->nodes.resize((size_t)m->W * m->N);
    for (int t = 0; t < m->W; ++t) {
        for (int k = 0; k < m->N; ++k) {
            Node& nd = m->nodes[m->gid(t, k)];
            nd.tree = t; nd.local = k; nd.dist = m->tpl.dist[k];
            nd.phylo_root = (k == 0);
            nd.phylo_leaf = m->tpl.children[k].empty();
            if (m->tpl.parent[k] >= 0) nd.par.push_back(m->gid(t, m->tpl.parent[k]));
            for (int c : m->tpl.children[k]) nd.chi.push_back(m->gid(t, c));
        }
        if (kind == 0) {
            if (t > 0 && order == 1) connect_trees(*m, t - 1, t);      // PhyloBayesModel.java:363-365
        } else {
            for (int j = 0; j < order; ++j)                             // PhyloBackground.java:389-393
                if (t - j > 0) connect_trees(*m, t - j - 1, t);
        }
    }
    m->pi.resize(m->W);
    for (int t = 0; t < m->W; ++t) {
        int dim = 1;
        for (size_t p = 0; p < m->nodes[m->gid(t, 0)].par.size(); ++p) dim *= ALPHA;  // PhyloBayesModel.java:338-339
        m->pi[t].assign((size_t)dim * 4, 0.25);
        for (int k = 0; k < m->N; ++k) {
            Node& nd = m->nodes[m->gid(t, k)];
            nd.size = 1;
            for (size_t p = 0; p < nd.par.size(); ++p) nd.size *= ALPHA;
        }
        init_tree_params(*m, t);
    }
    return m;
}

// evolution/FS81alpha.java:76-92, FS81beta.java:81-103, HKY.java:77-106; getUnweightedTransitions
// (FS81alpha.java:52-64).  Output [dim*4][4], row index = a1 + 4*ctx.
static void evol_tables(const Model& m, const std::vector<double>& pi, int dim, double len,
                        std::vector<double>& trans, std::vector<double>& indep) {
    trans.assign((size_t)dim * 16, 0.0);
    indep.assign((size_t)dim * 16, 0.0);
    for (int i = 0; i < dim; ++i) {
        const double* p = &pi[(size_t)i * 4];
        for (int a1 = 0; a1 < 4; ++a1)
            for (int a2 = 0; a2 < 4; ++a2) indep[(size_t)(a1 + i * 4) * 4 + a2] = p[a2];
        if (m.evol == F81ALPHA) {
            double alpha = len;
            for (int a1 = 0; a1 < 4; ++a1)
                for (int a2 = 0; a2 < 4; ++a2)
                    trans[(size_t)(a1 + i * 4) * 4 + a2] = (a1 == a2) ? (1 - alpha + p[a2] * alpha) : (p[a2] * alpha);
        } else if (m.evol == F81BETA) {
            const double eps = 1e-4;
            double beta = 1. / (1. - p[0] * p[0] - p[1] * p[1] - p[2] * p[2] - p[3] * p[3]);
            double ta = std::exp(-len * beta);
            if (ta < 0 + eps) ta = 0; else if (ta > 1 - eps) ta = 1;
            for (int a1 = 0; a1 < 4; ++a1)
                for (int a2 = 0; a2 < 4; ++a2)
                    trans[(size_t)(a1 + i * 4) * 4 + a2] = (a1 == a2) ? (ta + p[a2] * (1. - ta)) : (p[a2] * (1. - ta));
        } else {  // HKY as implemented: alpha_beta_ratio field stays 0 (HKY.java:32) => beta = 0;
                  // HKY_RATIO: the same statements (HKY.java:77-106) with the field holding the ratio
            double alpha = len, beta = alpha * (m.evol == HKY_RATIO ? m.hky_ratio : 0.0);
            double aA = alpha * p[0], aC = alpha * p[1], aG = alpha * p[2], aT = alpha * p[3];
            double bA = beta * p[0], bC = beta * p[1], bG = beta * p[2], bT = beta * p[3];
            double* r = &trans[(size_t)(i * 4) * 4];
            r[0] = 1 - (aG + bT + bC); r[1] = bC; r[2] = aG; r[3] = bT;
            r[4] = bA; r[5] = 1 - (aT + bA + bG); r[6] = bG; r[7] = aT;
            r[8] = aA; r[9] = bC; r[10] = 1 - (aA + bT + bC); r[11] = bT;
            r[12] = bA; r[13] = aC; r[14] = bG; r[15] = 1 - (aC + bA + bG);
        }
    }
}

// bayesNet/VirtualTree.java:82-96 (TEMPERATURE == 1 branch).
static void init_tree_params(Model& m, int t) {
    int dim = (int)m.pi[t].size() / 4;
    std::vector<double> trans, indep;
    for (int k = 0; k < m.N; ++k) {
        Node& nd = m.nodes[m.gid(t, k)];
        if (nd.phylo_root) {
            nd.cpf = m.pi[t];
            nd.indep = m.pi[t];
        } else {
            evol_tables(m, m.pi[t], dim, nd.dist, trans, indep);
            nd.cpf = trans;
            nd.indep = indep;
        }
        nd.lcpf.resize(nd.cpf.size());
        nd.lindep.resize(nd.indep.size());
        for (size_t i = 0; i < nd.cpf.size(); ++i) nd.lcpf[i] = std::log(nd.cpf[i]);
        for (size_t i = 0; i < nd.indep.size(); ++i) nd.lindep[i] = std::log(nd.indep[i]);
    }
}

// VirtualTree.setObservation (:183-198) + BayesNetNodeProperties.setObservation (:88-99).
static inline void set_observation(Model& m, int t, const uint8_t* col, bool revcomp) {
    for (int s = 0; s < m.S; ++s) {
        Node& nd = m.nodes[m.gid(t, m.leaf_local[s])];
        int c = col[s];
        if (revcomp && c < ALPHA) c = 3 - c;  // jstacs!/de/jstacs/data/alphabets/UnobservableDNAAlphabet.java:88-95
        if (c >= ALPHA) nd.observed = false;
        else { nd.observed = true; nd.obs = c; }
    }
}

// ------------------------------------------------------------------------------------------
// Mean field: algorithm/MeanFieldForBayesNet.java
// ------------------------------------------------------------------------------------------
struct MeanField {
    Model* m;
    int len;                 // seq.getLength(): number of trees the free energy in the stop rule spans
    std::vector<int> hmap;   // node -> 4*hidden index, or -1   (:52-63)
    std::vector<double> q;
    int sweeps = 0;

    MeanField(Model* mm, int seqlen) : m(mm), len(seqlen) {}

    void init() {  // :52-69
        int hidden = 0;
        hmap.assign(m->nodes.size(), -1);
        for (size_t i = 0; i < m->nodes.size(); ++i)
            if (!m->nodes[i].observed) hmap[i] = 4 * hidden++;
        q.assign((size_t)hidden * 4, 1. / ALPHA);
    }

    inline double lg(const Node& nd, int l, int a) const {
        if (m->log_in_loop) return std::log(nd.table(m->allow_indep)[l * 4 + a]);
        return nd.ltable(m->allow_indep)[l * 4 + a];
    }
    static inline int digit(int l, int p) { return (l >> (2 * p)) & 3; }  // CPF.java:238-256 lookup table

    double free_energy(int start, int end) const {  // :232-365
        double part1 = 0, part2 = 0;
        if (m->count_terms) {   // same loops, counting the structurally non-zero terms
            long long fl = 0, tr = 0;
            for (int t = start; t <= end; ++t)
                for (int i = 0; i < m->N; ++i) {
                    const Node& nd = m->nodes[m->gid(t, i)];
                    int nhid = 0, rows = 1;
                    for (size_t p = 0; p < nd.par.size(); ++p) { if (!m->nodes[nd.par[p]].observed) { ++nhid; rows *= 4; } }
                    if (!nd.observed) { fl += 4 * 2; tr += 4; }
                    if (nd.is_root() && !nd.observed) fl += 4 * 2;
                    else if (nd.observed) fl += (long long)rows * (2 + std::max(nhid - 1, 0));
                    else fl += 4LL * rows * (2 + nhid);
                }
            m->c_fe_flops += fl; m->c_fe_transc += tr; m->c_fe_calls++;
        }
        for (int t = start; t <= end; ++t)
            for (int i = 0; i < m->N; ++i) {
                int g = m->gid(t, i);
                double tmp = 0;
                if (!m->nodes[g].observed) {
                    int hn = hmap[g];
                    for (int h = 0; h < 4; ++h)
                        if (q[hn + h] > 0) tmp += q[hn + h] * std::log(q[hn + h]);
                }
                part1 += tmp;
            }
        for (int t = start; t <= end; ++t)
            for (int i = 0; i < m->N; ++i) {
                int g = m->gid(t, i);
                const Node& nd = m->nodes[g];
                const double* cp = nd.table(m->allow_indep);
                double tmp = 0;
                int hn = hmap[g];
                int np = (int)nd.par.size();
                if (nd.is_root() && !nd.observed) {          // :282-291
                    for (int h = 0; h < 4; ++h)
                        if (cp[h] > 0) tmp += q[hn + h] * lg(nd, 0, h);
                } else if (nd.observed) {                    // :292-326
                    for (int l = 0; l < nd.size; ++l) {
                        double prod = 1;
                        for (int p = 0; p < np; ++p) {
                            const Node& pa = m->nodes[nd.par[p]];
                            int hp = digit(l, p);
                            if (!pa.observed) prod *= q[hmap[nd.par[p]] + hp];
                            else if (hp != pa.obs) prod = 0;
                        }
                        if (cp[l * 4 + nd.obs] > 0) prod *= lg(nd, l, nd.obs);
                        else prod = -std::numeric_limits<double>::infinity();
                        tmp += prod;
                    }
                } else {                                     // :327-358
                    for (int h2 = 0; h2 < 4; ++h2)
                        for (int l = 0; l < nd.size; ++l) {
                            double prod = 1;
                            for (int p = 0; p < np; ++p) {
                                const Node& pa = m->nodes[nd.par[p]];
                                int hp = digit(l, p);
                                if (!pa.observed) prod *= q[hmap[nd.par[p]] + hp];
                                else if (hp != pa.obs) prod = 0;
                            }
                            if (cp[l * 4 + h2] > 0) prod *= q[hn + h2] * lg(nd, l, h2);
                            else prod = -std::numeric_limits<double>::infinity();
                            tmp += prod;
                        }
                }
                part2 += tmp;
            }
        return part1 - part2;
    }
    double free_energy() const { return free_energy(0, len - 1); }  // :214-216

    void optimize() {  // optimizeByNormalisation, :92-211
        std::vector<double> newq(q.size());
        const int max_steps = 100, min_steps = 2;
        bool converged = false;
        int k = 0;
        double old_score = free_energy(), new_score;
        const int total = (int)m->nodes.size();
        while ((k < max_steps && !converged) || k < min_steps) {
            for (int i = 0; i < total; ++i) {
                const Node& nd = m->nodes[i];
                if (nd.observed) continue;
                int hi = hmap[i];
                int np = (int)nd.par.size();
                for (int a = 0; a < 4; ++a) {
                    double acc = 0;
                    for (int l = 0; l < nd.size; ++l) {          // own rows, :117-135 (observed parents NOT filtered)
                        double prod = 1;
                        int nfac = 0;
                        for (int p = 0; p < np; ++p)
                            if (!m->nodes[nd.par[p]].observed) { prod *= q[hmap[nd.par[p]] + digit(l, p)]; ++nfac; }
                        acc += prod * lg(nd, l, a);
                        if (m->count_terms) m->c_upd_flops += 2 + std::max(nfac - 1, 0);
                    }
                    for (size_t c = 0; c < nd.chi.size(); ++c) {  // children, :139-191
                        int ci = nd.chi[c];
                        const Node& ch = m->nodes[ci];
                        int idx = ch.parent_index(i);
                        int ncp = (int)ch.par.size();
                        for (int ac = 0; ac < 4; ++ac) {
                            if (ch.observed && ch.obs != ac) continue;
                            for (int l = 0; l < ch.size; ++l) {
                                bool true_obs = true;
                                double prod = ch.observed ? 1.0 : q[hmap[ci] + ac];
                                for (int p = 0; p < ncp; ++p) {
                                    const Node& pa = m->nodes[ch.par[p]];
                                    int qp = digit(l, p);
                                    if (pa.observed && pa.obs != qp) true_obs = false;
                                    if (ch.par[p] != i && !pa.observed) prod *= q[hmap[ch.par[p]] + qp];
                                }
                                if (digit(l, idx) == a && true_obs) {
                                    acc += prod * lg(ch, l, ac);
                                    if (m->count_terms) {
                                        int nfac = ch.observed ? 0 : 1;
                                        for (int p = 0; p < ncp; ++p) if (ch.par[p] != i && !m->nodes[ch.par[p]].observed) ++nfac;
                                        m->c_upd_flops += 2 + std::max(nfac - 1, 0);
                                    }
                                }
                            }
                        }
                    }
                    newq[hi + a] = std::exp(acc);               // :192 (no max shift)
                    if (m->count_terms) m->c_upd_transc += 1;
                }
                if (m->count_terms) m->c_upd_flops += 7;
                double sum = 0;
                for (int a = 0; a < 4; ++a) sum += newq[hi + a];
                for (int a = 0; a < 4; ++a) q[hi + a] = newq[hi + a] / sum;  // Gauss-Seidel, :195-201
            }
            new_score = free_energy();
            if (std::fabs(old_score - new_score) < 1e-5) converged = true;  // :204-207 (sticky)
            old_score = new_score;
            ++k;
            if (m->count_terms) m->c_sweeps++;
        }
        sweeps = k;
    }
};

// ------------------------------------------------------------------------------------------
// Exact likelihood.  (1) literal frontier elimination, algorithm/SimpleNodeElimination.java:63-230
// (limited to frontiers of <= 9 nodes like the reference, :33-45); (2) Felsenstein pruning per
// tree for order 0 (what the GPU exact kernel computes); (3) brute force over all hidden nodes.
// ------------------------------------------------------------------------------------------
static double exact_elimination(const Model& m) {
    const int total = (int)m.nodes.size();
    std::vector<int> queue;          // nodeQueue
    std::vector<double> sum{1.0};    // sumVector[step-1]
    auto qpos = [&](int g) { for (size_t k = 0; k < queue.size(); ++k) if (queue[k] == g) return (int)k; return -1; };
    static const int pows[11] = {1, 4, 16, 64, 256, 1024, 4096, 16384, 65536, 262144, 1048576};
    for (int i = total - 1; i >= 0; --i) {
        const Node& nd = m.nodes[i];
        int old_len = (int)queue.size();
        if (qpos(i) < 0) queue.push_back(i);
        for (int p : nd.par) if (qpos(p) < 0) queue.push_back(p);
        int new_len = (int)queue.size();
        if (new_len > 9) throw std::runtime_error("SimpleNodeElimination frontier > 9 nodes (reference limit)");
        int ai = qpos(i);
        int len_all = pows[new_len];
        std::vector<double> gen;
        bool root = nd.is_root();
        if (!root) {
            gen.assign(len_all, 0.0);
            std::vector<int> pq(nd.par.size());
            for (size_t p = 0; p < nd.par.size(); ++p) pq[p] = qpos(nd.par[p]);
            for (int l1 = 0; l1 < len_all; ++l1) {
                int a = (l1 / pows[ai]) % 4;
                int l = 0;
                for (size_t p = 0; p < nd.par.size(); ++p) l += pows[p] * ((l1 / pows[pq[p]]) % 4);
                int old_ind = 0;
                for (int k = 0; k < old_len; ++k) old_ind += ((l1 / pows[k]) % 4) * pows[old_len - 1 - k];
                gen[l1] = nd.cpf[(size_t)l * 4 + a] * sum[old_ind];   // CPF.get(query,a): never the independent table, CPF.java:193-201
            }
        } else {
            gen = sum;  // :171-173
        }
        std::vector<double> nsum(len_all / 4, 0.0);
        for (int l1 = 0; l1 < len_all; ++l1) {
            int a = (l1 / pows[ai]) % 4;
            int ns = 0;
            for (int k = 0; k < new_len; ++k) {
                if (k > ai) ns += ((l1 / pows[k]) % 4) * pows[new_len - 1 - k];
                else if (k < ai) ns += ((l1 / pows[k]) % 4) * pows[new_len - 1 - k - 1];
            }
            double onehot = (nd.observed && nd.obs == a) ? 1.0 : 0.0;  // props._observation[a]
            double val;
            // CPF.get(0,a) may return the independent table for an unobserved leaf; roots are never leaves here.
            if (nd.observed && root) val = gen[l1] * onehot * nd.cpf[a];
            else if (nd.observed) val = gen[l1] * onehot;
            else if (root) val = gen[l1] * nd.cpf[a];
            else val = gen[l1];
            nsum[ns] += val;
        }
        queue.erase(queue.begin() + ai);
        sum.swap(nsum);
    }
    return sum[0];
}

// Felsenstein pruning for one order-0 tree t; returns P(column).  Gap leaves are missing data
// (sum over the leaf state of the true transition row), matching elimination's treatment.
static double pruning_tree(const Model& m, int t) {
    std::vector<double> part((size_t)m.N * 4, 1.0);
    for (int k = m.N - 1; k >= 1; --k) {
        const Node& nd = m.nodes[m.gid(t, k)];
        int par = m.tpl.parent[k];
        for (int a = 0; a < 4; ++a) {
            double s = 0;
            if (nd.phylo_leaf) {
                if (nd.observed) s = nd.cpf[(size_t)a * 4 + nd.obs];
                else for (int b = 0; b < 4; ++b) s += nd.cpf[(size_t)a * 4 + b];
            } else {
                for (int b = 0; b < 4; ++b) s += nd.cpf[(size_t)a * 4 + b] * part[(size_t)k * 4 + b];
            }
            part[(size_t)par * 4 + a] *= s;
        }
    }
    const Node& rt = m.nodes[m.gid(t, 0)];
    double p = 0;
    for (int a = 0; a < 4; ++a) p += rt.cpf[a] * part[a];
    return p;
}

// Brute force: sum the joint over every assignment of the unobserved nodes (any order).  Pinning only.
static double brute_force(const Model& m) {
    const int total = (int)m.nodes.size();
    std::vector<int> hid;
    for (int i = 0; i < total; ++i) if (!m.nodes[i].observed) hid.push_back(i);
    if (hid.size() > 12) throw std::runtime_error("brute force: too many hidden nodes");
    std::vector<int> val(total, 0);
    for (int i = 0; i < total; ++i) if (m.nodes[i].observed) val[i] = m.nodes[i].obs;
    double acc = 0;
    long long combos = 1LL << (2 * hid.size());
    for (long long c = 0; c < combos; ++c) {
        for (size_t h = 0; h < hid.size(); ++h) val[hid[h]] = (int)((c >> (2 * h)) & 3);
        double p = 1;
        for (int i = 0; i < total; ++i) {
            const Node& nd = m.nodes[i];
            int l = 0;
            for (size_t pp = 0; pp < nd.par.size(); ++pp) l += val[nd.par[pp]] << (2 * pp);
            p *= nd.cpf[(size_t)l * 4 + val[i]];
        }
        acc += p;
    }
    return acc;
}

// ------------------------------------------------------------------------------------------
// Model-level scoring
// ------------------------------------------------------------------------------------------
// models/PhyloBayesModel.java:222-267 (+ AbstractPhyloModel.java:100-119: a fresh, uniformly
// initialised mean field for every call).  cols = W columns of S codes; revcomp scores the
// reverse complement of the window (jstacs!/de/jstacs/models/mixture/StrandModel.java:631-640).
static double score_window(Model& m, const uint8_t* cols, bool revcomp, int* sweeps) {
    for (int t = 0; t < m.W; ++t) {
        const uint8_t* col = revcomp ? cols + (size_t)(m.W - 1 - t) * m.S : cols + (size_t)t * m.S;
        set_observation(m, t, col, revcomp);
    }
    m.calls_total++;
    if (m.mode == EXACT_ELIMINATION) { if (sweeps) *sweeps = 0; return std::log(exact_elimination(m)); }
    if (m.mode == EXACT_PRUNING) {
        if (m.order != 0) throw std::runtime_error("pruning mode needs order 0");
        double s = 0;
        for (int t = 0; t < m.W; ++t) s += std::log(pruning_tree(m, t));
        if (sweeps) *sweeps = 0;
        return s;
    }
    MeanField mf(&m, m.W);
    mf.init();
    mf.optimize();
    m.sweeps_total += mf.sweeps;
    if (sweeps) *sweeps = mf.sweeps;
    return -mf.free_energy();
}

// models/PhyloBackground.java:241-305.  codes = whole alignment [L][S].
// Ranges shorter than order+1 with order >= 1 leave stale observations in the trailing trees in
// the reference (MeanFieldForBayesNet.java:84); here those trees keep whatever the previous call
// on this Model object left there, exactly like the Java object graph would.
static double bg_range(Model& m, const uint8_t* codes, int L, int start, int end) {
    if (end < start) return 0;
    if (start < 0 || end >= L) throw std::runtime_error("bg_range: bad range");
    const int order = m.order;
    int seq_len = end - start + 1;
    double log_sum = 0;
    int first_len = std::min(order + 1, seq_len);
    auto observe = [&](int u, int n_trees) {
        for (int t = 0; t < n_trees; ++t) set_observation(m, t, codes + (size_t)(u + t) * m.S, false);
    };
    auto eval = [&](MeanField& mf, int a, int b) -> double {
        if (m.mode == MEANFIELD) return mf.free_energy(a, b);
        if (m.mode == EXACT_PRUNING) {
            if (order != 0) throw std::runtime_error("pruning mode needs order 0");
            return -std::log(pruning_tree(m, 0));
        }
        return -std::log(exact_elimination(m));  // :274-277,288-291: whole-net likelihood (only meaningful for order 0)
    };
    int startpos = start;
    {   // first window, :260-283
        observe(start, first_len);
        MeanField mf(&m, first_len);
        mf.init();
        if (m.mode == MEANFIELD) { mf.optimize(); m.sweeps_total += mf.sweeps; }
        for (int i = 0; i < order && startpos <= end; ++i, ++startpos) log_sum += eval(mf, i, i);
    }
    // windows u = start, start+1, ..., end-order; window i scores absolute position startpos, :284-297
    int n_windows = 1 + std::max(0, (end - order) - (start + 1) + 1);
    for (int i = 0; i < n_windows && startpos <= end; ++i, ++startpos) {
        int u = start + i;
        int wl = (i == 0) ? first_len : order + 1;
        observe(u, wl);
        MeanField mf(&m, wl);
        mf.init();
        if (m.mode == MEANFIELD) { mf.optimize(); m.sweeps_total += mf.sweeps; }
        log_sum += eval(mf, order, order);
        m.calls_total++;
    }
    return -log_sum;
}

// jstacs!/de/jstacs/utils/Normalisation.java:241-252,352-375: max-shift, exp, sum, divide; returns offset+log(sum).
static double log_sum_normalisation(double* d, int start, int end) {
    double offset = -std::numeric_limits<double>::infinity();
    for (int i = start; i < end; ++i) offset = std::max(offset, d[i]);
    double sum = 0;
    for (int i = start; i < end; ++i) { d[i] = std::exp(d[i] - offset); sum += d[i]; }
    if (sum != 1.0) for (int i = start; i < end; ++i) d[i] /= sum;
    return offset + std::log(sum);
}
// Normalisation.getLogSum, :63-91
static double get_log_sum(const double* v, int n) {
    double offset = -std::numeric_limits<double>::infinity(), sum = 0;
    for (int i = 0; i < n; ++i) offset = std::max(offset, v[i]);
    if (std::isinf(offset)) return -std::numeric_limits<double>::infinity();
    for (int i = 0; i < n; ++i) sum += std::exp(v[i] - offset);
    return offset + std::log(sum);
}

struct EStepOut { double ll = 0, w0 = 0, w1 = 0; };

// One alignment of SingleHiddenMotifMixture.getNewWeights (jstacs!/de/jstacs/models/mixture/motif/
// SingleHiddenMotifMixture.java:520-564) + modify (:585-596) + UniformPositionPrior (:55-61).
// strand_logw != NULL wraps the motif in a StrandModel (StrandModel.java:631-640 via
// AbstractMixtureModel.java:1158-1170).  faithful_cost recomputes every background range like
// the Java does (its cache never hits); otherwise order-0 column scores are
// computed once and summed in the same order, which yields identical bits.
static void estep_alignment(Model& fg, Model& bg, const uint8_t* codes, int L, const double logw[2],
                            const double* strand_logw, double weight, bool faithful_cost,
                            double* gamma /*[L-W+1] or NULL*/, double* fwd /*NULL or [n]*/, double* rev,
                            EStepOut& out, int* arg_off, int* arg_strand, double* profile /*NULL or [n]*/) {
    const int W = fg.W, S = fg.S, ord = bg.order;
    const int n = L - W + 1;
    if (n <= 0) throw std::runtime_error("alignment shorter than the motif");
    std::vector<double> colF;
    auto bg_of = [&](int s, int e) -> double {
        if (e < s) return 0;
        if (faithful_cost || ord != 0) return bg_range(bg, codes, L, s, e);
        double ls = 0;
        for (int u = s; u <= e; ++u) ls += colF[u];
        return -ls;
    };
    if (!faithful_cost && ord == 0) {
        colF.resize(L);
        for (int u = 0; u < L; ++u) colF[u] = -bg_range(bg, codes, L, u, u);
    }
    double log_pseq = bg_of(0, L - 1);                                  // :526
    std::vector<double> a(n);
    double best = -std::numeric_limits<double>::infinity();
    int best_j = 0, best_strand = 0;
    int b1 = -ord, b2 = W + ord - 1;
    for (int j = 0; j < n; ++j, ++b1, ++b2) {
        int s = std::max(b1, 0), e = std::min(b2, L - 1);
        double motif, f0 = score_window(fg, codes + (size_t)j * S, false, nullptr), f1 = 0;
        if (strand_logw) {
            f1 = score_window(fg, codes + (size_t)j * S, true, nullptr);
            double comp[2] = {strand_logw[0] + f0, strand_logw[1] + f1};
            motif = get_log_sum(comp, 2);
        } else motif = f0;
        if (fwd) fwd[j] = f0;
        if (rev) rev[j] = f1;
        a[j] = -std::log((double)(L - W + 1)) + motif - bg_of(s, e) + bg_of(s, j - 1) + bg_of(j + W, e);  // :536-540
        if (profile) profile[j] = a[j];
        if (a[j] > best) {  // util/Util.java:528-538 strict '>' => first maximum
            best = a[j]; best_j = j;
            best_strand = (strand_logw && (strand_logw[1] + f1) > (strand_logw[0] + f0)) ? 1 : 0;  // SHM.java:737-754
        }
    }
    double help[2];
    help[0] = logw[0] + log_sum_normalisation(a.data(), 0, n);          // :588
    help[1] = logw[1];
    double lse = log_sum_normalisation(help, 0, 2);                     // :590
    out.ll += weight * (log_pseq + lse);                                // :544
    help[0] *= weight;                                                  // :547
    out.w0 += help[0];
    out.w1 += weight * help[1];
    if (gamma) for (int j = 0; j < n; ++j) gamma[j] = a[j] * help[0];   // :553-555
    if (arg_off) *arg_off = best_j;
    if (arg_strand) *arg_strand = best_strand;
}

// io/SequenceSpecificDataSelector.java:37-99 for ONE group (windows sharing a parent): sort by
// weight ascending, walk from the top, add (if raw weight >= q), stop after the running sum of
// normalised weights exceeds t.  The reference's AssociativeSort is an unstable quicksort, so the
// relative order of exactly equal weights is unspecified there; here ties are taken highest
// index first (stable ascending sort walked from the top).
static int select_group(const double* w, int n, double t, double q, int* idx_out, double* w_out) {
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return w[x] < w[y]; });
    double total = w[0];                                  // Normalisation.sumNormalisation, :410-418 (array order)
    for (int i = 1; i < n; ++i) total += w[i];
    double sum = 0;
    int cnt = 0;
    for (int k = n - 1; k >= 0; --k) {
        sum += w[order[k]] / total;
        if (w[order[k]] >= q) { idx_out[cnt] = order[k]; w_out[cnt] = w[order[k]]; ++cnt; }
        if (sum > t) break;
    }
    return cnt;
}

// util/MatrixLinearisation.java:15-19 via Normalisation.logSumNormalisation
static void lambda2pi(const double* lambda, double* pi, int n) {
    for (int i = 0; i < n; i += 4) {
        for (int k = 0; k < 4; ++k) pi[i + k] = lambda[i + k];
        log_sum_normalisation(pi, i, i + 4);
    }
}

// The fixed-q objective of the M-step: optimizing/AbstractFreeEnergyOptimizer.java:95-114 with
// FE_byParameter_ForBayesNodeJstacs.java:43-53 (one position) or MotifParamOptimizer.java:46-61.
struct FESet {
    Model* m;
    int n = 0;
    std::vector<uint8_t> cols;          // [n][W][S] (already reverse-complemented where asked)
    std::vector<double> weights;
    std::vector<MeanField> mfs;         // q fixed after prepare
};

static FESet* fe_prepare(Model& m, const uint8_t* cols, int n, const double* weights) {
    FESet* fs = new FESet();
    fs->m = &m; fs->n = n;
    fs->cols.assign(cols, cols + (size_t)n * m.W * m.S);
    fs->weights.assign(weights, weights + n);
    fs->mfs.reserve(n);
    for (int u = 0; u < n; ++u) {       // PhyloBayesModel.java:142-147 / PhyloBackground.java:127-134
        for (int t = 0; t < m.W; ++t) set_observation(m, t, &fs->cols[((size_t)u * m.W + t) * m.S], false);
        fs->mfs.emplace_back(&m, m.W);
        fs->mfs.back().init();
        fs->mfs.back().optimize();
    }
    return fs;
}

static double fe_sum(FESet& fs) {  // getWeightedLogLikelihoodSum, :95-114
    Model& m = *fs.m;
    double sum = 0;
    for (int u = 0; u < fs.n; ++u) {
        for (int t = 0; t < m.W; ++t) set_observation(m, t, &fs.cols[((size_t)u * m.W + t) * m.S], false);
        sum -= fs.mfs[u].free_energy() * fs.weights[u];
    }
    return sum;
}

// evaluateFunction(lambda): pos >= 0 -> FE_byParameter (one tree), pos < 0 -> all positions concatenated.
static double fe_eval(FESet& fs, int pos, const double* lambda, int dim) {
    Model& m = *fs.m;
    if (pos >= 0) {
        if ((int)m.pi[pos].size() != dim) throw std::runtime_error("fe_eval: dim mismatch");
        lambda2pi(lambda, m.pi[pos].data(), dim);
        init_tree_params(m, pos);
    } else {
        int off = 0;
        for (int t = 0; t < m.W; ++t) {
            int d = (int)m.pi[t].size();
            if (off + d > dim) throw std::runtime_error("fe_eval: dim mismatch");
            lambda2pi(lambda + off, m.pi[t].data(), d);
            init_tree_params(m, t);
            off += d;
        }
    }
    return fe_sum(fs);
}

// jstacs!/de/jstacs/algorithms/optimization/NumericalDifferentiableFunction.java:64-84:
// forward difference, x[i] perturbed in place and restored, dim+1 evaluations.
static void fe_grad(FESet& fs, int pos, double* lambda, int dim, double eps, double* f0_out, double* g) {
    double current = fe_eval(fs, pos, lambda, dim);
    for (int i = 0; i < dim; ++i) {
        double h = lambda[i];
        lambda[i] += eps;
        g[i] = (fe_eval(fs, pos, lambda, dim) - current) / eps;
        lambda[i] = h;
    }
    fe_eval(fs, pos, lambda, dim);  // leave the model at lambda (callers re-evaluate anyway)
    if (f0_out) *f0_out = current;
}

}  // namespace orc

// ------------------------------------------------------------------------------------------
// C interface for ctypes (tests, bench cpu_baseline).
// ------------------------------------------------------------------------------------------
using namespace orc;
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_model_new(int kind, int W, int order, const char* newick, int evol, double hky_ratio) {
    ORC_TRY return build_model(kind, W, order, newick, evol, hky_ratio); ORC_CATCH(nullptr)
}
void orc_model_free(void* h) { delete (Model*)h; }
void* orc_model_clone(void* h) { return new Model(*(Model*)h); }
int orc_model_nodes(void* h) { return ((Model*)h)->N; }
int orc_model_species(void* h) { return ((Model*)h)->S; }
int orc_model_trees(void* h) { return ((Model*)h)->W; }
int orc_model_pi_dim(void* h, int t) { return (int)((Model*)h)->pi[t].size(); }
void orc_model_topology(void* h, int* parent, int* leaf_species, double* dist) {
    Model* m = (Model*)h;
    for (int k = 0; k < m->N; ++k) { parent[k] = m->tpl.parent[k]; dist[k] = m->tpl.dist[k]; leaf_species[k] = -1; }
    for (int s = 0; s < m->S; ++s) leaf_species[m->leaf_local[s]] = s;
}
int orc_model_leaf_label(void* h, int s, char* buf, int n) {
    Model* m = (Model*)h;
    const std::string& l = m->tpl.label[m->leaf_local[s]];
    std::snprintf(buf, n, "%s", l.c_str());
    return (int)l.size();
}
int orc_model_set_pi(void* h, int t, const double* pi, int n) {
    Model* m = (Model*)h;
    if (t < 0 || t >= m->W || (int)m->pi[t].size() != n) { g_err = "set_pi: bad tree or size"; return -1; }
    m->pi[t].assign(pi, pi + n);
    init_tree_params(*m, t);
    return 0;
}
void orc_model_get_pi(void* h, int t, double* pi) { Model* m = (Model*)h; std::copy(m->pi[t].begin(), m->pi[t].end(), pi); }
int orc_model_set_branch(void* h, int t, int node, double len) {
    Model* m = (Model*)h;
    if (t < 0 || t >= m->W || node <= 0 || node >= m->N) { g_err = "set_branch: bad index"; return -1; }
    m->nodes[m->gid(t, node)].dist = len;
    init_tree_params(*m, t);
    return 0;
}
void orc_model_set_options(void* h, int mode, int allow_indep, int log_in_loop) {
    Model* m = (Model*)h; m->mode = mode; m->allow_indep = allow_indep != 0; m->log_in_loop = log_in_loop != 0;
}
// CPF table of (tree t, local node k): returns rows; out[rows*4]
int orc_model_cpf(void* h, int t, int k, int indep, double* out) {
    Model* m = (Model*)h; const Node& nd = m->nodes[m->gid(t, k)];
    const std::vector<double>& v = indep ? nd.indep : nd.cpf;
    std::copy(v.begin(), v.end(), out);
    return nd.size;
}
long long orc_model_sweeps_total(void* h) { return ((Model*)h)->sweeps_total; }
void orc_model_reset_stats(void* h) { ((Model*)h)->sweeps_total = 0; ((Model*)h)->calls_total = 0; }

double orc_score_window(void* h, const uint8_t* cols, int revcomp, int* sweeps) {
    ORC_TRY return score_window(*(Model*)h, cols, revcomp != 0, sweeps); ORC_CATCH(NAN)
}
// Term counter: scores one window with counting on and returns, per unit,
// out[0] = flop of one update sweep, out[1] = flop of one free-energy evaluation, out[2] / out[3] = their
// transcendental calls, out[4] = sweeps of this window.  Total work of a window = K*upd + (K+1)*fe.
int orc_count_terms(void* h, const uint8_t* cols, int revcomp, double* out) {
    ORC_TRY
    Model* m = (Model*)h;
    m->count_terms = true;
    m->c_upd_flops = m->c_upd_transc = m->c_fe_flops = m->c_fe_transc = m->c_sweeps = m->c_fe_calls = 0;
    int k = 0;
    const int mode = m->mode;
    m->mode = MEANFIELD;
    score_window(*m, cols, revcomp != 0, &k);
    m->mode = mode;
    m->count_terms = false;
    out[0] = (double)m->c_upd_flops / std::max<long long>(m->c_sweeps, 1);
    out[1] = (double)m->c_fe_flops / std::max<long long>(m->c_fe_calls, 1);
    out[2] = (double)m->c_upd_transc / std::max<long long>(m->c_sweeps, 1);
    out[3] = (double)m->c_fe_transc / std::max<long long>(m->c_fe_calls, 1);
    out[4] = (double)k;
    return 0;
    ORC_CATCH(-1)
}
// all offsets of one alignment; out[L-W+1]
int orc_score_windows(void* h, const uint8_t* codes, int L, int revcomp, double* out, int* sweeps) {
    ORC_TRY
    Model* m = (Model*)h;
    for (int j = 0; j + m->W <= L; ++j)
        out[j] = score_window(*m, codes + (size_t)j * m->S, revcomp != 0, sweeps ? sweeps + j : nullptr);
    return 0;
    ORC_CATCH(-1)
}
double orc_exact_bruteforce(void* h, const uint8_t* cols) {
    ORC_TRY
    Model* m = (Model*)h;
    for (int t = 0; t < m->W; ++t) set_observation(*m, t, cols + (size_t)t * m->S, false);
    return std::log(brute_force(*m));
    ORC_CATCH(NAN)
}
double orc_bg_range(void* h, const uint8_t* codes, int L, int start, int end) {
    ORC_TRY return bg_range(*(Model*)h, codes, L, start, end); ORC_CATCH(NAN)
}
int orc_bg_columns(void* h, const uint8_t* codes, int L, double* out) {  // order 0: -F of each column
    ORC_TRY
    Model* m = (Model*)h;
    for (int u = 0; u < L; ++u) out[u] = bg_range(*m, codes, L, u, u);
    return 0;
    ORC_CATCH(-1)
}

// Whole-sample E-step.  codes [sum cols][S]; col_offsets [n_aln+1]; outputs may be NULL.
// gamma / fwd / rev / profile are indexed by win_offsets (computed here: prefix sums of L_i-W+1).
// threads > 1 uses std::thread workers over alignments with a private copy of both models per thread (the Java
// object graph cannot be shared: observations live in the model).
int orc_estep(void* hfg, void* hbg, int n_aln, const int64_t* col_offsets, const uint8_t* codes,
              const double* logw, const double* strand_logw, const double* aln_weights, int faithful_cost,
              int threads, double* gamma, double* fwd, double* rev, double* profile, double* w_out,
              double* ll_out, int32_t* arg_off, uint8_t* arg_strand, long long* sweeps_out) {
    ORC_TRY
    Model* fg0 = (Model*)hfg; Model* bg0 = (Model*)hbg;
    const int W = fg0->W, S = fg0->S;
    std::vector<int64_t> woff(n_aln + 1, 0);
    for (int i = 0; i < n_aln; ++i) woff[i + 1] = woff[i] + std::max<int64_t>(0, (col_offsets[i + 1] - col_offsets[i]) - W + 1);
    std::vector<EStepOut> per(n_aln);
    std::string err;
    long long sweeps = 0, sweeps_bg = 0;
    if (threads < 1) threads = 1;
    std::atomic<int> next{0};
    std::mutex mu;
    auto worker = [&]() {
        Model fg = *fg0, bg = *bg0;   // private net state per thread
        fg.sweeps_total = bg.sweeps_total = 0;
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_aln) break;
            try {
                int L = (int)(col_offsets[i + 1] - col_offsets[i]);
                int ao = 0, as = 0;
                estep_alignment(fg, bg, codes + (size_t)col_offsets[i] * S, L, logw, strand_logw,
                                aln_weights ? aln_weights[i] : 1.0, faithful_cost != 0,
                                gamma ? gamma + woff[i] : nullptr, fwd ? fwd + woff[i] : nullptr,
                                rev ? rev + woff[i] : nullptr, per[i], &ao, &as,
                                profile ? profile + woff[i] : nullptr);
                if (arg_off) arg_off[i] = ao;
                if (arg_strand) arg_strand[i] = (uint8_t)as;
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(mu);
                err = e.what();
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        sweeps += fg.sweeps_total;
        sweeps_bg += bg.sweeps_total;
    };
    if (threads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    if (!err.empty()) { g_err = err; return -1; }
    // the Java accumulates ll, w[0], w[1] in alignment order (:544-550); w starts from initWithPrior = 0 here
    double ll = 0, w0 = 0, w1 = 0;
    for (int i = 0; i < n_aln; ++i) { ll += per[i].ll; w0 += per[i].w0; w1 += per[i].w1; }
    if (ll_out) *ll_out = ll;
    if (w_out) { w_out[0] = w0; w_out[1] = w1; }
    // [0] motif sweeps, [1] background sweeps.  NB the Java optimises the first window of every background range
    // twice (PhyloBackground.java:271-273 and :284-287), so a range of one order-0 column costs two optimisations.
    if (sweeps_out) { sweeps_out[0] = sweeps; sweeps_out[1] = sweeps_bg; }
    return 0;
    ORC_CATCH(-1)
}

// ------------------------------------------------------------------------------------------
// Non-phylogenetic alignment models, order 0 (the `model.evolution=false` branch, models/ModelUtil.java:84-90).
// models/AlignmentBasedModel.java:188-248: score of a window = sum over species rows o of (sum over motif positions
// i of log(condProb[i][x] / 1)), a gap scoring log(0.25) (:189-191 isCompletylUnObserved for a single position);
// the inner sum starts at 0 for every row (:170-175) and is then added to the running total (:241-243).
// models/AlignmentBasedBGModel.java:173-246: per position, per row, one running total (:236-240).
// Orders >= 1 use the ACCESS_TABLE expansion with its acknowledged bug (:128-141) and are not restated.
// ------------------------------------------------------------------------------------------
int orc_ab_score_windows(const double* condprob /*[W][4]*/, int W, const uint8_t* codes, int L, int S, int revcomp, double* out) {
    for (int j = 0; j + W <= L; ++j) {
        double log_sum = 0;
        for (int o = 0; o < S; ++o) {
            double sum = 0;
            for (int i = 0; i < W; ++i) {
                int c = revcomp ? codes[(size_t)(j + W - 1 - i) * S + o] : codes[(size_t)(j + i) * S + o];
                if (revcomp && c < 4) c = 3 - c;
                sum += (c >= 4) ? std::log(0.25) : std::log(condprob[i * 4 + c] / 1);
            }
            log_sum += sum;
        }
        out[j] = log_sum;
    }
    return 0;
}
int orc_ab_bg_range(const double* condprob /*[4]*/, const uint8_t* codes, int L, int S, int start, int end, double* out) {
    double log_sum = 0;
    for (int l = start; l <= end; ++l)
        for (int o = 0; o < S; ++o) {
            const int c = codes[(size_t)l * S + o];
            log_sum += (c >= 4) ? std::log(0.25) : std::log(condprob[c] / 1);
        }
    *out = (end < start) ? 0.0 : log_sum;
    return 0;
}
// AlignmentBasedModel.train :97-168 for order 0: counts[i][b] = PSEUDO_COUNT + sum over selected windows, rows:
// weight for the observed base, weight / 4 for each base under a gap; then each block of 4 is normalised by its sum
// (Normalisation.sumNormalisation, jstacs :410-418).  cols [n][W][S].
int orc_ab_train(const uint8_t* cols, int n, int W, int S, const double* weights, double pseudo, double* condprob_out) {
    std::vector<double> counts((size_t)W * 4, pseudo);
    for (int w = 0; w < n; ++w)
        for (int o = 0; o < S; ++o)
            for (int l = 0; l < W; ++l) {
                const int c = cols[((size_t)w * W + l) * S + o];
                if (c >= 4) for (int a = 0; a < 4; ++a) counts[l * 4 + a] += weights[w] / 4;
                else counts[l * 4 + c] += weights[w] / 1;
            }
    for (int i = 0; i < W; ++i) {
        double sum = counts[i * 4];
        for (int k = 1; k < 4; ++k) sum += counts[i * 4 + k];
        for (int k = 0; k < 4; ++k) condprob_out[i * 4 + k] = counts[i * 4 + k] / sum;
    }
    return 0;
}

int orc_select_group(const double* w, int n, double t, double q, int* idx_out, double* w_out) {
    ORC_TRY return select_group(w, n, t, q, idx_out, w_out); ORC_CATCH(-1)
}
void orc_lambda2pi(const double* lambda, double* pi, int n) { lambda2pi(lambda, pi, n); }

void* orc_fe_prepare(void* h, const uint8_t* cols, int n, const double* weights) {
    ORC_TRY return fe_prepare(*(Model*)h, cols, n, weights); ORC_CATCH(nullptr)
}
void orc_fe_free(void* fs) { delete (FESet*)fs; }
double orc_fe_sum(void* fs) { ORC_TRY return fe_sum(*(FESet*)fs); ORC_CATCH(NAN) }
double orc_fe_eval(void* fs, int pos, const double* lambda, int dim) {
    ORC_TRY return fe_eval(*(FESet*)fs, pos, lambda, dim); ORC_CATCH(NAN)
}
int orc_fe_grad(void* fs, int pos, double* lambda, int dim, double eps, double* f0, double* g) {
    ORC_TRY fe_grad(*(FESet*)fs, pos, lambda, dim, eps, f0, g); return 0; ORC_CATCH(-1)
}
// q of selected window u, in global node order, -1 rows for observed nodes; out[W*N*4]
int orc_fe_get_q(void* fsh, int u, double* out) {
    FESet* fs = (FESet*)fsh; const MeanField& mf = fs->mfs[u];
    for (size_t g = 0; g < mf.hmap.size(); ++g)
        for (int a = 0; a < 4; ++a) out[g * 4 + a] = mf.hmap[g] < 0 ? -1.0 : mf.q[mf.hmap[g] + a];
    return mf.sweeps;
}

// Column encoding: jstacs!/de/jstacs/data/alphabets/UnobservableDNAAlphabet.java:116 (A,C,G,T,- -> 0..4,
// case-insensitive via DiscreteAlphabet.getCode).  rows: S strings of length L -> codes [L][S].
int orc_encode_alignment(const char* const* rows, int S, int L, uint8_t* codes) {
    for (int s = 0; s < S; ++s)
        for (int u = 0; u < L; ++u) {
            int c;
            switch (rows[s][u]) {
                case 'A': case 'a': c = 0; break;
                case 'C': case 'c': c = 1; break;
                case 'G': case 'g': c = 2; break;
                case 'T': case 't': c = 3; break;
                case '-': c = 4; break;
                default: g_err = "WrongAlphabetException: symbol not in {A,C,G,T,-}"; return -1;
            }
            codes[(size_t)u * S + s] = (uint8_t)c;
        }
    return 0;
}
// io/PhyloSample.java:196-223: drop columns that are gaps in every species. Returns new length.
int orc_drop_allgap_columns(uint8_t* codes, int L, int S) {
    int w = 0;
    for (int u = 0; u < L; ++u) {
        int gaps = 0;
        for (int s = 0; s < S; ++s) gaps += codes[(size_t)u * S + s] == 4;
        if (gaps < S) { if (w != u) std::memmove(codes + (size_t)w * S, codes + (size_t)u * S, S); ++w; }
    }
    return w;
}
// The device layout the CUDA path must reproduce bit for bit (include/phyfoo.h "packed columns"):
// base plane word k of a column holds species 16k..16k+15, 2 bits each, species s at bits 2(s%16);
// a gap stores base bits 0.  gap plane word k holds species 32k..32k+31, 1 bit each.
void orc_pack_columns(const uint8_t* codes, int64_t n_cols, int S, uint32_t* bases, uint32_t* gaps) {
    int nb = (S + 15) / 16, ng = (S + 31) / 32;
    for (int64_t c = 0; c < n_cols; ++c) {
        for (int k = 0; k < nb; ++k) bases[(size_t)k * n_cols + c] = 0;
        for (int k = 0; k < ng; ++k) gaps[(size_t)k * n_cols + c] = 0;
        for (int s = 0; s < S; ++s) {
            uint8_t v = codes[(size_t)c * S + s];
            if (v >= 4) gaps[(size_t)(s / 32) * n_cols + c] |= 1u << (s % 32);
            else bases[(size_t)(s / 16) * n_cols + c] |= (uint32_t)v << (2 * (s % 16));
        }
    }
}

}  // extern "C"
