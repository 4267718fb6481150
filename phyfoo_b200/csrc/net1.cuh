// net1.cuh -- order-1 phylogenetic Bayesian network (motif with model.fg.order = 1, config 3): mean-field
// coordinate ascent + free energy for one window = one loopy net of W trees in which every node also has the
// same node of the previous position as a parent (bayesNet/BayesNetHandler.java:230-256,
// models/PhyloBayesModel.java:356-370).  Reference loops: algorithm/MeanFieldForBayesNet.java:92-211 (update),
// :232-365 (free energy).
//
// The update is sequential over the W*N nodes of a window (Gauss-Seidel in global node order), so the parallelism
// is: one QUAD of lanes per window (lane a owns symbol a of every node: its exponent, its exp, its share of the
// free energy), 8 windows per warp, and as many windows per SM as shared memory holds (the q of the hidden internal
// nodes of all W trees stays in shared memory for the whole window; the rarely touched q of gap leaves lives in a
// global scratch slot).  Warps never synchronise with each other: quads pull windows from a global counter.
//
// Same fixed-point iteration as the Java, sums re-associated:
//  * logs of the CPF tables are hoisted into per-parameter-set tables;
//  * the sweep and the free energy after it are fused: every CPF term involves a node and its parents only, and
//    parents precede the node in the global order, so the term is final when the node has been updated;
//  * a gap leaf is updated right after its tree parent (no node between the two positions in the reference's order
//    interacts with it);
//  * pass 0 (free energy at the uniform start) is the update pass with the exponent forced to 0.
//
// Table layout per position t (E1 = 16 + 32 (N-1) doubles, logs): root [ctx][4] at 0 (ctx = value of the root of
// position t-1; only row 0 is used at t = 0); node k >= 1: the reference's table [l][b], l = a_treeparent + 4 ctx
// (bayesNet/CPF.java:56-59,238-256: parent 0 is the least significant digit) stored as its off-diagonal values
// Lo[ctx][b] at 16 + 32 (k-1) and its diagonal values Ld[ctx][b] 16 further (see net1_pass).  The "independent" table
// of an unobserved leaf (CPF.java:146-152) has pi[ctx][b] in every row = the root table.
#pragma once

struct Net1Params {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, I, S, E1, prog_len;
    const int* prog;
    const double* tab;       // [W][E1]
    double* out; int32_t* sweeps;
    unsigned long long* counters;
    int max_sweeps, min_sweeps; double tol;
    double* gq;              // gap-leaf q: [(t*S + leaf slot)*4 + a][n_slots]
    int64_t n_slots;
    int nwb;                 // windows per block
    int tab_in_smem;         // copy the tables into shared memory (when they fit beside the window state)
    const int64_t* item_col0; const uint8_t* item_strand;   // nullable: explicit window list
    // M-step "prepare": converged q of every listed window, entry-major across the n_windows listed windows:
    // sink_qint[((t*I + i)*4 + a) * n + item], sink_qleaf[((t*S + leaf slot)*4 + a) * n + item], sink_ent[item]
    double* sink_qint; double* sink_qleaf; double* sink_ent;
};

struct Net1Ctx {
    // shared-memory views of this block
    double* q; int nwb; int wq;          // q[((t*I + i)*4 + a) * nwb + wq]
    const uint64_t* cb; const uint32_t* cg;   // column words [t][nwb]
    const int* par_internal; const int* node_of; const int* leaf_begin; const int* leaf_node; const int* leaf_species;
    const int* ichild_begin; const int* ichild;
    int W, I, S, E1;
    const double* tab;
    double* gq; int64_t n_slots; int64_t slot;
    unsigned qmask; int qb, a;
};

// table entry: tables live in shared memory when they fit (generic load), else in global memory behind L1
__device__ __forceinline__ double n1_ld(const double* p) { return *p; }
__device__ __forceinline__ double& n1_q(const Net1Ctx& c, int t, int i, int a) { return c.q[(size_t)((t * c.I + i) * 4 + a) * c.nwb + c.wq]; }
__device__ __forceinline__ double& n1_gq(const Net1Ctx& c, int t, int kk, int a) { return c.gq[(size_t)((t * c.S + kk) * 4 + a) * c.n_slots + c.slot]; }
__device__ __forceinline__ ColWords n1_col(const Net1Ctx& c, int t) {
    ColWords w; w.b = c.cb[(size_t)t * c.nwb + c.wq]; w.g = c.cg[(size_t)t * c.nwb + c.wq]; return w;
}

// exp of the four exponents of a node -> normalised q of this lane and log q; MeanFieldForBayesNet.java:192-201
__device__ __forceinline__ void n1_normalise(const Net1Ctx& c, double s, bool update, double& qn, double& lq) {
    const double ex = update ? s : 0.0;
    const double e = phy_exp(ex);                                               // :192, no max shift
    const double e0 = __shfl_sync(c.qmask, e, c.qb), e1 = __shfl_sync(c.qmask, e, c.qb + 1);
    const double e2 = __shfl_sync(c.qmask, e, c.qb + 2), e3 = __shfl_sync(c.qmask, e, c.qb + 3);
    const double z = ((e0 + e1) + e2) + e3;                                 // :195-198
    qn = e * __drcp_rn(z);                                                  // :199-201 (correctly rounded reciprocal, then one product)
    lq = ex - phy_log(z);
}

// sum of the other three entries of v (index order), i.e. the weight of the off-diagonal rows for own symbol a
__device__ __forceinline__ double n1_rest(const double* v, int a) {
    const double x0 = a == 0 ? v[1] : v[0], x1 = a <= 1 ? v[2] : v[1], x2 = a <= 2 ? v[3] : v[2];
    return (x0 + x1) + x2;
}
// v[a] without dynamic register indexing
__device__ __forceinline__ double n1_pick(const double* v, int a) { return a == 0 ? v[0] : a == 1 ? v[1] : a == 2 ? v[2] : v[3]; }

// One pass over the window's net; returns this lane's share of the free energy F = part1 - part2.
//
// The substitution tables have F81 structure (evolution/FS81alpha.java:76-92, FS81beta.java:81-103): for a given
// context c and target b the entry is the same for every parent symbol a != b, so a table is stored as its
// off-diagonal values Lo[c][b] and its diagonal values Ld[c][b] (model_host.hpp checks this bit for bit), and a sum
// over the 16 rows l = a + 4c collapses to  sum_c q_ctx(c) (q_par(b) Ld[c][b] + (sum_{a != b} q_par(a)) Lo[c][b]).
// Layout per position (E1 = 16 + 32 (N-1) logs): root [ctx][4] at 0; node k >= 1: Lo at 16 + 32 (k-1), Ld 16 further.
__device__ double net1_pass(const Net1Ctx& c, bool update) {
    const int a = c.a;
    double F = 0.0;
    for (int t = 0; t < c.W; ++t) {
        const double* tab = c.tab + (size_t)t * c.E1;
        const double* tabn = tab + c.E1;
        const bool has_next = t + 1 < c.W;
        const ColWords cw = n1_col(c, t);
        ColWords cwp; cwp.b = 0; cwp.g = 0;
        ColWords cwn; cwn.b = 0; cwn.g = 0;
        if (t > 0) cwp = n1_col(c, t - 1);
        if (has_next) cwn = n1_col(c, t + 1);
        for (int i = 0; i < c.I; ++i) {
            const int k = c.node_of[i];
            const int p = c.par_internal[i];
            const double* Lo = tab + 16 + 32 * (k - 1);
            const double* Ld = Lo + 16;
            // ---- own rows, :117-135 (all parents of an internal node are hidden)
            double own;
            if (i == 0) {
                if (t == 0) own = n1_ld(tab + a);
                else {
                    own = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) own += n1_q(c, t - 1, 0, cc) * n1_ld(tab + cc * 4 + a);
                }
            } else {
                double qp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) qp[j] = n1_q(c, t, p, j);
                double A, B;
                if (t == 0) { A = n1_ld(Ld + a); B = n1_ld(Lo + a); }
                else {
                    A = 0.0; B = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const double qh = n1_q(c, t - 1, i, cc);
                        A += qh * n1_ld(Ld + cc * 4 + a);
                        B += qh * n1_ld(Lo + cc * 4 + a);
                    }
                }
                own = n1_pick(qp, a) * A + n1_rest(qp, a) * B;
            }
            // ---- children, :139-191: internal tree children, leaf children, then the same node of position t+1
            double ch = 0.0;
            for (int j = c.ichild_begin[i]; j < c.ichild_begin[i + 1]; ++j) {
                const int ci = c.ichild[j];
                const double* Lco = tab + 16 + 32 * (c.node_of[ci] - 1);
                const double* Lcd = Lco + 16;
                // this lane handles the child's symbol ac = a: u = sum_cc q_prev(cc) Lo[cc][a], v likewise with Ld
                double u, v;
                if (t == 0) { u = n1_ld(Lco + a); v = n1_ld(Lcd + a); }
                else {
                    u = 0.0; v = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const double qpc = n1_q(c, t - 1, ci, cc);
                        u += qpc * n1_ld(Lco + cc * 4 + a);
                        v += qpc * n1_ld(Lcd + cc * 4 + a);
                    }
                }
                const double qc = n1_q(c, t, ci, a);
                const double wu = qc * u;                 // contribution of child symbol a to every OTHER own symbol
                double w[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) w[m] = __shfl_sync(c.qmask, wu, c.qb + m);
                ch += n1_rest(w, a) + qc * v;
            }
            double cobs = 0.0, cgap = 0.0;
            for (int kk = c.leaf_begin[i]; kk < c.leaf_begin[i + 1]; ++kk) {
                const int sp = c.leaf_species[kk];
                const int x = cw.obs(sp);
                const int y = t > 0 ? cwp.obs(sp) : 0;
                if (x < 4) {
                    const double* Ll = tab + 16 + 32 * (c.leaf_node[kk] - 1) + (a == x ? 16 : 0);
                    if (t == 0 || y < 4) cobs += n1_ld(Ll + y * 4 + x);                 // observed parents must match the row
                    else {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) cobs += n1_gq(c, t - 1, kk, cc) * n1_ld(Ll + cc * 4 + x);
                    }
                } else {
                    // unobserved leaf: every row of its table is pi[ctx] (CPF.java:146-152): the same value for every a
                    if (t == 0) {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac) cgap += n1_gq(c, 0, kk, ac) * n1_ld(tab + ac);
                    } else if (y < 4) {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac) cgap += n1_gq(c, t, kk, ac) * n1_ld(tab + y * 4 + ac);
                    } else {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) cgap += (n1_gq(c, t, kk, ac) * n1_gq(c, t - 1, kk, cc)) * n1_ld(tab + cc * 4 + ac);
                    }
                }
            }
            double nx = 0.0;
            if (has_next) {
                if (i == 0) {
#pragma unroll
                    for (int ac = 0; ac < 4; ++ac) nx += n1_q(c, t + 1, 0, ac) * n1_ld(tabn + a * 4 + ac);
                } else {
                    const double* Lno = tabn + 16 + 32 * (k - 1);
                    const double* Lnd = Lno + 16;
                    double qpn[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) qpn[m] = n1_q(c, t + 1, p, m);
#pragma unroll
                    for (int ac = 0; ac < 4; ++ac) {
                        // rows l = ap + 4a of the next node's table: diagonal for ap == ac, off-diagonal otherwise
                        const double rest = ac == 0 ? (qpn[1] + qpn[2]) + qpn[3] : ac == 1 ? (qpn[0] + qpn[2]) + qpn[3]
                                          : ac == 2 ? (qpn[0] + qpn[1]) + qpn[3] : (qpn[0] + qpn[1]) + qpn[2];
                        nx += n1_q(c, t + 1, i, ac) * (qpn[ac] * n1_ld(Lnd + a * 4 + ac) + rest * n1_ld(Lno + a * 4 + ac));
                    }
                }
            }
            const double s = (((own + ch) + cobs) + cgap) + nx;
            double qn, lq;
            n1_normalise(c, s, update, qn, lq);
            __syncwarp(c.qmask);                 // every lane of the quad has read the old q of this node's neighbours
            n1_q(c, t, i, a) = qn;
            __syncwarp(c.qmask);
            // entropy (:250-267), own expectation (:282-291 / :327-358) and the observed leaves' terms (:292-326)
            F += qn * ((lq - own) - cobs);

            // ---- unobserved leaf children of this node: update + free-energy terms
            for (int kk = c.leaf_begin[i]; kk < c.leaf_begin[i + 1]; ++kk) {
                const int sp = c.leaf_species[kk];
                if (cw.obs(sp) < 4) continue;
                const int y = t > 0 ? cwp.obs(sp) : 0;
                const double qs = ((n1_q(c, t, i, 0) + n1_q(c, t, i, 1)) + n1_q(c, t, i, 2)) + n1_q(c, t, i, 3);
                double ownl = 0.0, own_fe;
                if (t == 0) { ownl = qs * n1_ld(tab + a); own_fe = ownl; }
                else if (y < 4) {
                    // update: rows are NOT filtered by the observed parent (:129 TODO); free energy: they are (:327-358)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) ownl += qs * n1_ld(tab + cc * 4 + a);
                    own_fe = qs * n1_ld(tab + y * 4 + a);
                } else {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) ownl += (qs * n1_gq(c, t - 1, kk, cc)) * n1_ld(tab + cc * 4 + a);
                    own_fe = ownl;
                }
                double nl = 0.0;
                if (has_next) {
                    const int xn = cwn.obs(sp);
                    double qpn[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) qpn[m] = n1_q(c, t + 1, i, m);
                    if (xn < 4) {
                        const double* Lno = tabn + 16 + 32 * (c.leaf_node[kk] - 1);
                        nl = n1_pick(qpn, xn) * n1_ld(Lno + 16 + a * 4 + xn) + n1_rest(qpn, xn) * n1_ld(Lno + a * 4 + xn);
                    } else {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac)
#pragma unroll
                            for (int ap = 0; ap < 4; ++ap) nl += (n1_gq(c, t + 1, kk, ac) * qpn[ap]) * n1_ld(tabn + a * 4 + ac);
                    }
                }
                double ql, lql;
                n1_normalise(c, ownl + nl, update, ql, lql);
                __syncwarp(c.qmask);
                n1_gq(c, t, kk, a) = ql;
                __syncwarp(c.qmask);
                F += ql * (lql - own_fe);
            }
        }
    }
    return F;
}

__global__ void __launch_bounds__(256) score_o1_kernel(Net1Params p) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int nwb = p.nwb;
    double* s_q = smem;                                                  // [W*I*4][nwb]
    uint64_t* s_cb = (uint64_t*)(s_q + (size_t)p.W * p.I * 4 * nwb);     // [W][nwb]
    uint32_t* s_cg = (uint32_t*)(s_cb + (size_t)p.W * nwb);              // [W][nwb]
    int* s_prog = (int*)(s_cg + (size_t)p.W * nwb);
    // tables behind the program, 8-byte aligned
    double* s_tab = (double*)(s_prog + ((p.prog_len + 1) & ~1));
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    if (p.tab_in_smem)
        for (int i = tid; i < p.W * p.E1; i += T) s_tab[i] = p.tab[i];
    __syncthreads();

    Net1Ctx c;
    c.q = s_q; c.nwb = nwb; c.wq = tid >> 2;
    c.cb = s_cb; c.cg = s_cg;
    c.par_internal = s_prog;
    c.node_of = s_prog + p.I;
    c.leaf_begin = s_prog + 2 * p.I;
    c.leaf_node = s_prog + 3 * p.I + 1;
    c.leaf_species = s_prog + 3 * p.I + 1 + p.S;
    c.ichild_begin = s_prog + 3 * p.I + 1 + 2 * p.S;
    c.ichild = s_prog + 4 * p.I + 2 + 2 * p.S;
    c.W = p.W; c.I = p.I; c.S = p.S; c.E1 = p.E1;
    c.tab = p.tab_in_smem ? s_tab : p.tab;
    c.gq = p.gq; c.n_slots = p.n_slots; c.slot = (int64_t)blockIdx.x * nwb + c.wq;
    const int lane = tid & 31;
    c.qb = lane & ~3; c.a = lane & 3; c.qmask = 0xFu << c.qb;
    const int a = c.a;

    const long long n_items = (long long)p.n_windows * p.n_strands;
    bool need = true, active = true, upd = false, conv = false;
    int k = 0;
    long long item = -1;
    double old = 0.0;
    // a quad that never receives a window still walks the net (warp-uniform control flow): give it defined inputs
    for (int t = a; t < p.W; t += 4) { s_cb[(size_t)t * nwb + c.wq] = 0; s_cg[(size_t)t * nwb + c.wq] = 0; }

    for (;;) {
        if (need && active) {
            long long it = 0;
            if (a == 0) it = (long long)atomicAdd(&p.counters[0], 1ULL);
            item = __shfl_sync(c.qmask, it, c.qb);
            if (item >= n_items) active = false;
            else {
                int strand; int64_t c0;
                if (p.item_col0) { c0 = p.item_col0[item]; strand = p.item_strand ? p.item_strand[item] : 0; }
                else {
                    const int si = (int)(item / p.n_windows);
                    const int64_t w = item - (long long)si * p.n_windows;
                    strand = p.n_strands == 2 ? si : p.strand0;
                    const int al = find_alignment(p.win_off, p.n_aln, w);
                    c0 = p.col_off[al] + (w - p.win_off[al]);
                }
                for (int t = a; t < p.W; t += 4) {
                    const int64_t col = strand ? c0 + (p.W - 1 - t) : c0 + t;     // jstacs StrandModel.java:636-639
                    ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, col);
                    if (strand) cw = complement(cw);
                    cw.b &= p.S >= 32 ? ~0ULL : ((1ULL << (2 * p.S)) - 1);
                    s_cb[(size_t)t * nwb + c.wq] = cw.b; s_cg[(size_t)t * nwb + c.wq] = cw.g;
                }
                k = 0; conv = false; upd = false; need = false;
            }
            // uniform start, MeanFieldForBayesNet.java:52-69 (pass 0 rewrites it; neighbours are read before that)
            for (int j = a; j < p.W * p.I * 4; j += 4) s_q[(size_t)j * nwb + c.wq] = 0.25;
            for (int j = a; j < p.W * p.S * 4; j += 4) p.gq[(size_t)j * p.n_slots + c.slot] = 0.25;
            __syncwarp(c.qmask);
        }
        if (!__any_sync(0xffffffffu, active)) break;
        const double Fl = net1_pass(c, upd);
        const double f0 = __shfl_sync(c.qmask, Fl, c.qb), f1 = __shfl_sync(c.qmask, Fl, c.qb + 1);
        const double f2 = __shfl_sync(c.qmask, Fl, c.qb + 2), f3 = __shfl_sync(c.qmask, Fl, c.qb + 3);
        const double Fsum = ((f0 + f1) + f2) + f3;
        if (active) {
            bool done;
            if (!upd) { old = Fsum; upd = true; done = false; }               // F_0, MeanFieldForBayesNet.java:105
            else {
                if (fabs(old - Fsum) < p.tol) conv = true;                    // :204-207 (sticky)
                old = Fsum; ++k;
                done = !((k < p.max_sweeps && !conv) || k < p.min_sweeps);    // :107
            }
            if (done) {
                if (p.sink_qint) {
                    // entropy term sum_hidden sum_h q log q (MeanFieldForBayesNet.java:250-267), parameter independent
                    double ent = 0.0;
                    for (int j = a; j < p.W * p.I * 4; j += 4) {
                        const double v = s_q[(size_t)j * nwb + c.wq];
                        p.sink_qint[(size_t)j * p.n_windows + item] = v;
                        ent += v > 0.0 ? v * phy_log(v) : 0.0;
                    }
                    for (int t = 0; t < p.W; ++t) {
                        const ColWords cwt = n1_col(c, t);
                        for (int kk = 0; kk < p.S; ++kk) {
                            const double v = n1_gq(c, t, kk, a);
                            p.sink_qleaf[(size_t)((t * p.S + kk) * 4 + a) * p.n_windows + item] = v;
                            if (cwt.obs(c.leaf_species[kk]) >= 4) ent += v > 0.0 ? v * phy_log(v) : 0.0;
                        }
                    }
                    const double t0 = __shfl_sync(c.qmask, ent, c.qb), t1 = __shfl_sync(c.qmask, ent, c.qb + 1);
                    const double t2 = __shfl_sync(c.qmask, ent, c.qb + 2), t3 = __shfl_sync(c.qmask, ent, c.qb + 3);
                    if (a == 0) p.sink_ent[item] = ((t0 + t1) + t2) + t3;
                }
                if (a == 0) {
                    p.out[item] = -Fsum;                                      // PhyloBayesModel.java:259
                    if (p.sweeps) p.sweeps[item] = k;
                    atomicAdd(&p.counters[1], (unsigned long long)k);
                }
                need = true;
            }
        }
    }
}
