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
// Table layout per position t (E1v = 32 + 64 (N-1) doubles, logs; model_host.hpp::tables1v): the reference's table of
// node k >= 1 is [l][b], l = a_treeparent + 4 ctx (bayesNet/CPF.java:56-59,238-256: parent 0 is the least significant
// digit).  It has F81 structure (one value for every a != b), so it is stored as off-diagonal values Lo[ctx][b] and
// diagonal values Ld[ctx][b], each also transposed ([b][ctx]) so that the four values a lane needs are one 32-byte
// load: root block [Lpi | LpiT] at 0, node block [Lo | Ld | LoT | LdT] at 32 + 64 (k-1).  Rows of contexts that do not
// exist (t = 0 has one context) are 0 and the "previous position" of t = 0 is a virtual one-hot q = (1,0,0,0); one
// all-zero position is appended so that the last position's "next node" terms vanish without a branch.  The
// "independent" table of an unobserved leaf (CPF.java:146-152) has pi[ctx][b] in every row = the root table.
#pragma once

struct Net1Params {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, I, S, E1, prog_len;
    const int* prog;
    const double* tab;       // [W + 1][E1] (last position all zero)
    double* out; int32_t* sweeps;
    unsigned long long* counters;
    int max_sweeps, min_sweeps; double tol;
    double* gq;              // gap-leaf q: [(t*S + leaf slot)*4 + a][n_slots]
    int64_t n_slots;
    int nwb;                 // windows per block
    const int64_t* item_col0; const uint8_t* item_strand;   // nullable: explicit window list
    // M-step "prepare": converged q of every listed window, entry-major across the n_windows listed windows:
    // sink_qint[((t*I + i)*4 + a) * n + item], sink_qleaf[((t*S + leaf slot)*4 + a) * n + item], sink_ent[item]
    double* sink_qint; double* sink_qleaf; double* sink_ent;
};

struct Net1Ctx {
    double* q;                // this quad's window: q(t, i, a) at q[(t*I + i) * qs + a], qs = 4 * nwb
    const double* virt;       // block-shared one-hot (1,0,0,0): the "previous position" of t = 0
    int qs;
    const uint64_t* cb; const uint32_t* cg; int nwb; int wq;   // column words [t][nwb]
    const int* par_internal; const int* node_of; const int* leaf_begin; const int* leaf_node; const int* leaf_species;
    const int* ichild_begin; const int* ichild;
    int W, I, S, E1;
    const double* tab;
    double* gq; int64_t n_slots; int64_t slot;
    unsigned qmask; int qb, a;
};

struct __align__(32) N1Vec4 { double v[4]; };
// 32 bytes of a table in one request (LDG.E.256 on sm_100a), read-only path
__device__ __forceinline__ N1Vec4 n1_ld4(const double* p) {
    N1Vec4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
// the four q of a node from shared memory (two 16-byte loads)
__device__ __forceinline__ N1Vec4 n1_q4(const double* p) {
    N1Vec4 r;
    const double2 x = *reinterpret_cast<const double2*>(p), y = *reinterpret_cast<const double2*>(p + 2);
    r.v[0] = x.x; r.v[1] = x.y; r.v[2] = y.x; r.v[3] = y.y;
    return r;
}
__device__ __forceinline__ double n1_dot(const N1Vec4& x, const N1Vec4& y) {
    double r = x.v[0] * y.v[0];
    r = fma(x.v[1], y.v[1], r); r = fma(x.v[2], y.v[2], r); r = fma(x.v[3], y.v[3], r);
    return r;
}
__device__ __forceinline__ double n1_sum(const N1Vec4& x) { return ((x.v[0] + x.v[1]) + x.v[2]) + x.v[3]; }
__device__ __forceinline__ double& n1_gq(const Net1Ctx& c, int t, int kk, int a) { return c.gq[(size_t)((t * c.S + kk) * 4 + a) * c.n_slots + c.slot]; }
__device__ __forceinline__ ColWords n1_col(const Net1Ctx& c, int t) {
    ColWords w; w.b = c.cb[t * c.nwb + c.wq]; w.g = c.cg[t * c.nwb + c.wq]; return w;
}

// exp of the four exponents of a node -> normalised q of this lane and log q; MeanFieldForBayesNet.java:192-201
__device__ __forceinline__ void n1_normalise(unsigned mask, int qb, double s, bool update, double& qn, double& lq) {
    const double ex = update ? s : 0.0;
    const double e = phy_exp(ex);                                           // :192, no max shift
    const double e0 = __shfl_sync(mask, e, qb), e1 = __shfl_sync(mask, e, qb + 1);
    const double e2 = __shfl_sync(mask, e, qb + 2), e3 = __shfl_sync(mask, e, qb + 3);
    const double z = ((e0 + e1) + e2) + e3;                                 // :195-198
    qn = e * __drcp_rn(z);                                                  // :199-201 (correctly rounded reciprocal, then one product)
    lq = ex - phy_log(z);
}

// One pass over the window's net; returns this lane's share of the free energy F = part1 - part2.
// Control flow is warp-uniform except inside the leaf loops (gap leaves), so the node-level shuffles use the full mask.
__device__ double net1_pass(const Net1Ctx& c, bool update) {
    const int a = c.a;
    const unsigned full = 0xffffffffu;
    double F = 0.0;
    for (int t = 0; t < c.W; ++t) {
        const double* tab = c.tab + (size_t)t * c.E1;
        const double* tabn = tab + c.E1;                       // position W is all zero
        const bool has_next = t + 1 < c.W;
        double* qcur = c.q + t * c.I * c.qs;
        const double* qprv = t ? qcur - c.I * c.qs : c.virt;   // previous position (virtual one-hot at t = 0)
        const int sprv = t ? c.qs : 0;
        const double* qnxt = has_next ? qcur + c.I * c.qs : qcur;   // multiplied by the zero table at the last position
        const ColWords cw = n1_col(c, t);
        ColWords cwp; cwp.b = 0; cwp.g = 0;
        ColWords cwn; cwn.b = 0; cwn.g = 0;
        if (t > 0) cwp = n1_col(c, t - 1);
        if (has_next) cwn = n1_col(c, t + 1);
        for (int i = 0; i < c.I; ++i) {
            const int p = c.par_internal[i];
            const int nb = 32 + 64 * (c.node_of[i] - 1);       // this node's table block
            // ---- own rows, :117-135 (all parents of an internal node are hidden)
            double own;
            const N1Vec4 qh = n1_q4(qprv + i * sprv);
            if (i == 0) own = n1_dot(qh, n1_ld4(tab + 16 + 4 * a));
            else {
                const double A = n1_dot(qh, n1_ld4(tab + nb + 48 + 4 * a));     // diagonal rows
                const double B = n1_dot(qh, n1_ld4(tab + nb + 32 + 4 * a));     // off-diagonal rows
                const double tot = n1_sum(n1_q4(qcur + p * c.qs));
                const double qpa = qcur[p * c.qs + a];
                own = qpa * A + (tot - qpa) * B;
            }
            // ---- children, :139-191: internal tree children (one exchange for all of them), leaves, the next position
            double ch = 0.0, wu = 0.0;
            for (int j = c.ichild_begin[i]; j < c.ichild_begin[i + 1]; ++j) {
                const int ci = c.ichild[j];
                const int cb = 32 + 64 * (c.node_of[ci] - 1);
                const N1Vec4 qpc = n1_q4(qprv + ci * sprv);
                // this lane handles the child's symbol ac = a
                const double u = n1_dot(qpc, n1_ld4(tab + cb + 32 + 4 * a));    // parent symbol != a
                const double v = n1_dot(qpc, n1_ld4(tab + cb + 48 + 4 * a));    // parent symbol == a
                const double qc = qcur[ci * c.qs + a];
                wu = fma(qc, u, wu);
                ch = fma(qc, v, ch);
            }
            {
                const double w0 = __shfl_sync(full, wu, c.qb), w1 = __shfl_sync(full, wu, c.qb + 1);
                const double w2 = __shfl_sync(full, wu, c.qb + 2), w3 = __shfl_sync(full, wu, c.qb + 3);
                ch += (((w0 + w1) + w2) + w3) - wu;            // child symbols other than a reach own symbol a off-diagonally
            }
            double cobs = 0.0, cgap = 0.0;
            for (int kk = c.leaf_begin[i]; kk < c.leaf_begin[i + 1]; ++kk) {
                const int sp = c.leaf_species[kk];
                const int x = cw.obs(sp);
                const int y = t > 0 ? cwp.obs(sp) : 0;
                if (x < 4) {
                    const double* Ll = tab + 32 + 64 * (c.leaf_node[kk] - 1) + (a == x ? 16 : 0);
                    if (y < 4) cobs += __ldg(Ll + y * 4 + x);                   // observed parents must match the row
                    else {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) cobs += n1_gq(c, t - 1, kk, cc) * __ldg(Ll + cc * 4 + x);
                    }
                } else {
                    // unobserved leaf: every row of its table is pi[ctx] (CPF.java:146-152): the same value for every a
                    if (y < 4) {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac) cgap += n1_gq(c, t, kk, ac) * __ldg(tab + y * 4 + ac);
                    } else {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) cgap += (n1_gq(c, t, kk, ac) * n1_gq(c, t - 1, kk, cc)) * __ldg(tab + cc * 4 + ac);
                    }
                }
            }
            double nx;
            {
                const N1Vec4 qx = n1_q4(qnxt + i * c.qs);
                if (i == 0) nx = n1_dot(qx, n1_ld4(tabn + 4 * a));
                else {
                    const N1Vec4 qpn = n1_q4(qnxt + p * c.qs);
                    const N1Vec4 Lnd = n1_ld4(tabn + nb + 16 + 4 * a), Lno = n1_ld4(tabn + nb + 4 * a);
                    const double totn = n1_sum(qpn);
                    nx = 0.0;
#pragma unroll
                    for (int ac = 0; ac < 4; ++ac)   // rows l = ap + 4a of the next node's table: diagonal for ap == ac
                        nx = fma(qx.v[ac], fma(qpn.v[ac], Lnd.v[ac], (totn - qpn.v[ac]) * Lno.v[ac]), nx);
                }
            }
            const double s = (((own + ch) + cobs) + cgap) + nx;
            double qn, lq;
            n1_normalise(full, c.qb, s, update, qn, lq);
            __syncwarp();                        // every lane has read the old q of this node's neighbours
            qcur[i * c.qs + a] = qn;
            __syncwarp();
            // entropy (:250-267), own expectation (:282-291 / :327-358) and the observed leaves' terms (:292-326)
            F += qn * ((lq - own) - cobs);

            // ---- unobserved leaf children of this node: update + free-energy terms
            for (int kk = c.leaf_begin[i]; kk < c.leaf_begin[i + 1]; ++kk) {
                const int sp = c.leaf_species[kk];
                if (cw.obs(sp) < 4) continue;
                const int y = t > 0 ? cwp.obs(sp) : 0;
                const double qs = n1_sum(n1_q4(qcur + i * c.qs));
                double ownl = 0.0, own_fe;
                if (t == 0) { ownl = qs * __ldg(tab + a); own_fe = ownl; }
                else if (y < 4) {
                    // update: rows are NOT filtered by the observed parent (:129 TODO); free energy: they are (:327-358)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) ownl += qs * __ldg(tab + cc * 4 + a);
                    own_fe = qs * __ldg(tab + y * 4 + a);
                } else {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) ownl += (qs * n1_gq(c, t - 1, kk, cc)) * __ldg(tab + cc * 4 + a);
                    own_fe = ownl;
                }
                double nl = 0.0;
                if (has_next) {
                    const int xn = cwn.obs(sp);
                    const N1Vec4 qpn = n1_q4(qnxt + i * c.qs);
                    if (xn < 4) {
                        const double* Lno = tabn + 32 + 64 * (c.leaf_node[kk] - 1);
                        const double qd = xn == 0 ? qpn.v[0] : xn == 1 ? qpn.v[1] : xn == 2 ? qpn.v[2] : qpn.v[3];
                        nl = qd * __ldg(Lno + 16 + a * 4 + xn) + (n1_sum(qpn) - qd) * __ldg(Lno + a * 4 + xn);
                    } else {
                        const double totn = n1_sum(qpn);
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac) nl += (n1_gq(c, t + 1, kk, ac) * totn) * __ldg(tabn + a * 4 + ac);
                    }
                }
                double ql, lql;
                n1_normalise(c.qmask, c.qb, ownl + nl, update, ql, lql);
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
    double* s_q = smem;                                                  // [W*I][nwb][4]
    double* s_virt = s_q + (size_t)p.W * p.I * 4 * nwb;                  // [4]
    uint64_t* s_cb = (uint64_t*)(s_virt + 4);                            // [W][nwb]
    uint32_t* s_cg = (uint32_t*)(s_cb + (size_t)p.W * nwb);              // [W][nwb]
    int* s_prog = (int*)(s_cg + (size_t)p.W * nwb);
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    if (tid < 4) s_virt[tid] = tid == 0 ? 1.0 : 0.0;
    __syncthreads();

    Net1Ctx c;
    c.nwb = nwb; c.wq = tid >> 2; c.qs = 4 * nwb;
    c.q = s_q + 4 * c.wq; c.virt = s_virt;
    c.cb = s_cb; c.cg = s_cg;
    c.par_internal = s_prog;
    c.node_of = s_prog + p.I;
    c.leaf_begin = s_prog + 2 * p.I;
    c.leaf_node = s_prog + 3 * p.I + 1;
    c.leaf_species = s_prog + 3 * p.I + 1 + p.S;
    c.ichild_begin = s_prog + 3 * p.I + 1 + 2 * p.S;
    c.ichild = s_prog + 4 * p.I + 2 + 2 * p.S;
    c.W = p.W; c.I = p.I; c.S = p.S; c.E1 = p.E1;
    c.tab = p.tab;
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
    for (int t = a; t < p.W; t += 4) { s_cb[t * nwb + c.wq] = 0; s_cg[t * nwb + c.wq] = 0; }

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
                    const int al = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
                    c0 = p.col_off[al] + (w - p.win_off[al]);
                }
                for (int t = a; t < p.W; t += 4) {
                    const int64_t col = strand ? c0 + (p.W - 1 - t) : c0 + t;     // jstacs StrandModel.java:636-639
                    ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, col);
                    if (strand) cw = complement(cw);
                    s_cb[t * nwb + c.wq] = cw.b; s_cg[t * nwb + c.wq] = cw.g;
                }
                k = 0; conv = false; upd = false; need = false;
            }
            // uniform start, MeanFieldForBayesNet.java:52-69 (pass 0 rewrites it; neighbours are read before that)
            for (int j = 0; j < p.W * p.I; ++j) c.q[j * c.qs + a] = 0.25;
            for (int j = a; j < p.W * p.S * 4; j += 4) p.gq[(size_t)j * p.n_slots + c.slot] = 0.25;
            __syncwarp(c.qmask);
        }
        if (!__any_sync(0xffffffffu, active)) break;
        const double Fl = net1_pass(c, upd);
        const double f0 = __shfl_sync(0xffffffffu, Fl, c.qb), f1 = __shfl_sync(0xffffffffu, Fl, c.qb + 1);
        const double f2 = __shfl_sync(0xffffffffu, Fl, c.qb + 2), f3 = __shfl_sync(0xffffffffu, Fl, c.qb + 3);
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
                    for (int j = 0; j < p.W * p.I; ++j) {
                        const double v = c.q[j * c.qs + a];
                        p.sink_qint[(size_t)(j * 4 + a) * p.n_windows + item] = v;
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
