// net1w.cuh -- wavefront version of the order-1 network kernel (net1.cuh has the model and the arithmetic).
//
// The reference updates the W*I hidden internal nodes of a window one after the other (Gauss-Seidel in global node
// order).  What a node update reads is: NEW values of its tree parent, of the same node at the previous position and of
// its children at the previous position; OLD values of its children and of the same node / its parent at the next
// position.  All "new" reads are of nodes earlier in the global order, all "old" reads of later ones, and none of the
// rest of the net is read.  Hence any schedule that runs a node after the three kinds of nodes it needs new gives
// bit-for-bit the reference's sweep -- e.g. step = depth + 2 * position -- and nodes of one step are independent.
// The host list-schedules the W*I nodes into steps of at most four (model_host.hpp::wave_schedule); a window is
// processed by 16 lanes = 4 node slots x 4 symbols, two windows per warp, so the dependent chain per sweep is
// ~W*I/4 + depth steps instead of W*I node updates and a warp scheduler sees 4x the warps of the quad kernel.
//
// Node code is uniform across the slots of a warp: the root is a node whose tree parent is a virtual one-hot q with
// duplicated tables, child and leaf lists are padded with null entries (zero tables), the last position's "next node"
// reads an all-zero position.  Only gap leaves branch.
// Table layout per position (E1w = 64 N logs): node k block [Lo | Ld | LoT | LdT] at 64 k (see net1.cuh); root block
// [Lpi | Lpi | LpiT | LpiT]; W + 1 positions, the last all zero; one more zero block at the very end for null entries.
#pragma once

struct Net1wParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, I, S, E1, prog_len;
    const int* prog;         // par_internal[I] node_of[I] leaf_node[S] leaf_species[S] | MC ML n_steps | ich[I][MC] lf[I][ML] sched[n_steps][4]
    int MC, ML, n_steps;
    const double* tab;       // [W + 1][E1] + 64 zeros
    double* out; int32_t* sweeps;
    double* out_last;        // nullable: free energy of the LAST tree alone (background ranges: calcFreeEnergy(order, order))
    unsigned long long* counters;
    int max_sweeps, min_sweeps; double tol;
    double* gq; int64_t n_slots;
    double* gc;              // [n_slots][W * I * 4]: the observed-leaf sums of a window without gaps (they do not change from
                             // sweep to sweep; computed in the pass at the uniform start, read back afterwards)
    int nwb;
    const int64_t* item_col0; const uint8_t* item_strand;
    double* sink_qint; double* sink_qleaf; double* sink_ent;
    const int* n_items_dev;           // nullable: length of the item list lives on the device (windows with gaps listed by net1n.cuh)
    const int64_t* item_dst;          // nullable: out / sweeps / sink index of listed item (default: the item's own index)
    int64_t sink_stride;              // windows of the FE set the sinks belong to (0: n_windows)
    unsigned long long* counter_next; // work counter (counters[0] unless the launch continues another kernel's sweep sum)
    // background ranges of a single column (order 1, W = 2; models/PhyloBackground.java:260-283 with a range shorter than
    // order + 1): the second tree of the net keeps the observation of an EARLIER call (MeanFieldForBayesNet.java:84), i.e. an
    // arbitrary column item_col1[item], and the stop rule spans the first tree only (the sequence has one position)
    const int64_t* item_col1;         // nullable: column observed by tree 1 of listed item (default: the column after col0)
    int stop_first;                   // 1: stop rule on the free energy of the first tree alone (needs W = 2)
};

__global__ void __launch_bounds__(256, 3) score_o1w_kernel(Net1wParams p) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int nwb = p.nwb;
    // q rows are padded by one 32-byte slot: the four node slots of a window read four different rows at once, and with a
    // row stride of a multiple of 128 bytes they would all hit the same banks
    const int qs = 4 * nwb + 4;
    double* s_q = smem;                                                  // [W*I][nwb (+1)][4]
    double* s_virt = s_q + (size_t)p.W * p.I * qs;                       // [4] one-hot
    int* s_prog = (int*)(s_virt + 4);
    // leaf codes of the window: observation of (position t, leaf slot kk) in the low nibble, of (t-1, kk) in the high
    // nibble (0 at t = 0), 4 = unobserved; [W*S][nwb] bytes behind the program
    uint8_t* s_lc = (uint8_t*)(s_prog + p.prog_len);                     // [W][S + 1][nwb], row S stays 0
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    if (tid < 4) s_virt[tid] = tid == 0 ? 1.0 : 0.0;
    __syncthreads();
    const int* leaf_species = s_prog + 2 * p.I + p.S;
    const int* ich = s_prog + 2 * p.I + 2 * p.S;
    const int* lf = ich + p.I * p.MC;
    const int* sched = lf + p.I * p.ML;
    // node records: everything a node update needs besides t, as table / row offsets with null entries already resolved
    // (null child: the node itself with weight 0; null leaf: leaf row S (code 0) and the zero table block N)
    const int R = 2 + 3 * p.MC + 2 * p.ML;
    const int* rec = sched + 4 * p.n_steps;

    const int lane = tid & 31;
    const int wq = tid >> 4;                 // window slot in the block
    const int l16 = tid & 15;                // lane within the window
    const int slot = l16 >> 2, a = l16 & 3;
    const int qb = lane & ~3;                // first lane of this quad
    const int wb = lane & ~15;               // first lane of this window
    const unsigned qmask = 0xFu << qb;
    const unsigned full = 0xffffffffu;
    double* qbase = s_q + 4 * wq;
    const int64_t gslot = (int64_t)blockIdx.x * nwb + wq;
    auto GQ = [&](int t, int kk, int s) -> double& { return p.gq[(size_t)((t * p.S + kk) * 4 + s) * p.n_slots + gslot]; };

    const long long n_items = p.n_items_dev ? (long long)*p.n_items_dev : (long long)p.n_windows * p.n_strands;
    bool need = true, active = true, upd = false, conv = false, nogap = false;
    int k = 0;
    long long item = -1;
    double old = 0.0;
    double* const gcw = p.gc + (size_t)gslot * (p.W * p.I * 4);
    for (int j = l16; j < p.W * (p.S + 1); j += 16) s_lc[j * nwb + wq] = 0;

    for (;;) {
        if (need && active) {
            long long it = 0;
            if (l16 == 0) it = (long long)atomicAdd(p.counter_next, 1ULL);
            item = __shfl_sync(0xFFFFu << wb, it, wb);
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
                unsigned any_gap_bits = 0;
                for (int t = l16; t < p.W; t += 16) {
                    int64_t col = strand ? c0 + (p.W - 1 - t) : c0 + t;           // jstacs StrandModel.java:636-639
                    int64_t colp = strand ? col + 1 : col - 1;
                    if (p.item_col1 && t == 1) { col = p.item_col1[item]; colp = c0; }
                    ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, col);
                    ColWords cwp; cwp.b = 0; cwp.g = 0;
                    if (t > 0) cwp = load_column(p.bases, p.gaps, p.n_cols, p.nb, colp);
                    if (strand) { cw = complement(cw); cwp = complement(cwp); }
                    any_gap_bits |= cw.g;
                    for (int kk = 0; kk < p.S; ++kk) {
                        const int sp = leaf_species[kk];
                        s_lc[(t * (p.S + 1) + kk) * nwb + wq] = (uint8_t)(cw.obs(sp) | ((t > 0 ? cwp.obs(sp) : 0) << 4));
                    }
                }
                nogap = !__any_sync(0xFFFFu << wb, (any_gap_bits & (p.S >= 32 ? 0xffffffffu : (1u << p.S) - 1u)) != 0);
                k = 0; conv = false; upd = false; need = false;
            }
            // uniform start, MeanFieldForBayesNet.java:52-69
            for (int j = l16; j < p.W * p.I * 4; j += 16) qbase[(j >> 2) * qs + (j & 3)] = 0.25;
            for (int j = l16; j < p.W * p.S * 4; j += 16) p.gq[(size_t)j * p.n_slots + gslot] = 0.25;
            __syncwarp(0xFFFFu << wb);
        }
        if (!__any_sync(full, active)) break;

        // ---- one sweep (or the pass at the uniform start), in wavefront steps
        double F = 0.0, Flast = 0.0;          // Flast: terms of the last tree only
        for (int step = 0; step < p.n_steps; ++step) {
            const int code = sched[step * 4 + slot];
            const bool live = code >= 0;
            const int t = live ? code >> 16 : 0, i = live ? code & 0xffff : 0;
            const double* tab = p.tab + (size_t)t * p.E1;
            const double* tabn = tab + p.E1;
            const bool has_next = t + 1 < p.W;
            double* qcur = qbase + t * p.I * qs;
            const double* qprv = t ? qcur - p.I * qs : s_virt;
            const int sprv = t ? qs : 0;
            const double* qnxt = has_next ? qcur + p.I * qs : qcur;
            const uint8_t* lct = s_lc + t * (p.S + 1) * nwb + wq;             // leaf codes of this position
            const int* r = rec + i * R;
            const int par = r[0], nb = r[1];
            // own rows (root: virtual one-hot parent, duplicated tables)
            const N1Vec4 qh = n1_q4(qprv + i * sprv);
            const double A = n1_dot(qh, n1_ld4(tab + nb + 48 + 4 * a));
            const double B = n1_dot(qh, n1_ld4(tab + nb + 32 + 4 * a));
            const double* qpp = par >= 0 ? qcur + par * qs : s_virt;
            const double tot = n1_sum(n1_q4(qpp));
            const double qpa = qpp[a];
            const double own = qpa * A + (tot - qpa) * B;
            // internal children (null entries: the node itself with weight 0), one exchange for all of them
            double ch = 0.0, wu = 0.0;
            for (int j = 0; j < p.MC; ++j) {
                const int ci = r[2 + 3 * j], cnb = r[3 + 3 * j];
                const N1Vec4 qpc = n1_q4(qprv + ci * sprv);
                const double u = n1_dot(qpc, n1_ld4(tab + cnb + 32 + 4 * a));
                const double v = n1_dot(qpc, n1_ld4(tab + cnb + 48 + 4 * a));
                const double qc = r[4 + 3 * j] ? qcur[ci * qs + a] : 0.0;
                wu = fma(qc, u, wu);
                ch = fma(qc, v, ch);
            }
            {
                const double w0 = __shfl_sync(full, wu, qb), w1 = __shfl_sync(full, wu, qb + 1);
                const double w2 = __shfl_sync(full, wu, qb + 2), w3 = __shfl_sync(full, wu, qb + 3);
                ch += (((w0 + w1) + w2) + w3) - wu;
            }
            // leaf children (null entries: code 0 and the zero table block).  In a window without gaps the sum over the
            // observed leaves is the same in every sweep: kept from the pass at the uniform start
            double cobs = 0.0, cgap = 0.0;
            const bool cached = nogap && upd;
            if (cached) cobs = gcw[(t * p.I + i) * 4 + a];
            else for (int j = 0; j < p.ML; ++j) {
                const int kk = r[2 + 3 * p.MC + 2 * j], lnb = r[3 + 3 * p.MC + 2 * j];
                const int lc = lct[kk * nwb];
                const int x = lc & 15, y = lc >> 4;
                if (x < 4) {
                    const double* Ll = tab + lnb + (a == x ? 16 : 0);
                    if (y < 4) cobs += __ldg(Ll + y * 4 + x);
                    else {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) cobs += GQ(t - 1, kk, cc) * __ldg(Ll + cc * 4 + x);
                    }
                } else {
                    if (y < 4) {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac) cgap += GQ(t, kk, ac) * __ldg(tab + y * 4 + ac);
                    } else {
#pragma unroll
                        for (int ac = 0; ac < 4; ++ac)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) cgap += (GQ(t, kk, ac) * GQ(t - 1, kk, cc)) * __ldg(tab + cc * 4 + ac);
                    }
                }
            }
            if (nogap && !upd && live) gcw[(t * p.I + i) * 4 + a] = cobs;
            // the same node at the next position (all-zero tables behind the last position)
            double nx = 0.0;
            {
                const N1Vec4 qx = n1_q4(qnxt + i * qs);
                const double* qpn_p = par >= 0 ? qnxt + par * qs : s_virt;
                const N1Vec4 qpn = n1_q4(qpn_p);
                const N1Vec4 Lnd = n1_ld4(tabn + nb + 16 + 4 * a), Lno = n1_ld4(tabn + nb + 4 * a);
                const double totn = n1_sum(qpn);
#pragma unroll
                for (int ac = 0; ac < 4; ++ac)
                    nx = fma(qx.v[ac], fma(qpn.v[ac], Lnd.v[ac], (totn - qpn.v[ac]) * Lno.v[ac]), nx);
            }
            const double s = (((own + ch) + cobs) + cgap) + nx;
            double qn, lq;
            n1_normalise(full, qb, s, upd, qn, lq);
            __syncwarp();
            if (live) qcur[i * qs + a] = qn;
            __syncwarp();
            if (live) { const double f = qn * ((lq - own) - cobs); F += f; if (!has_next) Flast += f; }

            // unobserved leaf children: update + free-energy terms (the only divergent part)
            if (live && !nogap) {
                for (int j = 0; j < p.ML; ++j) {
                    const int kk = r[2 + 3 * p.MC + 2 * j], lnb = r[3 + 3 * p.MC + 2 * j];
                    const int lc = lct[kk * nwb];
                    if ((lc & 15) < 4) continue;                  // observed (null entries carry code 0)
                    const int y = lc >> 4;
                    const double qsum = n1_sum(n1_q4(qcur + i * qs));
                    double ownl = 0.0, own_fe;
                    if (t == 0) { ownl = qsum * __ldg(tab + a); own_fe = ownl; }
                    else if (y < 4) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) ownl += qsum * __ldg(tab + cc * 4 + a);
                        own_fe = qsum * __ldg(tab + y * 4 + a);
                    } else {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) ownl += (qsum * GQ(t - 1, kk, cc)) * __ldg(tab + cc * 4 + a);
                        own_fe = ownl;
                    }
                    double nl = 0.0;
                    if (has_next) {
                        const int xn = lct[(p.S + 1 + kk) * nwb] & 15;
                        const N1Vec4 qpn = n1_q4(qnxt + i * qs);
                        if (xn < 4) {
                            const double* Lno = tabn + lnb;
                            const double qd = xn == 0 ? qpn.v[0] : xn == 1 ? qpn.v[1] : xn == 2 ? qpn.v[2] : qpn.v[3];
                            nl = qd * __ldg(Lno + 16 + a * 4 + xn) + (n1_sum(qpn) - qd) * __ldg(Lno + a * 4 + xn);
                        } else {
                            const double totn = n1_sum(qpn);
#pragma unroll
                            for (int ac = 0; ac < 4; ++ac) nl += (GQ(t + 1, kk, ac) * totn) * __ldg(tabn + a * 4 + ac);
                        }
                    }
                    double ql, lql;
                    n1_normalise(qmask, qb, ownl + nl, upd, ql, lql);
                    __syncwarp(qmask);
                    GQ(t, kk, a) = ql;
                    __syncwarp(qmask);
                    { const double f = ql * (lql - own_fe); F += f; if (!has_next) Flast += f; }
                }
            }
        }
        // free energy of the window: sum over its 16 lanes (fixed butterfly)
        F += __shfl_xor_sync(full, F, 1);
        F += __shfl_xor_sync(full, F, 2);
        F += __shfl_xor_sync(full, F, 4);
        F += __shfl_xor_sync(full, F, 8);
        const double Fsum = F;
        if (p.out_last || p.stop_first) {
            Flast += __shfl_xor_sync(full, Flast, 1);
            Flast += __shfl_xor_sync(full, Flast, 2);
            Flast += __shfl_xor_sync(full, Flast, 4);
            Flast += __shfl_xor_sync(full, Flast, 8);
        }
        if (active) {
            bool done;
            const double Fstop = p.stop_first ? Fsum - Flast : Fsum;          // calcFreeEnergy() spans the sequence's positions
            if (!upd) { old = Fstop; upd = true; done = false; }              // F_0, MeanFieldForBayesNet.java:105
            else {
                if (fabs(old - Fstop) < p.tol) conv = true;                   // :204-207 (sticky)
                old = Fstop; ++k;
                done = !((k < p.max_sweeps && !conv) || k < p.min_sweeps);    // :107
            }
            if (done) {
                if (p.sink_qint) {
                    double ent = 0.0;
                    const size_t sn = p.sink_stride ? (size_t)p.sink_stride : (size_t)p.n_windows;
                    const size_t su = p.item_dst ? (size_t)p.item_dst[item] : (size_t)item;
                    for (int j = l16; j < p.W * p.I * 4; j += 16) {
                        const double v = qbase[(j >> 2) * qs + (j & 3)];
                        p.sink_qint[(size_t)j * sn + su] = v;
                        ent += v > 0.0 ? v * phy_log(v) : 0.0;
                    }
                    for (int j = l16; j < p.W * p.S * 4; j += 16) {
                        const int t = j / (p.S * 4), kk = (j >> 2) % p.S;
                        const double v = p.gq[(size_t)j * p.n_slots + gslot];
                        p.sink_qleaf[(size_t)j * sn + su] = v;
                        if ((s_lc[(t * (p.S + 1) + kk) * nwb + wq] & 15) >= 4) ent += v > 0.0 ? v * phy_log(v) : 0.0;
                    }
                    const unsigned wm = 0xFFFFu << wb;
                    ent += __shfl_xor_sync(wm, ent, 1); ent += __shfl_xor_sync(wm, ent, 2);
                    ent += __shfl_xor_sync(wm, ent, 4); ent += __shfl_xor_sync(wm, ent, 8);
                    if (l16 == 0) p.sink_ent[su] = ent;
                }
                if (l16 == 0) {
                    const long long oi = p.item_dst ? p.item_dst[item] : item;
                    if (p.out_last) p.out_last[oi] = -Flast;
                    p.out[oi] = -Fsum;                                        // PhyloBayesModel.java:259
                    if (p.sweeps) p.sweeps[oi] = k;
                    atomicAdd(&p.counters[1], (unsigned long long)k);
                }
                need = true;
            }
        }
    }
}
