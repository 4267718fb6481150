// net1n.cuh -- order-1 phylogenetic Bayesian network, node-per-thread kernel (net1.cuh has the model, net1w.cuh the
// argument why a list schedule of the node updates reproduces the reference's Gauss-Seidel sweep).
//
// Mapping: ONE THREAD updates one hidden node with all four symbols in registers; a window is worked on by 4 node slots
// (the list schedule of model_host.hpp::node_program runs <= 4 independent nodes per step), a warp holds 8 windows
// (lane = 8 slot + window), a block a few warps that share the tables.  Compared with the 16-lanes-per-window kernel of
// net1w.cuh this removes every shuffle of the node update, loads each q vector once instead of once per symbol lane, and
// amortises the address arithmetic over four symbols -- the old kernel issued ~550 instructions per step of which a fifth
// were fp64.  What bounds this one is shared-memory bandwidth and the fp64 pipe (profiles/r02_c3_*).
//
// Per node update (t, i), reference loops algorithm/MeanFieldForBayesNet.java:108-202 (update) and :232-365 (free energy):
//   own(a)  = q_par(a) A(a) + (tot_par - q_par(a)) B(a)      A(a) = sum_c q_{t-1,i}(c) Ld[c][a], B(a) likewise with Lo
//   ch(a)   = sum_j [ q_j(a) A_j(a) + sum_{a' != a} q_j(a') B_j(a') ]      internal children j (old q, new context sums)
//   nx(a)   = sum_a' q_{t+1,i}(a') [ q_{t+1,par}(a') Ld'[a][a'] + (tot - q_{t+1,par}(a')) Lo'[a][a'] ]   (old values)
//   cobs(a) = sum over observed leaf children of log T_leaf[a, y][x]       (constant per window: kept from the set-up)
//   s = own + ch + cobs + nx;  q = exp(s) / z,  z = sum exp(s)             (no max shift, :192-201)
// The context sums A, B of a node are needed twice (by the node and by its tree parent), so the update of (t-1, i)
// produces them ("AB rows", two per node by position parity) right after it has the new q in registers; the tables of
// (t+1, i) it needs for that are the ones the nx term reads anyway.
// Free energy without a logarithm per node: the node's terms are sum_a q(a) (log q(a) - own(a) - cobs(a)), and
// log q(a) = s(a) - log z, so they equal sum_a q(a) (ch(a) + nx(a)) - log z (pass at the uniform start: -(own + cobs) and
// z = 4).  The log z of a thread's nodes are summed as ONE logarithm of the product of their mantissas plus ln 2 times the
// sum of their exponents (integer adds), one phy_log per thread and sweep instead of one per node.
// Windows with an unobserved leaf go to a device-side list and are scored by the net1w.cuh kernel afterwards (the bundled
// data and the headline configurations have no gaps).
//
// The N1N_HD parts (n1n_update) compile for the host as well: tests/emul runs them on the CPU without a GPU.
#pragma once
#include "fp64_math.cuh"

#include "n1n_math.cuh"

namespace phyfoo {

// One node update.  reg: this thread's window in the warp's region (AB and message rows), qreg: the same for the q rows
// (the same pointer when the whole region is in shared memory); tab: tables [W][I][34] (Lo | D = Ld - Lo | pad); exptab:
// 2^(j/16), 16 entries; r: the record (model_host.hpp::node_program); cobs: observed-leaf sums of this node; upd == false: the pass at the
// uniform start (exponents forced to 0).  All reads come before the row stores; the caller synchronises the warp afterwards.
// A slot without a node in this step (record flags 0) does nothing.
// qx, qpn: the (old) q of the same node and of its parent at the next position, rows r[2] and r[3] -- no update of the
// previous step can have written them, so the caller loads them a step ahead.
// Arithmetic (D rows and tables hold diagonal minus off-diagonal):
//   own(a) = q_par(a) Dm(a) + tot_par B(a)                                  [= q_par(a) A(a) + (tot_par - q_par(a)) B(a)]
//   nx(a)  = sum_c (qx(c) qpn(c)) D'[a][c] + tot_next sum_c qx(c) Lo'[a][c]
//   s = (own + cobs) + (ch + nx);  the node's free-energy terms use g = ch + nx (pass at the uniform start: -(own + cobs))
template <int MC>
N1N_HD void n1n_update(double* qreg, double* reg, const double* tab, const double* exptab, const int* r, const double (&cobs)[4],
                       const double (&qx)[4], const double (&qpn)[4], bool upd, N1nAcc& acc) {
    const int flags = r[7];
    if (!(flags & 1)) return;
    const bool last = (flags & 2) != 0;
    double qp[4], Dm[4], B[4], oc[4], cn[4];
    n1n_ld4(qreg + r[1], qp);
    n1n_ld4(reg + r[4], Dm);
    n1n_ld4(reg + r[4] + 32, B);
    const double totp = n1n_sum4(qp);
#pragma unroll
    for (int a = 0; a < 4; ++a) oc[a] = fma(qp[a], Dm[a], totp * B[a]) + cobs[a];
    // internal children: their messages (null entries read the all-zero row)
    n1n_ld4(reg + r[8], cn);
#pragma unroll
    for (int j = 1; j < MC; ++j) {
        double mj[4];
        n1n_ld4(reg + r[8 + j], mj);
#pragma unroll
        for (int a = 0; a < 4; ++a) cn[a] += mj[a];
    }
    // the same node at the next position (old values); its tables also serve the context sums produced below
    double T[32];
    {
        const double* tp = tab + r[6];
#pragma unroll
        for (int j = 0; j < 16; ++j) { const N1nD2 v = *reinterpret_cast<const N1nD2*>(tp + 2 * j); T[2 * j] = v.x; T[2 * j + 1] = v.y; }
    }
    if (!last) {
        double rr[4];
        const double totn = n1n_sum4(qpn);
#pragma unroll
        for (int c = 0; c < 4; ++c) rr[c] = qx[c] * qpn[c];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double u = qx[0] * T[a * 4], v = rr[0] * T[16 + a * 4];
#pragma unroll
            for (int c = 1; c < 4; ++c) { u = fma(qx[c], T[a * 4 + c], u); v = fma(rr[c], T[16 + a * 4 + c], v); }
            cn[a] += fma(totn, u, v);
        }
    }
    double e[4], g[4], qn[4], ex[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        ex[a] = upd ? oc[a] + cn[a] : 0.0;
        g[a] = upd ? cn[a] : -oc[a];
    }
    n1n_exp4_t16(ex, exptab, e);                                                    // MeanFieldForBayesNet.java:192, no max shift
    const double z = n1n_sum4(e);                                                   // :195-198
    // The reciprocal of z is a chain of dependent operations; what is linear in q is computed on the unnormalised
    // exponentials next to it and scaled afterwards: the node's free-energy terms, the context sums of the same node at
    // the next position (the last position produces those of position 0: one-hot, scale 1) and the message to the parent.
    double fu = e[0] * g[0];
    fu = fma(e[1], g[1], fu); fu = fma(e[2], g[2], fu); fu = fma(e[3], g[3], fu);
    double Dn[4], Bn[4], msg[4];
    {
        double ea[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ea[c] = last ? (c == 0 ? 1.0 : 0.0) : e[c];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double u = ea[0] * T[a], v = ea[0] * T[16 + a];
#pragma unroll
            for (int c = 1; c < 4; ++c) { u = fma(ea[c], T[c * 4 + a], u); v = fma(ea[c], T[16 + c * 4 + a], v); }
            Bn[a] = u; Dn[a] = v;
        }
    }
    // message to the tree parent at the next position, with the old q of (t+1, i) -- of (0, i) at the last position
    n1n_message(qx, Dn, Bn, msg);
    // z is a normal number well inside the exponent range unless the exponents were extreme: one test for the reciprocal
    // and for the (mantissa, exponent) accumulation of log z
    const int zhi = n1n_hi(z);
    const unsigned zef = (unsigned)zhi >> 20;
    const bool znormal = zef - 24u < 2000u;
    double rz = n1n_rcp_normal(z);
    if (znormal) { acc.E += (int)zef - 1023; acc.M *= n1n_with_hi(z, (zhi & 0x000fffff) | 0x3ff00000); }
    else {
        const N1nZR zr = n1n_z_rare(z);
        rz = zr.rz;
        acc.Fslow += zr.lz;
    }
    const double sc = last ? 1.0 : rz;
#pragma unroll
    for (int a = 0; a < 4; ++a) { qn[a] = e[a] * rz; Dn[a] *= sc; Bn[a] *= sc; msg[a] *= sc; }   // :199-201
    acc.Fq = fma(fu, rz, acc.Fq);
    if (flags & 4) n1n_message(qn, Dn, Bn, msg);        // a motif of a single position: the message carries the node's own new q
    n1n_st4(qreg + r[0], qn);
    n1n_st4(reg + r[4], Dn);
    n1n_st4(reg + r[4] + 32, Bn);
    n1n_st4(reg + r[5], msg);
}

// window set-up of the rows that do not come out of a node update: context sums of position 0 (the virtual predecessor is
// the one-hot (1, 0, 0, 0): row 0 of the tables) and the message of the uniform start to the parent at position 0
N1N_HD void n1n_setup_node(double* reg, const double* tab, int ab_row, int msg_row, int tab_row) {
    double D[4], B[4], msg[4];
    const double q0[4] = {0.25, 0.25, 0.25, 0.25};
#pragma unroll
    for (int a = 0; a < 4; ++a) { B[a] = tab[tab_row + a]; D[a] = tab[tab_row + 16 + a]; }
    n1n_message(q0, D, B, msg);
    n1n_st4(reg + ab_row, D);
    n1n_st4(reg + ab_row + 32, B);
    n1n_st4(reg + msg_row, msg);
}

}  // namespace phyfoo

#if defined(__CUDACC__)
// ----------------------------------------------------------------------------------------------------------------------
// device side
// ----------------------------------------------------------------------------------------------------------------------
struct Net1nParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, I, S;
    const int* prog; int prog_len;       // model_host.hpp::node_program
    const double* tab;                   // [W][I][34]
    const double* leaf_tab;              // [W][S][32]
    double* out; int32_t* sweeps;
    unsigned long long* counters;        // [0] work counter, [1] sweep sum
    int max_sweeps, min_sweeps; double tol;
    double* gc;                          // observed-leaf sums: one region of `region` doubles per warp of the grid
    double* gq;                          // q rows in global memory (QG variant): one region per warp of the grid
    int region, sregion, ab0, msg0, q0, onehot_row;
    double f0_const;                     // model_host.hpp::n1n_f0_const(): free energy of the uniform start minus the window's part
    int check_gaps;                      // 0: the data set has no unobserved symbol at all
    int* fb_count; int64_t* fb_col0; uint8_t* fb_strand; int64_t* fb_dst;   // windows with gaps, for the net1w.cuh kernel
    // M-step "prepare" (SINK instantiation): the windows are listed (first column, strand) and the converged q of the hidden
    // nodes goes to sink_qint[((t I + i) 4 + a) n_windows + item], sum q log q to sink_ent[item] (mstep1.cuh reads them)
    const int64_t* item_col0; const uint8_t* item_strand;
    double* sink_qint; double* sink_ent;
};

// QG: the q rows live in global memory (read through L1, written through to L2) and shared memory holds the AB and
// message rows only -- twice the warps per SM, i.e. two per scheduler
template <int MC, bool QG, bool SINK = false>
__global__ void __launch_bounds__(256, 1) score_o1n_kernel(Net1nParams p) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, T = blockDim.x, NW = T >> 5;
    const int W = p.W, I = p.I, S = p.S;
    const int TS = phyfoo::ModelHost::N1N_TAB_STRIDE;
    double* s_tab = smem;                                   // [W][I][34]
    double* s_exp = s_tab + (size_t)W * I * TS;             // [64]
    double* s_reg = s_exp + 64;                             // [NW][region] (QG: [NW][sregion])
    const int sreg = QG ? p.sregion : p.region;
    int* s_prog = (int*)(s_reg + (size_t)NW * sreg);
    for (int i = tid; i < W * I * TS; i += T) s_tab[i] = p.tab[i];
    for (int i = tid; i < 64; i += T) s_exp[i] = phyfoo::d_phy_exp_tab[i < 16 ? 4 * i : i];     // [0..15] = 2^(j/16) (n1n_exp4_t16)
    for (int i = tid; i < NW * sreg; i += T) s_reg[i] = 0.0;
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    double* const q_all = QG ? p.gq + (size_t)blockIdx.x * NW * p.region : s_reg;      // q rows of the block's warps
    if (QG) for (int i = tid; i < NW * p.region; i += T) q_all[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < NW * 8; i += T) q_all[(size_t)(i >> 3) * (QG ? p.region : sreg) + p.onehot_row + 2 * (i & 7)] = 1.0;
    __syncthreads();
    const int n_steps = s_prog[0], RL = s_prog[2];
    const int* leaf_begin = s_prog + 4;
    const int* leaf_species = leaf_begin + I + 1;
    const int* s_rec = s_prog + s_prog[3];

    const int lane = tid & 31, warp = tid >> 5;
    const int w = lane & 7, slot = lane >> 3;
    const unsigned full = 0xffffffffu;
    double* regw = s_reg + (size_t)warp * sreg;             // the warp's region
    double* reg = regw + 2 * w;                             // this thread's window
    double* qregw = q_all + (size_t)warp * (QG ? p.region : sreg);
    double* qreg = qregw + 2 * w;
    double* gcw = p.gc + (size_t)(blockIdx.x * NW + warp) * p.region;
    const double* gct = gcw + 2 * w;
    const long long n_items = (long long)p.n_windows * p.n_strands;
    const unsigned gap_mask = S >= 32 ? 0xffffffffu : (1u << S) - 1u;

    bool need = true, active = true, upd = false, conv = false;
    int k = 0;
    long long item = -1;
    double old = 0.0;

    for (;;) {
        // ---- windows that finished pull the next one; the whole warp sets a new window up
        unsigned nm = __ballot_sync(full, need && active) & 0xffu;
        while (nm) {
            const int wi = __ffs(nm) - 1;
            nm &= nm - 1;
            long long it = -1; int64_t c0 = 0; int strand = 0;
            double f0 = 0.0;
            for (;;) {
                long long v = 0;
                if (lane == 0) v = (long long)atomicAdd(&p.counters[0], 1ULL);
                it = __shfl_sync(full, v, 0);
                if (it >= n_items) { it = -1; break; }
                if (SINK) { c0 = p.item_col0[it]; strand = p.item_strand ? p.item_strand[it] : 0; }
                else {
                    const int si = (int)(it / p.n_windows);
                    const int64_t wdx = it - (long long)si * p.n_windows;
                    strand = p.n_strands == 2 ? si : p.strand0;
                    const int al = find_alignment_hint(p.win_off, p.n_aln, wdx, p.n_windows);
                    c0 = p.col_off[al] + (wdx - p.win_off[al]);
                }
                if (!p.check_gaps) break;
                unsigned any = 0;
                for (int t = lane; t < W; t += 32) {
                    const int64_t col = c0 + t;
                    for (int gpl = 0; gpl < (S + 31) / 32; ++gpl) any |= p.gaps[(size_t)gpl * p.n_cols + col] & gap_mask;
                }
                if (!__any_sync(full, any != 0)) break;
                if (lane == 0) {
                    const int idx = atomicAdd(p.fb_count, 1);
                    p.fb_col0[idx] = c0; p.fb_strand[idx] = (uint8_t)strand; p.fb_dst[idx] = it;
                }
            }
            if (it >= 0) {
                // observed-leaf sums of every node (constant over the sweeps), uniform start (MeanFieldForBayesNet.java:52-69):
                // one lane per (position, node), all four symbols.  Their total gives the free energy of the uniform start
                // (F_0, :105) in closed form, so the first pass over the net is already the first sweep.
                double csum = 0.0;
                for (int idx = lane; idx < W * I; idx += 32) {
                    const int t = idx / I, i = idx - t * I;
                    const int64_t col = strand ? c0 + (W - 1 - t) : c0 + t;       // jstacs StrandModel.java:636-639
                    ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, col);
                    ColWords cwp; cwp.b = 0; cwp.g = 0;
                    if (t > 0) cwp = load_column(p.bases, p.gaps, p.n_cols, p.nb, strand ? col + 1 : col - 1);
                    if (strand) { cw = complement(cw); cwp = complement(cwp); }
                    double c[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int kk = leaf_begin[i]; kk < leaf_begin[i + 1]; ++kk) {
                        const int sp = leaf_species[kk];
                        const int x = cw.obs(sp) & 3, y = t > 0 ? (cwp.obs(sp) & 3) : 0;
                        const double* lt = p.leaf_tab + ((size_t)(t * S + kk) * 32 + y * 4 + x);
                        const double off = __ldg(lt), diag = __ldg(lt + 16);
#pragma unroll
                        for (int a = 0; a < 4; ++a) c[a] += a == x ? diag : off;
                    }
                    const int o = p.q0 + idx * 32 + 2 * wi;
                    gcw[o] = c[0]; gcw[o + 1] = c[1]; gcw[o + 16] = c[2]; gcw[o + 17] = c[3];
                    qregw[o] = 0.25; qregw[o + 1] = 0.25; qregw[o + 16] = 0.25; qregw[o + 17] = 0.25;
                    csum += (c[0] + c[1]) + (c[2] + c[3]);
                }
#pragma unroll
                for (int sh = 16; sh >= 1; sh >>= 1) csum += __shfl_xor_sync(full, csum, sh);
                f0 = fma(-0.25, csum, p.f0_const);
                // context sums and messages of position 0
                for (int i = lane; i < I; i += 32)
                    phyfoo::n1n_setup_node(regw + 2 * wi, s_tab, p.ab0 + i * 64, p.msg0 + i * 32, i * TS);
            }
            if (w == wi) {
                if (it >= 0) { item = it; k = 0; conv = false; upd = true; old = f0; need = false; }     // F_0, :105
                else active = false;
            }
        }
        __syncwarp();
        if (!__any_sync(full, active)) break;

        // ---- one sweep (or the pass at the uniform start)
        phyfoo::N1nAcc acc; acc.Fq = 0.0; acc.M = 1.0; acc.E = 0; acc.Fslow = 0.0;
        // the record of the next step and its observed-leaf sums (L2) are fetched a step ahead
        constexpr int RLC = (8 + MC + 3) & ~3;
        const int* r = s_rec + slot * RL;
        int rr[RLC];
#pragma unroll
        for (int j = 0; j < RLC / 4; ++j) {
            const int4 v = reinterpret_cast<const int4*>(r)[j];
            rr[4 * j] = v.x; rr[4 * j + 1] = v.y; rr[4 * j + 2] = v.z; rr[4 * j + 3] = v.w;
        }
        double cn[4];
        {
            const double2 c01 = __ldcg(reinterpret_cast<const double2*>(gct + rr[0]));
            const double2 c23 = __ldcg(reinterpret_cast<const double2*>(gct + rr[0] + 16));
            cn[0] = c01.x; cn[1] = c01.y; cn[2] = c23.x; cn[3] = c23.y;
        }
        double qxn[4], qpnn[4];
        phyfoo::n1n_ld4(qreg + rr[2], qxn);
        phyfoo::n1n_ld4(qreg + rr[3], qpnn);
        for (int step = 0; step < n_steps; ++step) {
            const double cobs[4] = {cn[0], cn[1], cn[2], cn[3]};
            const double qx[4] = {qxn[0], qxn[1], qxn[2], qxn[3]}, qpn[4] = {qpnn[0], qpnn[1], qpnn[2], qpnn[3]};
            r += 4 * RL;
            int rn[RLC];
            const int4* rp = reinterpret_cast<const int4*>(step + 1 < n_steps ? r : r - 4 * RL);
#pragma unroll
            for (int j = 0; j < RLC / 4; ++j) {
                const int4 v = rp[j];
                rn[4 * j] = v.x; rn[4 * j + 1] = v.y; rn[4 * j + 2] = v.z; rn[4 * j + 3] = v.w;
            }
            {
                const double2 c01 = __ldcg(reinterpret_cast<const double2*>(gct + rn[0]));
                const double2 c23 = __ldcg(reinterpret_cast<const double2*>(gct + rn[0] + 16));
                cn[0] = c01.x; cn[1] = c01.y; cn[2] = c23.x; cn[3] = c23.y;
            }
            phyfoo::n1n_ld4(qreg + rn[2], qxn);
            phyfoo::n1n_ld4(qreg + rn[3], qpnn);
            phyfoo::n1n_update<MC>(qreg, reg, s_tab, s_exp, rr, cobs, qx, qpn, upd, acc);
            if ((step & 63) == 63) phyfoo::n1n_fold(acc);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < RLC; ++j) rr[j] = rn[j];
        }
        double F = phyfoo::n1n_free_energy(acc);
        F += __shfl_xor_sync(full, F, 8);
        F += __shfl_xor_sync(full, F, 16);
        bool fin = false;
        if (active) {
            if (fabs(old - F) < p.tol) conv = true;                               // MeanFieldForBayesNet.java:204-207 (sticky)
            old = F; ++k;
            bool done = !((k < p.max_sweeps && !conv) || k < p.min_sweeps);       // :107
            if (done) {
                if (slot == 0) {
                    p.out[item] = -F;                                             // PhyloBayesModel.java:259
                    if (p.sweeps) p.sweeps[item] = k;
                    atomicAdd(&p.counters[1], (unsigned long long)k);
                }
                need = true;
                fin = true;
            }
        }
        if (SINK && __any_sync(full, fin)) {
            // the four slot lanes of a finished window copy its q rows out (node idx = t I + i by lane slot) and add up q log q
            double ent = 0.0;
            if (fin) {
                const size_t n = (size_t)p.n_windows;
                for (int idx = slot; idx < W * I; idx += 4) {
                    double qv[4];
                    phyfoo::n1n_ld4(qreg + p.q0 + idx * 32, qv);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        p.sink_qint[(size_t)(idx * 4 + a) * n + (size_t)item] = qv[a];
                        ent += qv[a] > 0.0 ? qv[a] * phyfoo::phy_log(qv[a]) : 0.0;
                    }
                }
            }
            ent += __shfl_xor_sync(full, ent, 8);
            ent += __shfl_xor_sync(full, ent, 16);
            if (fin && slot == 0) p.sink_ent[item] = ent;
        }
    }
}
#endif  // __CUDACC__
