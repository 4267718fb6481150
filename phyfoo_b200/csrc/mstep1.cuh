// mstep1.cuh -- fixed-q free-energy objective of the ORDER-1 motif network (included by mstep.cuh).
//
// With q fixed, the parameters of position t enter only the CPFs of tree t, whose terms involve the q of trees t and
// t-1 (MeanFieldForBayesNet.java:269-364):  f(theta) = sum_t G_t(theta_t) - ENT,  G_t = sum_u w_u sum_{n in tree t} E_q[log CPF_n].
// A forward-difference gradient at position t needs dim + 1 parameter sets (dim = 4 at t = 0, 16 at t >= 1: four
// context rows, models/PhyloBayesModel.java:336-341), evaluated in one pass over the selected windows with shared q
// loads; the tables of all sets sit in shared memory (F81-structured: 16 + 32 (N-1) logs per set, see net1.cuh).
#pragma once

static const int FE1_KMAX = 17;

struct Fe1EvalParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col0; const uint8_t* strand; const double* weight;
    const double* q_int; const double* q_leaf;     // [((t*I+i)*4+a)*n + u], [((t*S+kk)*4+a)*n + u]
    int64_t n;
    int W, I, S, E1, prog_len;
    const int* prog;
    const double* tabs;        // [K][E1] parameter sets of position `pos`
    double* part;              // [gridDim.x][FE1_KMAX]
    int pos;
};

template <int K>
__global__ void __launch_bounds__(128) fe1_eval_kernel(Fe1EvalParams p) {
    extern __shared__ double smem[];
    __shared__ double sh[32];
    const int T = blockDim.x, tid = threadIdx.x;
    double* s_tab = smem;                                   // [K][E1]
    int* s_prog = (int*)(s_tab + (size_t)K * p.E1);
    for (int i = tid; i < K * p.E1; i += T) s_tab[i] = p.tabs[i];
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    __syncthreads();
    const int* par_internal = s_prog;
    const int* node_of = s_prog + p.I;
    const int* leaf_begin = s_prog + 2 * p.I;
    const int* leaf_node = s_prog + 3 * p.I + 1;
    const int* leaf_species = s_prog + 3 * p.I + 1 + p.S;
    const int t = p.pos;
    const size_t n = (size_t)p.n;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int64_t u = (int64_t)blockIdx.x * T + tid; u < p.n; u += (int64_t)gridDim.x * T) {
        const int strand = p.strand ? p.strand[u] : 0;
        const int64_t c = strand ? p.col0[u] + (p.W - 1 - t) : p.col0[u] + t;
        ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
        ColWords cwp; cwp.b = 0; cwp.g = 0;
        if (t > 0) cwp = load_column(p.bases, p.gaps, p.n_cols, p.nb, strand ? c + 1 : c - 1);
        if (strand) { cw = complement(cw); cwp = complement(cwp); }
        auto Q = [&](int tt, int i, int a) { return p.q_int[(size_t)((tt * p.I + i) * 4 + a) * n + u]; };
        auto GQ = [&](int tt, int kk, int a) { return p.q_leaf[(size_t)((tt * p.S + kk) * 4 + a) * n + u]; };
        double e[K];
#pragma unroll
        for (int k = 0; k < K; ++k) e[k] = 0.0;
        for (int i = 0; i < p.I; ++i) {
            double qi[4], qp[4], qh[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                qi[a] = Q(t, i, a);
                qh[a] = t > 0 ? Q(t - 1, i, a) : (a == 0 ? 1.0 : 0.0);     // t = 0: a single context row
                qp[a] = i > 0 ? Q(t, par_internal[i], a) : 0.0;
            }
            const int ncc = t > 0 ? 4 : 1;
            if (i == 0) {
                // root: sum_{c,a} q_prev(c) q(a) log pi[c][a]   (:282-291 at t = 0, :327-358 otherwise)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double* tb = s_tab + (size_t)k * p.E1;
                    double r = 0.0;
                    for (int cc = 0; cc < ncc; ++cc)
#pragma unroll
                        for (int a = 0; a < 4; ++a) r += (qh[cc] * qi[a]) * tb[cc * 4 + a];
                    e[k] += r;
                }
            } else {
                const int off = 16 + 32 * (node_of[i] - 1);
                double rest[4];
                // weight of the off-diagonal rows: the sum of the other three parent symbols (no cancellation)
                rest[0] = (qp[1] + qp[2]) + qp[3]; rest[1] = (qp[0] + qp[2]) + qp[3];
                rest[2] = (qp[0] + qp[1]) + qp[3]; rest[3] = (qp[0] + qp[1]) + qp[2];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double* lo = s_tab + (size_t)k * p.E1 + off;
                    const double* ld = lo + 16;
                    double r = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        double A = 0.0, B = 0.0;
                        for (int cc = 0; cc < ncc; ++cc) { A += qh[cc] * ld[cc * 4 + a]; B += qh[cc] * lo[cc * 4 + a]; }
                        r += qi[a] * (qp[a] * A + rest[a] * B);
                    }
                    e[k] += r;
                }
            }
            const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
            for (int kk = leaf_begin[i]; kk < leaf_begin[i + 1]; ++kk) {
                const int sp = leaf_species[kk];
                const int x = cw.obs(sp);
                const int y = t > 0 ? cwp.obs(sp) : 0;
                double wc[4];                                   // weight of context row cc
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wc[cc] = t == 0 ? (cc == 0 ? 1.0 : 0.0) : (y < 4 ? (cc == y ? 1.0 : 0.0) : GQ(t - 1, kk, cc));
                if (x < 4) {
                    // observed leaf: sum_a q_par(a) sum_cc w(cc) log T[a + 4cc][x]   (:292-326)
                    const int off = 16 + 32 * (leaf_node[kk] - 1);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double* lo = s_tab + (size_t)k * p.E1 + off;
                        const double* ld = lo + 16;
                        double r = 0.0;
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const double* tb = a == x ? ld : lo;
                            double L = 0.0;
                            for (int cc = 0; cc < ncc; ++cc) L += wc[cc] * tb[cc * 4 + x];
                            r += qi[a] * L;
                        }
                        e[k] += r;
                    }
                } else {
                    // unobserved leaf ("independent" table = pi[ctx]): sum_a q_leaf(a) qs sum_cc w(cc) log pi[cc][a]  (:327-358)
                    double ql[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) ql[a] = GQ(t, kk, a);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double* tb = s_tab + (size_t)k * p.E1;
                        double r = 0.0;
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            double L = 0.0;
                            for (int cc = 0; cc < ncc; ++cc) L += (qs * wc[cc]) * tb[cc * 4 + a];
                            r += ql[a] * L;
                        }
                        e[k] += r;
                    }
                }
            }
        }
        const double w = p.weight[u];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += w * e[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double s = block_sum(acc[k], sh);
        if (tid == 0) p.part[(size_t)blockIdx.x * FE1_KMAX + k] = s;
        __syncthreads();
    }
}

// Expected counts of the table entries of every position (sufficient statistics of the objective above): with q fixed,
// G_t(theta_t) = sum_e C[t][e] tab_theta,t[e] with the E1 entries of model_host.hpp::build_tables1_into -- the coefficients
// fe1_eval_kernel multiplies the entries with, summed over the selected windows.  Shared-memory atomics per block (blockIdx.y =
// position), block partials reduced in a fixed order by fe_count_reduce_kernel.
struct Fe1CountParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col0; const uint8_t* strand; const double* weight;
    const double* q_int; const double* q_leaf;
    int64_t n;
    int W, I, S, E1, prog_len;
    const int* prog;
    double* part;              // [W][gridDim.x][E1]
};

__global__ void __launch_bounds__(128) fe1_count_kernel(Fe1CountParams p) {
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int t = blockIdx.y;
    double* s_c = smem;                                     // [E1]
    int* s_prog = (int*)(s_c + p.E1);
    for (int i = tid; i < p.E1; i += T) s_c[i] = 0.0;
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    __syncthreads();
    const int* par_internal = s_prog;
    const int* node_of = s_prog + p.I;
    const int* leaf_begin = s_prog + 2 * p.I;
    const int* leaf_node = s_prog + 3 * p.I + 1;
    const int* leaf_species = s_prog + 3 * p.I + 1 + p.S;
    const size_t n = (size_t)p.n;
    const int ncc = t > 0 ? 4 : 1;
    for (int64_t u = (int64_t)blockIdx.x * T + tid; u < p.n; u += (int64_t)gridDim.x * T) {
        const int strand = p.strand ? p.strand[u] : 0;
        const int64_t c = strand ? p.col0[u] + (p.W - 1 - t) : p.col0[u] + t;
        ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
        ColWords cwp; cwp.b = 0; cwp.g = 0;
        if (t > 0) cwp = load_column(p.bases, p.gaps, p.n_cols, p.nb, strand ? c + 1 : c - 1);
        if (strand) { cw = complement(cw); cwp = complement(cwp); }
        auto Q = [&](int tt, int i, int a) { return p.q_int[(size_t)((tt * p.I + i) * 4 + a) * n + u]; };
        auto GQ = [&](int tt, int kk, int a) { return p.q_leaf[(size_t)((tt * p.S + kk) * 4 + a) * n + u]; };
        const double w = p.weight[u];
        for (int i = 0; i < p.I; ++i) {
            double qi[4], qp[4], qh[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                qi[a] = Q(t, i, a);
                qh[a] = t > 0 ? Q(t - 1, i, a) : (a == 0 ? 1.0 : 0.0);
                qp[a] = i > 0 ? Q(t, par_internal[i], a) : 0.0;
            }
            if (i == 0) {
                for (int cc = 0; cc < ncc; ++cc)
#pragma unroll
                    for (int a = 0; a < 4; ++a) atomicAdd(&s_c[cc * 4 + a], w * (qh[cc] * qi[a]));
            } else {
                const int off = 16 + 32 * (node_of[i] - 1);
                double rest[4];
                rest[0] = (qp[1] + qp[2]) + qp[3]; rest[1] = (qp[0] + qp[2]) + qp[3];
                rest[2] = (qp[0] + qp[1]) + qp[3]; rest[3] = (qp[0] + qp[1]) + qp[2];
                for (int cc = 0; cc < ncc; ++cc)
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        atomicAdd(&s_c[off + cc * 4 + a], w * (qi[a] * rest[a] * qh[cc]));          // off-diagonal rows
                        atomicAdd(&s_c[off + 16 + cc * 4 + a], w * (qi[a] * qp[a] * qh[cc]));       // diagonal
                    }
            }
            const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
            for (int kk = leaf_begin[i]; kk < leaf_begin[i + 1]; ++kk) {
                const int sp = leaf_species[kk];
                const int x = cw.obs(sp);
                const int y = t > 0 ? cwp.obs(sp) : 0;
                double wc[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wc[cc] = t == 0 ? (cc == 0 ? 1.0 : 0.0) : (y < 4 ? (cc == y ? 1.0 : 0.0) : GQ(t - 1, kk, cc));
                if (x < 4) {
                    const int off = 16 + 32 * (leaf_node[kk] - 1);
                    for (int cc = 0; cc < ncc; ++cc) {
                        if (wc[cc] == 0.0) continue;
#pragma unroll
                        for (int a = 0; a < 4; ++a) atomicAdd(&s_c[(a == x ? off + 16 : off) + cc * 4 + x], w * (qi[a] * wc[cc]));
                    }
                } else {
                    double ql[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) ql[a] = GQ(t, kk, a);
                    for (int cc = 0; cc < ncc; ++cc) {
                        if (wc[cc] == 0.0) continue;
#pragma unroll
                        for (int a = 0; a < 4; ++a) atomicAdd(&s_c[cc * 4 + a], w * (ql[a] * (qs * wc[cc])));
                    }
                }
            }
        }
    }
    __syncthreads();
    double* out = p.part + ((size_t)t * gridDim.x + blockIdx.x) * p.E1;
    for (int i = tid; i < p.E1; i += T) out[i] = s_c[i];
}

// Finishes the evaluation of ONE position: G_k = sum over blocks (fixed order); G[pos] = G_0;
// mode 1: values[0] = sum_t G_t - ENT, values[value_off + k - 1] = the same with G_k in place of G_0 (k >= 1).
__global__ void __launch_bounds__(256) fe1_final_kernel(const double* __restrict__ part, int blocks, int pos, int K, int W, int mode,
                                                        int value_off, double* __restrict__ G, double* __restrict__ values) {
    __shared__ double gk[FE1_KMAX];
    __shared__ double sh[32];
    for (int k = 0; k < K; ++k) {
        double v = 0.0;
        for (int b = threadIdx.x; b < blocks; b += blockDim.x) v += part[(size_t)b * FE1_KMAX + k];
        const double s = block_sum(v, sh);
        if (threadIdx.x == 0) gk[k] = s;
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    G[pos] = gk[0];
    if (mode == 0) return;
    double total = 0.0, rest = 0.0;
    for (int t = 0; t < W; ++t) { total += G[t]; if (t != pos) rest += G[t]; }
    values[0] = total - G[W];
    for (int k = 1; k < K; ++k) values[value_off + k - 1] = (rest + gk[k]) - G[W];
}
