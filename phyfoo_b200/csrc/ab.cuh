// ab.cuh -- the non-phylogenetic alignment models of order 0 (the `model.evolution=false` branch of the reference's
// factory, models/ModelUtil.java:84-90): a product over the species rows of one PWM,
// gaps scoring log(0.25).  Pure table gathers over the packed columns.
//   models/AlignmentBasedModel.java:188-248    motif windows   -> ab_windows_kernel   (sum order: rows outer, positions inner)
//   models/AlignmentBasedBGModel.java:173-246  background      -> ab_columns_kernel   (one term per column)
//   models/AlignmentBasedModel.java:97-168     count M-step    -> ab_count_kernel
// Orders >= 1 go through the reference's ACCESS_TABLE expansion, whose gap handling is acknowledged buggy in the source
// (:133 "here is an error"); they are not implemented.
#pragma once

struct AbWindowParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int strand;
    int W, S;
    const double* ltab;      // [W][5] logs; entry 4 = log(0.25)
    double* out;
};

__global__ void __launch_bounds__(256) ab_windows_kernel(AbWindowParams p) {
    extern __shared__ double s_tab[];                      // [W][5]
    for (int i = threadIdx.x; i < p.W * 5; i += blockDim.x) s_tab[i] = p.ltab[i];
    __syncthreads();
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < p.n_windows; w += (int64_t)gridDim.x * blockDim.x) {
        const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
        const int64_t c0 = p.col_off[a] + (w - p.win_off[a]);
        double log_sum = 0.0;
        for (int o = 0; o < p.S; ++o) {                    // AlignmentBasedModel.java:241-243
            double sum = 0.0;
            for (int i = 0; i < p.W; ++i) {                // :170-175
                const int64_t c = p.strand ? c0 + (p.W - 1 - i) : c0 + i;
                ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                if (p.strand) cw = complement(cw);
                sum += s_tab[i * 5 + cw.obs(o)];
            }
            log_sum += sum;
        }
        p.out[w] = log_sum;
    }
}

// per-column background term: sum over rows (AlignmentBasedBGModel.java:236-240 adds them to one running total; per
// column the partial sum is exact in the same order, ranges are prefix sums on the caller's side)
__global__ void __launch_bounds__(256) ab_columns_kernel(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ gaps, int64_t n_cols, int nb,
                                                         int S, double l0, double l1, double l2, double l3, double lgap, double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const ColWords cw = load_column(bases, gaps, n_cols, nb, c);
    double sum = 0.0;
    for (int o = 0; o < S; ++o) {
        const int x = cw.obs(o);
        sum += x == 0 ? l0 : x == 1 ? l1 : x == 2 ? l2 : x == 3 ? l3 : lgap;
    }
    out[c] = sum;
}

struct AbCountParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col0; const uint8_t* strand; const double* weight; int64_t n;
    int W, S;
    double* part;            // [gridDim.x][W*4]
};

// weighted base counts of the selected windows (shared-memory atomics per block, fixed-order reduction of the partials)
__global__ void __launch_bounds__(256) ab_count_kernel(AbCountParams p) {
    extern __shared__ double s_c[];                        // [W][4]
    for (int i = threadIdx.x; i < p.W * 4; i += blockDim.x) s_c[i] = 0.0;
    __syncthreads();
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < p.n; u += (int64_t)gridDim.x * blockDim.x) {
        const int strand = p.strand ? p.strand[u] : 0;
        const double w = p.weight[u];
        for (int l = 0; l < p.W; ++l) {
            const int64_t c = strand ? p.col0[u] + (p.W - 1 - l) : p.col0[u] + l;
            ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
            if (strand) cw = complement(cw);
            for (int o = 0; o < p.S; ++o) {
                const int x = cw.obs(o);
                if (x < 4) atomicAdd(&s_c[l * 4 + x], w / 1);
                else for (int a = 0; a < 4; ++a) atomicAdd(&s_c[l * 4 + a], w / 4);     // AlignmentBasedModel.java:132-141
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.W * 4; i += blockDim.x) p.part[(size_t)blockIdx.x * p.W * 4 + i] = s_c[i];
}

extern "C" int phyfoo_ab_score_windows(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, const double* condprob, int32_t strand, double* out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob || !out || width < 1) return fail(ctx, PHYFOO_EINVAL, "ab_score_windows: NULL argument or width < 1");
    if (strand != 0 && strand != 1) return fail(ctx, PHYFOO_EINVAL, "ab_score_windows: strand must be 0 or 1");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, width, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t nw = h_woff->back();
    if (nw == 0) return PHYFOO_OK;
    std::vector<double> lt((size_t)width * 5);
    for (int i = 0; i < width; ++i) {
        for (int b = 0; b < 4; ++b) lt[(size_t)i * 5 + b] = std::log(condprob[i * 4 + b] / 1);    // AlignmentBasedModel.java:220
        lt[(size_t)i * 5 + 4] = std::log(0.25);                                                    // :190
    }
    CU(ctx->fg_scores.ensure((size_t)nw * sizeof(double)));
    CU(ctx->misc.ensure(lt.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->misc.p, lt.data(), lt.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AbWindowParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_woff; p.n_aln = ds->n_aln; p.n_windows = nw; p.strand = strand;
    p.W = width; p.S = ds->S; p.ltab = ctx->misc.as<double>(); p.out = ctx->fg_scores.as<double>();
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((nw + 255) / 256, 16LL * ctx->sm_count));
    {
        KernelTimer kt_(ctx, "ab_windows_kernel");
        ab_windows_kernel<<<blocks, 256, lt.size() * sizeof(double), ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->fg_scores.p, (size_t)nw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_ab_bg_columns(phyfoo_ctx* ctx, const phyfoo_dataset* ds, const double* condprob4, double* out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob4 || !out) return fail(ctx, PHYFOO_EINVAL, "ab_bg_columns: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    if (ds->n_cols == 0) return PHYFOO_OK;
    CU(ctx->bg_scores.ensure((size_t)ds->n_cols * sizeof(double)));
    {
        KernelTimer kt_(ctx, "ab_columns_kernel");
        ab_columns_kernel<<<(unsigned)((ds->n_cols + 255) / 256), 256, 0, ctx->stream>>>(
            ds->d_bases, ds->d_gaps, ds->n_cols, ds->nb, ds->S, std::log(condprob4[0] / 1), std::log(condprob4[1] / 1),
            std::log(condprob4[2] / 1), std::log(condprob4[3] / 1), std::log(0.25), ctx->bg_scores.as<double>());
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->bg_scores.p, (size_t)ds->n_cols * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_ab_train(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, int64_t n_sel, const int32_t* sel_aln,
                               const int32_t* sel_off, const uint8_t* sel_strand, const double* sel_weight, double pseudo_count,
                               double* condprob_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob_out || width < 1 || n_sel < 0 || (n_sel > 0 && (!sel_aln || !sel_off || !sel_weight)))
        return fail(ctx, PHYFOO_EINVAL, "ab_train: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const size_t WC = (size_t)width * 4;
    std::vector<double> counts(WC, 0.0);
    if (n_sel > 0) {
        std::vector<int64_t> col0((size_t)n_sel);
        for (int64_t u = 0; u < n_sel; ++u) {
            const int a = sel_aln[u];
            if (a < 0 || a >= ds->n_aln) return fail(ctx, PHYFOO_EINVAL, "ab_train: alignment index out of range");
            if (sel_off[u] < 0 || sel_off[u] + width > ds->col_off[a + 1] - ds->col_off[a]) return fail(ctx, PHYFOO_EINVAL, "ab_train: window leaves its alignment");
            col0[u] = ds->col_off[a] + sel_off[u];
        }
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n_sel + 255) / 256, (int64_t)ctx->sm_count));
        const size_t bytes = (size_t)n_sel * 17 + 64 + ((size_t)blocks + 1) * WC * sizeof(double);
        CU(ctx->misc.ensure(bytes));
        double* d_part = ctx->misc.as<double>();
        double* d_w = d_part + ((size_t)blocks + 1) * WC;
        int64_t* d_col0 = (int64_t*)(d_w + n_sel);
        uint8_t* d_strand = (uint8_t*)(d_col0 + n_sel);
        CU(cudaMemcpyAsync(d_w, sel_weight, (size_t)n_sel * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_col0, col0.data(), (size_t)n_sel * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (sel_strand) CU(cudaMemcpyAsync(d_strand, sel_strand, (size_t)n_sel, cudaMemcpyHostToDevice, ctx->stream));
        AbCountParams p;
        p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
        p.col0 = d_col0; p.strand = sel_strand ? d_strand : nullptr; p.weight = d_w; p.n = n_sel; p.W = width; p.S = ds->S; p.part = d_part;
        {
            KernelTimer kt_(ctx, "ab_count_kernel");
            ab_count_kernel<<<blocks, 256, WC * sizeof(double), ctx->stream>>>(p);
        }
        CU(cudaGetLastError());
        {
            KernelTimer kt_(ctx, "fe_count_reduce_kernel");
            fe_count_reduce_kernel<<<(unsigned)((WC + 127) / 128), 128, 0, ctx->stream>>>(d_part, blocks, (int)WC, (int)WC, d_part + (size_t)blocks * WC);
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(counts.data(), d_part + (size_t)blocks * WC, WC * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));   // col0 is stack-scoped
    }
    for (int i = 0; i < width; ++i) {             // :150-158: pseudo count, then sumNormalisation per block of 4
        double c[4], sum = 0.0;
        for (int k = 0; k < 4; ++k) { c[k] = pseudo_count + counts[(size_t)i * 4 + k]; sum = k ? sum + c[k] : c[k]; }
        for (int k = 0; k < 4; ++k) condprob_out[i * 4 + k] = c[k] / sum;
    }
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Orders 1 and 2 (models/AlignmentBasedModel.java:97-228, AlignmentBasedBGModel.java:40-222).  Per species row a Markov
// chain over the columns: the entry of position p is condProb[p][sum_k 4^k x(l-k)], k = 0 the column itself, k <= min(order,
// p) its predecessors (the motif: inside the window only; the background: back to the start of the alignment, whatever range
// is asked for, so ranges stay sums of per-column terms).  Gaps: the reference expands an unobserved symbol into the four
// bases and averages the PROBABILITIES of the expansions (motif :196-221 through its ACCESS_TABLE, background :173-222 through
// a list of digit arrays); the background's likelihood copies ALL earlier list entries at a later gap (:185-191), not only the
// current ones, so with gaps at two predecessors of one column it averages over stale entries too.  ab_expand restates both
// list procedures literally (entries as hashes; "set digit k of every entry" = add x 4^k, the digit was 0), so those values are
// reproduced as they are.  A context without any observed symbol scores log(0.25) (:189-191 / :174-176).
// Table layout (condprob / counts): position p at offset ab_off(p, order), 4^(min(p, order) + 1) entries.
// ------------------------------------------------------------------------------------------------------------------------
#include "ab_expand.hpp"

// log P of one row at one position given the codes of the column and its predecessors
__device__ __forceinline__ double ab_term(const int* sym, int kmax, int variant, const double* __restrict__ cond, const double* __restrict__ lcond) {
    bool any_gap = false, any_obs = false;
    int h = 0, pw = 1;
    for (int k = 0; k <= kmax; ++k, pw *= 4) {
        if (sym[k] == 4) any_gap = true; else { any_obs = true; h += sym[k] * pw; }
    }
    if (!any_obs) return -1.3862943611198906;              // Math.log(0.25)
    if (!any_gap) return lcond[h];                         // log(condProb / 1)
    int idx[AB_MAX_EXPAND], pre, n;
    ab_expand(sym, kmax, variant, idx, pre, n);
    const int size = n - pre;
    double sum = 0.0;
    if (variant) { for (int i = pre; i < n; ++i) sum += cond[idx[i]] / size; }          // background :213-215
    else { for (int i = pre; i < n; ++i) sum += cond[idx[i]]; sum /= size; }            // motif :224-228
    return log(sum);
}

struct AbOrderParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int strand;
    int W, S, order;
    const double* cond; const double* lcond;               // concatenated tables (ab_off), probabilities and their logs
    double* out;
};

__global__ void __launch_bounds__(128) ab_windows_order_kernel(AbOrderParams p) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < p.n_windows; w += (int64_t)gridDim.x * blockDim.x) {
        const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
        const int64_t c0 = p.col_off[a] + (w - p.win_off[a]);
        double log_sum = 0.0;
        for (int o = 0; o < p.S; ++o) {                    // AlignmentBasedModel.java:241-243
            double sum = 0.0;
            int off = 0;
            for (int i = 0; i < p.W; ++i) {                // :170-175
                const int kmax = i < p.order ? i : p.order;
                int sym[AB_MAX_ORDER + 1];
                for (int k = 0; k <= kmax; ++k) {
                    const int64_t c = p.strand ? c0 + (p.W - 1 - (i - k)) : c0 + (i - k);
                    ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                    if (p.strand) cw = complement(cw);
                    sym[k] = cw.obs(o);
                }
                sum += ab_term(sym, kmax, 0, p.cond + off, p.lcond + off);
                off += ab_size(i, p.order);
            }
            log_sum += sum;
        }
        p.out[w] = log_sum;
    }
}

// background: one term per column = sum over the rows (AlignmentBasedBGModel.java:236-240)
__global__ void __launch_bounds__(128) ab_columns_order_kernel(AbOrderParams p) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_cols) return;
    const int a = find_alignment(p.col_off, p.n_aln, c);
    const int64_t l = c - p.col_off[a];                    // position in its alignment
    const int kmax = l < p.order ? (int)l : p.order;
    const int off = ab_off(kmax, p.order);                 // condProb[Math.min(posSeq, order)]
    ColWords cw[AB_MAX_ORDER + 1];
    for (int k = 0; k <= kmax; ++k) cw[k] = load_column(p.bases, p.gaps, p.n_cols, p.nb, c - k);
    double sum = 0.0;
    for (int o = 0; o < p.S; ++o) {
        int sym[AB_MAX_ORDER + 1];
        for (int k = 0; k <= kmax; ++k) sym[k] = cw[k].obs(o);
        sum += ab_term(sym, kmax, 1, p.cond + off, p.lcond + off);
    }
    p.out[c] = sum;
}

struct AbCountOrderParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; int n_aln;
    const int64_t* col0; const uint8_t* strand; const double* weight; int64_t n;      // motif: selected windows
    const double* aln_weight;                                                          // background: per alignment (nullable = 1)
    int W, S, order, total;
    double* part;            // [gridDim.x][total]
};

// AlignmentBasedModel.java:106-147: weighted counts of the selected windows, a gap's weight split over its expansions
__global__ void __launch_bounds__(128) ab_count_order_kernel(AbCountOrderParams p) {
    extern __shared__ double s_c[];
    for (int i = threadIdx.x; i < p.total; i += blockDim.x) s_c[i] = 0.0;
    __syncthreads();
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < p.n; u += (int64_t)gridDim.x * blockDim.x) {
        const int strand = p.strand ? p.strand[u] : 0;
        const double w = p.weight[u];
        for (int o = 0; o < p.S; ++o) {
            int off = 0;
            for (int l = 0; l < p.W; ++l) {
                const int kmax = l < p.order ? l : p.order;
                int sym[AB_MAX_ORDER + 1];
                for (int k = 0; k <= kmax; ++k) {
                    const int64_t c = strand ? p.col0[u] + (p.W - 1 - (l - k)) : p.col0[u] + (l - k);
                    ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                    if (strand) cw = complement(cw);
                    sym[k] = cw.obs(o);
                }
                int idx[AB_MAX_EXPAND], pre, n;
                ab_expand(sym, kmax, 0, idx, pre, n);
                for (int i = pre; i < n; ++i) atomicAdd(&s_c[off + idx[i]], w / (n - pre));
                off += ab_size(l, p.order);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.total; i += blockDim.x) p.part[(size_t)blockIdx.x * p.total + i] = s_c[i];
}

// AlignmentBasedBGModel.java:56-122: every column of every alignment, table min(l, order)
__global__ void __launch_bounds__(128) ab_bg_count_order_kernel(AbCountOrderParams p) {
    extern __shared__ double s_c[];
    for (int i = threadIdx.x; i < p.total; i += blockDim.x) s_c[i] = 0.0;
    __syncthreads();
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < p.n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        const int a = find_alignment(p.col_off, p.n_aln, c);
        const int64_t l = c - p.col_off[a];
        const int kmax = l < p.order ? (int)l : p.order;
        const int off = ab_off(kmax, p.order);
        const double w = p.aln_weight ? p.aln_weight[a] : 1.0;
        ColWords cw[AB_MAX_ORDER + 1];
        for (int k = 0; k <= kmax; ++k) cw[k] = load_column(p.bases, p.gaps, p.n_cols, p.nb, c - k);
        for (int o = 0; o < p.S; ++o) {
            int sym[AB_MAX_ORDER + 1];
            for (int k = 0; k <= kmax; ++k) sym[k] = cw[k].obs(o);
            int idx[AB_MAX_EXPAND], pre, n;
            ab_expand(sym, kmax, 0, idx, pre, n);
            for (int i = pre; i < n; ++i) atomicAdd(&s_c[off + idx[i]], w / (n - pre));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.total; i += blockDim.x) p.part[(size_t)blockIdx.x * p.total + i] = s_c[i];
}

static int ab_tables_to_device(phyfoo_ctx* ctx, const double* condprob, int total, const double** d_cond, const double** d_lcond) {
    std::vector<double> t((size_t)2 * total);
    for (int i = 0; i < total; ++i) { t[i] = condprob[i]; t[(size_t)total + i] = std::log(condprob[i] / 1); }
    CU(ctx->misc.ensure(t.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->misc.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));                // t is stack-scoped
    *d_cond = ctx->misc.as<double>();
    *d_lcond = ctx->misc.as<double>() + total;
    return PHYFOO_OK;
}

// sumNormalisation of every block of four (current symbol fastest) after the pseudo count (AlignmentBasedModel.java:150-158)
static void ab_normalise(const std::vector<double>& counts, double pseudo, double* out) {
    for (size_t k = 0; k < counts.size(); k += 4) {
        double c[4], sum = 0.0;
        for (int j = 0; j < 4; ++j) { c[j] = pseudo + counts[k + j]; sum = j ? sum + c[j] : c[j]; }
        for (int j = 0; j < 4; ++j) out[k + j] = c[j] / sum;
    }
}

extern "C" int phyfoo_ab_table_size(int32_t width, int32_t order) {
    if (width < 0 || order < 0 || order > AB_MAX_ORDER) return -1;
    return width == 0 ? ab_off(order + 1, order) : ab_off(width, order);       // width 0: the background (order + 1 tables)
}

extern "C" int phyfoo_ab_score_windows_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, int32_t order, const double* condprob,
                                             int32_t strand, double* out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob || !out || width < 1) return fail(ctx, PHYFOO_EINVAL, "ab_score_windows_order: NULL argument or width < 1");
    if (order < 0 || order > AB_MAX_ORDER) return fail(ctx, PHYFOO_EUNSUPPORTED, "ab_score_windows_order: orders 0..2 are supported");
    if (strand != 0 && strand != 1) return fail(ctx, PHYFOO_EINVAL, "ab_score_windows_order: strand must be 0 or 1");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, width, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t nw = h_woff->back();
    if (nw == 0) return PHYFOO_OK;
    AbOrderParams p;
    if ((rc = ab_tables_to_device(ctx, condprob, ab_off(width, order), &p.cond, &p.lcond))) return rc;
    CU(ctx->fg_scores.ensure((size_t)nw * sizeof(double)));
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_woff; p.n_aln = ds->n_aln; p.n_windows = nw; p.strand = strand;
    p.W = width; p.S = ds->S; p.order = order; p.out = ctx->fg_scores.as<double>();
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((nw + 127) / 128, 32LL * ctx->sm_count));
    {
        KernelTimer kt_(ctx, "ab_windows_order_kernel");
        ab_windows_order_kernel<<<blocks, 128, 0, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->fg_scores.p, (size_t)nw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_ab_bg_columns_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t order, const double* condprob, double* out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob || !out) return fail(ctx, PHYFOO_EINVAL, "ab_bg_columns_order: NULL argument");
    if (order < 0 || order > AB_MAX_ORDER) return fail(ctx, PHYFOO_EUNSUPPORTED, "ab_bg_columns_order: orders 0..2 are supported");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    if (ds->n_cols == 0) return PHYFOO_OK;
    AbOrderParams p;
    int rc = ab_tables_to_device(ctx, condprob, ab_off(order + 1, order), &p.cond, &p.lcond);
    if (rc) return rc;
    CU(ctx->bg_scores.ensure((size_t)ds->n_cols * sizeof(double)));
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = nullptr; p.n_aln = ds->n_aln; p.n_windows = 0; p.strand = 0;
    p.W = 0; p.S = ds->S; p.order = order; p.out = ctx->bg_scores.as<double>();
    {
        KernelTimer kt_(ctx, "ab_columns_order_kernel");
        ab_columns_order_kernel<<<(unsigned)((ds->n_cols + 127) / 128), 128, 0, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->bg_scores.p, (size_t)ds->n_cols * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

static int ab_run_counts(phyfoo_ctx* ctx, AbCountOrderParams& p, bool background, int64_t work, size_t head_bytes, std::vector<double>& counts) {
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((work + 127) / 128, (int64_t)ctx->sm_count));
    double* d_part = (double*)((char*)ctx->misc.p + head_bytes);
    p.part = d_part;
    {
        KernelTimer kt_(ctx, background ? "ab_bg_count_order_kernel" : "ab_count_order_kernel");
        if (background) ab_bg_count_order_kernel<<<blocks, 128, (size_t)p.total * sizeof(double), ctx->stream>>>(p);
        else ab_count_order_kernel<<<blocks, 128, (size_t)p.total * sizeof(double), ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    {
        KernelTimer kt_(ctx, "fe_count_reduce_kernel");
        fe_count_reduce_kernel<<<(unsigned)((p.total + 127) / 128), 128, 0, ctx->stream>>>(d_part, blocks, p.total, p.total, d_part + (size_t)blocks * p.total);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts.data(), d_part + (size_t)blocks * p.total, (size_t)p.total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_ab_train_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, int32_t order, int64_t n_sel, const int32_t* sel_aln,
                                     const int32_t* sel_off, const uint8_t* sel_strand, const double* sel_weight, double pseudo_count,
                                     double* condprob_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob_out || width < 1 || n_sel < 0 || (n_sel > 0 && (!sel_aln || !sel_off || !sel_weight)))
        return fail(ctx, PHYFOO_EINVAL, "ab_train_order: NULL argument");
    if (order < 0 || order > AB_MAX_ORDER) return fail(ctx, PHYFOO_EUNSUPPORTED, "ab_train_order: orders 0..2 are supported");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int total = ab_off(width, order);
    std::vector<double> counts((size_t)total, 0.0);
    if (n_sel > 0) {
        std::vector<int64_t> col0((size_t)n_sel);
        for (int64_t u = 0; u < n_sel; ++u) {
            const int a = sel_aln[u];
            if (a < 0 || a >= ds->n_aln) return fail(ctx, PHYFOO_EINVAL, "ab_train_order: alignment index out of range");
            if (sel_off[u] < 0 || sel_off[u] + width > ds->col_off[a + 1] - ds->col_off[a]) return fail(ctx, PHYFOO_EINVAL, "ab_train_order: window leaves its alignment");
            col0[u] = ds->col_off[a] + sel_off[u];
        }
        const size_t head = ((size_t)n_sel * 17 + 63) / 64 * 64;
        CU(ctx->misc.ensure(head + ((size_t)ctx->sm_count + 1) * total * sizeof(double)));
        double* d_w = ctx->misc.as<double>();
        int64_t* d_col0 = (int64_t*)(d_w + n_sel);
        uint8_t* d_strand = (uint8_t*)(d_col0 + n_sel);
        CU(cudaMemcpyAsync(d_w, sel_weight, (size_t)n_sel * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_col0, col0.data(), (size_t)n_sel * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (sel_strand) CU(cudaMemcpyAsync(d_strand, sel_strand, (size_t)n_sel, cudaMemcpyHostToDevice, ctx->stream));
        AbCountOrderParams p;
        p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb; p.col_off = ds->d_col_off; p.n_aln = ds->n_aln;
        p.col0 = d_col0; p.strand = sel_strand ? d_strand : nullptr; p.weight = d_w; p.n = n_sel; p.aln_weight = nullptr;
        p.W = width; p.S = ds->S; p.order = order; p.total = total;
        const int rc = ab_run_counts(ctx, p, false, n_sel, head, counts);      // synchronises: col0 is stack-scoped
        if (rc) return rc;
    }
    ab_normalise(counts, pseudo_count, condprob_out);
    return PHYFOO_OK;
}

extern "C" int phyfoo_ab_bg_train_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t order, const double* aln_weights, double pseudo_count,
                                        double* condprob_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !condprob_out) return fail(ctx, PHYFOO_EINVAL, "ab_bg_train_order: NULL argument");
    if (order < 0 || order > AB_MAX_ORDER) return fail(ctx, PHYFOO_EUNSUPPORTED, "ab_bg_train_order: orders 0..2 are supported");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int total = ab_off(order + 1, order);
    std::vector<double> counts((size_t)total, 0.0);
    if (ds->n_cols > 0) {
        const size_t head = aln_weights ? ((size_t)ds->n_aln * sizeof(double) + 63) / 64 * 64 : 0;
        CU(ctx->misc.ensure(head + ((size_t)ctx->sm_count + 1) * total * sizeof(double)));
        if (aln_weights) CU(cudaMemcpyAsync(ctx->misc.p, aln_weights, (size_t)ds->n_aln * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        AbCountOrderParams p;
        p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb; p.col_off = ds->d_col_off; p.n_aln = ds->n_aln;
        p.col0 = nullptr; p.strand = nullptr; p.weight = nullptr; p.n = 0; p.aln_weight = aln_weights ? ctx->misc.as<double>() : nullptr;
        p.W = 0; p.S = ds->S; p.order = order; p.total = total;
        const int rc = ab_run_counts(ctx, p, true, ds->n_cols, head, counts);
        if (rc) return rc;
    }
    ab_normalise(counts, pseudo_count, condprob_out);
    return PHYFOO_OK;
}
