// mstep.cuh -- M-step entry points (included at the end of phyfoo.cu): data selection and the fixed-q
// free-energy objective with its forward-difference gradient.
//
// Reference:
//   io/SequenceSpecificDataSelector.java:37-99                     -> select_kernel
//   models/PhyloBayesModel.java:122-219 (train), :142-147 (q once)  -> phyfoo_fe_prepare
//   models/PhyloBackground.java:111-176 (train), :117-134           -> phyfoo_fe_prepare_all
//   optimizing/AbstractFreeEnergyOptimizer.java:95-114              -> fe_eval_kernel (sum_u w_u F(q_u; theta))
//   optimizing/FE_byParameter_ForBayesNodeJstacs.java:43-53, MotifParamOptimizer.java:46-61 (theta from lambda)
//   jstacs NumericalDifferentiableFunction.java:64-84               -> phyfoo_fe_grad (dim + 1 parameter sets, one pass)
//
// With q fixed the objective splits over positions:
//   f(theta) = -sum_u w_u F_u = sum_t G_t(theta_t) - ENT,   G_t = sum_u w_u E_q[log CPF of tree t],  ENT = sum_u w_u sum q log q
// The reference re-evaluates every tree of every selected window for each of the dim+1 parameter sets; here the
// trees of the position(s) under test are evaluated for all dim+1 sets in ONE pass (shared q loads), the other
// positions' G_t are kept from the previous evaluation, and ENT is computed once in prepare.  Same sums, re-associated.

struct phyfoo_feset {
    const phyfoo_dataset* ds = nullptr;
    phyfoo_model* m = nullptr;
    int64_t n = 0;                 // windows
    int W = 0, I = 0, S = 0, E = 0;
    int64_t* d_col0 = nullptr;     // [n] first column of window u
    uint8_t* d_strand = nullptr;   // [n] or null
    cudaStream_t stream = nullptr; // the context's stream: buffers come from its stream-ordered pool (cudaMalloc / cudaFree of
                                   // ~150 MB per M-step took 1..400 ms each on the B200 boxes; the pool keeps what was freed)
    double* d_weight = nullptr;    // [n]
    double* d_qint = nullptr;      // [4I][W*n]  entry (i,a) of tree (t,u) at ((i*4+a) * W*n + t*n + u)
    double* d_qleaf = nullptr;     // [4S][W*n]  gap leaves only
    double* d_ent = nullptr;       // [W*n]
    double* d_G = nullptr;         // [W+1]: G_t under the model's current parameters; [W] = ENT
    double* d_tabs = nullptr;      // [W][KMAX][E] log tables of the parameter sets under test
    double* d_part = nullptr;      // [W][blocks][KMAX]
    double* d_values = nullptr;    // [1 + 4W]
    int blocks = 0;
    uint64_t g_version = 0;        // model version d_G belongs to (0 = never)
    int order = 0, E1 = 0;         // order 1: see mstep1.cuh (q layout [((t*I+i)*4+a)*n + u], one entropy per window)
};

static const int FE_KMAX = 5;      // order 0: dim = 4 per position -> 5 parameter sets

#include "mstep1.cuh"

// ------------------------------------------------------------------------------------------------
// selection: one block per alignment
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_kernel(const double* __restrict__ w, const int64_t* __restrict__ win_off, double t, double q,
                                                     int32_t* __restrict__ tmp_idx, double* __restrict__ tmp_w, int32_t* __restrict__ count) {
    __shared__ double sv[8];
    __shared__ int si[8];
    __shared__ double s_total, s_lv;
    __shared__ int s_lj, s_stop;
    const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int64_t w0 = win_off[a];
    const int n = (int)(win_off[a + 1] - w0);
    if (n <= 0) { if (tid == 0) count[a] = 0; return; }
    const double* wa = w + w0;
    if (tid == 0) {
        double total = wa[0];                       // Normalisation.sumNormalisation (jstacs :410-418): array order
        for (int i = 1; i < n; ++i) total += wa[i];
        s_total = total; s_lv = INFINITY; s_lj = n; s_stop = 0;
    }
    __syncthreads();
    double sum = 0.0;
    int cnt = 0;
    for (int it = 0; it < n; ++it) {
        const double lv = s_lv; const int lj = s_lj;
        // next element of the descending walk: largest (value, index) strictly below the previous pick
        double bv = -INFINITY; int bj = -1;
        for (int j = tid; j < n; j += T) {
            const double v = wa[j];
            if (!(v < lv || (v == lv && j < lj))) continue;
            if (bj < 0 || v > bv || (v == bv && j > bj)) { bv = v; bj = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int oj = __shfl_down_sync(0xffffffffu, bj, o);
            if (oj >= 0 && (bj < 0 || ov > bv || (ov == bv && oj > bj))) { bv = ov; bj = oj; }
        }
        if ((tid & 31) == 0) { sv[tid >> 5] = bv; si[tid >> 5] = bj; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < (T + 31) / 32; ++i)
                if (si[i] >= 0 && (bj < 0 || sv[i] > bv || (sv[i] == bv && si[i] > bj))) { bv = sv[i]; bj = si[i]; }
            if (bj < 0) s_stop = 1;                 // nothing left (or only NaNs)
            else {
                sum += bv / s_total;                // :85-95
                if (bv >= q) { tmp_idx[w0 + cnt] = bj; tmp_w[w0 + cnt] = bv; ++cnt; }
                if (sum > t) s_stop = 1;
                s_lv = bv; s_lj = bj;
            }
        }
        __syncthreads();
        if (s_stop) break;
    }
    if (tid == 0) count[a] = cnt;
}

__global__ void select_compact_kernel(const int32_t* __restrict__ tmp_idx, const double* __restrict__ tmp_w, const int64_t* __restrict__ win_off,
                                      const int64_t* __restrict__ out_off, int32_t* __restrict__ sel_aln, int32_t* __restrict__ sel_off,
                                      double* __restrict__ sel_w) {
    const int a = blockIdx.x;
    const int64_t o = out_off[a];
    const int cnt = (int)(out_off[a + 1] - o);
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        sel_aln[o + k] = a;
        sel_off[o + k] = tmp_idx[win_off[a] + k];
        sel_w[o + k] = tmp_w[win_off[a] + k];
    }
}

extern "C" int phyfoo_select(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, const double* weights, double t, double q,
                             int32_t* sel_aln, int32_t* sel_off, double* sel_weight, int64_t capacity, int64_t* n_selected) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !n_selected || width < 1) return fail(ctx, PHYFOO_EINVAL, "select: NULL argument or width < 1");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    *n_selected = 0;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, width, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t nw = h_woff->back();
    if (nw == 0 || ds->n_aln == 0) return PHYFOO_OK;
    const double* d_w;
    if (weights) {
        CU(ctx->selw.ensure((size_t)nw * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->selw.p, weights, (size_t)nw * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        d_w = ctx->selw.as<double>();
    } else {
        // weights == NULL: gamma of the most recent phyfoo_zoops_estep on this context, still on the device
        if (ctx->gamma_windows != nw || ctx->gamma_ds != ds || ctx->gamma_width != width)
            return fail(ctx, PHYFOO_EINVAL, "select: weights is NULL but no gamma of this data set and motif width is resident (run phyfoo_zoops_estep with gamma_out first)");
        d_w = ctx->gamma.as<double>();
    }
    const size_t na = (size_t)ds->n_aln;
    CU(ctx->sel_tmp.ensure((size_t)nw * (sizeof(int32_t) + sizeof(double)) + (na + 1) * (sizeof(int32_t) + sizeof(int64_t)) + 64));
    double* tmp_w = ctx->sel_tmp.as<double>();
    int64_t* d_out_off = (int64_t*)(tmp_w + nw);
    int32_t* tmp_idx = (int32_t*)(d_out_off + na + 1);
    int32_t* d_count = tmp_idx + nw;
    {
        KernelTimer kt_(ctx, "select_kernel");
        select_kernel<<<(unsigned)na, 256, 0, ctx->stream>>>(d_w, d_woff, t, q, tmp_idx, tmp_w, d_count);
    }
    CU(cudaGetLastError());
    std::vector<int32_t> count(na);
    CU(cudaMemcpyAsync(count.data(), d_count, na * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> out_off(na + 1, 0);
    for (size_t a = 0; a < na; ++a) out_off[a + 1] = out_off[a] + count[a];
    const int64_t total = out_off[na];
    *n_selected = total;
    if (total > capacity) return fail(ctx, PHYFOO_ENOMEM, "select: capacity too small for the selection (n_selected holds the required size)");
    if (total == 0) return PHYFOO_OK;
    if (!sel_aln || !sel_off || !sel_weight) return fail(ctx, PHYFOO_EINVAL, "select: NULL output array");
    CU(ctx->sel_out.ensure((size_t)total * (2 * sizeof(int32_t) + sizeof(double))));
    double* d_sw = ctx->sel_out.as<double>();
    int32_t* d_sa = (int32_t*)(d_sw + total);
    int32_t* d_so = d_sa + total;
    CU(cudaMemcpyAsync(d_out_off, out_off.data(), (na + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    {
        KernelTimer kt_(ctx, "select_compact_kernel");
        select_compact_kernel<<<(unsigned)na, 64, 0, ctx->stream>>>(tmp_idx, tmp_w, d_woff, d_out_off, d_sa, d_so, d_sw);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sel_aln, d_sa, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(sel_off, d_so, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(sel_weight, d_sw, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// fixed-q objective
// ------------------------------------------------------------------------------------------------
struct FeEvalParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col0; const uint8_t* strand; const double* weight;
    const double* q_int; const double* q_leaf;
    int64_t n;                  // windows
    int W, I, S, E, prog_len;
    const int* prog;
    const double* tabs;         // [W][FE_KMAX][E]
    double* part;               // [W][gridDim.x][FE_KMAX]
    int pos0;                   // blockIdx.y + pos0 = position
};

// grid (blocks, positions); a thread walks the windows u = global thread id, + grid stride, ... of its position
template <int K>
__global__ void __launch_bounds__(256) fe_eval_kernel(FeEvalParams p) {
    extern __shared__ double smem[];
    __shared__ double sh[32];
    const int T = blockDim.x, tid = threadIdx.x;
    const int t = p.pos0 + blockIdx.y;
    double* s_tab = smem;                                   // [K][E]
    int* s_prog = (int*)(s_tab + (size_t)FE_KMAX * p.E);
    const double* gt = p.tabs + (size_t)t * FE_KMAX * p.E;
    for (int i = tid; i < K * p.E; i += T) s_tab[i] = gt[i];
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    __syncthreads();
    TreeProg tp;
    tp.n_internal = p.I; tp.n_leaves = p.S; tp.n_nodes = 0;
    tp.par_internal = s_prog;
    tp.node_of = s_prog + p.I;
    tp.leaf_begin = s_prog + 2 * p.I;
    tp.leaf_node = s_prog + 3 * p.I + 1;
    tp.leaf_species = s_prog + 3 * p.I + 1 + p.S;

    const size_t n_trees = (size_t)p.n * p.W;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int64_t u = (int64_t)blockIdx.x * T + tid; u < p.n; u += (int64_t)gridDim.x * T) {
        const int strand = p.strand ? p.strand[u] : 0;
        const int64_t c = strand ? p.col0[u] + (p.W - 1 - t) : p.col0[u] + t;
        ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
        if (strand) cw = complement(cw);
        const size_t tree = (size_t)t * p.n + u;
        double e[K];
        expected_log_cpf_multi<K>(tp, s_tab, p.E, p.q_int + tree, p.q_leaf + tree, n_trees, cw, e);
        const double w = p.weight[u];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += w * e[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double s = block_sum(acc[k], sh);
        if (tid == 0) p.part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * FE_KMAX + k] = s;
        __syncthreads();
    }
}

// ENT = sum_u w_u sum_t ent(t,u): one block, fixed order
__global__ void __launch_bounds__(1024) fe_entropy_kernel(const double* __restrict__ ent, const double* __restrict__ weight, int64_t n, int W,
                                                          double* __restrict__ out) {
    __shared__ double sh[32];
    double v = 0.0;
    for (int64_t u = threadIdx.x; u < n; u += blockDim.x) {
        double e = 0.0;
        for (int t = 0; t < W; ++t) e += ent[(size_t)t * n + u];
        v += weight[u] * e;
    }
    const double s = block_sum(v, sh);
    if (threadIdx.x == 0) *out = s;
}

// Finishes an evaluation: G_{t,k} = sum over blocks of part (fixed order), then
//   mode 0 (refresh):   G[t] = G_{t,0} for the n_pos positions from pos0
//   mode 1 (evaluate):  values[0] = sum_t G_t - ENT with G[pos] replaced by G_{pos,0};
//                       values[1 + 4*(t - pos0) + j] = the same with G_{t,1+j} in place of G_{t,0}
// G (positions under test) is updated to the new G_{t,0}: the model is left at lambda.
__global__ void __launch_bounds__(256) fe_final_kernel(const double* __restrict__ part, int blocks, int pos0, int n_pos, int K, int W, int mode,
                                                       double* __restrict__ G, double* __restrict__ values) {
    extern __shared__ double gk[];   // [n_pos][FE_KMAX]
    __shared__ double sh[32];
    for (int tp = 0; tp < n_pos; ++tp)
        for (int k = 0; k < K; ++k) {
            double v = 0.0;
            for (int b = threadIdx.x; b < blocks; b += blockDim.x) v += part[((size_t)tp * blocks + b) * FE_KMAX + k];
            const double s = block_sum(v, sh);
            if (threadIdx.x == 0) gk[tp * FE_KMAX + k] = s;
            __syncthreads();
        }
    if (threadIdx.x != 0) return;
    for (int tp = 0; tp < n_pos; ++tp) G[pos0 + tp] = gk[tp * FE_KMAX];
    if (mode == 0) return;
    double total = 0.0;
    for (int t = 0; t < W; ++t) total += G[t];
    values[0] = total - G[W];
    for (int tp = 0; tp < n_pos; ++tp) {
        double rest = 0.0;                       // every position but this one, same order as above
        for (int t = 0; t < W; ++t) if (t != pos0 + tp) rest += G[t];
        for (int k = 1; k < K; ++k) values[1 + (K - 1) * tp + (k - 1)] = (rest + gk[tp * FE_KMAX + k]) - G[W];
    }
}

static void fe_release(phyfoo_feset* fe) {
    void* bufs[] = {fe->d_col0, fe->d_strand, fe->d_weight, fe->d_qint, fe->d_qleaf, fe->d_ent, fe->d_G, fe->d_tabs, fe->d_part, fe->d_values};
    for (void* b : bufs) dev_free(fe->stream, b);
    delete fe;
}

// launches fe_eval_kernel + fe_final_kernel for positions [pos0, pos0 + n_pos) with the K tables per position already in d_tabs
static int fe_launch(phyfoo_ctx* ctx, phyfoo_feset* fe, int pos0, int n_pos, int K, int mode) {
    const phyfoo_dataset* ds = fe->ds;
    FeEvalParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col0 = fe->d_col0; p.strand = fe->d_strand; p.weight = fe->d_weight;
    p.q_int = fe->d_qint; p.q_leaf = fe->d_qleaf; p.n = fe->n;
    p.W = fe->W; p.I = fe->I; p.S = fe->S; p.E = fe->E; p.prog_len = fe->m->prog_len; p.prog = fe->m->d_prog;
    p.tabs = fe->d_tabs; p.part = fe->d_part; p.pos0 = pos0;
    const size_t smem = (size_t)FE_KMAX * fe->E * sizeof(double) + (size_t)fe->m->prog_len * sizeof(int) + 16;
    if (K == 1) {
        CU(cudaFuncSetAttribute(fe_eval_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt_(ctx, "fe_eval_kernel<1>");
        fe_eval_kernel<1><<<dim3(fe->blocks, n_pos), 256, smem, ctx->stream>>>(p);
    } else {
        CU(cudaFuncSetAttribute(fe_eval_kernel<FE_KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt_(ctx, "fe_eval_kernel<5>");
        fe_eval_kernel<FE_KMAX><<<dim3(fe->blocks, n_pos), 256, smem, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    {
        KernelTimer kt_(ctx, "fe_final_kernel");
        fe_final_kernel<<<1, 256, (size_t)n_pos * FE_KMAX * sizeof(double), ctx->stream>>>(fe->d_part, fe->blocks, pos0, n_pos, K, fe->W, mode, fe->d_G, fe->d_values);
    }
    CU(cudaGetLastError());
    return PHYFOO_OK;
}

// tables of the model's CURRENT parameters for all positions (set 0), then G_t for all t
// ---- order 1: one position per launch (mstep1.cuh) ----
static int fe1_launch(phyfoo_ctx* ctx, phyfoo_feset* fe, int pos, int K, int mode, int value_off, const std::vector<double>& tabs) {
    CU(cudaMemcpyAsync(fe->d_tabs, tabs.data(), (size_t)K * fe->E1 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const phyfoo_dataset* ds = fe->ds;
    Fe1EvalParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col0 = fe->d_col0; p.strand = fe->d_strand; p.weight = fe->d_weight;
    p.q_int = fe->d_qint; p.q_leaf = fe->d_qleaf; p.n = fe->n;
    p.W = fe->W; p.I = fe->I; p.S = fe->S; p.E1 = fe->E1; p.prog_len = fe->m->prog_len; p.prog = fe->m->d_prog;
    p.tabs = fe->d_tabs; p.part = fe->d_part; p.pos = pos;
    const size_t smem = (size_t)K * fe->E1 * sizeof(double) + (size_t)fe->m->prog_len * sizeof(int) + 16;
    if (smem > std::min<size_t>(ctx->smem_optin, 227 * 1024))
        return fail(ctx, PHYFOO_EUNSUPPORTED, "fe_eval: the parameter sets of an order-1 position do not fit in shared memory for this tree");
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((fe->n + 127) / 128, (int64_t)fe->blocks));
    {
        KernelTimer kt_(ctx, "fe1_eval_kernel");
        if (K == 1) { CU(cudaFuncSetAttribute(fe1_eval_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); fe1_eval_kernel<1><<<blocks, 128, smem, ctx->stream>>>(p); }
        else if (K == 5) { CU(cudaFuncSetAttribute(fe1_eval_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); fe1_eval_kernel<5><<<blocks, 128, smem, ctx->stream>>>(p); }
        else { CU(cudaFuncSetAttribute(fe1_eval_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); fe1_eval_kernel<17><<<blocks, 128, smem, ctx->stream>>>(p); }
    }
    CU(cudaGetLastError());
    {
        KernelTimer kt_(ctx, "fe1_final_kernel");
        fe1_final_kernel<<<1, 256, 0, ctx->stream>>>(fe->d_part, blocks, pos, K, fe->W, mode, value_off, fe->d_G, fe->d_values);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));   // d_tabs is reused by the next position; `tabs` belongs to the caller
    return PHYFOO_OK;
}

static int fe1_refresh(phyfoo_ctx* ctx, phyfoo_feset* fe) {
    ModelHost& h = fe->m->h;
    if (fe->g_version == h.version) return PHYFOO_OK;
    std::vector<double> tabs((size_t)fe->E1);
    for (int t = 0; t < fe->W; ++t) {
        h.build_tables1_into(t, h.pi[t].data(), tabs.data());
        int rc = fe1_launch(ctx, fe, t, 1, 0, 0, tabs);
        if (rc) return rc;
    }
    fe->g_version = h.version;
    return PHYFOO_OK;
}

// parameter sets of position t: set 0 = softmax(lambda_t), set 1 + i = softmax(lambda_t + eps e_i); launches the evaluation
static int fe1_position(phyfoo_ctx* ctx, phyfoo_feset* fe, int t, const double* lambda_t, bool with_grad, double eps, int value_off,
                        double* pi_out) {
    ModelHost& h = fe->m->h;
    const int dim = h.ctx_rows(t) * 4;
    const int K = with_grad ? dim + 1 : 1;
    std::vector<double> tabs((size_t)K * fe->E1), x(lambda_t, lambda_t + dim), pi(dim);
    lambda_to_pi(x.data(), pi_out, dim);
    h.build_tables1_into(t, pi_out, &tabs[0]);
    for (int i = 0; with_grad && i < dim; ++i) {
        const double keep = x[i];
        x[i] += eps;                               // NumericalDifferentiableFunction.java:73-79: perturb in place, restore
        lambda_to_pi(x.data(), pi.data(), dim);
        h.build_tables1_into(t, pi.data(), &tabs[(size_t)(1 + i) * fe->E1]);
        x[i] = keep;
    }
    for (double v : tabs) if (!std::isfinite(v)) return fail(ctx, PHYFOO_EUNSUPPORTED, "a CPF entry is 0 under the parameters under test: the free energy is not finite");
    return fe1_launch(ctx, fe, t, K, 1, value_off, tabs);
}

static int fe1_enqueue(phyfoo_ctx* ctx, phyfoo_feset* fe, int position, const double* lambda, int dim, bool with_grad, double eps) {
    ModelHost& h = fe->m->h;
    const int W = fe->W;
    if (position >= W) return fail(ctx, PHYFOO_EINVAL, "fe_eval: position out of range");
    int want = 0;
    if (position >= 0) want = h.ctx_rows(position) * 4;
    else for (int t = 0; t < W; ++t) want += h.ctx_rows(t) * 4;
    if (!lambda || dim != want) return fail(ctx, PHYFOO_EINVAL, "fe_eval: dim must be 4 x (context rows) per position under test");
    if (fe->n == 0) {
        CU(cudaMemsetAsync(fe->d_values, 0, (size_t)(1 + want) * sizeof(double), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<double> pi(16);
    if (position >= 0) {
        int rc = fe1_refresh(ctx, fe);
        if (rc) return rc;
        if (fe->n > 0 && (rc = fe1_position(ctx, fe, position, lambda, with_grad, eps, 1, pi.data()))) return rc;
        if (fe->n == 0) lambda_to_pi(lambda, pi.data(), want);
        h.set_pi(position, pi.data(), want);
    } else {
        // all positions (MotifParamOptimizer): move the model to lambda first so that every G_t is at the new base point,
        // then evaluate the perturbed sets position by position
        int off = 0;
        for (int t = 0; t < W; ++t) {
            const int d = h.ctx_rows(t) * 4;
            lambda_to_pi(lambda + off, pi.data(), d);
            h.set_pi(t, pi.data(), d);
            off += d;
        }
        int rc = fe1_refresh(ctx, fe);
        if (rc) return rc;
        off = 0;
        for (int t = 0; t < W && fe->n > 0; ++t) {
            const int d = h.ctx_rows(t) * 4;
            if (with_grad || t == W - 1) {
                if ((rc = fe1_position(ctx, fe, t, lambda + off, with_grad, eps, 1 + off, pi.data()))) return rc;
            }
            off += d;
        }
    }
    fe->g_version = h.version;
    return PHYFOO_OK;
}

static int fe_refresh(phyfoo_ctx* ctx, phyfoo_feset* fe) {
    ModelHost& h = fe->m->h;
    if (fe->order == 1) return fe1_refresh(ctx, fe);
    if (fe->g_version == h.version) return PHYFOO_OK;
    std::vector<double> tabs((size_t)fe->W * FE_KMAX * fe->E, 0.0);
    for (int t = 0; t < fe->W; ++t) h.build_tables_into(t, h.pi[t].data(), &tabs[(size_t)t * FE_KMAX * fe->E], nullptr);
    CU(cudaMemcpyAsync(fe->d_tabs, tabs.data(), tabs.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int rc = fe_launch(ctx, fe, 0, fe->W, 1, 0);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));   // tabs is stack-scoped
    fe->g_version = h.version;
    return PHYFOO_OK;
}

static int fe_create(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, int64_t n, phyfoo_feset** out) {
    phyfoo_feset* fe = new phyfoo_feset();
    fe->ds = ds; fe->m = m; fe->n = n;
    fe->W = m->h.W; fe->I = m->h.I; fe->S = m->h.S; fe->E = m->h.E;
    fe->order = m->h.order; fe->E1 = m->h.E1();
    fe->blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 2LL * ctx->sm_count));
    const size_t tab_doubles = fe->order == 1 ? (size_t)FE1_KMAX * fe->E1 : (size_t)fe->W * FE_KMAX * fe->E;
    const size_t part_doubles = fe->order == 1 ? (size_t)fe->blocks * FE1_KMAX : (size_t)fe->W * fe->blocks * FE_KMAX;
    const size_t value_doubles = fe->order == 1 ? (size_t)(1 + 4 + 16 * (fe->W - 1)) : (size_t)(1 + 4 * fe->W);
    const size_t n_trees = (size_t)std::max<int64_t>(n, 1) * fe->W;
    cudaError_t e;
    fe->stream = ctx->stream;
    const cudaStream_t st = ctx->stream;
    if ((e = dev_malloc(st, &fe->d_col0, std::max<int64_t>(n, 1) * sizeof(int64_t))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_weight, std::max<int64_t>(n, 1) * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_qint, n_trees * 4 * fe->I * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_qleaf, n_trees * 4 * fe->S * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_ent, n_trees * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_G, (fe->W + 1) * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_tabs, tab_doubles * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_part, part_doubles * sizeof(double))) != cudaSuccess ||
        (e = dev_malloc(st, &fe->d_values, value_doubles * sizeof(double))) != cudaSuccess) {
        fe_release(fe);
        ctx->err = std::string("fe_prepare: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? PHYFOO_ENOMEM : PHYFOO_ECUDA;
    }
    *out = fe;
    return PHYFOO_OK;
}

// runs the mean field of every listed window to convergence and stores q, the entropy terms, ENT and G_t
static int fe_run_prepare(phyfoo_ctx* ctx, phyfoo_feset* fe) {
    if (fe->n == 0) {
        CU(cudaMemsetAsync(fe->d_G, 0, (fe->W + 1) * sizeof(double), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        fe->g_version = fe->m->h.version;
        return PHYFOO_OK;
    }
    CU(ctx->fg_scores.ensure((size_t)fe->n * sizeof(double)));
    ScoreItems items{fe->d_col0, fe->d_strand, fe->d_qint, fe->d_qleaf, fe->d_ent};
    int rc = launch_score(ctx, fe->ds, fe->m, nullptr, fe->n, 0, ctx->fg_scores.as<double>(), nullptr, 2, &items);
    if (rc) return rc;
    {
        KernelTimer kt_(ctx, "fe_entropy_kernel");
        // order 0 stores one entropy per tree, order 1 one per window
        fe_entropy_kernel<<<1, 1024, 0, ctx->stream>>>(fe->d_ent, fe->d_weight, fe->n, fe->order == 1 ? 1 : fe->W, fe->d_G + fe->W);
    }
    CU(cudaGetLastError());
    fe->g_version = 0;
    return fe_refresh(ctx, fe);
}

static int fe_check_model(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const char* who) {
    if (m->h.order != 0 && !(m->h.order == 1 && m->h.kind == PHYFOO_MODEL_FG))
        return fail(ctx, PHYFOO_EUNSUPPORTED, std::string(who) + ": background models of order >= 1 are not implemented in this build");
    if (m->h.S != ds->S) return fail(ctx, PHYFOO_EINVAL, std::string(who) + ": model and dataset disagree on the number of species");
    return PHYFOO_OK;
}

extern "C" int phyfoo_fe_prepare(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, int64_t n_sel, const int32_t* sel_aln,
                                 const int32_t* sel_off, const uint8_t* sel_strand, const double* sel_weight, phyfoo_feset** out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !m || !out || n_sel < 0 || (n_sel > 0 && (!sel_aln || !sel_off || !sel_weight)))
        return fail(ctx, PHYFOO_EINVAL, "fe_prepare: NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    int rc = fe_check_model(ctx, ds, m, "fe_prepare");
    if (rc) return rc;
    const int W = m->h.W;
    std::vector<int64_t> col0((size_t)n_sel);
    for (int64_t u = 0; u < n_sel; ++u) {
        const int a = sel_aln[u];
        if (a < 0 || a >= ds->n_aln) return fail(ctx, PHYFOO_EINVAL, "fe_prepare: alignment index out of range");
        const int64_t L = ds->col_off[a + 1] - ds->col_off[a];
        if (sel_off[u] < 0 || sel_off[u] + W > L) return fail(ctx, PHYFOO_EINVAL, "fe_prepare: window leaves its alignment");
        if (sel_strand && sel_strand[u] > 1) return fail(ctx, PHYFOO_EINVAL, "fe_prepare: strand must be 0 or 1");
        col0[u] = ds->col_off[a] + sel_off[u];
    }
    phyfoo_feset* fe = nullptr;
    rc = fe_create(ctx, ds, m, n_sel, &fe);
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    if (n_sel > 0) {
        e = cudaMemcpyAsync(fe->d_col0, col0.data(), (size_t)n_sel * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(fe->d_weight, sel_weight, (size_t)n_sel * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && sel_strand) {
            e = dev_malloc(ctx->stream, &fe->d_strand, (size_t)n_sel);
            if (e == cudaSuccess) e = cudaMemcpyAsync(fe->d_strand, sel_strand, (size_t)n_sel, cudaMemcpyHostToDevice, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // col0 is stack-scoped
    }
    if (e != cudaSuccess) { fe_release(fe); ctx->err = std::string("fe_prepare: ") + cudaGetErrorString(e); return PHYFOO_ECUDA; }
    rc = fe_run_prepare(ctx, fe);
    if (rc) { fe_release(fe); return rc; }
    *out = fe;
    return PHYFOO_OK;
}

// item u = window u of width W (alignment-major): first column and the alignment's weight
__global__ void fe_all_items_kernel(const int64_t* __restrict__ win_off, const int64_t* __restrict__ col_off, int n_aln, int64_t n,
                                    const double* __restrict__ aln_w, int64_t* __restrict__ col0, double* __restrict__ weight) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int a = find_alignment(win_off, n_aln, u);
    col0[u] = col_off[a] + (u - win_off[a]);
    weight[u] = aln_w ? aln_w[a] : 1.0;
}

extern "C" int phyfoo_fe_prepare_all(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const double* aln_weights, phyfoo_feset** out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !m || !out) return fail(ctx, PHYFOO_EINVAL, "fe_prepare_all: NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    int rc = fe_check_model(ctx, ds, m, "fe_prepare_all");
    if (rc) return rc;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    rc = dataset_win_off(ctx, ds, m->h.W, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t n = h_woff->back();
    phyfoo_feset* fe = nullptr;
    rc = fe_create(ctx, ds, m, n, &fe);
    if (rc) return rc;
    if (n > 0) {
        const double* d_alnw = nullptr;
        cudaError_t e = cudaSuccess;
        if (aln_weights) {
            e = ctx->alnw.ensure((size_t)ds->n_aln * sizeof(double));
            if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->alnw.p, aln_weights, (size_t)ds->n_aln * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
            d_alnw = ctx->alnw.as<double>();
        }
        if (e == cudaSuccess) {
            {
                KernelTimer kt_(ctx, "fe_all_items_kernel");
                fe_all_items_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_woff, ds->d_col_off, ds->n_aln, n, d_alnw, fe->d_col0, fe->d_weight);
            }
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) { fe_release(fe); ctx->err = std::string("fe_prepare_all: ") + cudaGetErrorString(e); return PHYFOO_ECUDA; }
    }
    rc = fe_run_prepare(ctx, fe);
    if (rc) { fe_release(fe); return rc; }
    *out = fe;
    return PHYFOO_OK;
}

// Builds the parameter sets of one evaluation into d_tabs and enqueues the kernels; values land in fe->d_values:
// [f(lambda)] (with_grad = false) or [f(lambda), f(lambda + eps e_0), ...] (with_grad = true).
// pi_out receives softmax(lambda) per position under test.
static int fe_enqueue(phyfoo_ctx* ctx, phyfoo_feset* fe, int position, const double* lambda, int dim, bool with_grad, double eps,
                      std::vector<double>* pi_out) {
    ModelHost& h = fe->m->h;
    const int W = fe->W;
    if (fe->order == 1) return fe1_enqueue(ctx, fe, position, lambda, dim, with_grad, eps);
    const int pos0 = position >= 0 ? position : 0;
    const int n_pos = position >= 0 ? 1 : W;
    if (position >= W) return fail(ctx, PHYFOO_EINVAL, "fe_eval: position out of range");
    if (!lambda || dim != 4 * n_pos) return fail(ctx, PHYFOO_EINVAL, "fe_eval: dim must be 4 per position under test (order 0)");
    if (h.evol == PHYFOO_EVOL_HKY_AS_IMPLEMENTED)
        return fail(ctx, PHYFOO_EUNSUPPORTED, "HKY as implemented in the reference has zero transversion probabilities (HKY.java:32)");
    int rc = fe_refresh(ctx, fe);
    if (rc) return rc;
    const int K = with_grad ? 5 : 1;
    std::vector<double> tabs((size_t)n_pos * FE_KMAX * fe->E, 0.0), x(lambda, lambda + dim);
    pi_out->assign(dim, 0.0);
    for (int tp = 0; tp < n_pos; ++tp) {
        double* xt = &x[(size_t)4 * tp];
        double pi[4];
        lambda_to_pi(xt, &(*pi_out)[(size_t)4 * tp], 4);
        h.build_tables_into(pos0 + tp, &(*pi_out)[(size_t)4 * tp], &tabs[(size_t)tp * FE_KMAX * fe->E], nullptr);
        for (int i = 0; with_grad && i < 4; ++i) {
            const double keep = xt[i];
            xt[i] += eps;                           // NumericalDifferentiableFunction.java:73-79: perturb in place, restore
            lambda_to_pi(xt, pi, 4);
            h.build_tables_into(pos0 + tp, pi, &tabs[((size_t)tp * FE_KMAX + 1 + i) * fe->E], nullptr);
            xt[i] = keep;
        }
    }
    for (double v : tabs) if (!std::isfinite(v)) return fail(ctx, PHYFOO_EUNSUPPORTED, "a CPF entry is 0 under the parameters under test: the free energy is not finite");
    CU(cudaMemcpyAsync(fe->d_tabs + (size_t)pos0 * FE_KMAX * fe->E, tabs.data(), tabs.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (fe->n == 0) {
        CU(cudaMemsetAsync(fe->d_values, 0, (size_t)(1 + 4 * W) * sizeof(double), ctx->stream));
    } else {
        rc = fe_launch(ctx, fe, pos0, n_pos, K, 1);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));       // tabs is stack-scoped
    // leave the model at lambda (FE_byParameter_ForBayesNodeJstacs.java:43-53 writes the parameters into the net)
    for (int tp = 0; tp < n_pos; ++tp) h.set_pi(pos0 + tp, &(*pi_out)[(size_t)4 * tp], 4);
    fe->g_version = h.version;                     // fe_final_kernel stored the new G_t of the positions under test
    return PHYFOO_OK;
}

extern "C" int phyfoo_fe_eval(phyfoo_ctx* ctx, phyfoo_feset* fe, int32_t position, const double* lambda, int32_t dim, double* f_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!fe || !f_out) return fail(ctx, PHYFOO_EINVAL, "fe_eval: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    std::vector<double> pi;
    int rc = fe_enqueue(ctx, fe, position, lambda, dim, false, 0.0, &pi);
    if (rc) return rc;
    CU(cudaMemcpyAsync(f_out, fe->d_values, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_fe_grad(phyfoo_ctx* ctx, phyfoo_feset* fe, int32_t position, const double* lambda, int32_t dim, double eps,
                              double* f_out, double* g_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!fe || !g_out) return fail(ctx, PHYFOO_EINVAL, "fe_grad: NULL argument");
    if (!(eps > 0)) return fail(ctx, PHYFOO_EINVAL, "fe_grad: eps must be > 0");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    std::vector<double> pi;
    int rc = fe_enqueue(ctx, fe, position, lambda, dim, true, eps, &pi);
    if (rc) return rc;
    std::vector<double> v((size_t)dim + 1);
    CU(cudaMemcpyAsync(v.data(), fe->d_values, v.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (f_out) *f_out = v[0];
    for (int i = 0; i < dim; ++i) g_out[i] = (v[(size_t)i + 1] - v[0]) / eps;   // :80
    return PHYFOO_OK;
}

extern "C" int phyfoo_fe_values_device(phyfoo_ctx* ctx, phyfoo_feset* fe, int32_t position, const double* lambda, int32_t dim, double eps,
                                       double* values_device, void* stream) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!fe || !values_device) return fail(ctx, PHYFOO_EINVAL, "fe_values_device: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    std::vector<double> pi;
    int rc = fe_enqueue(ctx, fe, position, lambda, dim, eps > 0, eps, &pi);
    if (rc) return rc;
    const size_t nv = eps > 0 ? (size_t)dim + 1 : 1;
    cudaStream_t user = (cudaStream_t)stream;
    // fe_enqueue has synchronised the library's stream, so the values are complete: a plain copy on the caller's stream
    CU(cudaMemcpyAsync(values_device, fe->d_values, nv * sizeof(double), cudaMemcpyDeviceToDevice, user ? user : ctx->stream));
    if (!user) CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// sufficient statistics of the fixed-q objective (order 0): with q fixed,
//   f(theta) = sum_t sum_e C[t][e] * logtab_theta,t[e] - ENT,
// where C[t][e] are the expected counts of table entry e of position t (entry layout of model_host.hpp: 4 root
// entries, then 16 per non-root node [parent symbol][own symbol]) summed over the selected windows with their weights.
// One pass over the windows, then every later evaluation -- any parameter of any position, pi or branch length -- is
// a host-side dot product of W*E terms (SURVEY.md section 7 hard part 3; the reference re-sweeps all windows per evaluation).
// Accumulation uses shared-memory atomics per block (summation order within a block is not fixed: results agree
// to fp64 re-association run to run), block partials are reduced in a fixed order.
// ------------------------------------------------------------------------------------------------
struct FeCountParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col0; const uint8_t* strand; const double* weight;
    const double* q_int; const double* q_leaf;
    int64_t n;
    int W, I, S, E, prog_len;
    const int* prog;
    double* part;               // [W][gridDim.x][E]
};

__global__ void __launch_bounds__(256) fe_count_kernel(FeCountParams p) {
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int t = blockIdx.y;
    double* s_c = smem;                                     // [E]
    int* s_prog = (int*)(s_c + p.E);
    for (int i = tid; i < p.E; i += T) s_c[i] = 0.0;
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];
    __syncthreads();
    const int* par_internal = s_prog;
    const int* node_of = s_prog + p.I;
    const int* leaf_begin = s_prog + 2 * p.I;
    const int* leaf_node = s_prog + 3 * p.I + 1;
    const int* leaf_species = s_prog + 3 * p.I + 1 + p.S;
    const size_t n_trees = (size_t)p.n * p.W;
    for (int64_t u = (int64_t)blockIdx.x * T + tid; u < p.n; u += (int64_t)gridDim.x * T) {
        const int strand = p.strand ? p.strand[u] : 0;
        const int64_t c = strand ? p.col0[u] + (p.W - 1 - t) : p.col0[u] + t;
        ColWords cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
        if (strand) cw = complement(cw);
        const size_t tree = (size_t)t * p.n + u;
        const double w = p.weight[u];
        for (int i = 0; i < p.I; ++i) {
            double qi[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) qi[a] = p.q_int[(size_t)(i * 4 + a) * n_trees + tree];
            if (i == 0) {
#pragma unroll
                for (int a = 0; a < 4; ++a) atomicAdd(&s_c[a], w * qi[a]);                       // :282-291
            } else {
                const int e0 = tab_of(node_of[i]);
                const int pi_ = par_internal[i];
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const double wl = w * p.q_int[(size_t)(pi_ * 4 + l) * n_trees + tree];
#pragma unroll
                    for (int h = 0; h < 4; ++h) atomicAdd(&s_c[e0 + l * 4 + h], wl * qi[h]);     // :327-358
                }
            }
            const double qs = ((qi[0] + qi[1]) + qi[2]) + qi[3];
            for (int k = leaf_begin[i]; k < leaf_begin[i + 1]; ++k) {
                const int x = cw.obs(leaf_species[k]);
                if (x < 4) {
                    const int e0 = tab_of(leaf_node[k]);
#pragma unroll
                    for (int a = 0; a < 4; ++a) atomicAdd(&s_c[e0 + a * 4 + x], w * qi[a]);      // :292-326
                } else {
                    // "independent" table: every row is log pi -> counts go to the root entries
#pragma unroll
                    for (int a = 0; a < 4; ++a) atomicAdd(&s_c[a], (w * qs) * p.q_leaf[(size_t)(k * 4 + a) * n_trees + tree]);
                }
            }
        }
    }
    __syncthreads();
    double* out = p.part + ((size_t)t * gridDim.x + blockIdx.x) * p.E;
    for (int i = tid; i < p.E; i += T) out[i] = s_c[i];
}

// C[t][e] = sum over blocks of part (fixed order): one thread per (t, e)
__global__ void fe_count_reduce_kernel(const double* __restrict__ part, int blocks, int WE, int E, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= WE) return;
    const int t = i / E, e = i - t * E;
    double v = 0.0;
    for (int b = 0; b < blocks; ++b) v += part[((size_t)t * blocks + b) * E + e];
    out[i] = v;
}

extern "C" int phyfoo_fe_suffstats(phyfoo_ctx* ctx, phyfoo_feset* fe, double* counts_out, double* entropy_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!fe || !counts_out) return fail(ctx, PHYFOO_EINVAL, "fe_suffstats: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int W = fe->W, E = fe->order == 1 ? fe->E1 : fe->E;          // order 1: the entries of build_tables1_into (mstep1.cuh)
    const size_t WE = (size_t)W * E;
    if (fe->n == 0) {
        for (size_t i = 0; i < WE; ++i) counts_out[i] = 0.0;
        if (entropy_out) *entropy_out = 0.0;
        return PHYFOO_OK;
    }
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((fe->n + 255) / 256, (int64_t)ctx->sm_count));
    CU(ctx->misc.ensure(((size_t)blocks + 1) * WE * sizeof(double)));
    double* d_part = ctx->misc.as<double>();
    double* d_c = d_part + (size_t)blocks * WE;
    const phyfoo_dataset* ds = fe->ds;
    FeCountParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col0 = fe->d_col0; p.strand = fe->d_strand; p.weight = fe->d_weight;
    p.q_int = fe->d_qint; p.q_leaf = fe->d_qleaf; p.n = fe->n;
    p.W = W; p.I = fe->I; p.S = fe->S; p.E = E; p.prog_len = fe->m->prog_len; p.prog = fe->m->d_prog;
    p.part = d_part;
    const size_t smem = (size_t)E * sizeof(double) + (size_t)fe->m->prog_len * sizeof(int) + 16;
    if (fe->order == 1) {
        Fe1CountParams p1;
        p1.bases = p.bases; p1.gaps = p.gaps; p1.n_cols = p.n_cols; p1.nb = p.nb; p1.col0 = p.col0; p1.strand = p.strand; p1.weight = p.weight;
        p1.q_int = p.q_int; p1.q_leaf = p.q_leaf; p1.n = p.n; p1.W = W; p1.I = fe->I; p1.S = fe->S; p1.E1 = E; p1.prog_len = p.prog_len;
        p1.prog = p.prog; p1.part = d_part;
        KernelTimer kt_(ctx, "fe1_count_kernel");
        fe1_count_kernel<<<dim3(blocks, W), 128, smem, ctx->stream>>>(p1);
    } else {
        KernelTimer kt_(ctx, "fe_count_kernel");
        fe_count_kernel<<<dim3(blocks, W), 256, smem, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    {
        KernelTimer kt_(ctx, "fe_count_reduce_kernel");
        fe_count_reduce_kernel<<<(unsigned)((WE + 127) / 128), 128, 0, ctx->stream>>>(d_part, blocks, (int)WE, E, d_c);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts_out, d_c, WE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (entropy_out) CU(cudaMemcpyAsync(entropy_out, fe->d_G + W, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int64_t phyfoo_fe_size(const phyfoo_feset* fe) { return fe ? fe->n : 0; }

extern "C" void phyfoo_fe_free(phyfoo_ctx* ctx, phyfoo_feset* fe) {
    if (!fe) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    fe_release(fe);
}
