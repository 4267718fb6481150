// phyfoo.cu -- libphyfoo.so: the C ABI of include/phyfoo.h on one B200 (sm_100a).
//
// Kernels (all fp64, no tensor cores: the contractions are 4x4 per tree edge with data-dependent
// leaf selection -- SURVEY.md section 2.2):
//   pack_columns_kernel      K0  codes [cols][S] u8 -> 2-bit base planes + gap planes (bit-exact layout)
//   score_o0_kernel<MODE>    K1/K2/K4  persistent kernel; W consecutive threads = one window (one
//                            thread per motif position = one tree); the groups of a block pull windows
//                            from a global counter, so a window that needs 14 sweeps never holds back
//                            one that needs 3; tables + tree program + per-thread state in shared memory
//   zoops_kernel             K3  one block per alignment: a_j, max, log-sum-exp, gamma, w, ll, arg-max
//                            (the block that finishes last sums the per-alignment partials -> [ll, w0, w1])
//   select_kernel, fe_*      M-step (data selection, fixed-q free energy sums, forward differences)
//   dfma_peak_kernel         fp64 FMA roofline probe
//
// No CPU fallback anywhere: every entry point needs the CUDA device the context was created on.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "model_host.hpp"
#include "tree_pass.cuh"

using namespace phyfoo;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
// stream-ordered allocations from the device's default memory pool (kept warm: release threshold = unlimited), so that
// creating and freeing a data set per call costs microseconds instead of synchronous cudaMalloc / cudaFree
template <class T> static cudaError_t dev_malloc(cudaStream_t st, T** p, size_t bytes) { return cudaMallocAsync((void**)p, bytes ? bytes : 1, st); }
static void dev_free(cudaStream_t st, void* p) { if (p) cudaFreeAsync(p, st); }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct JitKernel;      // jit_o0.hpp

struct phyfoo_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t total_mem = 0, smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t kev[4] = {nullptr, nullptr, nullptr, nullptr};   // brackets of the bg / fg / zoops kernels of the last E-step
    bool kev_valid = false;
    std::string err;
    // scratch (grown on demand, reused across calls)
    DevBuf counters;                 // unsigned long long [8]
    DevBuf fg_scores, bg_scores;     // window scores (both strands) / column scores
    DevBuf sweeps_buf;               // int32 per item when the caller wants K
    DevBuf gamma, profile, part, argoff, argstrand, alnw, out3, gstate, misc;
    DevBuf selw, sel_tmp, sel_out;   // M-step selection scratch
    int64_t gamma_windows = -1;      // number of windows whose gamma of the last E-step is resident in `gamma` ...
    const phyfoo_dataset* gamma_ds = nullptr; int gamma_width = 0;   // ... and the data set and motif width it belongs to
    DevBuf memo_traj;                // pattern-table path: free-energy trajectories (memo.cuh)
    int memo_mode = 1;               // 0 = never, 1 = when it pays off, 2 = whenever it applies
    int memo_sweep_cap = 64;
    int memo_wave = 1;               // trajectories by the wavefront kernel when the tree allows it (0: one quad per tree)
    int memo_split = 0;              // > 0: first trajectory phase (passes); the rest runs on `aux` under the first window pass
    cudaStream_t aux = nullptr;      // second trajectory phase of the pattern-table path
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    DevBuf memo_state;               // mean-field state of every (pattern, position) tree between the two phases
    int bg_generic = 0;              // 1: an order-1 background goes through the generic net kernel too (netg.cuh; cross-checks)
    int order1_kernel = 2;           // 0 = one quad per window (net1.cuh), 1 = wavefront, 16 lanes per window (net1w.cuh),
                                     // 2 = one thread per node (net1n.cuh; windows with gaps through the wavefront kernel)
    DevBuf memo_pending;             // fallback list of the pattern-table path
    int order0_kernel = -1;          // order-0 mean-field scoring: -1 = choose (F81-structured kernel for more than 6 species, where the
                                     // pattern table does not apply), 0 = generic kernel (tree_pass.cuh), 1 = F81-structured kernel
    std::map<std::string, std::shared_ptr<JitKernel>> jit_cache;   // generated source -> loaded kernel (jit_o0.hpp)
    std::string jit_note;            // why the last request for a run-time compiled kernel was not served
    int jit_max_threads = 512;       // block size cap of the run-time compiled kernel (tuning)
    int l2_persist = 1;              // per-window state kept in global memory (order-1 q rows, large-tree state): L2 access-policy window
    size_t l2_persist_max = 0, l2_window_max = 0;
    int order1_qglobal = -1;         // node-per-thread kernel: -1 = choose, 0 = q rows in shared memory, n > 0 = q rows in global memory
                                     // behind L1 with n warps per SM (<= 8)
    DevBuf o1n_cache, o1n_fb;
    DevBuf multi_scores[8], multi_gamma[8];   // K-motif mixture: window scores and gamma per component (resident between E- and M-step)
    int64_t multi_windows[8] = {0}; int multi_width[8] = {0}; int multi_k = 0; const phyfoo_dataset* multi_ds = nullptr;        // node-per-thread order-1 kernel: observed-leaf sums per warp; list of windows with gaps
    // per-kernel device times of the most recent call (CUDA events on the library's stream)
    struct KTime { const char* name; cudaEvent_t e0, e1; };
    std::vector<KTime> ktimes;
    int n_ktimes = 0;
    int last_path = 0;               // scoring path of the most recent motif/background launch: 0 direct, 1 pattern table, 2 order-1 net
    int launches = 0;                // kernels launched by the current call
};

struct MemoSet;
struct phyfoo_dataset {
    MemoSet* memo = nullptr;                     // pattern set (memo.cuh), built on first use
    int n_aln = 0, S = 0, nb = 0, ng = 0;
    int64_t n_cols = 0;
    int64_t n_gap_cols = 0;                      // columns with at least one unobserved symbol
    int min_len = 0, max_len = 0;
    std::vector<int64_t> col_off;
    int64_t* d_col_off = nullptr;
    uint32_t* d_bases = nullptr;
    uint32_t* d_gaps = nullptr;
    std::map<int, int64_t*> d_win_off;           // per motif width
    std::map<int, std::vector<int64_t>> win_off;
};

struct phyfoo_model {
    ModelHost h;
    uint64_t uploaded = 0;
    int* d_prog = nullptr;
    int prog_len = 0;
    double* d_logtab = nullptr;   // [E][W]
    double* d_probtab = nullptr;  // [E][W]
    double* d_tab1 = nullptr;     // order 1: [W][E1] log tables (net1.cuh)
    double* d_tab1w = nullptr;    // order 1, wavefront kernel (net1w.cuh)
    int* d_prog1w = nullptr;
    int prog1w_len = 0, MC = 1, ML = 1, n_steps = 0;
    double* d_tabf = nullptr;     // order 0, compact tables of the F81-structured kernel (tree_pass_f81.cuh): [Ef][W]
    bool f81_ok = false;          // the tables have F81 structure and are finite (else the generic kernel scores)
    double* d_tab1n = nullptr;    // order 1, node-per-thread kernel (net1n.cuh): internal-node tables [W][I][32]
    double* d_leaftab1n = nullptr;//                                              leaf tables [W][S][32]
    int* d_prog1n = nullptr;
    int prog1n_len = 0, MC1n = 1, n_steps1n = 0;
    double f0_const1n = 0.0;
    // generic net (netg.cuh): background models of order >= 2
    int* d_netg_rec = nullptr; int* d_netg_lists = nullptr; double* d_netg_tab = nullptr;
    size_t netg_lists_len = 0, netg_tab_len = 0, netg_tab_used = 0;
    uint64_t netg_uploaded = 0;
};

static thread_local std::string g_init_err;

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return PHYFOO_ECUDA;                                                                   \
        }                                                                                          \
    } while (0)

static void memo_release(cudaStream_t st, MemoSet* ms);

// brackets one kernel launch with events so that callers (bench.py) can attribute device time per kernel
struct KernelTimer {
    phyfoo_ctx* ctx; int slot;
    cudaStream_t stream;
    KernelTimer(phyfoo_ctx* c, const char* name, cudaStream_t on = nullptr) : ctx(c), slot(-1), stream(on ? on : c->stream) {
        if (ctx->n_ktimes >= 64) return;
        if ((int)ctx->ktimes.size() <= ctx->n_ktimes) {
            phyfoo_ctx::KTime kt{name, nullptr, nullptr};
            if (cudaEventCreate(&kt.e0) != cudaSuccess || cudaEventCreate(&kt.e1) != cudaSuccess) return;
            ctx->ktimes.push_back(kt);
        }
        slot = ctx->n_ktimes++;
        ctx->ktimes[slot].name = name;
        cudaEventRecord(ctx->ktimes[slot].e0, stream);
    }
    ~KernelTimer() {
        if (slot >= 0) cudaEventRecord(ctx->ktimes[slot].e1, stream);
        ctx->launches++;
    }
};

static int fail(phyfoo_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_init_err = msg;
    return code;
}

// ------------------------------------------------------------------------------------------------
// K0: packing.  One thread per column; bit-exact contract in include/phyfoo.h.
// Replaces the int[pos][species] container of jstacs MultiDimensionalDiscreteSequence.java:46-52.
// ------------------------------------------------------------------------------------------------
__global__ void pack_columns_kernel(const uint8_t* __restrict__ codes, int64_t n_cols, int S, int nb, int ng,
                                    uint32_t* __restrict__ bases, uint32_t* __restrict__ gaps, int* __restrict__ bad) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const uint8_t* col = codes + c * S;
    for (int k = 0; k < nb; ++k) {
        uint32_t w = 0;
        for (int s = 16 * k; s < min(S, 16 * k + 16); ++s) {
            uint8_t v = col[s];
            if (v < 4) w |= (uint32_t)v << (2 * (s & 15));
            else if (v > 4) *bad = 1;  // WrongAlphabetException territory; caller reports EINVAL
        }
        bases[(int64_t)k * n_cols + c] = w;
    }
    uint32_t any = 0;
    for (int k = 0; k < ng; ++k) {
        uint32_t w = 0;
        for (int s = 32 * k; s < min(S, 32 * k + 32); ++s)
            if (col[s] >= 4) w |= 1u << (s & 31);
        gaps[(int64_t)k * n_cols + c] = w;
        any |= w;
    }
    // bad[1]: number of columns with an unobserved symbol (one atomic per warp); the order-1 kernels choose their path by it
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, any != 0);
    if (m && (threadIdx.x & 31) == __ffs(act) - 1) atomicAdd(bad + 1, __popc(m));
}

// ------------------------------------------------------------------------------------------------
// K1/K2/K4: order-0 window scoring
// ------------------------------------------------------------------------------------------------
struct ScoreParams {
    const uint32_t* bases; const uint32_t* gaps; int64_t n_cols; int nb;
    const int64_t* col_off; const int64_t* win_off; int n_aln;
    int64_t n_windows; int n_strands; int strand0;
    int W, I, S, E, prog_len;
    const double* tab;   // [E][W]
    const int* prog;
    double* out;         // [n_strands][n_windows]
    int32_t* sweeps;     // nullable, same shape
    unsigned long long* counter_next;    // work counter (next item)
    unsigned long long* counter_sweeps;  // running sum of the sweep counts
    int max_sweeps, min_sweeps; double tol;
    double* gstate;      // nullable: per-thread state in global memory instead of shared
    int groups;          // windows in flight per block
    // M-step "prepare": explicit window list instead of all windows, and a sink for the converged q
    const int64_t* item_col0;     // nullable: first column of window `item`
    const uint8_t* item_strand;   // nullable (all forward)
    double* q_int; double* q_leaf; double* ent;   // nullable; strided by n_windows * W (one slot per tree)
    const int* n_items_dev;       // nullable: number of listed windows lives on the device (pattern-table fallback list)
    const int64_t* item_dst;      // nullable: out/sweeps index of listed window `item` (default: item)
};

#include "score_common.cuh"

static const int SCORE_CHUNK = 4;   // windows a group takes from the global counter at a time

// ITEMS = true: M-step "prepare" variant (explicit window list, converged q written out)
template <int MODE, bool ITEMS>
__global__ void __launch_bounds__(256, 2) score_o0_kernel(ScoreParams p) {
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    double* s_tab = smem;
    double* s_F = s_tab + (size_t)p.E * p.W;
    long long* s_item = (long long*)(s_F + T);
    double* s_state = (double*)(s_item + p.groups);
    int* s_prog = (int*)(s_state + (p.gstate ? 0 : (size_t)p.I * 8 * T));

    for (int i = tid; i < p.E * p.W; i += T) s_tab[i] = p.tab[i];
    for (int i = tid; i < p.prog_len; i += T) s_prog[i] = p.prog[i];

    TreeProg tp;
    tp.n_internal = p.I; tp.n_leaves = p.S; tp.n_nodes = 0;
    tp.par_internal = s_prog;
    tp.node_of = s_prog + p.I;
    tp.leaf_begin = s_prog + 2 * p.I;
    tp.leaf_node = s_prog + 3 * p.I + 1;
    tp.leaf_species = s_prog + 3 * p.I + 1 + p.S;

    const int g = tid / p.W, t = tid - g * p.W;
    const bool lane_ok = g < p.groups;
    const Tab tb{s_tab + t, p.W};
    State st;
    if (p.gstate) { st.base = p.gstate + ((size_t)blockIdx.x * T + tid); st.stride = gridDim.x * T; }
    else { st.base = s_state + tid; st.stride = T; }

    const long long n_items = (ITEMS && p.n_items_dev) ? (long long)*p.n_items_dev : (long long)p.n_windows * p.n_strands;
    bool need = lane_ok, active = lane_ok, upd = false, conv = false;
    int k = 0;
    long long item = -1, chunk_end = 0;     // the group owns items [item + 1, chunk_end) of its current chunk
    double old = 0.0;
    ColWords cw; cw.b = 0; cw.g = 0;

    for (;;) {
        const bool refill = need && active && item + 1 >= chunk_end;      // uniform within a group
        if (refill && t == 0) s_item[g] = (long long)atomicAdd(p.counter_next, (unsigned long long)SCORE_CHUNK);
        // one barrier publishes the fetched chunk AND votes on the exit; a thread that runs out of items in this iteration
        // is counted as active once more, which costs the block one idle pass at the very end
        const int any_active = __syncthreads_or(active ? 1 : 0);
        if (need && active) {
            if (refill) { item = s_item[g]; chunk_end = item + SCORE_CHUNK; }
            else ++item;
            if (item >= n_items) active = false;
            else {
                int strand;
                int64_t c0;
                if (ITEMS) {
                    c0 = p.item_col0[item];
                    strand = p.item_strand ? p.item_strand[item] : 0;
                } else {
                    const int si = (int)(item / p.n_windows);
                    const int64_t w = item - (long long)si * p.n_windows;
                    strand = p.n_strands == 2 ? si : p.strand0;
                    const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
                    c0 = p.col_off[a] + (w - p.win_off[a]);
                }
                // jstacs StrandModel.java:636-639 / Sequence.java:1321-1333: reverse the columns, complement the symbols
                const int64_t c = strand ? c0 + (p.W - 1 - t) : c0 + t;
                cw = load_column(p.bases, p.gaps, p.n_cols, p.nb, c);
                if (strand) cw = complement(cw);
                k = 0; conv = false; upd = false; need = false;
            }
        }
        if (!any_active) break;
        double F = 0.0;
        if (active) {
            if (MODE == PHYFOO_MODE_EXACT) F = -log(pruning_likelihood(tp, tb, st, cw));
            else F = meanfield_pass(tp, tb, st, cw, upd);
        }
        s_F[tid] = F;
        __syncthreads();
        if (active) {
            double Fsum = 0.0;
            const double* fp = s_F + g * p.W;
            for (int m = 0; m < p.W; ++m) Fsum += fp[m];   // tree-major, like calcFreeEnergy(0, W-1)
            bool done;
            if (MODE == PHYFOO_MODE_EXACT) done = true;
            else if (!upd) { old = Fsum; upd = true; done = false; }          // F_0, MeanFieldForBayesNet.java:105
            else {
                if (fabs(old - Fsum) < p.tol) conv = true;                    // :204-207 (sticky)
                old = Fsum; ++k;
                done = !((k < p.max_sweeps && !conv) || k < p.min_sweeps);    // :107
            }
            if (done) {
                if (MODE != PHYFOO_MODE_EXACT && ITEMS && p.q_int) {
                    // FE-set layout: tree slot = position-major (t * n + window) so one position's trees are contiguous
                    const size_t n_trees = (size_t)p.n_windows * p.W, tree = (size_t)t * p.n_windows + item;
                    const QSink sink{p.q_int + tree, p.q_leaf + tree, n_trees};
                    p.ent[tree] = store_q(tp, tb, st, cw, sink);
                }
                if (t == 0) {
                    const long long dst = (ITEMS && p.item_dst) ? (long long)p.item_dst[item] : item;
                    p.out[dst] = -Fsum;                                       // PhyloBayesModel.java:259
                    if (p.sweeps) p.sweeps[dst] = k;
                    if (MODE != PHYFOO_MODE_EXACT) atomicAdd(p.counter_sweeps, (unsigned long long)k);
                }
                need = true;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3: ZOOPS posterior, one block per alignment.
// jstacs SingleHiddenMotifMixture.java:520-564 (getNewWeights), :585-596 (modify),
// UniformPositionPrior.java:55-61, StrandModel.java:631-640, Normalisation.java:63-91,352-375.
// ------------------------------------------------------------------------------------------------
struct ZoopsParams {
    const double* fwd; const double* rev;    // [n_windows] each (rev may be null)
    const double* bgcol;                     // [n_cols] background log-score per column (order 0); order 1: the conditional terms
    // order-1 background (nullable): start terms per column, single-column ranges bg(j + W, e) per column and bg(s, j - 1) per
    // window (bg1_terms_device); a_j = prior + motif - bg(s, e) + bg(s, j - 1) + bg(j + W, e), s = max(j - 1, 0), e = min(j + W, L - 1)
    const double* bg1_first; const double* bg1_single; const double* bg1_left;
    const double* bgk_right; int bg_order;   // generic path (any order k): the right flank per offset; s = max(j-k, 0), e = min(j+W+k-1, L-1)
    const int64_t* col_off; const int64_t* win_off; int n_aln; int W;
    double logw0, logw1; int has_strand; double slw0, slw1;
    const double* aln_w;                     // nullable
    double* gamma; double* profile;          // nullable, [n_windows]
    double* part;                            // [3][n_aln]: ll, w0, w1 per alignment
    int32_t* argoff; uint8_t* argstrand;     // nullable
    double* out3;                            // [ll, w0, w1] summed over the alignments by the block that finishes last
    unsigned int* ticket;                    // zero between launches
    int stage;                               // 1: dynamic shared memory holds [max_len] background terms + [max_len] a_j
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
    // deterministic: warp shuffles in a fixed pattern, then warp 0 adds the warp sums in order
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += sh[i];
    return r;
}

__device__ __forceinline__ double logsum2(double v0, double v1) {   // Normalisation.getLogSum, :63-91
    const double off = fmax(v0, v1);
    if (isinf(off) && off < 0) return off;
    return off + log(exp(v0 - off) + exp(v1 - off));
}

__global__ void __launch_bounds__(256) zoops_kernel(ZoopsParams p) {
    __shared__ double sh[32];
    __shared__ double sh_v[8];
    __shared__ int sh_i[8];
    const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int64_t c0 = p.col_off[a];
    const int L = (int)(p.col_off[a + 1] - c0);
    const int64_t w0 = p.win_off[a];
    const int n = (int)(p.win_off[a + 1] - w0);
    const double weight = p.aln_w ? p.aln_w[a] : 1.0;

    // the alignment's background terms (each is read W times) and its a_j (read twice more) staged in shared memory
    extern __shared__ double s_stage[];
    double* s_bg = s_stage;
    double* s_aj = s_stage + L;
    const double* bgc = p.bgcol + c0;
    if (p.stage) {
        for (int u = tid; u < L; u += T) s_bg[u] = p.bgcol[c0 + u];
        __syncthreads();
        bgc = s_bg;
    }
    // logPSeq = bg.getLogProbFor(seq, 0, l-1), :526 (order 1: start term of column 0 + the conditional terms of the others)
    const bool o1 = p.bg1_first != nullptr;
    double ps = 0.0;
    for (int u = tid; u < L; u += T) ps += (o1 && u == 0) ? p.bg1_first[c0] : bgc[u];
    const double log_pseq = block_sum(ps, sh);

    // a_j and its maximum (first index wins: util/Util.java:528-538)
    const double lprior = -log((double)n);
    double best = -INFINITY;
    int best_j = 0x7fffffff;
    for (int j = tid; j < n; j += T) {
        double motif = p.fwd[w0 + j];
        if (p.has_strand) motif = logsum2(p.slw0 + motif, p.slw1 + p.rev[w0 + j]);
        double bgw = 0.0;
        if (!o1) for (int m = 0; m < p.W; ++m) bgw += bgc[j + m];    // bg(s,e), order 0: s=j, e=j+W-1
        else {
            const int ko = p.bg_order;
            const int s = j > ko ? j - ko : 0, e = j + p.W + ko - 1 < L - 1 ? j + p.W + ko - 1 : L - 1;
            bgw = p.bg1_first[c0 + s];
            for (int u = s + ko; u <= e; ++u) bgw += bgc[u];
            if (j > 0) bgw -= p.bg1_left[w0 + j];                     // + bg(s, j-1)
            if (j + p.W <= L - 1) bgw -= p.bgk_right ? p.bgk_right[w0 + j] : p.bg1_single[c0 + j + p.W];  // + bg(j+W, e)
        }
        const double aj = lprior + motif - bgw;                       // :536-540
        if (p.profile) p.profile[w0 + j] = aj;
        if (p.stage) s_aj[j] = aj;
        else if (p.gamma) p.gamma[w0 + j] = aj;                       // overwritten with gamma_j below
        if (aj > best) { best = aj; best_j = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, o);
        const int oj = __shfl_down_sync(0xffffffffu, best_j, o);
        if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
    }
    if ((tid & 31) == 0) { sh_v[tid >> 5] = best; sh_i[tid >> 5] = best_j; }
    __syncthreads();
    best = sh_v[0]; best_j = sh_i[0];
    for (int i = 1; i < (T + 31) / 32; ++i)
        if (sh_v[i] > best || (sh_v[i] == best && sh_i[i] < best_j)) { best = sh_v[i]; best_j = sh_i[i]; }
    if (best_j == 0x7fffffff) best_j = 0;   // every a_j was NaN or -inf

    // sum_j exp(a_j - max)   (Normalisation.logSumNormalisation :352-375)
    double se = 0.0;
    for (int j = tid; j < n; j += T) {
        double aj;
        if (p.stage) aj = s_aj[j];                                    // (written by this thread in the first pass)
        else if (p.gamma) aj = p.gamma[w0 + j];
        else {
            double motif = p.fwd[w0 + j];
            if (p.has_strand) motif = logsum2(p.slw0 + motif, p.slw1 + p.rev[w0 + j]);
            double bgw = 0.0;
            if (!o1) for (int m = 0; m < p.W; ++m) bgw += p.bgcol[c0 + j + m];
            else {
                const int ko = p.bg_order;
                const int s = j > ko ? j - ko : 0, e = j + p.W + ko - 1 < L - 1 ? j + p.W + ko - 1 : L - 1;
                bgw = p.bg1_first[c0 + s];
                for (int u = s + ko; u <= e; ++u) bgw += p.bgcol[c0 + u];
                if (j > 0) bgw -= p.bg1_left[w0 + j];
                if (j + p.W <= L - 1) bgw -= p.bgk_right ? p.bgk_right[w0 + j] : p.bg1_single[c0 + j + p.W];
            }
            aj = lprior + motif - bgw;
        }
        const double e = exp(aj - best);
        if (p.gamma) p.gamma[w0 + j] = e;
        se += e;
    }
    const double sum = block_sum(se, sh);
    const double Z = best + log(sum);

    // modify(): help = {logW0 + Z, logW1} -> logSumNormalisation, :588-590
    double h0 = p.logw0 + Z, h1 = p.logw1;
    const double off = fmax(h0, h1);
    double e0 = exp(h0 - off), e1 = exp(h1 - off);
    const double hs = e0 + e1;
    if (hs != 1.0) { e0 /= hs; e1 /= hs; }
    const double lse = off + log(hs);
    const double pm = e0 * weight;                                    // :547
    if (tid == 0) {
        p.part[a] = weight * (log_pseq + lse);                        // :544
        p.part[(size_t)p.n_aln + a] = pm;                             // :548
        p.part[(size_t)2 * p.n_aln + a] = weight * e1;                // :549
        if (p.argoff) p.argoff[a] = best_j;
        if (p.argstrand) {
            int s = 0;
            if (p.has_strand) s = (p.slw1 + p.rev[w0 + best_j]) > (p.slw0 + p.fwd[w0 + best_j]) ? 1 : 0;  // :737-754
            p.argstrand[a] = (uint8_t)s;
        }
    }
    if (p.gamma) {
        for (int j = tid; j < n; j += T) {
            double e = p.gamma[w0 + j];
            if (sum != 1.0) e /= sum;
            p.gamma[w0 + j] = e * pm;                                 // :553-555
        }
    }
    // the block that finishes last adds the per-alignment partials in a fixed pattern (deterministic for a given n_aln):
    // no second launch for three numbers
    __shared__ int s_last;
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = tid; i < p.n_aln; i += T) {
#pragma unroll
            for (int r = 0; r < 3; ++r) v[r] += __ldcg(p.part + (size_t)r * p.n_aln + i);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double s3 = block_sum(v[r], sh);
            if (tid == 0) p.out3[r] = s3;
            __syncthreads();
        }
        if (tid == 0) *p.ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// fp64 roofline probe: 8 independent DFMA chains per thread
// ------------------------------------------------------------------------------------------------
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int phyfoo_init(int device, phyfoo_ctx** out) {
    if (!out) return fail(nullptr, PHYFOO_EINVAL, "phyfoo_init: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, PHYFOO_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                               " (libphyfoo has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, PHYFOO_EINVAL, "phyfoo_init: bad device index");
    phyfoo_ctx* ctx = new phyfoo_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->kev[0])) != cudaSuccess || (e = cudaEventCreate(&ctx->kev[1])) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->kev[2])) != cudaSuccess || (e = cudaEventCreate(&ctx->kev[3])) != cudaSuccess ||
        (e = ctx->counters.ensure(16 * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(ctx->counters.p, 0, 16 * sizeof(unsigned long long))) != cudaSuccess) {
        g_init_err = std::string("phyfoo_init: ") + cudaGetErrorString(e);
        delete ctx;
        return PHYFOO_ECUDA;
    }
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major; ctx->cc_minor = prop.minor;
    ctx->total_mem = prop.totalGlobalMem;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    // L2 set-aside for the kernels that keep per-window state in global memory: without it their ~40-80 MB of rows, rewritten
    // every sweep, keep falling out of L2 (ncu on C3: 350 GB written to DRAM per E-step against 128 MB of algorithmic bytes)
    ctx->l2_persist_max = (size_t)std::max(prop.persistingL2CacheMaxSize, 0);
    ctx->l2_window_max = (size_t)std::max(prop.accessPolicyMaxWindowSize, 0);
    if (getenv("PHYFOO_VERBOSE")) fprintf(stderr, "phyfoo: L2 %d bytes, persisting max %d, window max %d\n", prop.l2CacheSize, prop.persistingL2CacheMaxSize, prop.accessPolicyMaxWindowSize);
    if (ctx->l2_persist_max) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max);
    *out = ctx;
    return PHYFOO_OK;
}

extern "C" void phyfoo_destroy(phyfoo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->aux) cudaStreamSynchronize(ctx->aux);
    ctx->memo_state.release();
    for (DevBuf* b : {&ctx->counters, &ctx->fg_scores, &ctx->bg_scores, &ctx->sweeps_buf, &ctx->gamma, &ctx->profile, &ctx->part,
                      &ctx->argoff, &ctx->argstrand, &ctx->alnw, &ctx->out3, &ctx->gstate, &ctx->misc, &ctx->selw, &ctx->sel_tmp, &ctx->sel_out, &ctx->memo_traj, &ctx->memo_pending,
                      &ctx->memo_state, &ctx->o1n_cache, &ctx->o1n_fb})
        b->release();
    for (int k = 0; k < 8; ++k) { ctx->multi_scores[k].release(); ctx->multi_gamma[k].release(); }
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->kev) cudaEventDestroy(e);
    for (auto& kt : ctx->ktimes) { cudaEventDestroy(kt.e0); cudaEventDestroy(kt.e1); }
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* phyfoo_last_error(const phyfoo_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

extern "C" int phyfoo_device_info(phyfoo_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem) {
    if (!ctx) return PHYFOO_EINVAL;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (total_mem) *total_mem = (int64_t)ctx->total_mem;
    return PHYFOO_OK;
}

extern "C" int phyfoo_fp64_peak(phyfoo_ctx* ctx, double* tflops_out) {
    if (!ctx || !tflops_out) return PHYFOO_EINVAL;
    CU(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 20000;
    CU(ctx->misc.ensure((size_t)blocks * threads * sizeof(double)));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->misc.as<double>(), iters, 0.999999, 1e-6);
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev1));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    *tflops_out = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// dataset
// ------------------------------------------------------------------------------------------------
static int dataset_build(phyfoo_ctx* ctx, int32_t n_aln, int32_t S, const int64_t* col_offsets, const uint8_t* codes,
                         bool codes_on_device, phyfoo_dataset** out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!out || !col_offsets || (!codes && n_aln > 0)) return fail(ctx, PHYFOO_EINVAL, "dataset_create: NULL argument");
    *out = nullptr;
    if (n_aln < 0 || S < 1) return fail(ctx, PHYFOO_EINVAL, "dataset_create: need n_aln >= 0 and n_species >= 1");
    if (S > 32) return fail(ctx, PHYFOO_EUNSUPPORTED, "dataset_create: more than 32 species are not supported by the packed-column kernels");
    if (col_offsets[0] != 0) return fail(ctx, PHYFOO_EINVAL, "dataset_create: col_offsets[0] must be 0");
    for (int i = 0; i < n_aln; ++i)
        if (col_offsets[i + 1] < col_offsets[i]) return fail(ctx, PHYFOO_EINVAL, "dataset_create: col_offsets must be non-decreasing");
    CU(cudaSetDevice(ctx->device));
    phyfoo_dataset* ds = new phyfoo_dataset();
    ds->n_aln = n_aln; ds->S = S; ds->nb = (S + 15) / 16; ds->ng = (S + 31) / 32;
    ds->col_off.assign(col_offsets, col_offsets + n_aln + 1);
    ds->n_cols = col_offsets[n_aln];
    ds->min_len = n_aln ? INT32_MAX : 0; ds->max_len = 0;
    for (int i = 0; i < n_aln; ++i) {
        int64_t L = col_offsets[i + 1] - col_offsets[i];
        if (L > INT32_MAX / 2) { delete ds; return fail(ctx, PHYFOO_EINVAL, "dataset_create: alignment too long"); }
        ds->min_len = std::min<int>(ds->min_len, (int)L);
        ds->max_len = std::max<int>(ds->max_len, (int)L);
    }
    const int64_t nc = std::max<int64_t>(ds->n_cols, 1);
    cudaError_t e;
    uint8_t* d_codes = nullptr;
    int* d_bad = nullptr;
    auto cleanup = [&]() {
        if (!codes_on_device && d_codes) dev_free(ctx->stream, d_codes);
        if (d_bad) dev_free(ctx->stream, d_bad);
    };
    auto bail = [&](const char* what, cudaError_t err) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(err);
        cleanup();
        if (ds->d_col_off) dev_free(ctx->stream, ds->d_col_off);
        if (ds->d_bases) dev_free(ctx->stream, ds->d_bases);
        if (ds->d_gaps) dev_free(ctx->stream, ds->d_gaps);
        delete ds;
        return err == cudaErrorMemoryAllocation ? PHYFOO_ENOMEM : PHYFOO_ECUDA;
    };
    if ((e = dev_malloc(ctx->stream, &ds->d_col_off, (size_t)(n_aln + 1) * sizeof(int64_t))) != cudaSuccess) return bail("cudaMalloc col_off", e);
    if ((e = dev_malloc(ctx->stream, &ds->d_bases, (size_t)ds->nb * nc * 4)) != cudaSuccess) return bail("cudaMalloc bases", e);
    if ((e = dev_malloc(ctx->stream, &ds->d_gaps, (size_t)ds->ng * nc * 4)) != cudaSuccess) return bail("cudaMalloc gaps", e);
    if ((e = dev_malloc(ctx->stream, &d_bad, 2 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc flag", e);
    if ((e = cudaMemsetAsync(d_bad, 0, 2 * sizeof(int), ctx->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = cudaMemcpyAsync(ds->d_col_off, col_offsets, (size_t)(n_aln + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        return bail("copy col_off", e);
    if (ds->n_cols > 0) {
        if (codes_on_device) d_codes = const_cast<uint8_t*>(codes);
        else {
            if ((e = dev_malloc(ctx->stream, &d_codes, (size_t)ds->n_cols * S)) != cudaSuccess) return bail("cudaMalloc codes", e);
            if ((e = cudaMemcpyAsync(d_codes, codes, (size_t)ds->n_cols * S, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
                return bail("copy codes", e);
        }
        const int threads = 256;
        const int64_t blocks = (ds->n_cols + threads - 1) / threads;
        {
            KernelTimer kt_(ctx, "pack_columns_kernel");
            pack_columns_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(d_codes, ds->n_cols, S, ds->nb, ds->ng, ds->d_bases, ds->d_gaps, d_bad);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return bail("pack_columns_kernel", e);
    }
    int bad2[2] = {0, 0};
    if ((e = cudaMemcpyAsync(bad2, d_bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return bail("copy flag", e);
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail("pack_columns_kernel", e);
    cleanup();
    const int bad = bad2[0];
    ds->n_gap_cols = bad2[1];
    if (bad) {
        dev_free(ctx->stream, ds->d_col_off); dev_free(ctx->stream, ds->d_bases); dev_free(ctx->stream, ds->d_gaps);
        delete ds;
        return fail(ctx, PHYFOO_EINVAL, "dataset_create: code outside 0..4 (WrongAlphabetException in the reference)");
    }
    *out = ds;
    return PHYFOO_OK;
}

extern "C" int phyfoo_dataset_create(phyfoo_ctx* ctx, int32_t n_aln, int32_t n_species, const int64_t* col_offsets,
                                     const uint8_t* codes, phyfoo_dataset** out) {
    return dataset_build(ctx, n_aln, n_species, col_offsets, codes, false, out);
}
extern "C" int phyfoo_dataset_create_device(phyfoo_ctx* ctx, int32_t n_aln, int32_t n_species, const int64_t* col_offsets,
                                            const uint8_t* codes_device, phyfoo_dataset** out) {
    return dataset_build(ctx, n_aln, n_species, col_offsets, codes_device, true, out);
}

extern "C" int phyfoo_dataset_packed(phyfoo_ctx* ctx, const phyfoo_dataset* ds, uint32_t* bases_out, uint32_t* gaps_out) {
    if (!ctx || !ds) return PHYFOO_EINVAL;
    CU(cudaSetDevice(ctx->device));
    if (ds->n_cols == 0) return PHYFOO_OK;
    if (bases_out) CU(cudaMemcpyAsync(bases_out, ds->d_bases, (size_t)ds->nb * ds->n_cols * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (gaps_out) CU(cudaMemcpyAsync(gaps_out, ds->d_gaps, (size_t)ds->ng * ds->n_cols * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int64_t phyfoo_dataset_columns(const phyfoo_dataset* ds) { return ds ? ds->n_cols : 0; }

extern "C" int64_t phyfoo_dataset_windows(const phyfoo_dataset* ds, int32_t width) {
    if (!ds || width < 1) return 0;
    int64_t n = 0;
    for (int i = 0; i < ds->n_aln; ++i) n += std::max<int64_t>(0, ds->col_off[i + 1] - ds->col_off[i] - width + 1);
    return n;
}

extern "C" void phyfoo_dataset_free(phyfoo_ctx* ctx, phyfoo_dataset* ds) {
    if (!ds) return;
    // stream-ordered frees: everything enqueued on the library's stream that still reads the buffers runs first
    cudaStream_t st = ctx ? ctx->stream : nullptr;
    if (ctx) cudaSetDevice(ctx->device);
    if (ctx && ctx->gamma_ds == ds) { ctx->gamma_ds = nullptr; ctx->gamma_windows = -1; }   // the resident gamma dies with its data set
    if (ctx && ctx->multi_ds == ds) { ctx->multi_ds = nullptr; ctx->multi_k = 0; }
    dev_free(st, ds->d_col_off); dev_free(st, ds->d_bases); dev_free(st, ds->d_gaps);
    for (auto& kv : ds->d_win_off) dev_free(st, kv.second);
    memo_release(st, ds->memo);
    delete ds;
}

// window prefix offsets for motif width W (device + host), cached on the dataset
static int dataset_win_off(phyfoo_ctx* ctx, const phyfoo_dataset* cds, int W, const int64_t** d_out, const std::vector<int64_t>** h_out) {
    phyfoo_dataset* ds = const_cast<phyfoo_dataset*>(cds);
    auto it = ds->d_win_off.find(W);
    if (it == ds->d_win_off.end()) {
        std::vector<int64_t> off(ds->n_aln + 1, 0);
        for (int i = 0; i < ds->n_aln; ++i) off[i + 1] = off[i] + std::max<int64_t>(0, ds->col_off[i + 1] - ds->col_off[i] - W + 1);
        int64_t* d = nullptr;
        CU(dev_malloc(ctx->stream, &d, off.size() * sizeof(int64_t)));
        // (no synchronisation: the source is pageable memory, which the runtime stages before the call returns, and the
        // vector's buffer lives on in ds->win_off)
        cudaError_t e = cudaMemcpyAsync(d, off.data(), off.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { dev_free(ctx->stream, d); ctx->err = std::string("win_off copy: ") + cudaGetErrorString(e); return PHYFOO_ECUDA; }
        ds->d_win_off[W] = d;
        ds->win_off[W] = std::move(off);
        it = ds->d_win_off.find(W);
    }
    *d_out = it->second;
    if (h_out) *h_out = &ds->win_off[W];
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------------
extern "C" int phyfoo_model_create(phyfoo_ctx* ctx, const phyfoo_model_desc* desc, phyfoo_model** out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!desc || !out) return fail(ctx, PHYFOO_EINVAL, "model_create: NULL argument");
    *out = nullptr;
    phyfoo_model* m = new phyfoo_model();
    try { m->h.init(*desc); }
    catch (const std::exception& e) { delete m; return fail(ctx, PHYFOO_EINVAL, e.what()); }
    if (m->h.S > 32) { delete m; return fail(ctx, PHYFOO_EUNSUPPORTED, "model_create: more than 32 species are not supported"); }
    *out = m;
    return PHYFOO_OK;
}

extern "C" int phyfoo_model_set_pi(phyfoo_ctx* ctx, phyfoo_model* m, int32_t position, const double* pi, int32_t n) {
    if (!ctx || !m || !pi) return PHYFOO_EINVAL;
    try { m->h.set_pi(position, pi, n); } catch (const std::exception& e) { return fail(ctx, PHYFOO_EINVAL, e.what()); }
    return PHYFOO_OK;
}
extern "C" int phyfoo_model_get_pi(phyfoo_ctx* ctx, const phyfoo_model* m, int32_t position, double* pi_out, int32_t n) {
    if (!ctx || !m || !pi_out) return PHYFOO_EINVAL;
    if (position < 0 || position >= m->h.W || n != (int)m->h.pi[position].size()) return fail(ctx, PHYFOO_EINVAL, "get_pi: bad position or size");
    std::copy(m->h.pi[position].begin(), m->h.pi[position].end(), pi_out);
    return PHYFOO_OK;
}
extern "C" int phyfoo_model_set_branch(phyfoo_ctx* ctx, phyfoo_model* m, int32_t position, int32_t node, double length) {
    if (!ctx || !m) return PHYFOO_EINVAL;
    try { m->h.set_branch(position, node, length); } catch (const std::exception& e) { return fail(ctx, PHYFOO_EINVAL, e.what()); }
    return PHYFOO_OK;
}
extern "C" int phyfoo_model_is_degenerate(const phyfoo_model* m) { return m && m->h.evol == PHYFOO_EVOL_HKY_AS_IMPLEMENTED ? 1 : 0; }
extern "C" int phyfoo_model_set_mode(phyfoo_ctx* ctx, phyfoo_model* m, int32_t mode) {
    if (!ctx || !m) return PHYFOO_EINVAL;
    if (mode != PHYFOO_MODE_MEANFIELD && mode != PHYFOO_MODE_EXACT) return fail(ctx, PHYFOO_EINVAL, "set_mode: unknown mode");
    m->h.mode = mode;
    return PHYFOO_OK;
}
extern "C" void phyfoo_model_free(phyfoo_ctx* ctx, phyfoo_model* m) {
    if (!m) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(m->d_prog); cudaFree(m->d_logtab); cudaFree(m->d_probtab); cudaFree(m->d_tab1); cudaFree(m->d_tab1w); cudaFree(m->d_prog1w);
    cudaFree(m->d_tab1n); cudaFree(m->d_leaftab1n); cudaFree(m->d_prog1n); cudaFree(m->d_tabf);
    cudaFree(m->d_netg_rec); cudaFree(m->d_netg_lists); cudaFree(m->d_netg_tab);
    delete m;
}

// copy the order-0 tables to the device in [E][W] layout when the parameters changed
static int model_upload(phyfoo_ctx* ctx, phyfoo_model* m) {
    ModelHost& h = m->h;
    if (h.order != 0 && h.order != 1)
        return fail(ctx, PHYFOO_EUNSUPPORTED, "a background model of order >= 2 is scored through phyfoo_bg_range_terms / phyfoo_zoops_estep (generic net kernel) only");
    if (m->uploaded == h.version) return PHYFOO_OK;
    if (h.order == 1) {
        std::vector<double> t1, t1w, t1n, tl1n;
        h.tables1v(t1);
        h.tables1w(t1w);
        h.tables1n(t1n);
        h.leaf_tables1n(tl1n);
        m->f0_const1n = h.n1n_f0_const();
        if (!m->d_prog) {
            std::vector<int> pn;
            h.node_program(pn, m->MC1n, m->n_steps1n);
            m->prog1n_len = (int)pn.size();
            CU(cudaMalloc(&m->d_prog1n, pn.size() * sizeof(int)));
            CU(cudaMemcpyAsync(m->d_prog1n, pn.data(), pn.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMalloc(&m->d_tab1n, t1n.size() * sizeof(double)));
            CU(cudaMalloc(&m->d_leaftab1n, tl1n.size() * sizeof(double)));
            std::vector<int> pw;
            h.wave_program(pw, m->MC, m->ML, m->n_steps);
            m->prog1w_len = (int)pw.size();
            CU(cudaMalloc(&m->d_prog1w, pw.size() * sizeof(int)));
            CU(cudaMemcpyAsync(m->d_prog1w, pw.data(), pw.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMalloc(&m->d_tab1w, t1w.size() * sizeof(double)));
            std::vector<int> prog = h.prog_flat();
            m->prog_len = (int)prog.size();
            CU(cudaMalloc(&m->d_prog, prog.size() * sizeof(int)));
            CU(cudaMemcpyAsync(m->d_prog, prog.data(), prog.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMalloc(&m->d_tab1, t1.size() * sizeof(double)));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        CU(cudaMemcpyAsync(m->d_tab1, t1.data(), t1.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(m->d_tab1w, t1w.data(), t1w.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(m->d_tab1n, t1n.data(), t1n.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(m->d_leaftab1n, tl1n.data(), tl1n.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        m->uploaded = h.version;
        return PHYFOO_OK;
    }
    const size_t n = (size_t)h.E * h.W;
    if (!m->d_prog) {
        std::vector<int> prog = h.prog_flat();
        m->prog_len = (int)prog.size();
        CU(cudaMalloc(&m->d_prog, prog.size() * sizeof(int)));
        CU(cudaMemcpyAsync(m->d_prog, prog.data(), prog.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMalloc(&m->d_logtab, n * sizeof(double)));
        CU(cudaMalloc(&m->d_probtab, n * sizeof(double)));
        CU(cudaMalloc(&m->d_tabf, (size_t)h.Ef() * h.W * sizeof(double)));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<double> tf, tft((size_t)h.Ef() * h.W);
    m->f81_ok = h.tables0f(tf);
    if (m->f81_ok) {
        for (int t = 0; t < h.W; ++t)
            for (int e = 0; e < h.Ef(); ++e) tft[(size_t)e * h.W + t] = tf[(size_t)t * h.Ef() + e];
        CU(cudaMemcpyAsync(m->d_tabf, tft.data(), tft.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    std::vector<double> tl(n), tp(n);
    for (int t = 0; t < h.W; ++t)
        for (int e = 0; e < h.E; ++e) {
            tl[(size_t)e * h.W + t] = h.logtab[(size_t)t * h.E + e];
            tp[(size_t)e * h.W + t] = h.probtab[(size_t)t * h.E + e];
        }
    CU(cudaMemcpyAsync(m->d_logtab, tl.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(m->d_probtab, tp.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // tl/tp are stack-scoped
    m->uploaded = h.version;
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// scoring launches
// ------------------------------------------------------------------------------------------------
struct ScorePlan { int threads, groups, blocks; size_t smem; bool gstate; };

static int plan_score(phyfoo_ctx* ctx, const ModelHost& h, int prog_len, long long n_items, int mode, bool items, ScorePlan* pl) {
    const int W = h.W;
    if (W > 256) return fail(ctx, PHYFOO_EUNSUPPORTED, "motif width > 256 is not supported");
    const size_t tab_bytes = (size_t)h.E * W * sizeof(double);
    const size_t per_thread_state = (size_t)h.I * 8 * sizeof(double);
    const size_t budget = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    // block shape: the one that keeps the most threads resident per SM (shared memory: tables once per block + state per
    // thread; registers: 128 per thread).  Large trees used to get two tiny blocks per SM; one block of 256 threads runs
    // the 12-species configuration 1.8x faster.
    const int max_groups = std::max(1, 256 / W);
    int best_groups = 0; long long best_resident = 0;
    for (int groups = max_groups; groups >= 1; --groups) {
        const int threads = ((groups * W + 31) / 32) * 32;
        const size_t need = tab_bytes + (size_t)threads * 8 + (size_t)groups * 8 + (size_t)prog_len * 4 + 64 + per_thread_state * threads;
        if (need <= budget) {
            const long long by_smem = (long long)((228 * 1024) / (need + 1024));
            const long long by_regs = 65536 / (128LL * threads);
            const long long resident = std::min(by_smem, std::max(1LL, by_regs)) * groups;
            if (resident > best_resident) { best_resident = resident; best_groups = groups; }
        }
    }
    // state in global memory (coalesced, behind L2) when that at least doubles the windows in flight per SM: the state of a
    // 12-species tree (704 B) leaves room for one block of 21 windows, that of a 30-species tree (1.9 KB) for 5 windows,
    // while the tables alone allow two blocks of 256 threads.  Measured: 12 species 13.5 -> 18.4 M windows/s
    // (--config c3o0), 30 species 0.97 -> 3.30 M windows/s (--config c5 --n-aln 3000)
    long long g_resident = 0; int g_groups = 0;
    {
        const int groups = max_groups, threads = ((groups * W + 31) / 32) * 32;
        const size_t fixed = tab_bytes + (size_t)threads * 8 + (size_t)groups * 8 + (size_t)prog_len * 4 + 64;
        if (fixed <= budget) {
            const long long by_smem = (long long)((228 * 1024) / (fixed + 1024));
            const long long by_regs = 65536 / (128LL * threads);
            g_resident = std::min(by_smem, std::max(1LL, by_regs)) * groups; g_groups = groups;
        }
    }
    if (g_groups && g_resident >= 2 * best_resident) best_groups = 0;
    if (best_groups) {
        const int threads = ((best_groups * W + 31) / 32) * 32;
        pl->threads = threads; pl->groups = best_groups; pl->gstate = false;
        pl->smem = tab_bytes + (size_t)threads * 8 + (size_t)best_groups * 8 + (size_t)prog_len * 4 + 64 + per_thread_state * threads;
    } else {
        // state in global memory
        const int groups = g_groups ? g_groups : std::max(1, 128 / W);
        const int threads = ((groups * W + 31) / 32) * 32;
        const size_t fixed = tab_bytes + (size_t)threads * 8 + (size_t)groups * 8 + (size_t)prog_len * 4 + 64;
        if (fixed > budget) return fail(ctx, PHYFOO_EUNSUPPORTED, "model tables do not fit in shared memory");
        pl->threads = threads; pl->groups = groups; pl->smem = fixed; pl->gstate = true;
    }
    const void* fn = mode == PHYFOO_MODE_EXACT ? (const void*)score_o0_kernel<PHYFOO_MODE_EXACT, false>
                   : items ? (const void*)score_o0_kernel<PHYFOO_MODE_MEANFIELD, true> : (const void*)score_o0_kernel<PHYFOO_MODE_MEANFIELD, false>;
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, pl->threads, pl->smem));
    if (per_sm < 1) return fail(ctx, PHYFOO_ECUDA, "score kernel does not fit on an SM");
    long long want = (n_items + pl->groups - 1) / pl->groups;
    pl->blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)per_sm * ctx->sm_count));
    return PHYFOO_OK;
}

// explicit window list for the scoring kernels: M-step "prepare" (q sink) or the pattern-table fallback (device-side count)
struct ScoreItems {
    const int64_t* col0; const uint8_t* strand; double* q_int; double* q_leaf; double* ent;
    const int* n_dev = nullptr; const int64_t* dst = nullptr; bool fallback = false;
    double* out_last = nullptr;     // order-1 kernels: free energy of the last tree alone (background ranges)
    const int64_t* col1 = nullptr;  // order-1 wavefront kernel, W = 2: column of the second tree (single-column background ranges)
    bool stop_first = false;        //   ... and the stop rule on the first tree alone
    int64_t sink_stride = 0;        // listed fallback items of a "prepare" pass: windows of the FE set (the q sinks' stride)
};

#include "net1.cuh"
#include "net1w.cuh"
#include "net1n.cuh"
#include "netg.cuh"
#include "tree_pass_f81.cuh"
#include "jit_o0.hpp"
#include "memo.cuh"

static void memo_release(cudaStream_t st, MemoSet* ms) {
    if (!ms) return;
    dev_free(st, ms->d_flags); dev_free(st, ms->d_dense_of); dev_free(st, ms->d_code_of); dev_free(st, ms->d_count);
    delete ms;
}

// occurring column patterns of the data set (forward and complemented), built once
static int dataset_memo(phyfoo_ctx* ctx, const phyfoo_dataset* cds, MemoSet** out) {
    phyfoo_dataset* ds = const_cast<phyfoo_dataset*>(cds);
    if (!ds->memo) {
        MemoSet* ms = new MemoSet();
        ms->code_bits = 3 * ds->S; ms->n_codes = 1 << ms->code_bits;
        long long valid = 1;
        for (int s = 0; s < ds->S; ++s) valid *= 5;
        ms->cap = (int)std::max<long long>(1, std::min<long long>(valid, 2 * ds->n_cols));
        cudaError_t e;
        if ((e = dev_malloc(ctx->stream, &ms->d_flags, (size_t)ms->n_codes * sizeof(int32_t))) != cudaSuccess ||
            (e = dev_malloc(ctx->stream, &ms->d_dense_of, (size_t)ms->n_codes * sizeof(int32_t))) != cudaSuccess ||
            (e = dev_malloc(ctx->stream, &ms->d_code_of, (size_t)ms->cap * sizeof(int32_t))) != cudaSuccess ||
            (e = dev_malloc(ctx->stream, &ms->d_count, sizeof(int32_t))) != cudaSuccess ||
            (e = cudaMemsetAsync(ms->d_flags, 0, (size_t)ms->n_codes * sizeof(int32_t), ctx->stream)) != cudaSuccess) {
            memo_release(ctx->stream, ms);
            ctx->err = std::string("pattern set: ") + cudaGetErrorString(e);
            return PHYFOO_ECUDA;
        }
        if (ds->n_cols > 0) {
            {
                KernelTimer kt_(ctx, "memo_mark_kernel");
                memo_mark_kernel<<<(unsigned)((ds->n_cols + 255) / 256), 256, 0, ctx->stream>>>(ds->d_bases, ds->d_gaps, ds->n_cols, ds->nb, ds->S, ms->d_flags);
            }
        }
        {
            KernelTimer kt_(ctx, "memo_compact_kernel");
            memo_compact_kernel<<<1, 1024, 0, ctx->stream>>>(ms->d_flags, ms->n_codes, ds->S, ms->d_dense_of, ms->d_code_of, ms->cap, ms->d_count);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) { memo_release(ctx->stream, ms); ctx->err = std::string("pattern set kernels: ") + cudaGetErrorString(e); return PHYFOO_ECUDA; }
        ds->memo = ms;
    }
    *out = ds->memo;
    return PHYFOO_OK;
}

// number of distinct column patterns (forward + complemented) of the data set; -1 when the pattern table does not apply
extern "C" int64_t phyfoo_dataset_patterns(phyfoo_ctx* ctx, const phyfoo_dataset* ds) {
    if (!ctx || !ds || 3 * ds->S > 18) return -1;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    MemoSet* ms = nullptr;
    if (dataset_memo(ctx, ds, &ms) != PHYFOO_OK) return -1;
    int32_t d = 0;
    if (cudaMemcpyAsync(&d, ms->d_count, sizeof(d), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return d;
}

static int memo_K1(const phyfoo_ctx* ctx, const ModelHost& h, int mode) {
    return mode == PHYFOO_MODE_EXACT ? 1 : std::min(ctx->memo_sweep_cap, h.max_sweeps) + 1;
}

static bool memo_applies(const phyfoo_ctx* ctx, const phyfoo_dataset* ds, const ModelHost& h, int mode, long long n_items) {
    if (ctx->memo_mode == 0 || h.order != 0 || 3 * ds->S > 18 || h.W > 16) return false;
    if (ctx->memo_mode == 2) return true;
    long long valid = 1;
    for (int s = 0; s < ds->S; ++s) valid *= 5;
    const long long cap = std::min<long long>(valid, 2 * ds->n_cols);
    // trajectories cost cap * W * K1 tree passes, the direct kernel about n_items * W * 10
    return n_items * 10 >= 4 * cap * memo_K1(ctx, h, mode);
}

static int launch_score(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                        int strands, double* d_out, int32_t* d_sweeps, int counter_slot, const ScoreItems* items = nullptr);

struct MemoJob {
    phyfoo_model* m; int mode; const int64_t* d_win_off; int64_t n_windows; int strands;
    double* d_out; int32_t* d_sweeps; int counter_slot;
};

// pattern-table path (memo.cuh): trajectories of the occurring patterns for up to two models in one launch
// (background + motif of an E-step), then per model one gather pass over the windows and the fallback list
static int launch_memo_jobs(phyfoo_ctx* ctx, const phyfoo_dataset* ds, const MemoJob* jobs, int n_jobs) {
    MemoSet* ms = nullptr;
    int rc = dataset_memo(ctx, ds, &ms);
    if (rc) return rc;
    for (int j = 0; j < n_jobs; ++j) if ((rc = model_upload(ctx, jobs[j].m))) return rc;
    size_t traj_off[3] = {0, 0, 0}, pend_off[3] = {0, 0, 0}, state_off[3] = {0, 0, 0};
    long long n_items[2] = {0, 0};
    int K1[2] = {1, 1}, KA[2] = {1, 1};
    bool split = false;
    for (int j = 0; j < n_jobs; ++j) {
        const ModelHost& h = jobs[j].m->h;
        K1[j] = memo_K1(ctx, h, jobs[j].mode);
        // two phases: passes 0..KA-1 first, the rest on the second stream while the first gather pass runs (most windows stop
        // within the first few sweeps); KA a multiple of 4 so that the two never touch the same 32-byte sector of a row
        KA[j] = (ctx->memo_split > 0 && ctx->memo_split + 4 <= K1[j]) ? ctx->memo_split : K1[j];
        split = split || KA[j] < K1[j];
        n_items[j] = (long long)jobs[j].n_windows * (jobs[j].strands == 2 ? 2 : 1);
        traj_off[j + 1] = traj_off[j] + (size_t)ms->cap * h.W * ((K1[j] + 3) / 4 * 4);
        state_off[j + 1] = state_off[j] + (KA[j] < K1[j] ? (size_t)ms->cap * h.W * h.I * 12 : 0);   // wavefront: q, base, msg
        // per job, twice (list of the first gather pass, list for the direct kernel):
        // [count (8 bytes, int used)] [col0 int64 x n] [dst int64 x n] [strand u8 x n, padded to 8]
        pend_off[j + 1] = pend_off[j] + 2 * (8 + (size_t)n_items[j] * 16 + (((size_t)n_items[j] + 7) / 8) * 8);
    }
    // the counters of this call's jobs: [2 slot] work, [2 slot + 1] sweep sums, [4 + 2 j + l] lengths of the pending lists,
    // [8 + slot] work counters of the fallback launches ([12] is the posterior kernel's ticket, zero between launches).
    // Two jobs use every slot: one memset; a single job leaves the other model's sweep sum alone (e.g. the background's,
    // scored by the direct kernel just before, which phyfoo_estep_sweeps reports)
    if (n_jobs == 2) CU(cudaMemsetAsync(ctx->counters.p, 0, 12 * sizeof(unsigned long long), ctx->stream));
    else {
        unsigned long long* cnt = ctx->counters.as<unsigned long long>();
        const int sl = jobs[0].counter_slot;
        CU(cudaMemsetAsync(cnt + 2 * sl, 0, 2 * sizeof(unsigned long long), ctx->stream));
        CU(cudaMemsetAsync(cnt + 4, 0, 2 * sizeof(unsigned long long), ctx->stream));
        CU(cudaMemsetAsync(cnt + 8 + sl, 0, sizeof(unsigned long long), ctx->stream));
    }
    CU(ctx->memo_traj.ensure(traj_off[n_jobs] * sizeof(double)));
    CU(ctx->memo_pending.ensure(pend_off[n_jobs]));
    if (split) CU(ctx->memo_state.ensure(state_off[n_jobs] * sizeof(double)));
    MemoTrajJobs tj;
    const int threads = 128;
    size_t smem = 0;
    long long trees = 0;
    // wavefront schedule (memo_traj_wave_kernel): node slots = the larger of the even-depth and odd-depth internal nodes
    int lpt = 4;
    bool wave = ctx->memo_wave != 0;
    for (int j = 0; j < n_jobs && wave; ++j) {
        const ModelHost& h = jobs[j].m->h;
        MemoTrajParams& tp = tj.job[j];
        if (h.I < 2 || h.I > 8 || jobs[j].mode == PHYFOO_MODE_EXACT) { wave = false; break; }
        int cnt[2] = {0, 0};
        memset(tp.sched, -1, sizeof(tp.sched));
        tp.dmax = 0; tp.ML = 0; tp.MCH = 0;
        for (int i = 0; i < h.I && wave; ++i) {
            const int dep = i == 0 ? 0 : tp.depth[h.par_internal[i]] + 1;
            tp.depth[i] = (signed char)dep;
            tp.dmax = std::max(tp.dmax, dep);
            tp.ML = std::max(tp.ML, h.leaf_begin[i + 1] - h.leaf_begin[i]);
            tp.MCH = std::max(tp.MCH, h.ichild_begin[i + 1] - h.ichild_begin[i]);
            if (cnt[dep & 1] >= 4) { wave = false; break; }
            tp.sched[dep & 1][cnt[dep & 1]++] = (signed char)i;
        }
        if (tp.dmax > 6) wave = false;          // the ring keeps the terms of four sweeps
        lpt = std::max(lpt, std::max(cnt[0], cnt[1]) <= 2 ? 8 : 16);
    }
    if (!wave) lpt = 4;
    for (int j = 0; j < n_jobs; ++j) {
        phyfoo_model* m = jobs[j].m;
        const ModelHost& h = m->h;
        MemoTrajParams& tp = tj.job[j];
        tp.code_of = ms->d_code_of; tp.count = ms->d_count;
        tp.W = h.W; tp.I = h.I; tp.S = h.S; tp.E = h.E; tp.prog_len = m->prog_len; tp.cap = ms->cap; tp.K1 = K1[j]; tp.RS = (K1[j] + 3) / 4 * 4;
        tp.exact = jobs[j].mode == PHYFOO_MODE_EXACT ? 1 : 0;
        tp.tab = tp.exact ? m->d_probtab : m->d_logtab;
        tp.prog = m->d_prog; tp.traj = ctx->memo_traj.as<double>() + traj_off[j];
        tp.k0 = 0; tp.k1 = KA[j]; tp.state = split ? ctx->memo_state.as<double>() + state_off[j] : nullptr;
        const size_t per_tree = wave ? (size_t)(26 * h.I + 21) : (size_t)h.I * 8;
        smem = std::max(smem, (size_t)h.E * h.W * sizeof(double) + per_tree * (threads / lpt) * sizeof(double) + (size_t)m->prog_len * sizeof(int) + 16);
        trees = std::max(trees, (long long)ms->cap * h.W);
    }
    if (n_jobs == 1) tj.job[1] = tj.job[0];
    const int tpb = threads / lpt;   // lanes per tree: four symbols x node slots
    const int blocks = (int)std::max<long long>(1, std::min<long long>((trees + tpb - 1) / tpb, 16LL * ctx->sm_count));
    auto launch_traj = [&](cudaStream_t st, const char* name) -> cudaError_t {
        KernelTimer kt_(ctx, name, st);
        if (!wave) {
            cudaError_t e = cudaFuncSetAttribute(memo_traj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            memo_traj_kernel<<<dim3(blocks, n_jobs), threads, smem, st>>>(tj);
        } else {
            // the number of internal nodes as a template argument when both jobs agree on it (binary trees of <= 6 species)
            int it = tj.job[0].I;
            for (int j = 1; j < n_jobs; ++j) if (tj.job[j].I != it) it = 0;
            void (*fn)(MemoTrajJobs) = nullptr;
            if (lpt == 8) {
                switch (it) {
                    case 2: fn = memo_traj_wave_kernel<8, 2>; break;
                    case 3: fn = memo_traj_wave_kernel<8, 3>; break;
                    case 4: fn = memo_traj_wave_kernel<8, 4>; break;
                    case 5: fn = memo_traj_wave_kernel<8, 5>; break;
                    default: fn = memo_traj_wave_kernel<8, 0>;
                }
            } else {
                switch (it) {
                    case 4: fn = memo_traj_wave_kernel<16, 4>; break;
                    case 5: fn = memo_traj_wave_kernel<16, 5>; break;
                    default: fn = memo_traj_wave_kernel<16, 0>;
                }
            }
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            fn<<<dim3(blocks, n_jobs), threads, smem, st>>>(tj);
        }
        return cudaGetLastError();
    };
    CU(launch_traj(ctx->stream, "memo_traj_kernel"));
    if (split) {
        CU(cudaEventRecord(ctx->ev_a, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_a, 0));
        for (int j = 0; j < n_jobs; ++j) {
            tj.job[j].k0 = KA[j]; tj.job[j].k1 = K1[j];      // a job without a second phase gets an empty pass range
        }
        if (n_jobs == 1) tj.job[1] = tj.job[0];
        CU(launch_traj(ctx->aux, "memo_traj_kernel(second phase, concurrent)"));
        CU(cudaEventRecord(ctx->ev_b, ctx->aux));
    }
    MemoWindowParams wps[2];
    for (int j = 0; j < n_jobs; ++j) {
        phyfoo_model* m = jobs[j].m;
        const ModelHost& h = m->h;
        if (n_items[j] == 0) continue;
        unsigned long long* counters = ctx->counters.as<unsigned long long>();
        unsigned long long* c_sweeps = counters + 2 * jobs[j].counter_slot + 1;
        MemoWindowParams& wp = wps[j];
        wp.bases = ds->d_bases; wp.gaps = ds->d_gaps; wp.n_cols = ds->n_cols; wp.nb = ds->nb;
        wp.col_off = ds->d_col_off; wp.win_off = jobs[j].d_win_off; wp.n_aln = ds->n_aln;
        wp.n_windows = jobs[j].n_windows; wp.n_strands = jobs[j].strands == 2 ? 2 : 1; wp.strand0 = jobs[j].strands == 1 ? 1 : 0;
        wp.W = h.W; wp.S = h.S; wp.cap = ms->cap; wp.K1 = KA[j]; wp.RS = (K1[j] + 3) / 4 * 4;
        wp.dense_of = ms->d_dense_of; wp.traj = ctx->memo_traj.as<double>() + traj_off[j];
        wp.out = jobs[j].d_out; wp.sweeps = jobs[j].d_sweeps; wp.counter_sweeps = c_sweeps;
        wp.max_sweeps = h.max_sweeps; wp.min_sweeps = h.min_sweeps; wp.tol = h.tol; wp.exact = jobs[j].mode == PHYFOO_MODE_EXACT ? 1 : 0;
        // list 0: left by the first gather pass when the trajectories come in two phases; list 1: for the direct kernel
        const size_t list_bytes = (pend_off[j + 1] - pend_off[j]) / 2;
        const int first_list = KA[j] < K1[j] ? 0 : 1;
        char* pb = ctx->memo_pending.as<char>() + pend_off[j] + first_list * list_bytes;
        wp.pend_count = (int*)(counters + 4 + 2 * j + first_list);
        wp.pend_col0 = (int64_t*)(pb + 8);
        wp.pend_dst = wp.pend_col0 + n_items[j];
        wp.pend_strand = (uint8_t*)(wp.pend_dst + n_items[j]);
        wp.in_count = nullptr; wp.in_col0 = nullptr; wp.in_strand = nullptr; wp.in_dst = nullptr;
        const int wblocks = (int)std::max<long long>(1, std::min<long long>((n_items[j] + 255) / 256, 16LL * ctx->sm_count));
        {
            KernelTimer kt_(ctx, "memo_window_kernel");
            memo_window_kernel<false><<<wblocks, 256, 0, ctx->stream>>>(wp);
        }
        CU(cudaGetLastError());
    }
    if (split) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    for (int j = 0; j < n_jobs; ++j) {
        phyfoo_model* m = jobs[j].m;
        const ModelHost& h = m->h;
        if (n_items[j] == 0) continue;
        MemoWindowParams& wp = wps[j];
        if (KA[j] < K1[j]) {
            // second gather pass over the windows the first one left, now with the whole trajectories
            const size_t list_bytes = (pend_off[j + 1] - pend_off[j]) / 2;
            wp.in_count = wp.pend_count; wp.in_col0 = wp.pend_col0; wp.in_strand = wp.pend_strand; wp.in_dst = wp.pend_dst;
            char* pb = ctx->memo_pending.as<char>() + pend_off[j] + list_bytes;
            wp.pend_count = (int*)(ctx->counters.as<unsigned long long>() + 4 + 2 * j + 1);
            wp.pend_col0 = (int64_t*)(pb + 8);
            wp.pend_dst = wp.pend_col0 + n_items[j];
            wp.pend_strand = (uint8_t*)(wp.pend_dst + n_items[j]);
            wp.K1 = K1[j];
            const int wblocks = (int)std::max<long long>(1, std::min<long long>((n_items[j] + 255) / 256, 2LL * ctx->sm_count));
            {
                KernelTimer kt_(ctx, "memo_window_kernel(second pass)");
                memo_window_kernel<true><<<wblocks, 256, 0, ctx->stream>>>(wp);
            }
            CU(cudaGetLastError());
        }
        if (!wp.exact && K1[j] - 1 < h.max_sweeps) {
            // windows that need more sweeps than the stored trajectories hold: direct kernel over the (device-side) list
            ScoreItems items{wp.pend_col0, wp.pend_strand, nullptr, nullptr, nullptr, wp.pend_count, wp.pend_dst, true};
            rc = launch_score(ctx, ds, m, nullptr, n_items[j], 0, jobs[j].d_out, jobs[j].d_sweeps, jobs[j].counter_slot, &items);
            if (rc) return rc;
        }
    }
    ctx->last_path = 1;
    return PHYFOO_OK;
}

// order-1 motif network, wavefront kernel (net1w.cuh): 16 lanes per window, two windows per warp
static int launch_net1w(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                        int strands, double* d_out, int32_t* d_sweeps, int counter_slot, const ScoreItems* items) {
    ModelHost& h = m->h;
    const int n_strands = strands == 2 ? 2 : 1;
    const long long n_items = (long long)n_windows * n_strands;
    // per window: q of the hidden internal nodes + leaf codes; fixed: one padding slot per q row, one-hot, program
    const size_t per_window = (size_t)h.W * h.I * 4 * sizeof(double) + (size_t)h.W * (h.S + 1);
    const size_t fixed = (size_t)h.W * h.I * 32 + (size_t)m->prog1w_len * sizeof(int) + 32 + 64;
    const size_t budget = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    int nwb = 16;
    while (nwb > 2 && fixed + per_window * nwb > budget / 3) nwb >>= 1;     // three blocks per SM when possible
    while (nwb > 2 && fixed + per_window * nwb > budget) nwb >>= 1;
    if (fixed + per_window * nwb > budget) return fail(ctx, PHYFOO_EUNSUPPORTED, "order-1 window state does not fit in shared memory (motif too wide or tree too large)");
    // (measured on C3: 3 blocks per SM 2.89, 2 blocks 2.41, 1 block 1.42 M windows/s -- warps in flight beat a larger L1)
    // (again after the observed-leaf cache: 3 blocks per SM 3.54, 2 blocks 2.90 M windows/s)
    const size_t smem = fixed + per_window * nwb;
    CU(cudaFuncSetAttribute(score_o1w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_o1w_kernel, 16 * nwb, smem));
    if (per_sm < 1) return fail(ctx, PHYFOO_ECUDA, "order-1 wavefront kernel does not fit on an SM");
    const long long want = (n_items + nwb - 1) / nwb;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)per_sm * ctx->sm_count));
    const int64_t n_slots = (int64_t)blocks * nwb;
    CU(ctx->gstate.ensure(((size_t)h.W * h.S * 4 + (size_t)h.W * h.I * 4) * n_slots * sizeof(double)));
    Net1wParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_win_off; p.n_aln = ds->n_aln;
    p.n_windows = n_windows; p.n_strands = n_strands; p.strand0 = strands == 1 ? 1 : 0;
    p.W = h.W; p.I = h.I; p.S = h.S; p.E1 = h.E1w(); p.prog_len = m->prog1w_len; p.prog = m->d_prog1w;
    p.MC = m->MC; p.ML = m->ML; p.n_steps = m->n_steps;
    p.tab = m->d_tab1w;
    p.out = d_out; p.sweeps = d_sweeps; p.out_last = items ? items->out_last : nullptr;
    p.counters = ctx->counters.as<unsigned long long>() + 2 * counter_slot;
    p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
    p.gq = ctx->gstate.as<double>(); p.n_slots = n_slots; p.nwb = nwb;
    p.gc = p.gq + (size_t)h.W * h.S * 4 * n_slots;
    p.item_col0 = items ? items->col0 : nullptr; p.item_strand = items ? items->strand : nullptr;
    p.sink_qint = items ? items->q_int : nullptr; p.sink_qleaf = items ? items->q_leaf : nullptr; p.sink_ent = items ? items->ent : nullptr;
    p.n_items_dev = items ? items->n_dev : nullptr; p.item_dst = items ? items->dst : nullptr;
    p.item_col1 = items ? items->col1 : nullptr; p.stop_first = items && items->stop_first ? 1 : 0;
    p.sink_stride = items ? items->sink_stride : 0;
    const bool fallback = items && items->fallback;
    // a fallback launch (windows with gaps listed by the node-per-thread kernel) pulls its work from a counter of its own
    // ([8 + slot]) and continues the sweep sum of that kernel
    p.counter_next = fallback ? ctx->counters.as<unsigned long long>() + 8 + counter_slot : p.counters;
    if (fallback) CU(cudaMemsetAsync(p.counter_next, 0, sizeof(unsigned long long), ctx->stream));
    else CU(cudaMemsetAsync(p.counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
    {
        KernelTimer kt_(ctx, fallback ? "score_o1w_kernel(windows with gaps)" : "score_o1w_kernel");
        score_o1w_kernel<<<blocks, 16 * nwb, smem, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    return PHYFOO_OK;
}

// L2 access-policy window over a scratch region for the launches that follow on the library's stream (bytes = 0: off)
static void l2_window(phyfoo_ctx* ctx, void* base, size_t bytes) {
    if (!ctx->l2_persist || !ctx->l2_persist_max || !ctx->l2_window_max) return;
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.base_ptr = base;
    a.accessPolicyWindow.num_bytes = std::min(bytes, ctx->l2_window_max);
    a.accessPolicyWindow.hitRatio = bytes ? (float)std::min(1.0, (double)ctx->l2_persist_max / (double)std::max<size_t>(a.accessPolicyWindow.num_bytes, 1)) : 0.f;
    a.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    a.accessPolicyWindow.missProp = bytes ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &a);
}

// order-1 motif network: quad per window (net1.cuh)
// order-1 motif network, node-per-thread kernel (net1n.cuh): 4 threads per window, 8 windows per warp; returns 1 when the
// configuration is not one it takes (the caller then runs the wavefront kernel), 0 when launched
static int launch_net1n(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                        int strands, double* d_out, int32_t* d_sweeps, int counter_slot, const ScoreItems* items = nullptr) {
    ModelHost& h = m->h;
    if (m->MC1n > 3 || !m->d_prog1n) return 1;
    // mostly gapped data: nearly every window would go through the list anyway
    if ((double)ds->n_gap_cols * h.W >= 0.5 * (double)std::max<int64_t>(ds->n_cols, 1)) return 1;
    const int n_strands = strands == 2 ? 2 : 1;
    const long long n_items = (long long)n_windows * n_strands;
    const ModelHost::N1nLayout L = h.n1n_layout();
    const size_t fixed = ((size_t)h.W * h.I * ModelHost::N1N_TAB_STRIDE + 64) * sizeof(double) + (size_t)m->prog1n_len * sizeof(int) + 16;
    const size_t budget = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    // q rows in shared memory: as many warps per SM as fit (4 at C3).  q rows in global memory behind L1 (order1_qglobal):
    // shared memory holds the AB and message rows only, 8 warps per SM (registers: ~200 x 256 threads)
    // (measured on C3, 2000 alignments: shared memory, 4 warps 192 ms; global memory, 4 / 6 / 7 / 8 warps 263 / 198 / 164 / 153 ms:
    // a lone warp per scheduler cannot hide its own latencies, two keep the fp64 pipe ~88 % busy)
    int want_qg = ctx->order1_qglobal;
    if (want_qg < 0) want_qg = (fixed + 8 * (size_t)L.region * sizeof(double) <= budget) ? 0 : 8;
    const bool qg = want_qg != 0;
    const size_t per_warp = (size_t)(qg ? L.sregion : L.region) * sizeof(double);
    if (fixed + per_warp > budget) return 1;
    int nw = (int)std::min<size_t>(qg ? want_qg : 8, (budget - fixed) / per_warp);
    // (12 / 16 warps per SM with launch bounds of 384 / 512 threads, i.e. 168 / 128 registers and a few spills: 174 / 286 ms
    // against 133 ms at 8 warps on 2000 alignments of C3)
    nw = std::max(1, std::min(nw, 8));
    const long long want_warps = (n_items + 7) / 8;
    const void* fn = nullptr;
    const bool sink = items != nullptr;          // M-step "prepare": listed windows, converged q and entropy stored
    if (sink) {
        if (qg) fn = m->MC1n <= 1 ? (const void*)score_o1n_kernel<1, true, true> : m->MC1n == 2 ? (const void*)score_o1n_kernel<2, true, true> : (const void*)score_o1n_kernel<3, true, true>;
        else fn = m->MC1n <= 1 ? (const void*)score_o1n_kernel<1, false, true> : m->MC1n == 2 ? (const void*)score_o1n_kernel<2, false, true> : (const void*)score_o1n_kernel<3, false, true>;
    } else if (qg) fn = m->MC1n <= 1 ? (const void*)score_o1n_kernel<1, true> : m->MC1n == 2 ? (const void*)score_o1n_kernel<2, true> : (const void*)score_o1n_kernel<3, true>;
    else fn = m->MC1n <= 1 ? (const void*)score_o1n_kernel<1, false> : m->MC1n == 2 ? (const void*)score_o1n_kernel<2, false> : (const void*)score_o1n_kernel<3, false>;
    const size_t smem = fixed + per_warp * nw;
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)std::max<long long>(1, std::min<long long>((want_warps + nw - 1) / nw, ctx->sm_count));
    const size_t cache_bytes = (size_t)blocks * nw * L.region * sizeof(double) * (qg ? 2 : 1);
    const bool fresh = cache_bytes > ctx->o1n_cache.cap;
    CU(ctx->o1n_cache.ensure(cache_bytes));
    if (fresh) CU(cudaMemsetAsync(ctx->o1n_cache.p, 0, ctx->o1n_cache.cap, ctx->stream));   // idle slots read their (unused) rows
    const bool check = ds->n_gap_cols > 0;
    int* fb_count = nullptr;
    if (check) {
        // list of windows with gaps: [count (16 bytes)] [col0: int64 x n] [dst: int64 x n] [strand: uint8 x n]
        CU(ctx->o1n_fb.ensure(16 + (size_t)n_items * 17));
        fb_count = ctx->o1n_fb.as<int>();
        CU(cudaMemsetAsync(fb_count, 0, 16, ctx->stream));
    }
    Net1nParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_win_off; p.n_aln = ds->n_aln;
    p.n_windows = n_windows; p.n_strands = n_strands; p.strand0 = strands == 1 ? 1 : 0;
    p.W = h.W; p.I = h.I; p.S = h.S;
    p.prog = m->d_prog1n; p.prog_len = m->prog1n_len; p.tab = m->d_tab1n; p.leaf_tab = m->d_leaftab1n;
    p.out = d_out; p.sweeps = d_sweeps;
    p.counters = ctx->counters.as<unsigned long long>() + 2 * counter_slot;
    p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
    p.gc = ctx->o1n_cache.as<double>();
    p.gq = qg ? p.gc + (size_t)blocks * nw * L.region : nullptr;
    p.region = L.region; p.sregion = L.sregion; p.ab0 = L.ab0; p.msg0 = L.msg0; p.q0 = L.q0; p.onehot_row = L.onehot_row;
    p.f0_const = m->f0_const1n;
    p.check_gaps = check ? 1 : 0;
    p.fb_count = fb_count;
    p.fb_col0 = check ? (int64_t*)((char*)ctx->o1n_fb.p + 16) : nullptr;
    p.fb_dst = check ? p.fb_col0 + n_items : nullptr;
    p.fb_strand = check ? (uint8_t*)(p.fb_dst + n_items) : nullptr;
    p.item_col0 = sink ? items->col0 : nullptr; p.item_strand = sink ? items->strand : nullptr;
    p.sink_qint = sink ? items->q_int : nullptr; p.sink_ent = sink ? items->ent : nullptr;
    CU(cudaMemsetAsync(p.counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
    // (window over the q rows alone, leaving the observed-leaf sums to the normal policy: measured worse -- 31 instead of 18 GB of
    // DRAM traffic per 2 M windows)
    if (qg) l2_window(ctx, ctx->o1n_cache.p, cache_bytes);
    {
        KernelTimer kt_(ctx, "score_o1n_kernel");
        void* args[] = {&p};
        CU(cudaLaunchKernel(fn, dim3(blocks), dim3(32 * nw), args, smem, ctx->stream));
    }
    if (qg) l2_window(ctx, nullptr, 0);
    CU(cudaGetLastError());
    if (check) {
        ScoreItems it{p.fb_col0, p.fb_strand, nullptr, nullptr, nullptr};
        it.n_dev = fb_count; it.dst = p.fb_dst; it.fallback = true;
        if (sink) { it.q_int = items->q_int; it.q_leaf = items->q_leaf; it.ent = items->ent; it.sink_stride = n_items; }
        // (n_windows of the launch: the list's capacity bound for the grid; the kernel reads the real length on the device)
        int rc = launch_net1w(ctx, ds, m, d_win_off, std::min<long long>(n_items, 64LL * ctx->sm_count), 0, d_out, d_sweeps, counter_slot, &it);
        if (rc) return rc;
    }
    return 0;
}

static int launch_net1(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                       int strands, double* d_out, int32_t* d_sweeps, int counter_slot, const ScoreItems* items) {
    ModelHost& h = m->h;
    if (h.mode == PHYFOO_MODE_EXACT && !items)
        return fail(ctx, PHYFOO_EUNSUPPORTED, "exact likelihood of an order-1 network is infeasible in the reference too "
                                              "(SimpleNodeElimination's 9-node frontier, SimpleNodeElimination.java:33-45)");
    // the M-step "prepare" pass of a motif model (listed windows, q sinks, nothing background-specific) runs on the
    // node-per-thread kernel too
    const bool prepare = items && items->q_int && !items->col1 && !items->out_last && !items->stop_first && !items->fallback && strands == 0;
    if (ctx->order1_kernel == 2 && (!items || prepare)) {
        const int rc = launch_net1n(ctx, ds, m, d_win_off, n_windows, strands, d_out, d_sweeps, counter_slot, prepare ? items : nullptr);
        if (rc <= 0) return rc;
    }
    if (ctx->order1_kernel >= 1) return launch_net1w(ctx, ds, m, d_win_off, n_windows, strands, d_out, d_sweeps, counter_slot, items);
    const int n_strands = strands == 2 ? 2 : 1;
    const long long n_items = (long long)n_windows * n_strands;
    // windows per block: as many windows per SM as shared memory holds, in multiples of 8 (whole warps).  The kernel is
    // latency-bound (one quad per window), so windows in flight are what matters; the tables stay in global memory
    // behind L1/L2 (measured on C3: 48 windows/SM with the tables in L2 beat 32 windows/SM with the tables in shared
    // memory by 25 %).
    const size_t per_window = (size_t)h.W * h.I * 4 * sizeof(double) + (size_t)h.W * 12;
    const size_t fixed = (size_t)((m->prog_len + 1) & ~1) * sizeof(int) + 32 + 64;
    const size_t budget = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    int best_nwb = 0, best_per_sm = 0;
    for (int nwb = 8; nwb <= 64; nwb += 8) {
        const size_t need = fixed + per_window * nwb;
        if (need > budget) break;
        const int blocks_sm = (int)std::min<size_t>((228 * 1024) / (need + 1024), 2048 / (4 * nwb));
        if (blocks_sm * nwb > best_per_sm) { best_per_sm = blocks_sm * nwb; best_nwb = nwb; }
    }
    if (!best_nwb) return fail(ctx, PHYFOO_EUNSUPPORTED, "order-1 window state does not fit in shared memory (motif too wide or tree too large)");
    const size_t smem = fixed + per_window * best_nwb;
    CU(cudaFuncSetAttribute(score_o1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_o1_kernel, 4 * best_nwb, smem));
    if (per_sm < 1) return fail(ctx, PHYFOO_ECUDA, "order-1 kernel does not fit on an SM");
    const long long want = (n_items + best_nwb - 1) / best_nwb;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)per_sm * ctx->sm_count));
    const int64_t n_slots = (int64_t)blocks * best_nwb;
    CU(ctx->gstate.ensure((size_t)h.W * h.S * 4 * n_slots * sizeof(double)));
    Net1Params p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_win_off; p.n_aln = ds->n_aln;
    p.n_windows = n_windows; p.n_strands = n_strands; p.strand0 = strands == 1 ? 1 : 0;
    p.W = h.W; p.I = h.I; p.S = h.S; p.E1 = h.E1v(); p.prog_len = m->prog_len; p.prog = m->d_prog;
    p.tab = m->d_tab1;
    p.out = d_out; p.sweeps = d_sweeps;
    p.counters = ctx->counters.as<unsigned long long>() + 2 * counter_slot;
    p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
    p.gq = ctx->gstate.as<double>(); p.n_slots = n_slots; p.nwb = best_nwb;
    p.item_col0 = items ? items->col0 : nullptr; p.item_strand = items ? items->strand : nullptr;
    p.sink_qint = items ? items->q_int : nullptr; p.sink_qleaf = items ? items->q_leaf : nullptr; p.sink_ent = items ? items->ent : nullptr;
    CU(cudaMemsetAsync(p.counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
    {
        KernelTimer kt_(ctx, "score_o1_kernel");
        score_o1_kernel<<<blocks, 4 * best_nwb, smem, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    return PHYFOO_OK;
}

// order-0 mean field with F81-structured tables (tree_pass_f81.cuh): groups of a few warps whose lane count is a multiple of
// the motif width exchange the trees' free energies; returns 1 when the configuration does not fit (generic kernel then)
static int launch_score_f81(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                            int strands, double* d_out, int32_t* d_sweeps, int counter_slot) {
    ModelHost& h = m->h;
    const int W = h.W;
    const int n_strands = strands == 2 ? 2 : 1;
    const long long n_items = (long long)n_windows * n_strands;
    // group: the fewest warps whose lanes the windows fill best (W = 12: 3 warps = 8 windows, no idle lane)
    int best_nwg = 0; double best_use = 0.0;
    for (int nwg = 1; nwg <= 8; ++nwg) {
        if (32 * nwg < W) continue;
        const double use = (double)((32 * nwg) / W * W) / (32 * nwg);
        if (use > best_use + 1e-9) { best_use = use; best_nwg = nwg; }
    }
    if (!best_nwg) return 1;
    const int GL = 32 * best_nwg;
    int n_slots = 0;
    for (int i = 0; i < h.I; ++i) n_slots += h.ichild_begin[i + 1] > h.ichild_begin[i];
    const size_t state_per_thread = (size_t)(n_slots * 8 + h.I) * sizeof(double);
    const size_t budget = std::min<size_t>(ctx->smem_optin, 227 * 1024);
    auto fixed_for = [&](int T, int ng) {
        return ((size_t)h.Ef() * W + 64 + 2 * (size_t)T) * sizeof(double) + (size_t)ng * (GL / W) * sizeof(long long) +
               (size_t)(m->prog_len + h.I) * sizeof(int) + 64;
    };
    // state in shared memory when at least 8 warps per SM fit that way (a 12-species tree: 472 bytes per thread, 12 warps);
    // otherwise in global memory (coalesced, behind L1/L2) with as many groups as the register file allows
    int n_groups = std::max(1, std::min(15, 512 / GL));
    bool gstate = true;
    for (int ng = n_groups; ng >= 1; --ng) {
        const int T = ng * GL;
        if (fixed_for(T, ng) + state_per_thread * T <= budget) {
            if (T >= 256 || ng == n_groups) { n_groups = ng; gstate = false; }
            break;
        }
    }
    const int T = n_groups * GL;
    const size_t fixed = fixed_for(T, n_groups);
    if (fixed > budget) return 1;
    const size_t smem = fixed + (gstate ? 0 : state_per_thread * T);
    CU(cudaFuncSetAttribute(score_o0f_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long per_block = (long long)n_groups * (GL / W);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n_items + per_block - 1) / per_block, ctx->sm_count));
    ScoreFParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_win_off; p.n_aln = ds->n_aln;
    p.n_windows = n_windows; p.n_strands = n_strands; p.strand0 = strands == 1 ? 1 : 0;
    p.W = W; p.I = h.I; p.S = h.S; p.Ef = h.Ef(); p.prog_len = m->prog_len;
    p.tab = m->d_tabf; p.prog = m->d_prog;
    p.out = d_out; p.sweeps = d_sweeps;
    unsigned long long* counters = ctx->counters.as<unsigned long long>();
    p.counter_next = counters + 2 * counter_slot; p.counter_sweeps = counters + 2 * counter_slot + 1;
    p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
    p.group_lanes = GL; p.n_groups = n_groups; p.n_slots = n_slots;
    p.gstate = nullptr;
    if (gstate) {
        CU(ctx->gstate.ensure((size_t)blocks * T * state_per_thread));
        p.gstate = ctx->gstate.as<double>();
    }
    CU(cudaMemsetAsync(p.counter_next, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (gstate) l2_window(ctx, ctx->gstate.p, (size_t)blocks * T * state_per_thread);
    {
        KernelTimer kt_(ctx, "score_o0f_kernel");
        score_o0f_kernel<<<blocks, T, smem, ctx->stream>>>(p);
    }
    if (gstate) l2_window(ctx, nullptr, 0);
    CU(cudaGetLastError());
    return 0;
}

// The kernel compiled for this model's tree (jit_o0.hpp).  Returns 0 when launched, 1 when the caller should take the next
// kernel (no NVRTC / driver library, compilation failed -- ctx->jit_note says why), < 0... never: errors are PHYFOO_* > 1.
static int jit_get(phyfoo_ctx* ctx, const ModelHost& h, bool gaps, int nb, std::shared_ptr<JitKernel>* out) {
    JitApi& api = jit_api();
    if (!api.ok || !api.have_driver) { ctx->jit_note = api.why; return 1; }
    JitShape sh;
    // 1024 threads of 64 registers or 512 of 128: the generated pass needs ~100+ registers, so at most 512 threads
    if (!jit_o0_shape(h, gaps, std::min<size_t>(ctx->smem_optin, 227 * 1024), ctx->jit_max_threads, &sh)) { ctx->jit_note = "no block shape fits"; return 1; }
    const std::string src = jit_o0_source(h, sh, gaps, nb);
    auto it = ctx->jit_cache.find(src);
    if (it != ctx->jit_cache.end()) {
        if (!it->second) { return 1; }          // failed before: do not try again
        *out = it->second;
        return 0;
    }
    ctx->jit_cache[src] = nullptr;
    std::string log;
    std::vector<char> cubin;
    if (jit_compile(src, log, &cubin)) { ctx->jit_note = log; return 1; }
    auto k = std::make_shared<JitKernel>();
    int rc = api.ModuleLoadData(&k->module, cubin.data());
    if (rc) { ctx->jit_note = "cuModuleLoadData failed: " + std::to_string(rc); return 1; }
    rc = api.ModuleGetFunction(&k->fn, k->module, "score_o0j_kernel");
    if (rc) { ctx->jit_note = "cuModuleGetFunction failed: " + std::to_string(rc); return 1; }
    rc = api.FuncSetAttribute(k->fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)sh.smem);
    if (rc) { ctx->jit_note = "cuFuncSetAttribute(shared memory) failed: " + std::to_string(rc); return 1; }
    api.FuncGetAttribute(&k->regs, 4 /* CU_FUNC_ATTRIBUTE_NUM_REGS */, k->fn);
    k->shape = sh;
    ctx->jit_cache[src] = k;
    *out = k;
    return 0;
}

static int launch_score_jit(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                            int strands, double* d_out, int32_t* d_sweeps, int counter_slot) {
    ModelHost& h = m->h;
    std::shared_ptr<JitKernel> k;
    const int got = jit_get(ctx, h, ds->n_gap_cols > 0, ds->nb, &k);
    if (got) return got;
    const JitShape& sh = k->shape;
    const int n_strands = strands == 2 ? 2 : 1;
    const long long n_items = (long long)n_windows * n_strands;
    const long long per_block = (long long)sh.NG * (sh.GL / h.W);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n_items + per_block - 1) / per_block, ctx->sm_count));
    ScoreFParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_win_off; p.n_aln = ds->n_aln;
    p.n_windows = n_windows; p.n_strands = n_strands; p.strand0 = strands == 1 ? 1 : 0;
    p.W = h.W; p.I = h.I; p.S = h.S; p.Ef = h.Ef(); p.prog_len = m->prog_len;
    p.tab = m->d_tabf; p.prog = m->d_prog;
    p.out = d_out; p.sweeps = d_sweeps;
    unsigned long long* counters = ctx->counters.as<unsigned long long>();
    p.counter_next = counters + 2 * counter_slot; p.counter_sweeps = counters + 2 * counter_slot + 1;
    p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
    p.group_lanes = sh.GL; p.n_groups = sh.NG; p.n_slots = 0;
    p.gstate = nullptr;
    if (sh.gstate) {
        CU(ctx->gstate.ensure((size_t)blocks * sh.T * sh.state_doubles * sizeof(double)));
        p.gstate = ctx->gstate.as<double>();
    }
    CU(cudaMemsetAsync(p.counter_next, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (sh.gstate) l2_window(ctx, ctx->gstate.p, (size_t)blocks * sh.T * sh.state_doubles * sizeof(double));
    int rc;
    {
        KernelTimer kt_(ctx, "score_o0j_kernel");
        void* args[] = {&p};
        rc = jit_api().LaunchKernel(k->fn, blocks, 1, 1, sh.T, 1, 1, (unsigned)sh.smem, ctx->stream, args, nullptr);
    }
    if (sh.gstate) l2_window(ctx, nullptr, 0);
    if (rc) return fail(ctx, PHYFOO_ECUDA, "cuLaunchKernel(score_o0j_kernel) failed: " + std::to_string(rc));
    CU(cudaGetLastError());
    return 0;
}

// HKY as implemented (evolution/HKY.java:32,77-106): beta == 0, so a child symbol has non-zero probability only when it
// equals the parent's or is its transition partner (A<->G, C<->T), and every table has zero entries.  What the reference's
// mean field returns then (algorithm/MeanFieldForBayesNet.java; the restatement runs these cases, tests/test_gpu_parity.py):
//  * the stop rule never fires (the free energy is NaN or +inf in every sweep, and |x - x| < 1e-5 is false for both):
//    max_sweeps sweeps for every window;
//  * a tree with a hidden non-root node: the exponents of the root are -inf for every symbol in the very first sweep
//    (a hidden child at uniform q meets a zero table entry for every parent symbol), z = 0, q = 0/0: the score is NaN;
//  * a star tree: a column whose observed leaves are all purines or all pyrimidines keeps z > 0, and the free energy is +inf
//    through the zero-entry guard of the observed-leaf terms (:315-324): score -inf; a column with both classes has z = 0:
//    NaN.  Unobserved leaves use the "independent" table (bayesNet/CPF.java:146-152), which has no zeros, and change nothing.
//    (The reverse complement swaps the two classes as a whole, so both strands score alike.)
__global__ void hky_as_implemented_kernel(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ gaps, int64_t n_cols, int nb,
                                          const int64_t* __restrict__ col_off, const int64_t* __restrict__ win_off, int n_aln,
                                          int64_t n_windows, int n_strands, int W, int S, int star, int max_sweeps,
                                          double* __restrict__ out, int32_t* __restrict__ sweeps, unsigned long long* __restrict__ sweep_sum) {
    const long long n_items = (long long)n_windows * n_strands;
    unsigned long long local = 0;
    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (long long)gridDim.x * blockDim.x) {
        const int64_t w = item % n_windows;
        double v = NAN;
        if (star) {
            const int a = find_alignment_hint(win_off, n_aln, w, n_windows);
            const int64_t c0 = col_off[a] + (w - win_off[a]);
            bool mixed = false;
            for (int t = 0; t < W; ++t) {
                const ColWords cw = load_column(bases, gaps, n_cols, nb, c0 + t);
                bool pur = false, pyr = false;
                for (int sp = 0; sp < S; ++sp) { const int x = cw.obs(sp); if (x < 4) { if (x & 1) pyr = true; else pur = true; } }
                mixed = mixed || (pur && pyr);
            }
            v = mixed ? NAN : -INFINITY;
        }
        out[item] = v;
        if (sweeps) sweeps[item] = max_sweeps;
        local += (unsigned long long)max_sweeps;
    }
    if (local) atomicAdd(sweep_sum, local);
}

// scores into d_out[n_strands][n_windows]; strands: 0 = forward only, 1 = reverse only, 2 = both.
// items != NULL: score the listed windows (n_windows of them) instead and store their converged q.
static int launch_score(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const int64_t* d_win_off, int64_t n_windows,
                        int strands, double* d_out, int32_t* d_sweeps, int counter_slot, const ScoreItems* items) {
    int rc = model_upload(ctx, m);
    if (rc) return rc;
    ModelHost& h = m->h;
    // the M-step always runs the mean field, whatever CALC_LOGLIKELIHOOD says (PhyloBayesModel.java:142-147)
    const int mode = items ? PHYFOO_MODE_MEANFIELD : h.mode;
    const bool fallback = items && items->fallback;
    if (h.S != ds->S) return fail(ctx, PHYFOO_EINVAL, "model and dataset disagree on the number of species");
    if (h.evol == PHYFOO_EVOL_HKY_AS_IMPLEMENTED && mode == PHYFOO_MODE_MEANFIELD) {
        // HKY as the reference implements it (evolution/HKY.java:32: the transition/transversion ratio is never stored, so
        // every transversion has probability 0): the mean-field score of EVERY window is degenerate, and the library returns
        // the reference's degenerate values rather than an error (phyfoo_model_is_degenerate() flags the model)
        if (items) return fail(ctx, PHYFOO_EUNSUPPORTED, "HKY as implemented: the M-step objective is not finite for any parameter set (HKY.java:32)");
        if (h.order != 0) return fail(ctx, PHYFOO_EUNSUPPORTED, "HKY as implemented is offered for order-0 models only");
        if (h.S != ds->S) return fail(ctx, PHYFOO_EINVAL, "model and dataset disagree on the number of species");
        const int n_strands = strands == 2 ? 2 : 1;
        const long long n_items = (long long)n_windows * n_strands;
        if (n_items == 0) return PHYFOO_OK;
        ctx->last_path = 0;
        unsigned long long* counters = ctx->counters.as<unsigned long long>() + 2 * counter_slot;
        CU(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
        {
            KernelTimer kt_(ctx, "hky_as_implemented_kernel");
            hky_as_implemented_kernel<<<(unsigned)std::min<long long>((n_items + 255) / 256, 65535LL * 16), 256, 0, ctx->stream>>>(
                ds->d_bases, ds->d_gaps, ds->n_cols, ds->nb, ds->d_col_off, d_win_off, ds->n_aln, n_windows, n_strands, h.W, h.S, h.I == 1 ? 1 : 0,
                h.max_sweeps, d_out, d_sweeps, counters + 1);
        }
        CU(cudaGetLastError());
        return PHYFOO_OK;
    }
    if (mode == PHYFOO_MODE_MEANFIELD && !h.tables_finite())
        return fail(ctx, PHYFOO_EUNSUPPORTED, "a CPF entry is 0: the reference's free energy is not finite for this parameter set");
    const int n_strands = strands == 2 ? 2 : 1;
    const long long n_items = (long long)n_windows * n_strands;
    if (n_items == 0) return PHYFOO_OK;
    if (h.order == 1) { ctx->last_path = 2; return launch_net1(ctx, ds, m, d_win_off, n_windows, strands, d_out, d_sweeps, counter_slot, items); }
    if (!items && memo_applies(ctx, ds, h, mode, n_items)) {
        const MemoJob job{m, mode, d_win_off, n_windows, strands, d_out, d_sweeps, counter_slot};
        return launch_memo_jobs(ctx, ds, &job, 1);
    }
    if (!fallback) ctx->last_path = 0;
    if (!items && mode == PHYFOO_MODE_MEANFIELD && m->f81_ok && h.W <= 256 &&
        ctx->order0_kernel != 0) {
        if (ctx->order0_kernel != 1) {
            // the kernel compiled for this model's tree (any number of species)
            rc = launch_score_jit(ctx, ds, m, d_win_off, n_windows, strands, d_out, d_sweeps, counter_slot);
            if (rc != 1) return rc;
        }
        // precompiled kernels: F81-structured for more than 6 species (or on request), the generic one below otherwise
        if (ctx->order0_kernel >= 1 || 3 * h.S > 18) {
            rc = launch_score_f81(ctx, ds, m, d_win_off, n_windows, strands, d_out, d_sweeps, counter_slot);
            if (rc <= 0) return rc;
        }
    }
    ScorePlan pl;
    // the fallback list of the pattern-table path is short (its length lives on the device): one block per SM
    rc = plan_score(ctx, h, m->prog_len, fallback ? std::min<long long>(n_items, 64LL * ctx->sm_count) : n_items, mode, items != nullptr, &pl);
    if (rc) return rc;
    ScoreParams p;
    p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
    p.col_off = ds->d_col_off; p.win_off = d_win_off; p.n_aln = ds->n_aln;
    p.n_windows = n_windows; p.n_strands = n_strands; p.strand0 = strands == 1 ? 1 : 0;
    p.W = h.W; p.I = h.I; p.S = h.S; p.E = h.E; p.prog_len = m->prog_len;
    p.tab = mode == PHYFOO_MODE_EXACT ? m->d_probtab : m->d_logtab;
    p.prog = m->d_prog;
    p.out = d_out; p.sweeps = d_sweeps;
    // counters: [2 slot] work counter, [2 slot + 1] sweep sum; a fallback launch continues the sweep sum of the
    // pattern-table pass and pulls its work from a counter of its own ([8 + slot])
    unsigned long long* counters = ctx->counters.as<unsigned long long>();
    p.counter_next = fallback ? counters + 8 + counter_slot : counters + 2 * counter_slot;
    p.counter_sweeps = counters + 2 * counter_slot + 1;
    p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
    p.groups = pl.groups;
    p.item_col0 = nullptr; p.item_strand = nullptr; p.q_int = nullptr; p.q_leaf = nullptr; p.ent = nullptr;
    p.n_items_dev = nullptr; p.item_dst = nullptr;
    if (items) {
        p.item_col0 = items->col0; p.item_strand = items->strand; p.q_int = items->q_int; p.q_leaf = items->q_leaf; p.ent = items->ent;
        p.n_items_dev = items->n_dev; p.item_dst = items->dst;
    }
    p.gstate = nullptr;
    if (pl.gstate) {
        CU(ctx->gstate.ensure((size_t)pl.blocks * pl.threads * h.I * 8 * sizeof(double)));
        p.gstate = ctx->gstate.as<double>();
    }
    if (!fallback) {            // (a fallback launch finds its counters cleared by the pattern-table pass)
        CU(cudaMemsetAsync(p.counter_next, 0, sizeof(unsigned long long), ctx->stream));
        CU(cudaMemsetAsync(p.counter_sweeps, 0, sizeof(unsigned long long), ctx->stream));
    }
    {
        KernelTimer kt_(ctx, fallback ? "score_o0_kernel(fallback list)" : "score_o0_kernel");
        if (mode == PHYFOO_MODE_EXACT) score_o0_kernel<PHYFOO_MODE_EXACT, false><<<pl.blocks, pl.threads, pl.smem, ctx->stream>>>(p);
        else if (items) score_o0_kernel<PHYFOO_MODE_MEANFIELD, true><<<pl.blocks, pl.threads, pl.smem, ctx->stream>>>(p);
        else score_o0_kernel<PHYFOO_MODE_MEANFIELD, false><<<pl.blocks, pl.threads, pl.smem, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    return PHYFOO_OK;
}

struct CallTimer {
    phyfoo_ctx* ctx;
    explicit CallTimer(phyfoo_ctx* c) : ctx(c) { ctx->launches = 0; ctx->n_ktimes = 0; cudaEventRecord(ctx->ev0, ctx->stream); }
    float stop() {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        return ms;
    }
};

static int read_sweeps(phyfoo_ctx* ctx, int slot, int64_t* out) {
    unsigned long long v[2];
    CU(cudaMemcpyAsync(v, ctx->counters.as<unsigned long long>() + 2 * slot, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out = (int64_t)v[1];
    return PHYFOO_OK;
}

extern "C" int phyfoo_score_windows(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int32_t strand,
                                    double* out, int32_t* sweeps_out, phyfoo_stats* stats) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !fg || !out) return fail(ctx, PHYFOO_EINVAL, "score_windows: NULL argument");
    if (strand != 0 && strand != 1) return fail(ctx, PHYFOO_EINVAL, "score_windows: strand must be 0 or 1");
    CU(cudaSetDevice(ctx->device));
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, fg->h.W, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t nw = h_woff->back();
    CU(ctx->fg_scores.ensure((size_t)std::max<int64_t>(nw, 1) * sizeof(double)));
    if (sweeps_out) CU(ctx->sweeps_buf.ensure((size_t)std::max<int64_t>(nw, 1) * sizeof(int32_t)));
    CallTimer tm(ctx);
    rc = launch_score(ctx, ds, fg, d_woff, nw, strand, ctx->fg_scores.as<double>(), sweeps_out ? ctx->sweeps_buf.as<int32_t>() : nullptr, 0);
    if (rc) return rc;
    const float ms = tm.stop();
    CU(cudaGetLastError());
    if (nw > 0) {
        CU(cudaMemcpyAsync(out, ctx->fg_scores.p, (size_t)nw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (sweeps_out) CU(cudaMemcpyAsync(sweeps_out, ctx->sweeps_buf.p, (size_t)nw * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_windows = nw; stats->kernel_launches = ctx->launches; stats->gpu_ms = ms;
        if (nw > 0 && (rc = read_sweeps(ctx, 0, &stats->sweeps_fg))) return rc;
    }
    return PHYFOO_OK;
}

// per-column background scores (order 0): the same kernel with windows of width 1
static int launch_bg_columns(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* d_out, int32_t* d_sweeps, int slot) {
    if (bg->h.kind != PHYFOO_MODEL_BG) return fail(ctx, PHYFOO_EINVAL, "bg_columns: model is not a background model");
    if (bg->h.order != 0) return fail(ctx, PHYFOO_EUNSUPPORTED, "bg_columns is for order 0: a background of order >= 1 has range terms (phyfoo_bg_range_terms)");
    return launch_score(ctx, ds, bg, ds->d_col_off, ds->n_cols, 0, d_out, d_sweeps, slot);
}

// order-1 background (models/PhyloBackground.java:241-305): every 2-column window (u-1, u) of every alignment through the
// order-1 network kernel (W = 2); its free energy restricted to the last tree is the conditional term of column u
// (:284-297 calcFreeEnergy(order, order)), restricted to the first tree the start term of a range beginning at u-1
// (:273-283 calcFreeEnergy(0, 0)).
__global__ void bg1_items_kernel(const int64_t* __restrict__ win_off, const int64_t* __restrict__ col_off, int n_aln, int64_t n,
                                 int64_t* __restrict__ col0) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int a = find_alignment(win_off, n_aln, u);
    col0[u] = col_off[a] + (u - win_off[a]);
}
// cond[c0 + 1] = last-tree term of window c0; first[c0] = whole-window term minus the last-tree term
__global__ void bg1_scatter_kernel(const int64_t* __restrict__ col0, const double* __restrict__ total, const double* __restrict__ last,
                                   int64_t n, double* __restrict__ cond, double* __restrict__ first) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    cond[col0[u] + 1] = last[u];
    if (first) first[col0[u]] = total[u] - last[u];
}

// pairs of columns for the single-column ranges: (c, c) for every column
__global__ void bg1_pairs_same_kernel(int64_t n, int64_t* __restrict__ colA, int64_t* __restrict__ colB) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    colA[u] = u; colB[u] = u;
}
// ... and (j - 1, min(j + W, L - 1)) for offset j of every alignment (the range left of the motif, whose second tree still
// holds the last column of the range bg(s, e) scored just before it; j = 0 has no such range: a dummy pair)
__global__ void bg1_pairs_left_kernel(const int64_t* __restrict__ win_off, const int64_t* __restrict__ col_off, int n_aln, int64_t n, int W,
                                      int64_t* __restrict__ colA, int64_t* __restrict__ colB) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int a = find_alignment(win_off, n_aln, u);
    const int64_t c0 = col_off[a], L = col_off[a + 1] - c0, j = u - win_off[a];
    colA[u] = c0 + (j > 0 ? j - 1 : 0);
    colB[u] = c0 + (j + W < L - 1 ? j + W : L - 1);
}
__global__ void bg1_first_term_kernel(const double* __restrict__ total, const double* __restrict__ last, int64_t n, double* __restrict__ out) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) out[u] = total[u] - last[u];
}

// Order-1 background terms on the device, in ctx->bg_scores: cond[nc], first[nc] (every 2-column window of every alignment
// through the order-1 network kernel, W = 2) and, for the ZOOPS E-step (W_motif > 0), single[nc] and left[windows]: the
// single-column ranges bg(j + W, e) and bg(s, j - 1) of jstacs SingleHiddenMotifMixture.java:536-540, whose second tree the
// reference leaves at the observation of the previous call (MeanFieldForBayesNet.java:84) -- column e = min(j + W, L - 1) in
// both cases, because bg(s, e) is scored just before them.
struct Bg1Terms { double* cond; double* first; double* single; double* left; };
static int bg1_terms_device(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, int W_motif, const int64_t* d_woff_motif,
                            int64_t nw_motif, Bg1Terms* out) {
    if (bg->h.kind != PHYFOO_MODEL_BG || bg->h.order != 1) return fail(ctx, PHYFOO_EINVAL, "needs a background model of order 1");
    if (bg->h.mode != PHYFOO_MODE_MEANFIELD) return fail(ctx, PHYFOO_EUNSUPPORTED, "exact mode is not defined for order-1 networks");
    if (bg->h.S != ds->S) return fail(ctx, PHYFOO_EINVAL, "model and dataset disagree on the number of species");
    if (ds->n_aln > 0 && ds->min_len < 2) return fail(ctx, PHYFOO_EINVAL, "order-1 background: an alignment is shorter than order + 1 columns (the reference's score of such a range is not a function of the inputs, MeanFieldForBayesNet.java:84)");
    int rc = model_upload(ctx, bg);
    if (rc) return rc;
    if (!bg->h.tables_finite()) return fail(ctx, PHYFOO_EUNSUPPORTED, "a CPF entry is 0: the reference's free energy is not finite for this parameter set");
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    rc = dataset_win_off(ctx, ds, 2, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t n = h_woff->back(), nc = ds->n_cols;
    const int64_t n_extra = W_motif > 0 ? std::max(nc, nw_motif) : 0;
    const int64_t n_items = std::max(n, n_extra);
    // [total | last] (scratch of the current launch) | cond | first | single | left
    CU(ctx->bg_scores.ensure((size_t)(2 * n_items + 3 * nc + (W_motif > 0 ? nw_motif : 0) + 4) * sizeof(double)));
    CU(ctx->misc.ensure((size_t)2 * std::max<int64_t>(n_items, 1) * sizeof(int64_t)));
    double* d_total = ctx->bg_scores.as<double>();
    double* d_last = d_total + n_items;
    out->cond = d_last + n_items; out->first = out->cond + nc; out->single = out->first + nc; out->left = out->single + nc;
    int64_t* d_colA = ctx->misc.as<int64_t>();
    int64_t* d_colB = d_colA + std::max<int64_t>(n_items, 1);
    CU(cudaMemsetAsync(out->cond, 0, (size_t)2 * nc * sizeof(double), ctx->stream));
    const int keep = ctx->order1_kernel;
    auto run = [&](int64_t cnt, const int64_t* colB, bool stop_first) {
        ScoreItems items{d_colA, nullptr, nullptr, nullptr, nullptr};
        items.out_last = d_last; items.col1 = colB; items.stop_first = stop_first;
        ctx->order1_kernel = 1;                   // the per-tree output lives in the wavefront kernel
        const int r = launch_net1(ctx, ds, bg, nullptr, cnt, 0, d_total, nullptr, 1, &items);
        ctx->order1_kernel = keep;
        return r;
    };
    if (n > 0) {
        {
            KernelTimer kt_(ctx, "bg1_items_kernel");
            bg1_items_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_woff, ds->d_col_off, ds->n_aln, n, d_colA);
        }
        CU(cudaGetLastError());
        if ((rc = run(n, nullptr, false))) return rc;
        {
            KernelTimer kt_(ctx, "bg1_scatter_kernel");
            bg1_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_colA, d_total, d_last, n, out->cond, out->first);
        }
        CU(cudaGetLastError());
    }
    if (W_motif > 0 && nc > 0) {
        bg1_pairs_same_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(nc, d_colA, d_colB);
        CU(cudaGetLastError());
        if ((rc = run(nc, d_colB, true))) return rc;
        bg1_first_term_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(d_total, d_last, nc, out->single);
        CU(cudaGetLastError());
        if (nw_motif > 0) {
            bg1_pairs_left_kernel<<<(unsigned)((nw_motif + 255) / 256), 256, 0, ctx->stream>>>(d_woff_motif, ds->d_col_off, ds->n_aln, nw_motif, W_motif, d_colA, d_colB);
            CU(cudaGetLastError());
            if ((rc = run(nw_motif, d_colB, true))) return rc;
            bg1_first_term_kernel<<<(unsigned)((nw_motif + 255) / 256), 256, 0, ctx->stream>>>(d_total, d_last, nw_motif, out->left);
            CU(cudaGetLastError());
        }
    }
    return PHYFOO_OK;
}

// ---- background models of any order >= 1 through the generic net kernel (netg.cuh) ------------------------------------------
// Items name the column of every tree.  Full windows (order + 1 consecutive columns of an alignment) give, per column c,
//   first[c] = - sum of the free energies of the single trees 0 .. order-1 of the window starting at c   (PhyloBackground.java:273-283)
//   cond[c]  = - free energy of the last tree of the window ending at c                                    (:284-297)
// so that getLogProbFor(seq, s, e) = first[s] + sum cond[s+order .. e] for ranges of at least order + 1 columns.  The ZOOPS E-step
// (jstacs SingleHiddenMotifMixture.java:536-540) also asks, per offset j, for the flanks bg(s, j-1) and bg(j+W, e) with
// s = max(j - order, 0), e = min(j + W + order - 1, L - 1): both shorter than order + 1 columns, so their trailing trees hold what
// the call before them left there (MeanFieldForBayesNet.java:84): bg(s, e) ends on the window (e-order .. e); the left flank
// then observes its n1 = j - s columns in the trees 0 .. n1-1, the right flank its n2 = e - j - W + 1 columns in 0 .. n2-1.
static int netg_upload(phyfoo_ctx* ctx, phyfoo_model* m) {
    if (m->netg_uploaded == m->h.version) return PHYFOO_OK;
    std::vector<int> rec, lists; std::vector<double> tab;
    try { m->h.netg_build(rec, lists, tab); } catch (const std::exception& e) { return fail(ctx, PHYFOO_EINVAL, std::string("generic net: ") + e.what()); }
    for (double v : tab) if (!std::isfinite(v)) return fail(ctx, PHYFOO_EUNSUPPORTED, "a CPF entry is 0: the reference's free energy is not finite for this parameter set");
    if (!m->d_netg_rec || lists.size() > m->netg_lists_len || tab.size() > m->netg_tab_len) {
        cudaFree(m->d_netg_rec); cudaFree(m->d_netg_lists); cudaFree(m->d_netg_tab);
        m->d_netg_rec = nullptr; m->d_netg_lists = nullptr; m->d_netg_tab = nullptr;
        CU(cudaMalloc(&m->d_netg_rec, rec.size() * sizeof(int)));
        CU(cudaMalloc(&m->d_netg_lists, std::max<size_t>(lists.size(), 1) * sizeof(int)));
        CU(cudaMalloc(&m->d_netg_tab, tab.size() * sizeof(double)));
        m->netg_lists_len = lists.size(); m->netg_tab_len = tab.size();
    }
    CU(cudaMemcpyAsync(m->d_netg_rec, rec.data(), rec.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(m->d_netg_lists, lists.data(), lists.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(m->d_netg_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));      // the vectors are stack-scoped
    m->netg_tab_used = tab.size();
    m->netg_uploaded = m->h.version;
    return PHYFOO_OK;
}

// full windows: item u = window u of width T (alignment-major); cols[t][u] = first column + t
__global__ void bgk_items_full_kernel(const int64_t* __restrict__ win_off, const int64_t* __restrict__ col_off, int n_aln, int64_t n, int T,
                                      int64_t* __restrict__ cols) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int a = find_alignment(win_off, n_aln, u);
    const int64_t c = col_off[a] + (u - win_off[a]);
    for (int t = 0; t < T; ++t) cols[(size_t)t * n + u] = c + t;
}
__global__ void bgk_scatter_full_kernel(const int64_t* __restrict__ cols, const double* __restrict__ ofirst, const double* __restrict__ olast,
                                        int64_t n, int k, double* __restrict__ cond, double* __restrict__ first) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    cond[cols[u] + k] = -olast[u];
    first[cols[u]] = -ofirst[u];
}
// the two flank ranges of offset j (item u = motif window u): right == 0: bg(s, j-1); right == 1: bg(j+W, e).  A flank without
// columns is a dummy item (len 1, its value is not used).
__global__ void bgk_items_flank_kernel(const int64_t* __restrict__ win_off, const int64_t* __restrict__ col_off, int n_aln, int64_t n, int W, int k,
                                       int right, int64_t* __restrict__ cols, int32_t* __restrict__ len) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int a = find_alignment(win_off, n_aln, u);
    const int64_t c0 = col_off[a], L = col_off[a + 1] - c0, j = u - win_off[a];
    const int64_t s = j > k ? j - k : 0, e = j + W + k - 1 < L - 1 ? j + W + k - 1 : L - 1;
    const int64_t n1 = j - s, n2 = e - (j + W) + 1;
    for (int t = 0; t <= k; ++t) {
        const int64_t after_full = e - k + t;                              // what bg(s, e) left in tree t
        const int64_t after_left = (n1 > 0 && t < n1) ? s + t : after_full;
        int64_t c;
        if (!right) c = after_left;
        else c = t < n2 ? j + W + t : after_left;
        cols[(size_t)t * n + u] = c0 + c;
    }
    const int64_t ln = right ? n2 : n1;
    len[u] = (int32_t)(ln > 0 ? ln : 1);
}
__global__ void bgk_negate_kernel(const double* __restrict__ in, int64_t n, double* __restrict__ out) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) out[u] = -in[u];
}

struct BgkTerms { double* cond; double* first; double* left; double* right; };
static int bgk_terms_device(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, int W_motif, const int64_t* d_woff_motif,
                            int64_t nw_motif, BgkTerms* out, int32_t* d_sweeps_full = nullptr) {
    ModelHost& h = bg->h;
    const int k = h.order, T = k + 1;
    if (h.kind != PHYFOO_MODEL_BG || k < 1) return fail(ctx, PHYFOO_EINVAL, "needs a background model of order >= 1");
    if (T > phyfoo::NETG_MAX_TREES) return fail(ctx, PHYFOO_EUNSUPPORTED, "background order above 7");
    if (h.mode != PHYFOO_MODE_MEANFIELD) return fail(ctx, PHYFOO_EUNSUPPORTED, "exact mode is not defined for networks of order >= 1");
    if (h.evol == PHYFOO_EVOL_HKY_AS_IMPLEMENTED) return fail(ctx, PHYFOO_EUNSUPPORTED, "HKY as implemented: background orders >= 1 are not supported");
    if (h.S != ds->S) return fail(ctx, PHYFOO_EINVAL, "model and dataset disagree on the number of species");
    if (ds->n_aln > 0 && ds->min_len < T) return fail(ctx, PHYFOO_EINVAL, "background of order k: an alignment is shorter than k + 1 columns (the reference's score of such a range is not a function of the inputs, MeanFieldForBayesNet.java:84)");
    int rc = netg_upload(ctx, bg);
    if (rc) return rc;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    rc = dataset_win_off(ctx, ds, T, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t n = h_woff->back(), nc = ds->n_cols;
    const int64_t n_items = std::max<int64_t>(std::max(n, W_motif > 0 ? nw_motif : 0), 1);
    // [ofirst | olast] (scratch of the current launch) | cond | first | left | right
    CU(ctx->bg_scores.ensure((size_t)(2 * n_items + 2 * nc + 2 * (W_motif > 0 ? nw_motif : 0) + 4) * sizeof(double)));
    CU(ctx->misc.ensure((size_t)T * n_items * sizeof(int64_t) + (size_t)n_items * sizeof(int32_t)));
    double* d_ofirst = ctx->bg_scores.as<double>();
    double* d_olast = d_ofirst + n_items;
    out->cond = d_olast + n_items; out->first = out->cond + nc; out->left = out->first + nc; out->right = out->left + (W_motif > 0 ? nw_motif : 0);
    int64_t* d_cols = ctx->misc.as<int64_t>();
    int32_t* d_len = (int32_t*)(d_cols + (size_t)T * n_items);
    CU(cudaMemsetAsync(out->cond, 0, (size_t)2 * nc * sizeof(double), ctx->stream));
    const int G = h.W * h.N;
    unsigned long long* sweep_sum = ctx->counters.as<unsigned long long>() + 2 * 1 + 1;          // slot 1 = background
    CU(cudaMemsetAsync(ctx->counters.as<unsigned long long>() + 2, 0, 2 * sizeof(unsigned long long), ctx->stream));
    auto run = [&](int64_t cnt, const int32_t* len, int32_t* sweeps, const char* name) -> int {
        // q in global memory, one slot per thread of the grid: at most ~512 MB
        const size_t per_thread = (size_t)G * 4 * sizeof(double);
        long long blocks = std::min<long long>((cnt + 127) / 128, 8LL * ctx->sm_count);
        blocks = std::max<long long>(1, std::min<long long>(blocks, (long long)((512ull << 20) / (per_thread * 128))));
        const size_t stride = (size_t)blocks * 128;
        CU(ctx->gstate.ensure(per_thread * stride));
        NetgParams p;
        p.bases = ds->d_bases; p.gaps = ds->d_gaps; p.n_cols = ds->n_cols; p.nb = ds->nb;
        p.view.n_nodes = G; p.view.n_trees = h.W; p.view.nodes_per_tree = h.N;
        p.view.rec = bg->d_netg_rec; p.view.lists = bg->d_netg_lists; p.view.tab = bg->d_netg_tab;
        p.item_cols = d_cols; p.item_len = len; p.n_items = cnt;
        p.q = ctx->gstate.as<double>(); p.q_stride = stride;
        p.out_first = d_ofirst; p.out_last = d_olast; p.sweeps = sweeps; p.sweep_sum = sweep_sum;
        p.max_sweeps = h.max_sweeps; p.min_sweeps = h.min_sweeps; p.tol = h.tol;
        {
            KernelTimer kt_(ctx, name);
            score_netg_kernel<<<(unsigned)blocks, 128, 0, ctx->stream>>>(p);
        }
        CU(cudaGetLastError());
        return PHYFOO_OK;
    };
    if (n > 0) {
        bgk_items_full_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_woff, ds->d_col_off, ds->n_aln, n, T, d_cols);
        CU(cudaGetLastError());
        if ((rc = run(n, nullptr, d_sweeps_full, "score_netg_kernel"))) return rc;
        bgk_scatter_full_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_cols, d_ofirst, d_olast, n, k, out->cond, out->first);
        CU(cudaGetLastError());
    }
    if (W_motif > 0 && nw_motif > 0) {
        for (int right = 0; right < 2; ++right) {
            bgk_items_flank_kernel<<<(unsigned)((nw_motif + 255) / 256), 256, 0, ctx->stream>>>(d_woff_motif, ds->d_col_off, ds->n_aln, nw_motif, W_motif, k,
                                                                                                right, d_cols, d_len);
            CU(cudaGetLastError());
            if ((rc = run(nw_motif, d_len, nullptr, right ? "score_netg_kernel(right flanks)" : "score_netg_kernel(left flanks)"))) return rc;
            bgk_negate_kernel<<<(unsigned)((nw_motif + 255) / 256), 256, 0, ctx->stream>>>(d_ofirst, nw_motif, right ? out->right : out->left);
            CU(cudaGetLastError());
        }
    }
    return PHYFOO_OK;
}

extern "C" int phyfoo_bg_range_terms(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* cond_out, double* first_out,
                                     int32_t* sweeps_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !bg || !cond_out) return fail(ctx, PHYFOO_EINVAL, "bg_range_terms: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int64_t nc = ds->n_cols;
    if (bg->h.kind != PHYFOO_MODEL_BG || bg->h.order < 1) return fail(ctx, PHYFOO_EINVAL, "bg_range_terms: needs a background model of order >= 1 (order 0: phyfoo_bg_columns)");
    const int T = bg->h.order + 1;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, T, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t n = h_woff->back();
    if (sweeps_out) CU(ctx->sweeps_buf.ensure((size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)));
    BgkTerms t;
    rc = bgk_terms_device(ctx, ds, bg, 0, nullptr, 0, &t, sweeps_out ? ctx->sweeps_buf.as<int32_t>() : nullptr);
    if (rc) return rc;
    if (nc == 0) return PHYFOO_OK;
    CU(cudaMemcpyAsync(cond_out, t.cond, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (first_out) CU(cudaMemcpyAsync(first_out, t.first, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sweeps_out && n > 0) CU(cudaMemcpyAsync(sweeps_out, ctx->sweeps_buf.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_bg_columns_order1(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* cond_out, double* first_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !bg || !cond_out) return fail(ctx, PHYFOO_EINVAL, "bg_columns_order1: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    Bg1Terms t;
    const int rc = bg1_terms_device(ctx, ds, bg, 0, nullptr, 0, &t);
    if (rc) return rc;
    const int64_t nc = ds->n_cols;
    if (nc == 0) return PHYFOO_OK;
    CU(cudaMemcpyAsync(cond_out, t.cond, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (first_out) CU(cudaMemcpyAsync(first_out, t.first, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_bg_columns(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* out,
                                 int32_t* sweeps_out, phyfoo_stats* stats) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !bg || !out) return fail(ctx, PHYFOO_EINVAL, "bg_columns: NULL argument");
    CU(cudaSetDevice(ctx->device));
    const int64_t nc = ds->n_cols;
    CU(ctx->bg_scores.ensure((size_t)std::max<int64_t>(nc, 1) * sizeof(double)));
    if (sweeps_out) CU(ctx->sweeps_buf.ensure((size_t)std::max<int64_t>(nc, 1) * sizeof(int32_t)));
    CallTimer tm(ctx);
    int rc = launch_bg_columns(ctx, ds, bg, ctx->bg_scores.as<double>(), sweeps_out ? ctx->sweeps_buf.as<int32_t>() : nullptr, 1);
    if (rc) return rc;
    const float ms = tm.stop();
    CU(cudaGetLastError());
    if (nc > 0) {
        CU(cudaMemcpyAsync(out, ctx->bg_scores.p, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (sweeps_out) CU(cudaMemcpyAsync(sweeps_out, ctx->sweeps_buf.p, (size_t)nc * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_columns = nc; stats->kernel_launches = ctx->launches; stats->gpu_ms = ms;
        if (nc > 0 && (rc = read_sweeps(ctx, 1, &stats->sweeps_bg))) return rc;
    }
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// E-step
// ------------------------------------------------------------------------------------------------
// Enqueues the whole E-step on ctx->stream; leaves [ll, w0, w1] in d_out3.  aln_w_dev nullable (device).
static int enqueue_estep(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, phyfoo_model* bg, const double* logw,
                         const double* strand_logw, const double* aln_w_dev, double* d_gamma, double* d_profile,
                         int32_t* d_argoff, uint8_t* d_argstrand, double* d_out3, int64_t* nw_out) {
    if (fg->h.kind != PHYFOO_MODEL_FG) return fail(ctx, PHYFOO_EINVAL, "zoops_estep: fg is not a motif model");
    const int W = fg->h.W;
    if (ds->n_aln == 0) return fail(ctx, PHYFOO_EINVAL, "zoops_estep: empty sample");
    if (ds->min_len < W) return fail(ctx, PHYFOO_EINVAL, "zoops_estep: an alignment is shorter than the motif");
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, W, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t nw = h_woff->back();
    *nw_out = nw;
    const int n_strands = strand_logw ? 2 : 1;
    CU(ctx->fg_scores.ensure((size_t)nw * n_strands * sizeof(double)));
    CU(ctx->bg_scores.ensure((size_t)ds->n_cols * sizeof(double)));
    CU(ctx->part.ensure((size_t)3 * ds->n_aln * sizeof(double)));
    ctx->kev_valid = false;
    CU(cudaEventRecord(ctx->kev[0], ctx->stream));
    Bg1Terms bg1{nullptr, nullptr, nullptr, nullptr};
    BgkTerms bgk{nullptr, nullptr, nullptr, nullptr};
    bool have_bg1 = false, have_bgk = false;
    const bool both_memo = bg->h.kind == PHYFOO_MODEL_BG && bg->h.order == 0 && fg->h.S == ds->S && bg->h.S == ds->S &&
                           fg->h.evol != PHYFOO_EVOL_HKY_AS_IMPLEMENTED && bg->h.evol != PHYFOO_EVOL_HKY_AS_IMPLEMENTED &&
                           fg->h.order == 0 && fg->h.tables_finite() && bg->h.tables_finite() &&
                           memo_applies(ctx, ds, fg->h, fg->h.mode, (long long)nw * n_strands) &&
                           memo_applies(ctx, ds, bg->h, bg->h.mode, (long long)ds->n_cols);
    if (both_memo) {
        // background and motif trajectories in ONE launch (both are small, latency-bound grids), then the two gathers
        const MemoJob jobs[2] = {{bg, bg->h.mode, ds->d_col_off, ds->n_cols, 0, ctx->bg_scores.as<double>(), nullptr, 1},
                                 {fg, fg->h.mode, d_woff, nw, strand_logw ? 2 : 0, ctx->fg_scores.as<double>(), nullptr, 0}};
        rc = launch_memo_jobs(ctx, ds, jobs, 2);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->kev[1], ctx->stream));   // phases are not separable on this path: see phyfoo_last_kernel_times
    } else {
        if (bg->h.kind == PHYFOO_MODEL_BG && bg->h.order >= 1 && (bg->h.order >= 2 || ctx->bg_generic)) {
            rc = bgk_terms_device(ctx, ds, bg, W, d_woff, nw, &bgk);
            bg1.cond = bgk.cond; bg1.first = bgk.first; bg1.left = bgk.left; bg1.single = nullptr;
            have_bg1 = true; have_bgk = true;
        } else if (bg->h.kind == PHYFOO_MODEL_BG && bg->h.order == 1) {
            rc = bg1_terms_device(ctx, ds, bg, W, d_woff, nw, &bg1);
            have_bg1 = true;
        } else rc = launch_bg_columns(ctx, ds, bg, ctx->bg_scores.as<double>(), nullptr, 1);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->kev[1], ctx->stream));
        rc = launch_score(ctx, ds, fg, d_woff, nw, strand_logw ? 2 : 0, ctx->fg_scores.as<double>(), nullptr, 0);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ctx->kev[2], ctx->stream));
    ZoopsParams z;
    z.fwd = ctx->fg_scores.as<double>();
    z.rev = strand_logw ? ctx->fg_scores.as<double>() + nw : nullptr;
    z.bgcol = have_bg1 ? bg1.cond : ctx->bg_scores.as<double>();
    z.bg1_first = have_bg1 ? bg1.first : nullptr; z.bg1_single = bg1.single; z.bg1_left = bg1.left;
    z.bgk_right = have_bgk ? bgk.right : nullptr; z.bg_order = have_bg1 ? bg->h.order : 0;
    z.col_off = ds->d_col_off; z.win_off = d_woff; z.n_aln = ds->n_aln; z.W = W;
    z.logw0 = logw[0]; z.logw1 = logw[1];
    z.has_strand = strand_logw ? 1 : 0;
    z.slw0 = strand_logw ? strand_logw[0] : 0.0; z.slw1 = strand_logw ? strand_logw[1] : 0.0;
    z.aln_w = aln_w_dev;
    z.gamma = d_gamma; z.profile = d_profile;
    z.part = ctx->part.as<double>();
    z.argoff = d_argoff; z.argstrand = d_argstrand;
    z.out3 = d_out3;
    z.ticket = (unsigned int*)(ctx->counters.as<unsigned long long>() + 12);
    size_t zs = (size_t)2 * ds->max_len * sizeof(double);
    z.stage = zs <= 96 * 1024 ? 1 : 0;
    if (!z.stage) zs = 0;
    if (zs > 48 * 1024) CU(cudaFuncSetAttribute(zoops_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zs));
    if (ds->n_aln > 0) {
        KernelTimer kt_(ctx, "zoops_kernel");
        zoops_kernel<<<ds->n_aln, 256, zs, ctx->stream>>>(z);
    } else {
        CU(cudaMemsetAsync(d_out3, 0, 3 * sizeof(double), ctx->stream));
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->kev[3], ctx->stream));
    ctx->kev_valid = true;
    return PHYFOO_OK;
}

extern "C" int phyfoo_zoops_estep(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, phyfoo_model* bg,
                                  const double* logw, const double* strand_logw, const double* aln_weights,
                                  double* gamma_out, double* profile_out, double* w_out, double* ll_out,
                                  int32_t* argmax_off, uint8_t* argmax_strand, phyfoo_stats* stats) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !fg || !bg || !logw) return fail(ctx, PHYFOO_EINVAL, "zoops_estep: NULL argument");
    CU(cudaSetDevice(ctx->device));
    const int64_t nw_guess = phyfoo_dataset_windows(ds, fg->h.W);
    if (gamma_out) CU(ctx->gamma.ensure((size_t)std::max<int64_t>(nw_guess, 1) * sizeof(double)));
    if (profile_out) CU(ctx->profile.ensure((size_t)std::max<int64_t>(nw_guess, 1) * sizeof(double)));
    if (argmax_off) CU(ctx->argoff.ensure((size_t)std::max(ds->n_aln, 1) * sizeof(int32_t)));
    if (argmax_strand) CU(ctx->argstrand.ensure((size_t)std::max(ds->n_aln, 1)));
    CU(ctx->out3.ensure(3 * sizeof(double)));
    const double* d_alnw = nullptr;
    if (aln_weights) {
        CU(ctx->alnw.ensure((size_t)std::max(ds->n_aln, 1) * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->alnw.p, aln_weights, (size_t)ds->n_aln * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        d_alnw = ctx->alnw.as<double>();
    }
    CallTimer tm(ctx);
    int64_t nw = 0;
    int rc = enqueue_estep(ctx, ds, fg, bg, logw, strand_logw, d_alnw, gamma_out ? ctx->gamma.as<double>() : nullptr,
                           profile_out ? ctx->profile.as<double>() : nullptr, argmax_off ? ctx->argoff.as<int32_t>() : nullptr,
                           argmax_strand ? ctx->argstrand.as<uint8_t>() : nullptr, ctx->out3.as<double>(), &nw);
    if (rc) return rc;
    ctx->gamma_windows = gamma_out ? nw : -1;
    ctx->gamma_ds = gamma_out ? ds : nullptr; ctx->gamma_width = fg->h.W;
    // one synchronisation for the whole call: end-of-kernels event, then every copy (results, sweep counters), then wait
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaGetLastError());
    double r3[3];
    unsigned long long cnt[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(r3, ctx->out3.p, sizeof(r3), cudaMemcpyDeviceToHost, ctx->stream));
    if (stats) CU(cudaMemcpyAsync(cnt, ctx->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
    if (argmax_off) CU(cudaMemcpyAsync(argmax_off, ctx->argoff.p, (size_t)ds->n_aln * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (argmax_strand) CU(cudaMemcpyAsync(argmax_strand, ctx->argstrand.p, (size_t)ds->n_aln, cudaMemcpyDeviceToHost, ctx->stream));
    if (gamma_out) CU(cudaMemcpyAsync(gamma_out, ctx->gamma.p, (size_t)nw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (profile_out) CU(cudaMemcpyAsync(profile_out, ctx->profile.p, (size_t)nw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ll_out) *ll_out = r3[0];
    if (w_out) { w_out[0] = r3[1]; w_out[1] = r3[2]; }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_windows = nw * (strand_logw ? 2 : 1); stats->n_columns = ds->n_cols;
        stats->kernel_launches = ctx->launches; stats->gpu_ms = ms;
        stats->sweeps_fg = (int64_t)cnt[1]; stats->sweeps_bg = (int64_t)cnt[3];      // counters: [2 slot + 1] = sweep sum
    }
    return PHYFOO_OK;
}

extern "C" int phyfoo_zoops_estep_device(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, phyfoo_model* bg,
                                         const double* logw, const double* strand_logw, const double* aln_weights_device,
                                         double* partial_device, void* stream) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !fg || !bg || !logw || !partial_device) return fail(ctx, PHYFOO_EINVAL, "zoops_estep_device: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    // The library's kernels run on its own stream; order them after the caller's stream and make the
    // caller's stream wait for the result, so the call is asynchronous for the host.
    cudaStream_t user = (cudaStream_t)stream;
    if (user) {
        CU(cudaEventRecord(ctx->ev0, user));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev0, 0));
    }
    int64_t nw = 0;
    int rc = enqueue_estep(ctx, ds, fg, bg, logw, strand_logw, aln_weights_device, nullptr, nullptr, nullptr, nullptr, partial_device, &nw);
    if (rc) return rc;
    if (user) {
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamWaitEvent(user, ctx->ev1, 0));
    } else {
        // no caller stream to order against (the legacy default stream does not synchronise with the library's non-blocking
        // stream): finish before returning, so that whatever the caller does next sees the three doubles
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return PHYFOO_OK;
}

// device time of the three phases of the most recent E-step on this context (waits for it to finish):
// background columns (incl. its counter memset), motif windows, ZOOPS posterior + final reduction
extern "C" int phyfoo_estep_kernel_ms(phyfoo_ctx* ctx, float* bg_ms, float* fg_ms, float* zoops_ms) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ctx->kev_valid) return fail(ctx, PHYFOO_EINVAL, "estep_kernel_ms: no E-step has run on this context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->kev[3]));
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, ctx->kev[0], ctx->kev[1]));
    CU(cudaEventElapsedTime(&b, ctx->kev[1], ctx->kev[2]));
    CU(cudaEventElapsedTime(&c, ctx->kev[2], ctx->kev[3]));
    if (bg_ms) *bg_ms = a;
    if (fg_ms) *fg_ms = b;
    if (zoops_ms) *zoops_ms = c;
    return PHYFOO_OK;
}
// sums of mean-field sweeps of the most recent E-step (waits for it to finish)
extern "C" int phyfoo_estep_sweeps(phyfoo_ctx* ctx, int64_t* sweeps_fg, int64_t* sweeps_bg) {
    if (!ctx) return PHYFOO_EINVAL;
    CU(cudaSetDevice(ctx->device));
    int rc;
    int64_t a = 0, b = 0;
    if ((rc = read_sweeps(ctx, 0, &a))) return rc;
    if ((rc = read_sweeps(ctx, 1, &b))) return rc;
    if (sweeps_fg) *sweeps_fg = a;
    if (sweeps_bg) *sweeps_bg = b;
    return PHYFOO_OK;
}

// page-locked host memory for callers that want the copies of a call to run at full PCIe speed
extern "C" void* phyfoo_host_alloc(phyfoo_ctx* ctx, int64_t bytes) {
    if (!ctx || bytes < 0) return nullptr;
    void* p = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaHostAlloc(&p, bytes ? (size_t)bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        ctx->err = "phyfoo_host_alloc: cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
extern "C" void phyfoo_host_free(phyfoo_ctx* ctx, void* p) { (void)ctx; if (p) cudaFreeHost(p); }

extern "C" int phyfoo_last_launches(const phyfoo_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int phyfoo_last_score_path(const phyfoo_ctx* ctx) { return ctx ? ctx->last_path : -1; }
// per-kernel device times of the most recent call on this context (waits for the call's kernels to finish)
extern "C" int phyfoo_last_kernel_times(phyfoo_ctx* ctx, int32_t max_entries, const char** names_out, float* ms_out) {
    if (!ctx) return PHYFOO_EINVAL;
    CU(cudaSetDevice(ctx->device));
    int n = 0;
    for (int i = 0; i < ctx->n_ktimes && n < max_entries; ++i) {
        CU(cudaEventSynchronize(ctx->ktimes[i].e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ktimes[i].e0, ctx->ktimes[i].e1));
        if (names_out) names_out[n] = ctx->ktimes[i].name;
        if (ms_out) ms_out[n] = ms;
        ++n;
    }
    return n;
}
extern "C" int phyfoo_set_option(phyfoo_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return PHYFOO_EINVAL;
    if (std::strcmp(name, "pattern_memo") == 0) {
        if (value < 0 || value > 2) return fail(ctx, PHYFOO_EINVAL, "set_option: pattern_memo must be 0 (off), 1 (auto) or 2 (always)");
        ctx->memo_mode = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "order0_kernel") == 0) {
        if (value < -1 || value > 2) return fail(ctx, PHYFOO_EINVAL, "set_option: order0_kernel must be -1 (choose), 0 (generic), 1 (F81-structured) or 2 (compiled for the tree at run time)");
        ctx->order0_kernel = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "jit_max_threads") == 0) {
        if (value < 32 || value > 1024) return fail(ctx, PHYFOO_EINVAL, "set_option: jit_max_threads must be 32..1024");
        ctx->jit_max_threads = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "l2_persist") == 0) {
        ctx->l2_persist = value != 0;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "order1_qglobal") == 0) {
        if (value < -1 || value > 8) return fail(ctx, PHYFOO_EINVAL, "set_option: order1_qglobal must be -1 (choose), 0 (q rows in shared memory) or 1..8 warps per SM");
        ctx->order1_qglobal = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "bg_generic") == 0) {
        if (value != 0 && value != 1) return fail(ctx, PHYFOO_EINVAL, "set_option: bg_generic must be 0 or 1");
        ctx->bg_generic = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "order1_kernel") == 0) {
        if (value < 0 || value > 2) return fail(ctx, PHYFOO_EINVAL, "set_option: order1_kernel must be 0 (quad per window), 1 (wavefront) or 2 (thread per node)");
        ctx->order1_kernel = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "pattern_memo_sweep_cap") == 0) {
        if (value < 1 || value > 1000) return fail(ctx, PHYFOO_EINVAL, "set_option: pattern_memo_sweep_cap must be in 1..1000");
        ctx->memo_sweep_cap = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "pattern_memo_wave") == 0) {
        if (value != 0 && value != 1) return fail(ctx, PHYFOO_EINVAL, "set_option: pattern_memo_wave must be 0 or 1");
        ctx->memo_wave = (int)value;
        return PHYFOO_OK;
    }
    if (std::strcmp(name, "pattern_memo_split") == 0) {
        if (value < 0 || value > 1000 || value % 4) return fail(ctx, PHYFOO_EINVAL, "set_option: pattern_memo_split must be 0 (one phase) or a multiple of 4");
        ctx->memo_split = (int)value;
        return PHYFOO_OK;
    }
    return fail(ctx, PHYFOO_EINVAL, std::string("set_option: unknown option ") + name);
}
extern "C" void* phyfoo_stream(phyfoo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int phyfoo_synchronize(phyfoo_ctx* ctx) {
    if (!ctx) return PHYFOO_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

#include "mstep.cuh"
#include "scan.cuh"
#include "ab.cuh"

// scores every window of the sample into ctx->fg_scores (device) without copying them out
static int scan_scores(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int strand, const int64_t** d_woff, int64_t* nw) {
    if (strand != 0 && strand != 1) return fail(ctx, PHYFOO_EINVAL, "scan: strand must be 0 or 1");
    const std::vector<int64_t>* h_woff;
    int rc = dataset_win_off(ctx, ds, fg->h.W, d_woff, &h_woff);
    if (rc) return rc;
    *nw = h_woff->back();
    CU(ctx->fg_scores.ensure((size_t)std::max<int64_t>(*nw, 1) * sizeof(double)));
    return launch_score(ctx, ds, fg, *d_woff, *nw, strand, ctx->fg_scores.as<double>(), nullptr, 0);
}

extern "C" int phyfoo_scan_hits(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int32_t strand, double threshold,
                                int32_t* n_hits, int32_t* first_hit, double* max_score, int32_t* arg_max) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !fg) return fail(ctx, PHYFOO_EINVAL, "scan_hits: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int64_t* d_woff; int64_t nw = 0;
    int rc = scan_scores(ctx, ds, fg, strand, &d_woff, &nw);
    if (rc) return rc;
    const size_t na = (size_t)ds->n_aln;
    if (na == 0) return PHYFOO_OK;
    CU(ctx->misc.ensure(na * (3 * sizeof(int32_t) + sizeof(double)) + 64));
    double* d_max = ctx->misc.as<double>();
    int32_t* d_hits = (int32_t*)(d_max + na);
    int32_t* d_first = d_hits + na;
    int32_t* d_arg = d_first + na;
    {
        KernelTimer kt_(ctx, "scan_hits_kernel");
        scan_hits_kernel<<<(unsigned)na, 256, 0, ctx->stream>>>(ctx->fg_scores.as<double>(), d_woff, threshold, d_hits, d_first, d_max, d_arg);
    }
    CU(cudaGetLastError());
    if (n_hits) CU(cudaMemcpyAsync(n_hits, d_hits, na * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (first_hit) CU(cudaMemcpyAsync(first_hit, d_first, na * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (max_score) CU(cudaMemcpyAsync(max_score, d_max, na * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (arg_max) CU(cudaMemcpyAsync(arg_max, d_arg, na * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

extern "C" int phyfoo_score_order_stats(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int32_t strand, int32_t n_ranks,
                                        const int64_t* ranks, double* values_out, int64_t* n_scores_out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !fg || n_ranks < 0 || (n_ranks > 0 && (!ranks || !values_out))) return fail(ctx, PHYFOO_EINVAL, "score_order_stats: NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    const int64_t* d_woff; int64_t nw = 0;
    int rc = scan_scores(ctx, ds, fg, strand, &d_woff, &nw);
    if (rc) return rc;
    if (n_scores_out) *n_scores_out = nw;
    if (n_ranks == 0) return PHYFOO_OK;
    if (nw == 0) return fail(ctx, PHYFOO_EINVAL, "score_order_stats: the sample has no windows");
    if (nw > INT32_MAX) return fail(ctx, PHYFOO_EUNSUPPORTED, "score_order_stats: more than 2^31 scores in one call");
    for (int i = 0; i < n_ranks; ++i)
        if (ranks[i] < 0 || ranks[i] >= nw) return fail(ctx, PHYFOO_EINVAL, "score_order_stats: rank out of range");
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const double*)nullptr, (double*)nullptr, (int)nw, 0, 64, ctx->stream));
    CU(ctx->sel_tmp.ensure(tmp_bytes + 256));
    CU(ctx->selw.ensure((size_t)nw * sizeof(double)));
    CU(cub::DeviceRadixSort::SortKeys(ctx->sel_tmp.p, tmp_bytes, ctx->fg_scores.as<double>(), ctx->selw.as<double>(), (int)nw, 0, 64, ctx->stream));
    ctx->launches += 1;
    CU(ctx->misc.ensure((size_t)n_ranks * (sizeof(int64_t) + sizeof(double))));
    double* d_out = ctx->misc.as<double>();
    int64_t* d_idx = (int64_t*)(d_out + n_ranks);
    CU(cudaMemcpyAsync(d_idx, ranks, (size_t)n_ranks * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    {
        KernelTimer kt_(ctx, "scan_pick_kernel");
        scan_pick_kernel<<<(n_ranks + 127) / 128, 128, 0, ctx->stream>>>(ctx->selw.as<double>(), d_idx, n_ranks, d_out);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(values_out, d_out, (size_t)n_ranks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHYFOO_OK;
}

// ------------------------------------------------------------------------------------------------
// ZOOPS mixture of SEVERAL motif models + "no motif" (BASELINE.json configs[3]; DESIGN.md section "multi-motif mixture").
// The reference's SingleHiddenMotifMixture has exactly one motif component (jstacs SingleHiddenMotifMixture.java:716-718);
// this is its direct generalisation: an alignment holds at most one site of at most one of K motifs,
//   a_{k,j} = -log n_k + motif_k(j) - bg(j, j + W_k - 1),  Z_k = logsumexp_j a_{k,j},
//   help = (logw_1 + Z_1, ..., logw_K + Z_K, logw_0) -> logSumNormalisation -> p_1..p_K, p_0,
//   ll += weight (logPSeq + lse),  gamma_{k,j} = exp(a_{k,j} - Z_k) p_k weight,  w_k += p_k weight,
// with the arithmetic of the one-motif kernel per component (K = 1 reproduces it bit for bit); order-0 background.
// ------------------------------------------------------------------------------------------------
static const int ZOOPS_KMAX = 8;
struct ZoopsMultiParams {
    int K;
    const double* score[ZOOPS_KMAX];         // [n_windows_k] window scores of motif k
    const int64_t* win_off[ZOOPS_KMAX];      // window offsets for its width
    int W[ZOOPS_KMAX];
    double logw[ZOOPS_KMAX + 1];             // motifs 0..K-1, then "no motif"
    const double* bgcol; const int64_t* col_off; int n_aln;
    const double* aln_w;
    double* gamma[ZOOPS_KMAX];               // nullable
    double* part;                            // [K + 2][n_aln]: ll, w_1..w_K, w_0
    int32_t* arg_motif; int32_t* arg_off;    // nullable: first maximum of logw_k + a_{k,j} in (k, j) order
    double* out;                             // [K + 2]
    unsigned int* ticket;
};

__global__ void __launch_bounds__(256) zoops_multi_kernel(ZoopsMultiParams p) {
    __shared__ double sh[32];
    __shared__ double sh_v[8];
    __shared__ int sh_i[8];
    __shared__ double s_Z[ZOOPS_KMAX], s_best[ZOOPS_KMAX];
    __shared__ int s_bestj[ZOOPS_KMAX];
    const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int64_t c0 = p.col_off[a];
    const int L = (int)(p.col_off[a + 1] - c0);
    const double weight = p.aln_w ? p.aln_w[a] : 1.0;
    const double* bgc = p.bgcol + c0;
    double ps = 0.0;
    for (int u = tid; u < L; u += T) ps += bgc[u];
    const double log_pseq = block_sum(ps, sh);                            // bg.getLogProbFor(seq, 0, l-1)
    for (int k = 0; k < p.K; ++k) {
        const int64_t w0 = p.win_off[k][a];
        const int n = (int)(p.win_off[k][a + 1] - w0), W = p.W[k];
        const double lprior = -log((double)n);
        double best = -INFINITY;
        int best_j = 0x7fffffff;
        for (int j = tid; j < n; j += T) {
            double bgw = 0.0;
            for (int m = 0; m < W; ++m) bgw += bgc[j + m];
            const double aj = lprior + p.score[k][w0 + j] - bgw;
            if (p.gamma[k]) p.gamma[k][w0 + j] = aj;
            if (aj > best) { best = aj; best_j = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oj = __shfl_down_sync(0xffffffffu, best_j, o);
            if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
        }
        __syncthreads();
        if ((tid & 31) == 0) { sh_v[tid >> 5] = best; sh_i[tid >> 5] = best_j; }
        __syncthreads();
        best = sh_v[0]; best_j = sh_i[0];
        for (int i = 1; i < (T + 31) / 32; ++i)
            if (sh_v[i] > best || (sh_v[i] == best && sh_i[i] < best_j)) { best = sh_v[i]; best_j = sh_i[i]; }
        if (best_j == 0x7fffffff) best_j = 0;
        double se = 0.0;
        for (int j = tid; j < n; j += T) {
            double aj;
            if (p.gamma[k]) aj = p.gamma[k][w0 + j];
            else {
                double bgw = 0.0;
                for (int m = 0; m < W; ++m) bgw += bgc[j + m];
                aj = lprior + p.score[k][w0 + j] - bgw;
            }
            const double e = exp(aj - best);
            if (p.gamma[k]) p.gamma[k][w0 + j] = e;
            se += e;
        }
        const double sum = block_sum(se, sh);
        if (tid == 0) { s_Z[k] = best + log(sum); s_best[k] = best; s_bestj[k] = best_j; sh[31] = sum; }
        __syncthreads();
        const double sumk = sh[31];
        if (p.gamma[k] && sumk != 1.0)
            for (int j = tid; j < n; j += T) p.gamma[k][w0 + j] /= sumk;  // normalised over j; scaled by p_k below
        __syncthreads();
    }
    // component posterior: logSumNormalisation over (logw_k + Z_k, logw_0)
    double help[ZOOPS_KMAX + 1];
    double off = p.logw[p.K];
    for (int k = 0; k < p.K; ++k) { help[k] = p.logw[k] + s_Z[k]; off = fmax(off, help[k]); }
    help[p.K] = p.logw[p.K];
    double hs = 0.0;
    for (int k = 0; k <= p.K; ++k) { help[k] = exp(help[k] - off); hs += help[k]; }
    if (hs != 1.0) for (int k = 0; k <= p.K; ++k) help[k] /= hs;
    const double lse = off + log(hs);
    if (tid == 0) {
        p.part[a] = weight * (log_pseq + lse);
        for (int k = 0; k <= p.K; ++k) p.part[(size_t)(1 + k) * p.n_aln + a] = weight * help[k];
        if (p.arg_off) {
            double bv = -INFINITY; int bk = 0;
            for (int k = 0; k < p.K; ++k) { const double v = p.logw[k] + s_best[k]; if (v > bv) { bv = v; bk = k; } }
            p.arg_motif[a] = bk; p.arg_off[a] = s_bestj[bk];
        }
    }
    for (int k = 0; k < p.K; ++k) {
        if (!p.gamma[k]) continue;
        const int64_t w0 = p.win_off[k][a];
        const int n = (int)(p.win_off[k][a + 1] - w0);
        const double pk = help[k] * weight;
        for (int j = tid; j < n; j += T) p.gamma[k][w0 + j] *= pk;
    }
    __shared__ int s_last;
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int r = 0; r < p.K + 2; ++r) {
            double v = 0.0;
            for (int i = tid; i < p.n_aln; i += T) v += __ldcg(p.part + (size_t)r * p.n_aln + i);
            const double s3 = block_sum(v, sh);
            if (tid == 0) p.out[r] = s3;
            __syncthreads();
        }
        if (tid == 0) *p.ticket = 0u;
    }
}

// One E-step of the K-motif mixture.  Window scores of motif k stay resident in ctx->multi_scores[k] / gamma in
// ctx->multi_gamma[k] (phyfoo_select_multi reads them for the M-step of component k).
extern "C" int phyfoo_zoops_estep_multi(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t n_motifs, phyfoo_model* const* fgs, phyfoo_model* bg,
                                        const double* logw, const double* aln_weights, double* const* gamma_out, double* w_out, double* ll_out,
                                        int32_t* argmax_motif, int32_t* argmax_off, phyfoo_stats* stats) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !fgs || !bg || !logw || n_motifs < 1 || n_motifs > ZOOPS_KMAX) return fail(ctx, PHYFOO_EINVAL, "zoops_estep_multi: NULL argument or more than 8 motifs");
    if (bg->h.kind != PHYFOO_MODEL_BG || bg->h.order != 0) return fail(ctx, PHYFOO_EUNSUPPORTED, "zoops_estep_multi: the background must be a PhyloBackground of order 0");
    if (ds->n_aln == 0) return fail(ctx, PHYFOO_EINVAL, "zoops_estep_multi: empty sample");
    CU(cudaSetDevice(ctx->device));
    CallTimer tm(ctx);
    ZoopsMultiParams z;
    z.K = n_motifs;
    int64_t nw[ZOOPS_KMAX];
    int64_t sweeps_fg = 0, n_windows = 0;
    const double* d_alnw = nullptr;
    if (aln_weights) {
        CU(ctx->alnw.ensure((size_t)ds->n_aln * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->alnw.p, aln_weights, (size_t)ds->n_aln * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        d_alnw = ctx->alnw.as<double>();
    }
    CU(ctx->bg_scores.ensure((size_t)std::max<int64_t>(ds->n_cols, 1) * sizeof(double)));
    int rc = launch_bg_columns(ctx, ds, bg, ctx->bg_scores.as<double>(), nullptr, 1);
    if (rc) return rc;
    for (int k = 0; k < n_motifs; ++k) {
        phyfoo_model* fg = fgs[k];
        if (!fg || fg->h.kind != PHYFOO_MODEL_FG) return fail(ctx, PHYFOO_EINVAL, "zoops_estep_multi: a motif model is missing or not a motif");
        if (ds->min_len < fg->h.W) return fail(ctx, PHYFOO_EINVAL, "zoops_estep_multi: an alignment is shorter than a motif");
        const int64_t* d_woff; const std::vector<int64_t>* h_woff;
        if ((rc = dataset_win_off(ctx, ds, fg->h.W, &d_woff, &h_woff))) return rc;
        nw[k] = h_woff->back();
        n_windows += nw[k];
        CU(ctx->multi_scores[k].ensure((size_t)nw[k] * sizeof(double)));
        CU(ctx->multi_gamma[k].ensure((size_t)nw[k] * sizeof(double)));
        if ((rc = launch_score(ctx, ds, fg, d_woff, nw[k], 0, ctx->multi_scores[k].as<double>(), nullptr, 0))) return rc;
        int64_t sk = 0;
        if (stats && (rc = read_sweeps(ctx, 0, &sk))) return rc;
        sweeps_fg += sk;
        z.score[k] = ctx->multi_scores[k].as<double>(); z.win_off[k] = d_woff; z.W[k] = fg->h.W; z.logw[k] = logw[k];
        z.gamma[k] = ctx->multi_gamma[k].as<double>();
        ctx->multi_windows[k] = nw[k]; ctx->multi_width[k] = fg->h.W;
    }
    ctx->multi_ds = ds; ctx->multi_k = n_motifs;
    z.logw[n_motifs] = logw[n_motifs];
    z.bgcol = ctx->bg_scores.as<double>(); z.col_off = ds->d_col_off; z.n_aln = ds->n_aln; z.aln_w = d_alnw;
    CU(ctx->part.ensure((size_t)(n_motifs + 2) * ds->n_aln * sizeof(double)));
    CU(ctx->out3.ensure((ZOOPS_KMAX + 2) * sizeof(double)));
    CU(ctx->argoff.ensure((size_t)2 * ds->n_aln * sizeof(int32_t)));
    z.part = ctx->part.as<double>(); z.out = ctx->out3.as<double>();
    z.arg_off = argmax_off ? ctx->argoff.as<int32_t>() : nullptr;
    z.arg_motif = argmax_off ? ctx->argoff.as<int32_t>() + ds->n_aln : nullptr;
    z.ticket = (unsigned int*)(ctx->counters.as<unsigned long long>() + 12);
    {
        KernelTimer kt_(ctx, "zoops_multi_kernel");
        zoops_multi_kernel<<<ds->n_aln, 256, 0, ctx->stream>>>(z);
    }
    CU(cudaGetLastError());
    const float ms = tm.stop();
    double r[ZOOPS_KMAX + 2];
    CU(cudaMemcpyAsync(r, ctx->out3.p, (size_t)(n_motifs + 2) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (argmax_off) CU(cudaMemcpyAsync(argmax_off, z.arg_off, (size_t)ds->n_aln * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (argmax_off && argmax_motif) CU(cudaMemcpyAsync(argmax_motif, z.arg_motif, (size_t)ds->n_aln * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (gamma_out)
        for (int k = 0; k < n_motifs; ++k)
            if (gamma_out[k]) CU(cudaMemcpyAsync(gamma_out[k], z.gamma[k], (size_t)nw[k] * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ll_out) *ll_out = r[0];
    if (w_out) for (int k = 0; k <= n_motifs; ++k) w_out[k] = r[1 + k];
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_windows = n_windows; stats->n_columns = ds->n_cols; stats->kernel_launches = ctx->launches; stats->gpu_ms = ms;
        stats->sweeps_fg = sweeps_fg;
    }
    return PHYFOO_OK;
}
// io/SequenceSpecificDataSelector.java:37-99 on the gamma of component k of the last phyfoo_zoops_estep_multi (still resident)
extern "C" int phyfoo_select_multi(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t motif, double t, double q, int32_t* sel_aln,
                                   int32_t* sel_off, double* sel_weight, int64_t capacity, int64_t* n_selected) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || motif < 0 || motif >= ctx->multi_k || ctx->multi_ds != ds) return fail(ctx, PHYFOO_EINVAL, "select_multi: no gamma of this data set and component is resident (run phyfoo_zoops_estep_multi first)");
    // hand the component's gamma to phyfoo_select as the resident one
    std::swap(ctx->gamma, ctx->multi_gamma[motif]);
    const int64_t kw = ctx->gamma_windows; const phyfoo_dataset* kd = ctx->gamma_ds; const int kwd = ctx->gamma_width;
    ctx->gamma_windows = ctx->multi_windows[motif]; ctx->gamma_ds = ds; ctx->gamma_width = ctx->multi_width[motif];
    const int rc = phyfoo_select(ctx, ds, ctx->multi_width[motif], nullptr, t, q, sel_aln, sel_off, sel_weight, capacity, n_selected);
    std::swap(ctx->gamma, ctx->multi_gamma[motif]);
    ctx->gamma_windows = kw; ctx->gamma_ds = kd; ctx->gamma_width = kwd;
    return rc;
}

// ---- the run-time compiled order-0 kernel (jit_o0.hpp): inspection without a device ----
// Generates the kernel of this model's tree (gaps: with the code for unobserved symbols) and compiles it for sm_100a.
// source_out / log_out (nullable, NUL-terminated, truncated to their capacities) receive the generated source and the
// compiler's log; *cubin_bytes the size of the image.  PHYFOO_EUNSUPPORTED: no NVRTC library, order != 0, or no block shape.
extern "C" int phyfoo_jit_check(const phyfoo_model_desc* desc, int32_t gaps, char* source_out, int64_t source_cap, char* log_out, int64_t log_cap,
                                int64_t* cubin_bytes) {
    if (!desc) return PHYFOO_EINVAL;
    auto put = [](char* dst, int64_t cap, const std::string& text) {
        if (!dst || cap <= 0) return;
        const size_t n = std::min<size_t>(text.size(), (size_t)cap - 1);
        memcpy(dst, text.data(), n);
        dst[n] = 0;
    };
    if (cubin_bytes) *cubin_bytes = 0;
    try {
        ModelHost h;
        h.init(*desc);
        if (h.order != 0) return PHYFOO_EUNSUPPORTED;
        JitShape sh;
        if (!jit_o0_shape(h, gaps != 0, 227 * 1024, 512, &sh)) { put(log_out, log_cap, "no block shape fits"); return PHYFOO_EUNSUPPORTED; }
        const std::string src = jit_o0_source(h, sh, gaps != 0, (h.S + 15) / 16);
        put(source_out, source_cap, src);
        std::string log;
        std::vector<char> cubin;
        if (!jit_api().ok) { put(log_out, log_cap, jit_api().why); return PHYFOO_EUNSUPPORTED; }
        if (jit_compile(src, log, &cubin)) { put(log_out, log_cap, log); return PHYFOO_ECUDA; }
        put(log_out, log_cap, log);
        if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    } catch (const std::exception& e) { put(log_out, log_cap, e.what()); return PHYFOO_EINVAL; }
    return PHYFOO_OK;
}
// why the last request for a run-time compiled kernel fell back to score_o0f_kernel ("" when it did not)
extern "C" const char* phyfoo_jit_note(const phyfoo_ctx* ctx) { return ctx ? ctx->jit_note.c_str() : ""; }

// ------------------------------------------------------------------------------------------------
// sufficient-statistics objective on the host inside the library (suffstat_host.hpp)
// ------------------------------------------------------------------------------------------------
#include "suffstat_host.hpp"

struct phyfoo_ssobj { SuffStatObjective o; };

extern "C" int phyfoo_ss_create(phyfoo_ctx* ctx, phyfoo_feset* fe, phyfoo_ssobj** out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!fe || !out) return fail(ctx, PHYFOO_EINVAL, "ss_create: NULL argument");
    *out = nullptr;
    const ModelHost& h = fe->m->h;
    std::vector<double> c((size_t)h.W * (h.order == 1 ? h.E1() : h.E));
    double ent = 0.0;
    const int rc = phyfoo_fe_suffstats(ctx, fe, c.data(), &ent);
    if (rc) return rc;
    phyfoo_ssobj* ss = new phyfoo_ssobj();
    ss->o.init(h, c.data(), ent);
    *out = ss;
    return PHYFOO_OK;
}
// from given counts (e.g. counts a caller reduced itself); no device involved
// The M-step objective of a BACKGROUND of order k >= 1 (models/PhyloBackground.java:111-176): every (k+1)-column window of every
// alignment is a mean field over the whole net; with q fixed, f(theta) = - sum_p w_p F_p(theta).  One pass of the generic net
// kernel (netg.cuh) fixes q and collects the expected counts of the net's table entries; everything after that is host work on
// the returned object (phyfoo_ss_eval / _grad / _eval_branches: position t has 4^min(t, k) x 4 parameters).
extern "C" int phyfoo_ss_create_bg(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, const double* aln_weights, phyfoo_ssobj** out) {
    if (!ctx) return PHYFOO_EINVAL;
    if (!ds || !bg || !out) return fail(ctx, PHYFOO_EINVAL, "ss_create_bg: NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    ctx->launches = 0; ctx->n_ktimes = 0;
    ModelHost& h = bg->h;
    const int k = h.order, T = k + 1;
    if (h.kind != PHYFOO_MODEL_BG || k < 1) return fail(ctx, PHYFOO_EINVAL, "ss_create_bg: needs a background model of order >= 1 (order 0: phyfoo_fe_prepare_all + phyfoo_ss_create)");
    if (T > phyfoo::NETG_MAX_TREES) return fail(ctx, PHYFOO_EUNSUPPORTED, "background order above 7");
    if (h.evol == PHYFOO_EVOL_HKY_AS_IMPLEMENTED) return fail(ctx, PHYFOO_EUNSUPPORTED, "HKY as implemented: the M-step objective is not finite for any parameter set (HKY.java:32)");
    if (h.S != ds->S) return fail(ctx, PHYFOO_EINVAL, "model and dataset disagree on the number of species");
    if (ds->n_aln > 0 && ds->min_len < T) return fail(ctx, PHYFOO_EINVAL, "background of order k: an alignment is shorter than k + 1 columns");
    int rc = netg_upload(ctx, bg);
    if (rc) return rc;
    const int64_t* d_woff; const std::vector<int64_t>* h_woff;
    rc = dataset_win_off(ctx, ds, T, &d_woff, &h_woff);
    if (rc) return rc;
    const int64_t n = h_woff->back();
    const size_t n_entries = bg->netg_tab_used;
    std::vector<double> c(n_entries + 1, 0.0);
    if (n > 0) {
        CU(ctx->misc.ensure((size_t)T * n * sizeof(int64_t)));
        CU(ctx->bg_scores.ensure((size_t)(2 * n + n_entries + 1) * sizeof(double)));
        int64_t* d_cols = ctx->misc.as<int64_t>();
        double* d_ofirst = ctx->bg_scores.as<double>();
        double* d_counts = d_ofirst + 2 * n;
        CU(cudaMemsetAsync(d_counts, 0, (n_entries + 1) * sizeof(double), ctx->stream));
        const double* d_alnw = nullptr;
        if (aln_weights) {
            CU(ctx->alnw.ensure((size_t)ds->n_aln * sizeof(double)));
            CU(cudaMemcpyAsync(ctx->alnw.p, aln_weights, (size_t)ds->n_aln * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            d_alnw = ctx->alnw.as<double>();
        }
        bgk_items_full_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_woff, ds->d_col_off, ds->n_aln, n, T, d_cols);
        CU(cudaGetLastError());
        const int G = h.W * h.N;
        const size_t per_thread = (size_t)G * 4 * sizeof(double);
        long long blocks = std::min<long long>((n + 127) / 128, 8LL * ctx->sm_count);
        blocks = std::max<long long>(1, std::min<long long>(blocks, (long long)((512ull << 20) / (per_thread * 128))));
        CU(ctx->gstate.ensure(per_thread * (size_t)blocks * 128));
        NetgCountParams p;
        p.s.bases = ds->d_bases; p.s.gaps = ds->d_gaps; p.s.n_cols = ds->n_cols; p.s.nb = ds->nb;
        p.s.view.n_nodes = G; p.s.view.n_trees = h.W; p.s.view.nodes_per_tree = h.N;
        p.s.view.rec = bg->d_netg_rec; p.s.view.lists = bg->d_netg_lists; p.s.view.tab = bg->d_netg_tab;
        p.s.item_cols = d_cols; p.s.item_len = nullptr; p.s.n_items = n;
        p.s.q = ctx->gstate.as<double>(); p.s.q_stride = (size_t)blocks * 128;
        p.s.out_first = d_ofirst; p.s.out_last = d_ofirst + n; p.s.sweeps = nullptr; p.s.sweep_sum = nullptr;
        p.s.max_sweeps = h.max_sweeps; p.s.min_sweeps = h.min_sweeps; p.s.tol = h.tol;
        p.aln_w = d_alnw; p.col_off = ds->d_col_off; p.n_aln = ds->n_aln;
        p.counts = d_counts; p.entropy = d_counts + n_entries;
        {
            KernelTimer kt_(ctx, "netg_count_kernel");
            netg_count_kernel<<<(unsigned)blocks, 128, 0, ctx->stream>>>(p);
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c.data(), d_counts, (n_entries + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    try {
        phyfoo_ssobj* ss = new phyfoo_ssobj();
        ss->o.init_generic(h, c.data(), n_entries, c[n_entries]);
        *out = ss;
    } catch (const std::exception& e) { return fail(ctx, PHYFOO_EINVAL, std::string("ss_create_bg: ") + e.what()); }
    return PHYFOO_OK;
}

extern "C" int phyfoo_ss_create_host(const phyfoo_model_desc* desc, const double* counts, double entropy, phyfoo_ssobj** out) {
    if (!desc || !counts || !out) return PHYFOO_EINVAL;
    *out = nullptr;
    try {
        ModelHost h;
        h.init(*desc);
        if (h.order > 1) return PHYFOO_EUNSUPPORTED;
        phyfoo_ssobj* ss = new phyfoo_ssobj();
        ss->o.init(h, counts, entropy);
        *out = ss;
    } catch (const std::exception&) { return PHYFOO_EINVAL; }
    return PHYFOO_OK;
}
extern "C" void phyfoo_ss_free(phyfoo_ssobj* ss) { delete ss; }
extern "C" double phyfoo_ss_value(const phyfoo_ssobj* ss) { return ss ? ss->o.value() : NAN; }
extern "C" int64_t phyfoo_ss_evaluations(const phyfoo_ssobj* ss) { return ss ? ss->o.evaluations : 0; }
extern "C" int phyfoo_ss_eval(phyfoo_ssobj* ss, int32_t position, const double* lambda, int32_t dim, double* f_out) {
    if (!ss || !lambda || !f_out || position < 0 || position >= ss->o.h.W || dim != ss->o.dim(position)) return PHYFOO_EINVAL;
    *f_out = ss->o.eval_lambda(position, lambda);
    return PHYFOO_OK;
}
extern "C" int phyfoo_ss_grad(phyfoo_ssobj* ss, int32_t position, const double* lambda, int32_t dim, double eps, double* f_out, double* g_out) {
    if (!ss || !lambda || !g_out || position < 0 || position >= ss->o.h.W || dim != ss->o.dim(position) || !(eps > 0)) return PHYFOO_EINVAL;
    std::vector<double> xv(lambda, lambda + dim);
    double* x = xv.data();
    const double f0 = ss->o.eval_lambda(position, x);
    for (int i = 0; i < dim; ++i) {                     // jstacs NumericalDifferentiableFunction.java:64-84
        const double keep = x[i];
        x[i] += eps;
        g_out[i] = (ss->o.eval_lambda(position, x) - f0) / eps;
        x[i] = keep;
    }
    ss->o.eval_lambda(position, x);
    if (f_out) *f_out = f0;
    return PHYFOO_OK;
}
extern "C" int phyfoo_ss_eval_branches(phyfoo_ssobj* ss, int32_t position, const double* branch, double* f_out) {
    if (!ss || !branch || !f_out || position >= ss->o.h.W) return PHYFOO_EINVAL;
    *f_out = ss->o.eval_branches(position, branch);
    return PHYFOO_OK;
}
extern "C" int phyfoo_ss_get_pi(const phyfoo_ssobj* ss, int32_t position, double* pi_out, int32_t n) {
    if (!ss || !pi_out || position < 0 || position >= ss->o.h.W || n != ss->o.dim(position)) return PHYFOO_EINVAL;
    for (int i = 0; i < n; ++i) pi_out[i] = ss->o.h.pi[position][i];
    return PHYFOO_OK;
}
// the M-step of PhyloBayesModel.train on this objective: the listed positions one after the other (the caller draws the
// order, util/PWMUtil.java:131-144), each from its current parameters, termination TC_MOTIF_SEPARATED_POSITIONS
extern "C" int phyfoo_ss_optimize(phyfoo_ssobj* ss, const int32_t* order, int32_t n_pos, double* f_before, double* f_after, int64_t* evaluations) {
    if (!ss || !order || n_pos < 0) return PHYFOO_EINVAL;
    const int W = ss->o.h.W;
    std::vector<char> seen(W, 0);
    bool distinct = true;
    for (int j = 0; j < n_pos; ++j) {
        if (order[j] < 0 || order[j] >= W) return PHYFOO_EINVAL;
        if (seen[order[j]]) distinct = false;
        seen[order[j]] = 1;
    }
    if (f_before) *f_before = ss->o.value();
    const long long e0 = ss->o.evaluations;
    auto one = [](SuffStatObjective& o, int t) {
        const int d = o.dim(t);
        std::vector<double> x(d);
        for (int i = 0; i < d; ++i) x[i] = std::log(o.h.pi[t][i]);
        // (an exception of the optimizer is caught and the run goes on from the last iterate, as in
        // optimizing/FE_byParameter_ForBayesNodeJstacs.java:61-83)
        try { bfgs_minimise([&](const double* z) { return -o.eval_lambda(t, z); }, x, tc_motif_separated_positions()); } catch (const std::exception&) {}
        o.eval_lambda(t, x.data());                               // initParameters(lambda), :78
    };
    // With q fixed the objective is separable, f = sum_t G_t(theta_t) - ENT: the optimisation of a position neither reads nor
    // changes the parameters of another one, only the constant the other terms add to f.  Distinct positions therefore run side by
    // side on host threads, each against the other terms as they were on entry (so the result does not depend on the number of
    // threads); PHYFOO_SS_THREADS=1 runs them one after the other, each seeing the terms the earlier ones left.
    int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char* e = std::getenv("PHYFOO_SS_THREADS")) nt = std::max(1, std::atoi(e));
    nt = std::min(nt, (int)n_pos);
    if (nt > 1 && distinct) {
        const std::vector<double> G0 = ss->o.G;
        std::vector<SuffStatObjective> copy(nt, ss->o);
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k)
            th.emplace_back([&, k]() {
                for (int j = k; j < n_pos; j += nt) {
                    copy[k].G = G0;                      // the other terms as they were on entry
                    one(copy[k], order[j]);
                }
            });
        for (auto& x : th) x.join();
        for (int k = 0; k < nt; ++k) {
            ss->o.evaluations += copy[k].evaluations - e0;
            for (int j = k; j < n_pos; j += nt) {
                const int t = order[j];
                ss->o.h.pi[t] = copy[k].h.pi[t];
                ss->o.G[t] = ss->o.term(t, ss->o.h.pi[t].data());
            }
        }
    } else {
        for (int j = 0; j < n_pos; ++j) one(ss->o, order[j]);
    }
    if (f_after) *f_after = ss->o.value();
    if (evaluations) *evaluations = ss->o.evaluations - e0;
    return PHYFOO_OK;
}

#include "multi.cuh"
