/* phyfoo.h -- C ABI of libphyfoo.so: the B200-native drop-in for PhyFoo's site-scoring /
 * free-energy hot path.  Plain pointers and sizes only; every pointer is caller-owned HOST memory
 * unless the name says "device" or it is an opaque handle.  A Java subclass of PhyFoo's models
 * binds these entry points through JNI or Panama (INTEGRATION.md shows the stubs); the Python
 * package phyfoo_b200 binds the identical symbols through ctypes.
 *
 * Reference interfaces replaced (X.java:N = /root/reference/main/java/src/X.java,
 * jstacs!/p:N = source inside /root/reference/libs/jstacs.jar):
 *   Model.getLogProbFor(Sequence,int,int)            jstacs!/de/jstacs/models/Model.java:223
 *     impl  models/PhyloBayesModel.java:222-267      -> phyfoo_score_windows   (whole sample per call)
 *     impl  models/PhyloBackground.java:241-305      -> phyfoo_bg_columns
 *   SingleHiddenMotifMixture.getNewWeights / modify / getProfileOfScoresFor / getStrandProbabilitiesFor
 *     jstacs!/de/jstacs/models/mixture/motif/SingleHiddenMotifMixture.java:499-566,585-596,643-688,737-754
 *     StrandModel  jstacs!/de/jstacs/models/mixture/StrandModel.java:561-583,631-640
 *                                                    -> phyfoo_zoops_estep
 *   SequenceSpecificDataSelector                     io/SequenceSpecificDataSelector.java:37-99
 *                                                    -> phyfoo_select
 *   AbstractFreeEnergyOptimizer.getWeightedLogLikelihoodSum  optimizing/AbstractFreeEnergyOptimizer.java:95-114
 *   FE_byParameter_ForBayesNodeJstacs.evaluateFunction       optimizing/FE_byParameter_ForBayesNodeJstacs.java:43-53
 *   MotifParamOptimizer.evaluateFunction                      optimizing/MotifParamOptimizer.java:46-61
 *   NumericalDifferentiableFunction.evaluateGradientOfFunction
 *     jstacs!/de/jstacs/algorithms/optimization/NumericalDifferentiableFunction.java:64-84
 *                                                    -> phyfoo_fe_prepare / phyfoo_fe_eval / phyfoo_fe_grad
 *   MultiDimensionalDiscreteSequence ctor (column tensor int[pos][species])
 *     jstacs!/de/jstacs/data/sequences/MultiDimensionalDiscreteSequence.java:46-52
 *                                                    -> phyfoo_dataset_create (packs on the device)
 *
 * Conventions: every function returns 0 on success or a negative PHYFOO_E* code; the message is
 * available from phyfoo_last_error().  One context belongs to one GPU and to one caller thread at a
 * time (the reference is single-threaded; this library does not lock); several GPUs behind one caller:
 * the phyfoo_multi_* entry points at the end of this file.  Calls are synchronous:
 * results are in the caller's buffers when the call returns.  There is NO CPU fallback: without a
 * CUDA device phyfoo_init fails with PHYFOO_ENODEVICE.
 */
#ifndef PHYFOO_H
#define PHYFOO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYFOO_OK 0
#define PHYFOO_EINVAL (-1)     /* IllegalArgumentException on the Java side */
#define PHYFOO_ENODEVICE (-2)  /* no CUDA device / driver */
#define PHYFOO_ECUDA (-3)      /* a CUDA call failed */
#define PHYFOO_ENOMEM (-4)
#define PHYFOO_EUNSUPPORTED (-5)

#define PHYFOO_MODEL_FG 0 /* models/PhyloBayesModel.java (motif, `width` trees) */
#define PHYFOO_MODEL_BG 1 /* models/PhyloBackground.java (order+1 trees)        */

#define PHYFOO_EVOL_F81ALPHA 0           /* evolution/FS81alpha.java (default, EvolModel.java:7) */
#define PHYFOO_EVOL_F81BETA 1            /* evolution/FS81beta.java */
#define PHYFOO_EVOL_HKY_AS_IMPLEMENTED 2 /* evolution/HKY.java with its beta == 0 (ctor bug :32) */
#define PHYFOO_EVOL_HKY_RATIO 3          /* opt-in: evolution/HKY.java:77-106 with the ratio in effect, beta = alpha * hky_ratio -- the
                                          * model the reference's bundled data/synthetic_data/HKY files were drawn from
                                          * (tests/test_generator_pins.py); order-0 models only (general 4x4 tables) */

#define PHYFOO_MODE_MEANFIELD 0 /* -calcFreeEnergy() after optimizeByNormalisation (default path) */
#define PHYFOO_MODE_EXACT 1     /* PhyloBayesModel.CALC_LOGLIKELIHOOD = true (order 0: pruning)     */

typedef struct phyfoo_ctx phyfoo_ctx;
typedef struct phyfoo_dataset phyfoo_dataset;
typedef struct phyfoo_model phyfoo_model;
typedef struct phyfoo_feset phyfoo_feset;

/* Tree and parameters of one model.  Nodes are numbered in DFS pre-order exactly as
 * io/NewickToBayesNet.java:106-147 numbers them (root = 0, parent[k] < k); leaf_species[k] is the
 * alignment row observed at leaf k (left-to-right leaf order of the Newick string = species order,
 * util/Util.java:39-53) and -1 at internal nodes. */
typedef struct {
    int32_t kind;                   /* PHYFOO_MODEL_FG / PHYFOO_MODEL_BG */
    int32_t width;                  /* motif length W (ignored for BG: order+1 trees) */
    int32_t order;                  /* FG: 0 or 1 (PhyloBayesModel.java:363); BG: Markov order */
    int32_t n_nodes;                /* N */
    const int32_t* parent;          /* [N] */
    const int32_t* leaf_species;    /* [N] */
    int32_t evol;                   /* PHYFOO_EVOL_* */
    double hky_ratio;               /* PHYFOO_EVOL_HKY_AS_IMPLEMENTED: accepted and ignored, like the reference does (HKY.java:32);
                                     * PHYFOO_EVOL_HKY_RATIO: beta / alpha, > 0 */
    const double* branch;           /* [trees][N] distance to parent per position (root entry unused) */
    const double* pi;               /* per tree t: [ctx_t][4] row-major, concatenated; ctx_t = 4^{#horizontal parents} */
    int32_t mode;                   /* PHYFOO_MODE_* */
    int32_t allow_independent_leaf; /* CPF.ALLOW_RETURN_INDEPENDENT_TRANSITION (CPF.java:14); must be 1 */
    int32_t max_sweeps;             /* 100  (MeanFieldForBayesNet.java:102); 0 = default */
    int32_t min_sweeps;             /* 2 */
    double tol;                     /* 1e-5 (:205); 0 = default */
} phyfoo_model_desc;

typedef struct {
    int64_t n_windows;      /* window scores produced (per strand) */
    int64_t n_columns;      /* background columns scored */
    int64_t sweeps_fg;      /* sum over windows of the mean-field sweep count K_w (integer parity check) */
    int64_t sweeps_bg;      /* sum over columns of K_c */
    int32_t kernel_launches;/* kernels launched by the call */
    float gpu_ms;           /* device time of the call's kernels (CUDA events on the library's stream) */
} phyfoo_stats;

/* ---- context ---- */
int phyfoo_init(int device, phyfoo_ctx** out);
void phyfoo_destroy(phyfoo_ctx* ctx);
const char* phyfoo_last_error(const phyfoo_ctx* ctx); /* ctx may be NULL: error of a failed phyfoo_init */
int phyfoo_device_info(phyfoo_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);
/* fp64 FMA roofline probe: runs a dependent-chain-free DFMA loop on every SM; returns TFLOP/s */
int phyfoo_fp64_peak(phyfoo_ctx* ctx, double* tflops_out);

/* ---- dataset: packed columns resident on the device ----
 * codes[sum cols][n_species], values 0..3 = A,C,G,T and 4 = gap/unobserved
 * (jstacs!/de/jstacs/data/alphabets/UnobservableDNAAlphabet.java:116); col_offsets[n_aln+1] are the
 * prefix sums of the alignment lengths.  Packed layout ("packed columns", bit-exact contract):
 *   base plane k, column c: uint32 at bases[k*n_cols + c], species 16k..16k+15, 2 bits each, species s
 *   in bits [2(s%16), 2(s%16)+1]; a gap stores base bits 00;
 *   gap plane k, column c: uint32 at gaps[k*n_cols + c], species 32k..32k+31, bit s%32 set = unobserved. */
int phyfoo_dataset_create(phyfoo_ctx* ctx, int32_t n_aln, int32_t n_species, const int64_t* col_offsets,
                          const uint8_t* codes, phyfoo_dataset** out);
/* same, but `codes` already lives on this context's device (no host->device copy) */
int phyfoo_dataset_create_device(phyfoo_ctx* ctx, int32_t n_aln, int32_t n_species, const int64_t* col_offsets,
                                 const uint8_t* codes_device, phyfoo_dataset** out);
int phyfoo_dataset_packed(phyfoo_ctx* ctx, const phyfoo_dataset* ds, uint32_t* bases_out, uint32_t* gaps_out);
int64_t phyfoo_dataset_columns(const phyfoo_dataset* ds);
int64_t phyfoo_dataset_windows(const phyfoo_dataset* ds, int32_t width); /* sum_i max(0, L_i - width + 1) */
/* distinct column patterns (forward and complemented) of the data set, or -1 when the pattern table does not apply (S > 6) */
int64_t phyfoo_dataset_patterns(phyfoo_ctx* ctx, const phyfoo_dataset* ds);
void phyfoo_dataset_free(phyfoo_ctx* ctx, phyfoo_dataset* ds);

/* ---- models ---- */
int phyfoo_model_create(phyfoo_ctx* ctx, const phyfoo_model_desc* desc, phyfoo_model** out);
/* PhyloPreparedAbstractModel.setCondProb(double[], position): models/AbstractPhyloModel.java:131-134 */
int phyfoo_model_set_pi(phyfoo_ctx* ctx, phyfoo_model* m, int32_t position, const double* pi, int32_t n);
int phyfoo_model_get_pi(phyfoo_ctx* ctx, const phyfoo_model* m, int32_t position, double* pi_out, int32_t n);
/* VirtualTree.reinitEdglengths (bayesNet/VirtualTree.java:237-255), one branch at a time */
int phyfoo_model_set_branch(phyfoo_ctx* ctx, phyfoo_model* m, int32_t position, int32_t node, double length);
int phyfoo_model_set_mode(phyfoo_ctx* ctx, phyfoo_model* m, int32_t mode);
/* 1 for PHYFOO_EVOL_HKY_AS_IMPLEMENTED: every substitution table has zero entries (the reference never stores the
 * transition/transversion ratio, evolution/HKY.java:32), so every mean-field score is degenerate.  The library returns what the
 * reference returns -- NaN, or -inf for purine-only / pyrimidine-only columns under a star tree, with max_sweeps sweeps per
 * window -- instead of an error (order 0; scoring and E-step; the M-step objective is refused).  Exact mode computes the
 * likelihood under these tables as usual (log 0 = -inf where a transversion would be needed). */
int phyfoo_model_is_degenerate(const phyfoo_model* m);
void phyfoo_model_free(phyfoo_ctx* ctx, phyfoo_model* m);

/* ---- scoring ----
 * out[w] for every window w of every alignment (alignment-major, offsets ascending; total =
 * phyfoo_dataset_windows(ds, W)) = PhyloBayesModel.getLogProbFor(seq, j, j+W-1).  strand = 1 scores the
 * reverse complement of every window (StrandModel.java:636-639).  sweeps_out (nullable) receives K_w. */
int phyfoo_score_windows(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int32_t strand,
                         double* out, int32_t* sweeps_out, phyfoo_stats* stats);
/* out[c] for every column c = PhyloBackground.getLogProbFor(seq, u, u) (order 0). sweeps_out nullable. */
int phyfoo_bg_columns(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* out,
                      int32_t* sweeps_out, phyfoo_stats* stats);

/* Order-1 background (models/PhyloBackground.java:241-305).  For a range [s, e] with e > s the reference returns
 *   -F(0,0) of the window (s, s+1)  +  sum_{u = s+1..e} -F(1,1) of the window (u-1, u)
 * (each window's mean field optimised on its own).  cond_out[c] = the conditional term of column c (0 at the first column
 * of an alignment), first_out[c] (nullable) = the start term of a range beginning at column c (0 at the last column of an
 * alignment); getLogProbFor(seq, s, e) = first[s] + sum cond[s+1..e].  A range of a single column has no value of its own in
 * the reference: its second tree keeps the observation of an earlier call (MeanFieldForBayesNet.java:84); this entry point
 * therefore covers ranges of >= 2 columns, and phyfoo_zoops_estep with an order-1 background reproduces the two single-column
 * flank ranges of every offset as the reference's call order produces them.  Alignments shorter than 2 columns are refused. */
int phyfoo_bg_columns_order1(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* cond_out, double* first_out);

/* Background of ANY order k >= 1 (models/PhyloBackground.java:241-305, structure :381-398) through the generic net kernel
 * (csrc/netg.cuh: parent / child lists and full CPF tables of up to 4^(k+1) rows, one thread per (k+1)-column window):
 *   first_out[c] = - sum of the free energies of the single trees 0 .. k-1 of the window starting at column c   (:273-283)
 *   cond_out[c]  = - free energy of the last tree of the window ending at column c                               (:284-297)
 * (0 where the window would leave its alignment), so that for a range of at least k + 1 columns
 *   getLogProbFor(seq, s, e) = first[s] + sum cond[s+k .. e].
 * Shorter ranges have no value of their own in the reference (their trailing trees keep the observation of an earlier call,
 * MeanFieldForBayesNet.java:84); phyfoo_zoops_estep with such a background reproduces the flank ranges of every offset as the
 * reference's call order produces them.  sweeps_out (nullable): mean-field sweeps per (k+1)-column window, alignment-major.
 * k = 1 gives the values of phyfoo_bg_columns_order1.  Alignments shorter than k + 1 columns are refused; k <= 7. */
int phyfoo_bg_range_terms(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, double* cond_out, double* first_out,
                          int32_t* sweeps_out);

/* ZOOPS E-step over the whole (local) sample: SingleHiddenMotifMixture.getNewWeights.
 *   logw[2]          log component weights (motif, no motif)
 *   strand_logw      NULL, or log strand weights [2] (motif wrapped in a StrandModel)
 *   aln_weights      NULL (all 1) or [n_aln]
 *   gamma_out        NULL or [windows]: seqweights[0] after the call (gamma_j * p_motif * weight)
 *   profile_out      NULL or [windows]: the unnormalised log profile a_j (getProfileOfScoresFor, UNNORMALIZED_CONDITIONAL)
 *   w_out[2], ll_out sufficient statistics of the component weights and the log-likelihood (local sums)
 *   argmax_off       NULL or [n_aln]: first offset of the maximal a_j (util/Util.java:528-538)
 *   argmax_strand    NULL or [n_aln]: strand call at that offset (0 wins ties) */
int phyfoo_zoops_estep(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, phyfoo_model* bg,
                       const double* logw, const double* strand_logw, const double* aln_weights,
                       double* gamma_out, double* profile_out, double* w_out, double* ll_out,
                       int32_t* argmax_off, uint8_t* argmax_strand, phyfoo_stats* stats);
/* Same E-step, but the three local sums [ll, w0, w1] are left in DEVICE memory at `partial_device`
 * (3 doubles) on `stream` (a cudaStream_t) so that the caller can all-reduce them over NCCL without a host round trip:
 * the library's work is ordered after what `stream` holds and `stream` waits for the result; asynchronous with respect to
 * the host.  stream == NULL (no caller stream; the legacy default stream does not order against the library's non-blocking
 * stream): the call returns when the three doubles are complete. */
int phyfoo_zoops_estep_device(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, phyfoo_model* bg,
                              const double* logw, const double* strand_logw, const double* aln_weights_device,
                              double* partial_device, void* stream);

/* ---- ZOOPS mixture of several motif models (BASELINE.json configs[3]: 3 motifs + homogeneous background) ----
 * The reference's mixture has exactly one motif component (jstacs!/.../SingleHiddenMotifMixture.java:716-718
 * getNumberOfMotifs() == 1); this is its direct generalisation (DESIGN.md "multi-motif mixture"): an alignment holds at most
 * one site of at most one of K <= 8 motifs,
 *   a_{k,j} = -log n_k + motif_k(j) - bg(j, j + W_k - 1),   Z_k = logsumexp_j a_{k,j},
 *   (p_1..p_K, p_0) = logSumNormalisation(logw_1 + Z_1, ..., logw_K + Z_K, logw_0),   ll += weight (logPSeq + lse),
 *   gamma_{k,j} = exp(a_{k,j} - Z_k) p_k weight,   w_k += p_k weight,
 * with the one-motif E-step's arithmetic per component (K = 1 gives phyfoo_zoops_estep's numbers).  logw[K + 1]: the motifs,
 * then "no motif"; gamma_out: NULL or K pointers (each NULL or [windows of that width]); w_out[K + 1]; argmax_motif / argmax_off
 * [n_aln]: first maximum of logw_k + a_{k,j} in (k, j) order.  Order-0 PhyloBackground.  The gamma of every component stays
 * on the device for phyfoo_select_multi (the M-step of component k: phyfoo_select_multi -> phyfoo_fe_prepare -> phyfoo_ss_* /
 * phyfoo_fe_*, exactly the one-motif M-step with gamma_k as weights). */
int phyfoo_zoops_estep_multi(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t n_motifs, phyfoo_model* const* fgs, phyfoo_model* bg,
                             const double* logw, const double* aln_weights, double* const* gamma_out, double* w_out, double* ll_out,
                             int32_t* argmax_motif, int32_t* argmax_off, phyfoo_stats* stats);
int phyfoo_select_multi(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t motif, double t, double q, int32_t* sel_aln,
                        int32_t* sel_off, double* sel_weight, int64_t capacity, int64_t* n_selected);

/* ---- genome scan / classification front-end (classification/ClassificationUtil.java) ----
 * Consumers of the per-offset motif scores that keep them on the device (config 5: 2*10^9 window scores).
 * scan_hits, per alignment: the number of offsets with score > threshold (determineSensitivityPerSeq :266-278 counts the
 * alignments where this is > 0; determinePositionSet :287-298 lists them), the first such offset (-1 if none), the
 * maximal score and its first offset (util/Util.java:528-538).  Any output pointer may be NULL. */
int phyfoo_scan_hits(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int32_t strand, double threshold,
                     int32_t* n_hits, int32_t* first_hit, double* max_score, int32_t* arg_max);
/* order statistics of all window scores of the sample: values_out[i] = sorted scores[ranks[i]] (ascending).
 * determineThresholdsForSpecifitiesPerPos :244-264 asks for rank (int)((n - 1) * specificity), determineThresholdForSpecifity
 * :161-180 for (int)(n * specificity); n is returned in n_scores_out (call with n_ranks = 0 to learn it). */
int phyfoo_score_order_stats(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* fg, int32_t strand, int32_t n_ranks,
                             const int64_t* ranks, double* values_out, int64_t* n_scores_out);

/* ---- non-phylogenetic alignment models, order 0 (the reference's `model.evolution=false` branch) ----
 * models/AlignmentBasedModel.java:188-248: window score = sum over species rows of sum over motif positions of
 * log condprob[i][base], a gap scoring log(0.25); condprob[width][4].  out[windows], alignment-major. */
int phyfoo_ab_score_windows(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, const double* condprob, int32_t strand, double* out);
/* models/AlignmentBasedBGModel.java:173-246: out[c] = sum over rows of log condprob4[base] (gap: log 0.25) for every column;
 * getLogProbFor(seq, s, e) is the sum of out over [s, e]. */
int phyfoo_ab_bg_columns(phyfoo_ctx* ctx, const phyfoo_dataset* ds, const double* condprob4, double* out);
/* models/AlignmentBasedModel.java:82-168 (count M-step over the selected windows, PSEUDO_COUNT = 1e-8 there): condprob_out[width][4] */
int phyfoo_ab_train(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, int64_t n_sel, const int32_t* sel_aln,
                    const int32_t* sel_off, const uint8_t* sel_strand, const double* sel_weight, double pseudo_count,
                    double* condprob_out);
/* Orders 0..2 of the same models (AlignmentBasedModel.java:97-228, AlignmentBasedBGModel.java:40-222): per species row a Markov
 * chain over the columns, entry of position p = condprob[p][sum_k 4^k x(l - k)] (k = 0 the column itself, k <= min(order, p) its
 * predecessors: inside the window for the motif, back to the start of the alignment for the background).  Tables are
 * concatenated: position p holds 4^(min(p, order) + 1) doubles, the current symbol fastest; the background has order + 1 tables
 * (column l of an alignment uses table min(l, order)).  phyfoo_ab_table_size(width, order) = doubles of a motif table set,
 * phyfoo_ab_table_size(0, order) = of a background's; -1 for an unsupported order.  Gaps follow the reference's expansion of an
 * unobserved symbol into the four bases, including the stale entries its background likelihood averages over when two
 * predecessors of a column are gaps (AlignmentBasedBGModel.java:185-191). */
int phyfoo_ab_table_size(int32_t width, int32_t order);
int phyfoo_ab_score_windows_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, int32_t order, const double* condprob,
                                  int32_t strand, double* out);
int phyfoo_ab_bg_columns_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t order, const double* condprob, double* out);
int phyfoo_ab_train_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, int32_t order, int64_t n_sel, const int32_t* sel_aln,
                          const int32_t* sel_off, const uint8_t* sel_strand, const double* sel_weight, double pseudo_count,
                          double* condprob_out);
/* AlignmentBasedBGModel.train (:40-137, PSEUDO_COUNT = 1e-2 there): counts over every column of every alignment weighted with
 * aln_weights[n_aln] (NULL = 1), a gap's weight split over its expansions; condprob_out as above */
int phyfoo_ab_bg_train_order(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t order, const double* aln_weights, double pseudo_count,
                             double* condprob_out);

/* ---- M-step data selection (io/SequenceSpecificDataSelector.java:37-99) ----
 * weights[windows] (e.g. gamma_out), one group per alignment: walk the windows by descending weight,
 * take those with raw weight >= q, stop after the running sum of normalised weights exceeds t.
 * sel_aln/sel_off/sel_weight must hold `capacity` entries; *n_selected receives the count
 * (PHYFOO_ENOMEM if capacity is too small). */
int phyfoo_select(phyfoo_ctx* ctx, const phyfoo_dataset* ds, int32_t width, const double* weights, double t,
                  double q, int32_t* sel_aln, int32_t* sel_off, double* sel_weight, int64_t capacity,
                  int64_t* n_selected);

/* ---- fixed-q free-energy objective (M-step) ----
 * prepare: runs the mean field to convergence for each listed window (fresh uniform start, like
 * PhyloBayesModel.train :142-147 / PhyloBackground.train :127-134) and keeps q on the device.
 * strand may be NULL (all forward). */
int phyfoo_fe_prepare(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, int64_t n_sel,
                      const int32_t* sel_aln, const int32_t* sel_off, const uint8_t* sel_strand,
                      const double* sel_weight, phyfoo_feset** out);
/* every (order+1)-column window of every alignment with per-alignment weights (PhyloBackground.train) */
int phyfoo_fe_prepare_all(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* m, const double* aln_weights,
                          phyfoo_feset** out);
/* f(lambda) = -sum_u w_u F(q_u; theta(lambda)); position >= 0: lambda holds that position's block(s)
 * (dim = 4*ctx), position < 0: all positions concatenated.  The model is left at lambda. */
int phyfoo_fe_eval(phyfoo_ctx* ctx, phyfoo_feset* fe, int32_t position, const double* lambda, int32_t dim,
                   double* f_out);
/* forward-difference gradient with the reference's perturbation order; g_out[dim] */
int phyfoo_fe_grad(phyfoo_ctx* ctx, phyfoo_feset* fe, int32_t position, const double* lambda, int32_t dim,
                   double eps, double* f_out, double* g_out);
/* device-resident variant for multi-GPU: values[dim+1] = [f(lambda), f(lambda+eps e_0), ...] local sums */
int phyfoo_fe_values_device(phyfoo_ctx* ctx, phyfoo_feset* fe, int32_t position, const double* lambda,
                            int32_t dim, double eps, double* values_device, void* stream);
/* Sufficient statistics of the objective: counts_out[W][E]: expected counts of the log-table entries of each position.  Order 0:
 * E = 4 + 16 (N - 1) (4 root entries, then [parent symbol][own symbol] per non-root node in tree node order); order-1 motif
 * network: E = 16 + 32 (N - 1) (root [context][symbol]; per non-root node the off-diagonal entries [context][symbol], then the
 * diagonal ones -- the F81-structured tables of csrc/model_host.hpp::build_tables1_into);
 * entropy_out: the parameter-independent term.  Then f(theta) = sum_t sum_e counts[t][e] * log table_theta,t[e] - entropy
 * for ANY parameters (pi or branch lengths) of any position: the optimizer needs no further device pass. */
int phyfoo_fe_suffstats(phyfoo_ctx* ctx, phyfoo_feset* fe, double* counts_out, double* entropy_out);
int64_t phyfoo_fe_size(const phyfoo_feset* fe);
void phyfoo_fe_free(phyfoo_ctx* ctx, phyfoo_feset* fe);

/* ---- the same objective on sufficient statistics, evaluated on the host INSIDE the library (motif orders 0 and 1; backgrounds of
 * any order through phyfoo_ss_create_bg below) ----
 * phyfoo_ss_create: one device pass (phyfoo_fe_suffstats) and a private copy of the model's parameters.  After that
 * evaluateFunction / evaluateGradientOfFunction cost a dot product of E terms and need neither the device nor a
 * collective: the sequential, data-dependent line search of the Jstacs optimizer (jstacs!/.../Optimizer.java:122-330) is no
 * longer a chain of kernel launches.  Same function as phyfoo_fe_eval (sums re-associated: 1e-12 relative,
 * tests/test_gpu_mstep.py follows a whole optimisation with both).  The object is left at the last lambda, like the model of
 * phyfoo_fe_eval; read the result with phyfoo_ss_get_pi and write it back with phyfoo_model_set_pi. */
typedef struct phyfoo_ssobj phyfoo_ssobj;
int phyfoo_ss_create(phyfoo_ctx* ctx, phyfoo_feset* fe, phyfoo_ssobj** out);
/* from counts the caller holds ([W][E], phyfoo_fe_suffstats layout); no device involved */
/* The same objective for a BACKGROUND of order k >= 1 (models/PhyloBackground.java:111-176: the mean fields of every
 * (k+1)-column window of every alignment, fixed once; f(theta) = - sum_p w_p F_p(theta) over the whole net of each window): one
 * pass of the generic net kernel (csrc/netg.cuh) collects the expected counts of the net's table entries.  aln_weights (nullable,
 * [n_aln]): window number p carries the weight of the alignment that holds COLUMN number p -- the reference indexes its per-column
 * weight vector with the window number (PhyloBackground.java:114,178-186; AbstractFreeEnergyOptimizer.java:104-108), kept.
 * Position t of the returned object has 4^min(t, k) x 4 parameters (phyfoo_ss_eval / _grad), phyfoo_ss_eval_branches works as
 * for the motif models.  (Order 0: phyfoo_fe_prepare_all + phyfoo_ss_create.) */
int phyfoo_ss_create_bg(phyfoo_ctx* ctx, const phyfoo_dataset* ds, phyfoo_model* bg, const double* aln_weights, phyfoo_ssobj** out);
int phyfoo_ss_create_host(const phyfoo_model_desc* desc, const double* counts, double entropy, phyfoo_ssobj** out);
int phyfoo_ss_eval(phyfoo_ssobj* ss, int32_t position, const double* lambda, int32_t dim, double* f_out);
int phyfoo_ss_grad(phyfoo_ssobj* ss, int32_t position, const double* lambda, int32_t dim, double eps, double* f_out, double* g_out);
/* branch lengths of one position (position < 0: of all positions) set to branch[N] (entry 0 unused): EdgeOptimizer /
 * GlobalEdgeOptimizer.evaluateFunction (optimizing/branchlengths/EdgeOptimizer.java:44-74) after the logit map */
int phyfoo_ss_eval_branches(phyfoo_ssobj* ss, int32_t position, const double* branch, double* f_out);
int phyfoo_ss_get_pi(const phyfoo_ssobj* ss, int32_t position, double* pi_out, int32_t n);
double phyfoo_ss_value(const phyfoo_ssobj* ss);
int64_t phyfoo_ss_evaluations(const phyfoo_ssobj* ss);
/* A whole M-step of the motif model in one call: the positions order[0..n_pos-1] one after the other (the caller draws the
 * order, util/PWMUtil.java:131-144), each by a BFGS on the forward-difference gradient with the termination rule of
 * optimizing/PreparedConditions.java:20-27.  The optimizer is Jstacs' quasiNewtonBFGS with its bracketing and Brent line search
 * restated statement by statement (csrc/suffstat_host.hpp; jstacs Optimizer.java:122-148,202-303,1000-1080); a Java caller can
 * just as well keep its own Optimizer and bind phyfoo_ss_eval / phyfoo_ss_grad instead.
 * With q fixed the objective is separable (f = sum_t G_t(theta_t) - ENT), so distinct positions are optimised side by side on up
 * to 16 host threads, each against the other positions' terms as they were on entry -- the result does not depend on the number
 * of threads; the environment variable PHYFOO_SS_THREADS=1 runs them strictly one after the other. */
int phyfoo_ss_optimize(phyfoo_ssobj* ss, const int32_t* order, int32_t n_pos, double* f_before, double* f_after, int64_t* evaluations);
void phyfoo_ss_free(phyfoo_ssobj* ss);

/* ---- options ----
 * "pattern_memo": 0 = never, 1 = when it pays off (default), 2 = whenever it applies.  For order-0 models and few
 * species (3 S <= 18 bits) the library scores each occurring column pattern once per position and parameter set and
 * gathers per window; results are bit-identical to the direct kernel (the reference's own, disabled, caches had the
 * same aim: models/AbstractPhyloModel.java:100-119, models/PhyloBackground.java:196). */
int phyfoo_set_option(phyfoo_ctx* ctx, const char* name, int64_t value);
/* "pattern_memo_sweep_cap" (default 64): sweeps stored per pattern; the (rare) windows whose stop rule needs more are
 * listed on the device and scored by the direct kernel -- same bits either way; 100 = max_sweeps serves every window
 * from the table. */
/* "pattern_memo_split" (default 0 = off; otherwise a multiple of 4): the trajectories are computed in two phases -- the
 * first `split` passes, then the rest on a second stream while the first gather pass over the windows runs (most windows
 * stop within a few sweeps); windows the first pass could not finish are gathered again once the second phase is done.
 * Results do not depend on the setting.  (Measured on C2: the two kernels compete for the same issue slots, so the
 * overlap gains 7 % with the four-lane trajectory kernel and nothing with the wavefront kernel; kept as a tuning knob.) */
/* "pattern_memo_wave" (default 1): the trajectory kernel runs the nodes of a tree as a wavefront over (node, sweep) --
 * node i of sweep k at step 2 k + depth(i), so nodes of alternating depth of consecutive sweeps run side by side on their
 * own lanes -- when the tree has 2..8 internal nodes and at most four per depth parity; 0 = four lanes per tree walking
 * the nodes one by one.  Same bits either way. */
/* "order1_kernel": 2 (default) = one thread per hidden node (all four symbols), 4 node slots x 8 windows per warp, windows
 * with an unobserved leaf through a device-side list to kernel 1; 1 = wavefront kernel (16 lanes per window, list-scheduled
 * node updates); 0 = one quad of lanes per window walking the nodes one by one.  Same fixed-point iteration; free energies
 * agree to re-association.
 * "order1_qglobal": kernel 2 keeps the q of a window's hidden nodes in shared memory (0) or in global memory behind L1 with
 * n warps per SM (1..8), which doubles the windows in flight for trees like C3's; -1 (default) chooses.
 * "order0_kernel": order-0 mean-field scoring outside the pattern table: 2 = the kernel compiled at run time for the model's
 * tree (below), 1 = its precompiled form for tables with F81 structure, 0 = the generic kernel; -1 (default) = 2, then 1 for
 * more than 6 species, else 0.  Kernels 1 and 2 are bit-identical; kernel 0 agrees with them to re-association.
 * "l2_persist" (default 1): per-window state kept in global memory (the q rows of the order-1 kernel, the state of very large
 * trees) is covered by an L2 access-policy window on the library's stream for the duration of the launch.
 * "bg_generic" (default 0): 1 = an order-1 background inside phyfoo_zoops_estep goes through the generic net kernel
 * (csrc/netg.cuh) that backgrounds of order >= 2 always use, instead of the order-1 wavefront kernel; same values to
 * re-association (a cross-check of the two statements of the mean field). */
/* Device time of every kernel the most recent call launched (CUDA events on the library's stream), in launch order:
 * returns the number of entries written (<= max_entries) or a negative error; names are static strings. */
int phyfoo_last_kernel_times(phyfoo_ctx* ctx, int32_t max_entries, const char** names_out, float* ms_out);
/* scoring path of the most recent motif/background launch: 0 direct kernel, 1 pattern table, 2 order-1 network */
int phyfoo_last_score_path(const phyfoo_ctx* ctx);

/* ---- order-0 scoring kernel compiled at run time for a model's tree (csrc/jit_o0.hpp; option "order0_kernel" = 2, and the
 * default choice for more than 6 species) ----
 * The tree of a model is fixed, so its mean-field pass (algorithm/MeanFieldForBayesNet.java:108-202, :232-365) is generated as
 * straight-line CUDA for that tree and compiled with NVRTC for sm_100a on first use (libnvrtc.so.12 and libcuda.so.1 are opened
 * with dlopen; without them, or when a compilation fails, scoring runs the precompiled F81-structured kernel and
 * phyfoo_jit_note says why).  phyfoo_jit_check generates and compiles without a device: source_out / log_out (nullable,
 * NUL-terminated, truncated) receive the generated source and the compiler log, *cubin_bytes the size of the image. */
int phyfoo_jit_check(const phyfoo_model_desc* desc, int32_t gaps, char* source_out, int64_t source_cap, char* log_out, int64_t log_cap,
                     int64_t* cubin_bytes);
const char* phyfoo_jit_note(const phyfoo_ctx* ctx);

/* ---- plumbing for callers that overlap the library with their own streams ---- */
/* Page-locked host memory (cudaHostAlloc): buffers handed to the library copy at full PCIe speed when they come from
 * here (a Java caller wraps the address in a MemorySegment / direct ByteBuffer).  NULL on failure. */
void* phyfoo_host_alloc(phyfoo_ctx* ctx, int64_t bytes);
void phyfoo_host_free(phyfoo_ctx* ctx, void* p);
int phyfoo_last_launches(const phyfoo_ctx* ctx); /* kernels launched by the most recent call on this context */
void* phyfoo_stream(phyfoo_ctx* ctx);            /* the library's cudaStream_t */
int phyfoo_synchronize(phyfoo_ctx* ctx);
/* device time (CUDA events on the library's stream) of the phases of the most recent E-step on this context:
 * background columns, motif windows, ZOOPS posterior + reduction.  Waits for that E-step to finish. */
int phyfoo_estep_kernel_ms(phyfoo_ctx* ctx, float* bg_ms, float* fg_ms, float* zoops_ms);
/* sums of mean-field sweep counts of the most recent E-step (integer parity check / roofline flop count) */
int phyfoo_estep_sweeps(phyfoo_ctx* ctx, int64_t* sweeps_fg, int64_t* sweeps_bg);

/* ---- several GPUs behind ONE caller (a JVM is one process) ----
 * the library takes a list of devices, shards the data set itself (contiguous blocks of alignments
 * balanced by column count; alignments are independent), replicates the models and owns the NCCL communicator
 * (ncclCommInitAll; libnccl.so.2 is bound at run time, PHYFOO_NCCL_LIB overrides the name).  The only exchange per consumer
 * is a sum all-reduce of a few doubles -- [ll, w0, w1] per E-step, [f, f_1..f_dim] per gradient -- issued with ncclAllReduce
 * on every device's stream right behind the kernel that produces them, before the one device-to-host copy.  One worker
 * thread per device runs the per-device part of each call; the calls themselves are synchronous like the single-GPU ones
 * and take the same arguments, with per-window / per-alignment arrays in the order of the WHOLE sample.  Results equal the
 * single-GPU ones up to the re-association of the all-reduced sums (1e-12 relative; scores, gamma, arg-max calls and
 * selected windows are bit-identical). */
typedef struct phyfoo_mctx phyfoo_mctx;
typedef struct phyfoo_mdataset phyfoo_mdataset;
typedef struct phyfoo_mmodel phyfoo_mmodel;
typedef struct phyfoo_mfeset phyfoo_mfeset;

int phyfoo_init_multi(const int* device_ids, int n_devices, phyfoo_mctx** out);
void phyfoo_destroy_multi(phyfoo_mctx* mctx);
const char* phyfoo_multi_last_error(const phyfoo_mctx* mctx); /* mctx may be NULL: error of a failed phyfoo_init_multi */
int phyfoo_multi_size(const phyfoo_mctx* mctx);
phyfoo_ctx* phyfoo_multi_ctx(phyfoo_mctx* mctx, int rank);    /* the device's own context (borrowed; e.g. for phyfoo_host_alloc) */
int phyfoo_multi_set_option(phyfoo_mctx* mctx, const char* name, int64_t value);
/* packs and shards: device r holds alignments bounds[r] .. bounds[r+1]-1 (bounds_out[n_devices + 1]) */
int phyfoo_multi_dataset_create(phyfoo_mctx* mctx, int32_t n_aln, int32_t n_species, const int64_t* col_offsets,
                                const uint8_t* codes, phyfoo_mdataset** out);
int phyfoo_multi_dataset_bounds(const phyfoo_mdataset* ds, int32_t* bounds_out);
void phyfoo_multi_dataset_free(phyfoo_mctx* mctx, phyfoo_mdataset* ds);
int phyfoo_multi_model_create(phyfoo_mctx* mctx, const phyfoo_model_desc* desc, phyfoo_mmodel** out);
int phyfoo_multi_model_set_pi(phyfoo_mctx* mctx, phyfoo_mmodel* m, int32_t position, const double* pi, int32_t n);
int phyfoo_multi_model_get_pi(phyfoo_mctx* mctx, const phyfoo_mmodel* m, int32_t position, double* pi_out, int32_t n);
int phyfoo_multi_model_set_branch(phyfoo_mctx* mctx, phyfoo_mmodel* m, int32_t position, int32_t node, double length);
int phyfoo_multi_model_set_mode(phyfoo_mctx* mctx, phyfoo_mmodel* m, int32_t mode);
void phyfoo_multi_model_free(phyfoo_mctx* mctx, phyfoo_mmodel* m);
/* SingleHiddenMotifMixture.getNewWeights over the whole sample: arguments as phyfoo_zoops_estep; w_out / ll_out are the sums
 * over all devices; stats: sums (gpu_ms: the slowest device) */
int phyfoo_multi_zoops_estep(phyfoo_mctx* mctx, const phyfoo_mdataset* ds, phyfoo_mmodel* fg, phyfoo_mmodel* bg,
                             const double* logw, const double* strand_logw, const double* aln_weights,
                             double* gamma_out, double* profile_out, double* w_out, double* ll_out,
                             int32_t* argmax_off, uint8_t* argmax_strand, phyfoo_stats* stats);
/* M-step: phyfoo_select + phyfoo_fe_prepare on every device.  weights: NULL = the gamma of the last phyfoo_multi_zoops_estep
 * (still resident on each device), or [windows of the whole sample].  n_selected (nullable): total number of selected windows. */
int phyfoo_multi_fe_prepare_selected(phyfoo_mctx* mctx, const phyfoo_mdataset* ds, phyfoo_mmodel* m, const double* weights,
                                     double t, double q, phyfoo_mfeset** out, int64_t* n_selected);
/* evaluateFunction / evaluateGradientOfFunction over all devices: [f, f_1..f_dim] all-reduced before the copy-out */
int phyfoo_multi_fe_eval(phyfoo_mctx* mctx, phyfoo_mfeset* fe, int32_t position, const double* lambda, int32_t dim, double* f_out);
int phyfoo_multi_fe_grad(phyfoo_mctx* mctx, phyfoo_mfeset* fe, int32_t position, const double* lambda, int32_t dim, double eps,
                         double* f_out, double* g_out);
/* sufficient statistics of all devices: counts and entropy all-reduced once (W * E + 1 doubles), then phyfoo_ss_* as above */
int phyfoo_multi_ss_create(phyfoo_mctx* mctx, phyfoo_mfeset* fe, phyfoo_ssobj** out);
int64_t phyfoo_multi_fe_size(const phyfoo_mfeset* fe);
void phyfoo_multi_fe_free(phyfoo_mctx* mctx, phyfoo_mfeset* fe);

#ifdef __cplusplus
}
#endif
#endif /* PHYFOO_H */
