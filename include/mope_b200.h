/*
 * mope_b200 -- C ABI of the B200-native ontogenetic-phylogeny log-likelihood.
 *
 * This is the drop-in boundary for ONE hot path of ammodramus/mope: what emcee calls once per
 * walker per step through mope's inference / likelihoods API.  Reference interfaces replaced
 * (paths relative to the reference tree):
 *
 *   mope_b200_create            <- Inference.__init__                       mope/inference.py:263-738
 *                                  (tables: TransitionData / TransitionDataBottleneck loaders,
 *                                   mope/transition_data_2d.py:78-111, transition_data_bottleneck.py:74-96;
 *                                   leaf emissions: _binom.get_binom_likelihoods_cython, mope/_binom.pyx:49-92)
 *   mope_b200_eval              <- Inference.loglike                        mope/inference.py:770-910
 *                                  = likelihoods.get_log_likelihood_somatic_newick  mope/likelihoods.py:308-403
 *                                  + ascertainment.get_locus_asc_probs              mope/ascertainment.py:132-236
 *                                  + Inference.poisson_log_like                     mope/inference.py:1113-1129
 *                                  (matrix fetch: TransitionData.get_transition_probabilities_2d,
 *                                   mope/transition_data_2d.py:133-238 incl. _util.check_transition_matrix
 *                                   mope/_util.pyx:129-148 and _interp.bilinear_interpolate mope/_interp.pyx:7-37;
 *                                   branch updates: mope/_likes.pyx:42-56,139-152,232-248; root: :312-322,426-442)
 *   mope_b200_transition_matrix <- TransitionData.get_transition_probabilities_2d / bottleneck twin
 *   mope_b200_leaf_likelihoods  <- _binom.get_binom_likelihoods_cython      mope/_binom.pyx:49-92
 *   mope_b200_branch_update     <- _likes.compute_{length,rate,bottleneck}_transition_likelihood
 *   mope_b200_non_asc_probs     <- _poisson_binom.get_non_asc_probs         mope/_poisson_binom.pyx:18-90
 *   mope_b200_root_prior        <- likelihoods.get_stationary_distribution_double_beta   mope/likelihoods.py:27-60
 *   mope_b200_eval_resident     <- Inference.loglike with the parameter vectors in HBM (planning by a kernel)
 *   mope_b200_mcmc_halfstep     <- one half-step of emcee's stretch move as Inference.run_mcmc drives it
 *                                  (mope/inference.py:1353-1356) + Inference.log_posterior (:1132-1161)
 *   mope_b200_gen_*             <- generate_transitions.py:57-67,194-228,386-414, _transition.pyx:6-33 (table generator)
 *   mope_b200_sel_loglike       <- likelihoods.get_log_l_and_asc_prob_selection          mope/likelihoods.py:404-525
 *   mope_b200_sel_branch_update <- _likes selection / masked operators      mope/_likes.pyx:59-136,155-229,251-307
 *
 * Conventions: plain C, no exceptions across the boundary.  Every function returns 0 on success or a
 * negative MOPE_E_* code; mope_b200_last_error() gives the message of the last failure on the calling
 * thread.  All arithmetic is fp64.  Device memory is OWNED BY THE CALLER (PyTorch in the Python host):
 * the caller asks for sizes with the *_bytes queries, allocates, and passes device pointers in.  (The operator-level entry
 * points -- transition_matrix, leaf_likelihoods, branch_update, non_asc_probs, sel_* -- take host arrays and stage them
 * through a grow-only scratch the context owns and mope_b200_destroy frees; mope_b200_sel_loglike also keeps its
 * per-pedigree row-group tables there.)
 * Functions taking a `stream` enqueue their work on that CUDA stream (cudaStream_t passed as void*);
 * unless documented otherwise they return after the work is enqueued.  A context is bound to one device
 * and is not re-entrant (external locking), matching the reference's process-private Inference object.
 */
#ifndef MOPE_B200_H
#define MOPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOPE_OK 0
#define MOPE_E_INVALID (-1)   /* bad argument / malformed description */
#define MOPE_E_CUDA (-2)      /* CUDA runtime error (message has the CUDA string) */
#define MOPE_E_NOMEM (-3)     /* arena / workspace too small */
#define MOPE_E_NODEVICE (-4)  /* no usable sm_100 device */

/* per-walker status bits written by mope_b200_eval (the reference raises ValueError for each) */
#define MOPE_ST_OK 0
#define MOPE_ST_TIME_TOO_LARGE 1    /* transition_data_2d.py:165-167 */
#define MOPE_ST_MUT_TOO_SMALL 2     /* transition_data_2d.py:168-170, transition_data_bottleneck.py:124-126 */
#define MOPE_ST_MUT_TOO_LARGE 4     /* transition_data_2d.py:171-173, transition_data_bottleneck.py:127-129 */
#define MOPE_ST_NB_TOO_SMALL 8      /* transition_data_bottleneck.py:118-120 */
#define MOPE_ST_NB_TOO_LARGE 16     /* transition_data_bottleneck.py:121-123 */
#define MOPE_ST_NOT_FINITE 32       /* transition_data_2d.py:135-136,158-161 */
#define MOPE_ST_NO_BOTTLENECK_TABLE 64 /* likelihoods.py:381-382 */
#define MOPE_ST_SKIPPED 128         /* device-resident evaluation only: prior = -inf, likelihood not evaluated (inference.py:1148-1149) */

typedef struct mope_ctx mope_ctx;

/* Transition tables, host pointers, row-major, complete sorted grids (the in-RAM state of
 * TransitionData(memory=True)).  drift[g][u] is the F x F matrix for gens[g], us[u]; P[i][j] =
 * Pr(child bin j | parent bin i).  bot may be NULL (no --bottlenecks file). */
typedef struct {
    int32_t F;
    int32_t n_gens, n_us;
    const double* drift;   /* [n_gens][n_us][F][F] */
    const double* gens;    /* [n_gens] ascending (generations, as float64 like _sorted_gens) */
    const double* us;      /* [n_us] ascending (per-generation mutation probability) */
    int32_t n_nbs, n_bot_us;
    const double* bot;     /* [n_nbs][n_bot_us][F][F] or NULL */
    const double* nbs;     /* [n_nbs] ascending bottleneck sizes */
    const double* bot_us;  /* [n_bot_us] ascending */
    double N;              /* Wright-Fisher population size of the tables */
    const double* freqs;   /* [F] bin frequencies */
    const int32_t* breaks; /* [F] first Wright-Fisher state (0..N) of every frequency bin, ascending from 0 (the file attribute
                            * `breaks`, generate_transitions.py:133-192); NULL = no device root prior (mope_b200_root_prior) */
} mope_tables_desc;

/* One pedigree type (data file + tree + ages file) after the reference's init-time processing
 * (fixed loci dropped, families counted).  Nodes are numbered in the reference's postorder
 * (newick.Node.walk('postorder')); node b < n_branches is also branch b; the root is node n_branches. */
typedef struct {
    int32_t n_leaves, n_branches, n_loci, n_fam, n_mult, n_asc;
    const double* counts;        /* [n_loci][n_leaves] focal-allele counts, NaN = missing tissue */
    const double* coverage;      /* [n_loci][n_leaves] */
    const double* row_mult;      /* [n_mult][n_loci] multiplier (age) value of every locus row */
    const double* asc_mult;      /* [n_mult][n_asc] multiplier values of every distinct ascertainment row pair */
    const int32_t* fam_asc;      /* [n_fam] index of the family's ascertainment row pair (0..n_asc-1) */
    const double* fam_count;     /* [n_fam] number of loci of the family (asc_ages['count']) */
    const double* fam_lgamma;    /* [n_fam] gammaln(count+1) */
    const int32_t* br_parent;    /* [n_branches] parent node id */
    const int32_t* br_leaf;      /* [n_branches] leaf column (0..n_leaves-1) or -1 for an internal node */
    const int32_t* br_var;       /* [n_branches] index of the branch's parameter name (0..n_var-1) */
    const int32_t* br_mult;      /* [n_branches] multiplier column (0..n_mult-1) or -1 */
    const int32_t* br_bottleneck;/* [n_branches] 1 for a `^` branch */
} mope_pedigree_desc;

typedef struct {
    int32_t n_var;                   /* distinct branch parameter names; ndim = 2*n_var + 2 */
    int32_t n_ped;
    const mope_pedigree_desc* peds;  /* [n_ped] */
    double min_phred_score;          /* -1 = none (mope/_binom.pyx:76-79) */
    double min_freq;                 /* ascertainment threshold (--min-het-freq) */
    double genome_size;
    double poisson_like_penalty;
} mope_model_desc;

/* Bytes of device memory the context needs for tables, emissions and model arrays. */
int mope_b200_arena_bytes(const mope_tables_desc* tables, const mope_model_desc* model, size_t* out_bytes);

/* Build a context on `device`.  Uploads the tables, runs the leaf-emission kernel, compiles the branch
 * schedule.  `arena` is caller-owned device memory of at least mope_b200_arena_bytes(); it must outlive
 * the context.  Synchronous. */
int mope_b200_create(const mope_tables_desc* tables, const mope_model_desc* model, int device,
                     void* arena, size_t arena_bytes, void* stream, mope_ctx** out);

int mope_b200_destroy(mope_ctx* ctx);

/* Workspace (caller-owned device scratch) needed to evaluate `n_walkers` in one pass.  Smaller
 * workspaces are accepted by mope_b200_eval as long as they hold one walker; it then runs in chunks. */
int mope_b200_workspace_bytes(const mope_ctx* ctx, int n_walkers, size_t* out_min_bytes, size_t* out_full_bytes);

/* Evaluate Inference.loglike for a batch of walkers.
 *   params  host [n_walkers][ndim]: log10 branch lengths (n_var), log10 scaled mutation rates (n_var),
 *           log10 ab, log10 ppoly -- the vector `loglike` receives (log_posterior's sign folding and the
 *           prior stay in the host wrapper, mope/inference.py:1132-1149).
 *   stat    host [n_walkers][F] root distribution (likelihoods.get_stationary_distribution_double_beta,
 *           mope/likelihoods.py:57-60; computed by the host wrapper with the reference's own SciPy call).
 *   out_ll  host [n_walkers].  Walkers whose parameters leave the tables get NaN and a non-zero status.
 *   out_status   host [n_walkers] MOPE_ST_* bits, may be NULL.
 *   out_root_ll  host [n_walkers][n_ped] tree log-likelihood before ascertainment, may be NULL.
 *   out_log_asc  host [n_walkers][sum_ped n_asc] per distinct ascertainment row pair, may be NULL.
 * Host<->device copies of params/stat/outputs happen inside; the call returns after the results are in
 * the host arrays (it synchronises `stream`). */
int mope_b200_eval(mope_ctx* ctx, const double* params, const double* stat, int n_walkers,
                   void* workspace, size_t workspace_bytes, void* stream,
                   double* out_ll, int32_t* out_status, double* out_root_ll, double* out_log_asc);

/* Same computation with everything already resident on the device (device pointers; params/stat/out as
 * above).  Enqueues only; no host synchronisation.  Per-walker interpolation plans are still derived on
 * the host from `params_host` (tiny), so the host copy of params is required too. */
int mope_b200_eval_device(mope_ctx* ctx, const double* params_host, const double* stat_dev, int n_walkers,
                          void* workspace, size_t workspace_bytes, void* stream,
                          double* out_ll_dev, int32_t* out_status_host);

/* Root (stationary) distribution of every walker on the device: likelihoods.get_stationary_distribution_double_beta
 * (mope/likelihoods.py:27-60) for alphabeta, polyprob = 10**x[-2:] (mope/inference.py:824-827).
 *   params_dev    device [n_walkers][ndim], the vectors mope_b200_eval receives (only the last two columns are read)
 *   out_stat_dev  device [n_walkers][F], usable as the `stat_dev` argument of mope_b200_eval_device
 * Every cell mass is integrated directly instead of being taken as a difference of two incomplete-beta values, so it
 * carries ~2e-15 relative error for any ab; the reference's own cell masses lose digits as ab -> 0 (relative error
 * ~1e-16 / (ab * cell width): 3e-8 at ab = 1e-5, 3e-4 at ab = 1e-9), which is why the host wrapper keeps SciPy's betainc
 * for walkers with ab < 1e-6 when bit-level agreement with the reference is wanted.  Enqueues only.  Needs
 * mope_tables_desc.breaks. */
int mope_b200_root_prior(mope_ctx* ctx, const double* params_dev, int n_walkers, void* stream, double* out_stat_dev);

/* Device-resident evaluation: the same computation as mope_b200_eval with NOTHING on the host -- the parameter vectors stay
 * in HBM and the per-walker interpolation plans (range checks, searchsorted, bilinear weights) are derived by a kernel.
 *   params_dev      device [n_walkers][ndim] (sign-folded log10 parameters, as for mope_b200_eval)
 *   stat_dev        device [n_walkers][F] root distributions, or NULL: computed here by the kernel of mope_b200_root_prior
 *   lnprior_dev     device [n_walkers] or NULL; walkers whose entry is -inf (or NaN) are skipped (status MOPE_ST_SKIPPED)
 *   out_ll_dev      device [n_walkers]; NaN for walkers with a non-zero status
 *   out_status_dev  device [n_walkers] MOPE_ST_* bits
 * The matrix pool is laid out statically (every walker owns the worst-case number of blocks), so the workspace must hold
 * at least one walker at that size: mope_b200_workspace_bytes gives it.  Enqueues only; no host synchronisation, no
 * host<->device copy. */
int mope_b200_eval_resident(mope_ctx* ctx, const double* params_dev, const double* stat_dev, const double* lnprior_dev,
                            int n_walkers, void* workspace, size_t workspace_bytes, void* stream,
                            double* out_ll_dev, int32_t* out_status_dev);

/* Ensemble sampler on the device: one half-step of emcee's stretch move (what mope's Inference.run_mcmc iterates,
 * mope/inference.py:1353-1356, with `a` = --chain-alpha), including Inference.log_posterior's sign folding and box prior
 * (mope/inference.py:1132-1149, 1055-1110) and the likelihood of the proposals (mope_b200_eval_resident).  All pointers
 * are device pointers owned by the caller; nothing is copied to or from the host and the call only enqueues.
 * The ensemble is split into walkers [0, n/2) and [n/2, n); which_half = 0 moves the first half against the second. */
typedef struct {
    double* pos;             /* [n_walkers][ndim] positions (log10 parameters as emcee holds them, not sign-folded) */
    double* lnprob;          /* [n_walkers] log posterior of pos */
    int64_t* naccept;        /* [n_walkers] accepted proposals so far */
    double* q;               /* scratch [n_walkers/2][ndim] proposals */
    double* folded;          /* scratch [n_walkers/2][ndim] sign-folded proposals */
    double* factor;          /* scratch [n_walkers/2] */
    double* lnprior;         /* scratch [n_walkers/2] */
    double* ll;              /* scratch [n_walkers/2] */
    int32_t* status;         /* scratch [n_walkers/2] MOPE_ST_* bits of the proposals' evaluation */
    const double* lower;     /* [ndim] prior box (Inference.lower / upper) */
    const double* upper;
    const double* prior_slope; /* [ndim] log prior inside the box = prior_const + sum_d prior_slope[d] * x[d] */
    double prior_const;
    const double* zu;        /* [n_walkers/2] uniforms for z, or NULL: counter-based Philox4x32-10 keyed by seed, step */
    const int32_t* partner;  /* [n_walkers/2] partner index within the other half (required when zu is given) */
    const double* au;        /* [n_walkers/2] uniforms for the accept test, or NULL: Philox */
    uint64_t seed, step;     /* Philox key / counter; step must differ for every half-step of a chain */
    double a;                /* stretch scale */
    int32_t n_walkers;
} mope_sampler_state;

int mope_b200_mcmc_halfstep(mope_ctx* ctx, const mope_sampler_state* state, int which_half,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels this library launched since the context was created. */
int64_t mope_b200_launch_count(const mope_ctx* ctx);

/* Host-only query (no device): where the pruning kernel keeps the parked messages of pedigree `ped` at grid size F.
 * The pass walks the tree in the reference's postorder (likelihoods.py:352-397); the message of every child but the last
 * waits ("is parked") until its siblings are done.  out[0..3] = parks per 64-row tile living in the chain buffer, in a
 * tensor-memory slot, in the L2 scratch, and the number of L2 scratch slots. */
int mope_b200_schedule_stats(const mope_model_desc* model, int F, int ped, int32_t* out);

/* Branch updates of the last eval, counted per (walker, branch): GEMMs executed, and branches that took the
 * reference's identity short-cut (N*t below the first table generation, transition_data_2d.py:184-187), for which
 * the message is forwarded without a GEMM.  For age-scaled branches every (walker, branch) counts as one GEMM. */
int mope_b200_last_eval_stats(const mope_ctx* ctx, int64_t* gemm_ops, int64_t* identity_ops);

/* Executed work of the last eval, counted by the pruning kernel itself: the sum over all GEMM units it ran through the
 * tensor pipe (a unit = one branch, or one generation segment of an age-scaled branch, on one tile) of the unit's real
 * rows.  2 * F^2 * unit_rows = executed flops.  Branches forwarded by the identity short-cut, generation segments no
 * row of a warp blends, the ascertainment short-cuts and the K chunks of a leaf branch in which a warp's eight emission
 * rows are exactly zero (skipped: they add nothing) do not count -- the kernels count rows x k columns and this returns
 * that sum / F.  Synchronises the stream of the last eval. */
int mope_b200_last_eval_unit_rows(mope_ctx* ctx, int64_t* unit_rows);

/* Age-scaled branches of the last eval: (warp, generation segment) pairs the pruning kernel went through -- a warp = 8
 * rows of a tile, a segment = one grid generation the tile's rows blend on that branch -- and how many of them the warp
 * actually multiplied in (at least one of its rows has a non-zero weight there).  The difference is tensor-pipe time the
 * warp's partner spends alone.  F <= 208 only (zeros otherwise).  Synchronises the stream of the last eval. */
int mope_b200_last_eval_segment_stats(mope_ctx* ctx, int64_t* warp_segments, int64_t* active_warp_segments);

/* Measurement support (bench.py).  With profiling on, every pruning-kernel launch is bracketed by CUDA events on
 * its stream; mope_b200_kernel_times synchronises, returns the accumulated device time (ms) and launch count of the
 * pruning kernel and of the interpolation kernel since the last call, and resets the accumulators. */
int mope_b200_set_profiling(mope_ctx* ctx, int on);
int mope_b200_kernel_times(mope_ctx* ctx, double* tree_ms, int64_t* tree_launches, double* interp_ms, int64_t* interp_launches);
/* fp64 tensor-pipe ceiling: register-only mma.sync.m8n8k4.f64 loop for about `millis` ms on `stream`; TFLOP/s out. */
int mope_b200_fp64_peak(mope_ctx* ctx, double millis, void* stream, double* out_tflops);

/* Operator-level entry points (parity tests mirror the reference's operator contracts). */

/* P = get_transition_probabilities_2d(x, scaled_mut): bottleneck=0 -> drift tables, x = scaled time,
 * result clamped/renormalised; bottleneck=1 -> bottleneck tables, x = Nb, no check.  out_P host [F][F].
 * Returns 0 and *out_status = MOPE_ST_* (out_P untouched when status != 0). */
int mope_b200_transition_matrix(mope_ctx* ctx, int bottleneck, double x, double scaled_mut,
                                void* stream, double* out_P, int32_t* out_status);

/* likes[l][s][f] = Binom(counts[l][s]; coverage[l][s], p_f) exactly as _binom.get_binom_likelihoods_cython;
 * host pointers; uses ctx for the device only. */
int mope_b200_leaf_likelihoods(mope_ctx* ctx, const double* counts, const double* coverage, int n_loci,
                               int n_leaves, const double* freqs, int F, double min_phred_score,
                               void* stream, double* out_likes);

/* parent[r][:] *= P_r @ child[r][:] for n_rows rows.  kind: 0 fixed length, 1 per-row lengths (rate),
 * 2 bottleneck.  lengths: host [1] (kind 0/2) or [n_rows] (kind 1).  child/parent host [n_rows][F]. */
int mope_b200_branch_update(mope_ctx* ctx, int kind, const double* child, double* parent, int n_rows,
                            const double* lengths, double scaled_mut, void* stream, int32_t* out_status);

/* Poisson-binomial non-ascertainment probabilities for n_sets read sets (ragged: qual_off[n_sets+1]
 * offsets into qualities).  out host [n_sets][F]. */
int mope_b200_non_asc_probs(mope_ctx* ctx, const double* qualities, const int64_t* qual_off, int n_sets,
                            const double* freqs, int F, double min_empirical_freq, void* stream, double* out);

/* Host helper of the root prior (likelihoods.discretize_beta_distn, mope/likelihoods.py:27-41: st.beta.cdf at the
 * cell edges).  out[i][j] = betainc(ab[i], ab[i], edges[j]) evaluated with `betainc_fn`, the C entry point of SciPy's
 * own regularised incomplete beta function (scipy.special.cython_special.__pyx_capi__["__pyx_fuse_0betainc"],
 * signature double(double a, double b, double x, int skip_dispatch)) -- the values are bit-identical to
 * scipy.stats.beta.cdf, which the CDF differences of the reference require; rows are split over n_threads native
 * threads (no interpreter lock involved).  Pure host code: no context, no device. */
int mope_b200_beta_cdf_table(void* betainc_fn, const double* ab, int n_ab, const double* edges, int n_edges,
                             double* out, int n_threads);

/* ---- transition-table generator (replaces the NumPy/SciPy loops of mope/generate_transitions.py) ----
 * Device pointers, fp64, row-major; no context (tables are generated before any model exists).  `stream` is a
 * cudaStream_t (0 = default); the calls are asynchronous on it. */

/* P[i][j] = Binomial(j; n_trials, p_host[i]) for i < rows_valid and j <= n_trials, 0 elsewhere in the rows x cols matrix
 * (leading dimension ld): the Wright-Fisher matrix (generate_transitions.py:57-67: n_trials = N, p = pstar) and the
 * bottleneck factors (_transition.pyx:6-33: rows_valid = N1 + 1, n_trials = N2).  lnchoose_host: ln C(n_trials, j),
 * j = 0..n_trials, from the caller's log-gamma (the reference's pmf is SciPy's: its gammaln and libm's lgamma differ in
 * the last bit, which a 10 000-generation matrix power amplifies to ~1e-8); NULL = libm lgamma. */
int mope_b200_gen_binom_matrix(double* P_dev, int ld, int rows, int cols, int rows_valid, int n_trials,
                               const double* p_host, const double* lnchoose_host, void* stream);

/* At[c][r] = A[r][c] (rows x cols -> cols x rows). */
int mope_b200_gen_transpose(const double* A_dev, double* At_dev, int rows, int cols, int lda, int ldt, void* stream);

/* C[M x N] = A[M x K] . Bt[N x K]^T on the fp64 tensor pipe (np.dot / np.linalg.matrix_power of
 * generate_transitions.py:48-53,217-228,386-414).  16-byte aligned operands, even leading dimensions. */
int mope_b200_gen_dgemm_tn(const double* A_dev, const double* Bt_dev, double* C_dev, int M, int N, int K,
                           int lda, int ldbt, int ldc, void* stream);

/* out[F][F] = bin_matrix(P, breaks) (generate_transitions.py:194-208); breaks_dev: int32 [F] first state of every bin. */
int mope_b200_gen_bin_matrix(const double* P_dev, int ld, int n_states, const int32_t* breaks_dev, int F,
                             double* out_dev, void* stream);

/* ---- selection model and masked-matrix variants ----
 * The context's drift grid is read as (generations x selection coefficient s): `us` of mope_tables_desc holds the s values
 * of a selection table set (TransitionData(selection=True), transition_data_2d.py:52-55); the scaled coefficient is
 * alpha = 2 N s, exactly like the scaled mutation rate of the neutral model. */

/* parent[r][:] *= Pmask_r . child[r][:] with P_r = get_transition_probabilities_2d(time_r, alphas[r]) for n_rows rows.
 *   mask 0  the full matrix: _likes.compute_rate_transition_likelihood_selection (_likes.pyx:59-79; per_row_times = 1,
 *           times[n_rows]) and compute_length_transition_likelihood_selection (:155-175; per_row_times = 0, times[1]);
 *   mask 1  Pzero[0,0] = P[0,0]: compute_leaf_zero_transition_likelihood / compute_length_transition_likelihood_zero
 *           (:82-98, :178-194);  mask 2  Pzero[0,1:-1]: ..._focal_node_zero / ..._zero_focalnode (:101-117, :197-212);
 *   mask 3  Pzero[1:-1,1:-1]: ..._zero_descendant (:120-136, :215-229).  For the masked variants pass the single mutation
 *           rate in every alphas[r].  mask | MOPE_MASK_BOTTLENECK: the matrix comes from the bottleneck tables and times[0]
 *           is the bottleneck size: compute_bottleneck_transition_likelihood_zero / _zero_focalnode / _zero_descendant
 *           (:251-307).
 * child / parent host [n_rows][F]; out_trans (may be NULL) host [n_rows][F] receives the products Pmask_r . child[r]
 * (the `ret` array of the selection variants).  *out_status = MOPE_ST_* of the first failing fetch (nothing is modified). */
#define MOPE_MASK_BOTTLENECK 4
int mope_b200_sel_branch_update(mope_ctx* ctx, int mask, const double* child, double* parent, int n_rows,
                                const double* times, int per_row_times, const double* alphas, double* out_trans,
                                void* stream, int32_t* out_status);

/* likelihoods.get_log_l_and_asc_prob_selection (likelihoods.py:404-525) for pedigree `ped` of the context, one parameter
 * vector: branch_lengths host [n_branches] (branch order), locus_alphas host [n_branches][n_pos] scaled selection
 * coefficient of every branch at every unique position, position_idx host [n_loci] unique-position index of every locus
 * row (in the row order of the pedigree description), stat host [F] root distribution.  Three rows per locus (likelihood,
 * all-zero, all-one: e0 = 1[freq <= min_freq], eN = 1[freq >= 1 - min_freq]) go through the tree; out_loglikes /
 * out_log_asc host [n_loci]. */
int mope_b200_sel_loglike(mope_ctx* ctx, int ped, const double* branch_lengths, const double* locus_alphas, int n_pos,
                          const int32_t* position_idx, const double* stat, void* stream, double* out_loglikes,
                          double* out_log_asc, int32_t* out_status);

const char* mope_b200_last_error(void);
const char* mope_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOPE_B200_H */
