/* ppcseq_b200 -- C ABI of the B200-native hot path of stemangiola/ppcseq.
 *
 * This is the drop-in boundary: what the reference reaches today through the compiled Stan model
 * object `stanmodels$negBinomial_MPI` (reference R/stanmodels.R:10-25, registered by
 * src/RcppExports.cpp:15-25) and the rstan calls made on it in do_inference()
 * (R/utilities.R:1482-1513, :256-264) plus the fit queries that follow
 * (R/utilities.R:689-691, :738-743, :1252-1255).  Plain pointers and sizes only; every function
 * returns 0 on success or a PPCSEQ_E* code, with a message in ppcseq_last_error().
 *
 * Conventions
 *   - counts are int32, dense, gene-major [G][S]; S follows the reference's S index
 *     (R/utilities.R:955-958), G its G index (checked genes are 0..K-1, R/utilities.R:949-952).
 *   - X is the model.matrix, row-major [S][C] (R/utilities.R:887-900).
 *   - theta is Stan's unconstrained vector in declaration order
 *     (inst/stan/negBinomial_MPI.stan:180-199), length D = 6 + 2G + K + max(0,C-2)K:
 *     [lambda_mu(offset), log lambda_sigma, lambda_skew, intercept[G], alpha_sub_1[K],
 *      alpha_2 (col-major (C-2)xK), sigma_raw[G], log(-sigma_slope), sigma_intercept, log sigma_sigma]
 *   - host-pointer entry points copy in/out; *_device entry points take device pointers and only
 *     enqueue work on `stream` (a cudaStream_t passed as void*; NULL = the model's own stream).
 *   - the library never keeps a host pointer after a call returns.
 *   - a model handle is NOT re-entrant: its default evaluation scratch (grid-reduction counters and lines) serves ONE
 *     evaluation stream at a time.  Issue ppcseq_log_prob_grad[_device] calls on one handle from one host thread and on
 *     one stream (or order the streams yourself); the samplers allocate their own per-chain scratch and may run their
 *     chains concurrently.  Different handles are independent.
 */
#ifndef PPCSEQ_B200_H
#define PPCSEQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPCSEQ_OK 0
#define PPCSEQ_EINVAL 1   /* bad argument */
#define PPCSEQ_ECUDA 2    /* CUDA runtime error (message has the cudaError string) */
#define PPCSEQ_ENOMEM 3
#define PPCSEQ_ESTATE 4   /* call not valid in the handle's current state */
#define PPCSEQ_ENCCL 5
#define PPCSEQ_EDIVERGED 6 /* the inference algorithm failed (non-finite gradient, no usable step size) */
#define PPCSEQ_ECOMM 7    /* a device-side wait timed out (gene shards out of step / a lost reduction line): the
                             evaluation's lp is NaN, the handle's status flag is sticky, discard the handle */

typedef struct ppcseq_model ppcseq_model;
typedef struct ppcseq_fit ppcseq_fit;     /* posterior draws resident in HBM: the stanfit stand-in */

/* thread-local message of the last failing call */
const char *ppcseq_last_error(void);
/* ABI version (bumped on any signature change) */
int ppcseq_abi_version(void);

/* ---- model = Stan data block (inst/stan/negBinomial_MPI.stan:142-173) resident in HBM --------
 * Replaces the implicit data lookup rstan does in do_inference's frame (R/utilities.R:1395-1473).
 * `device` is the CUDA ordinal.  Data are copied to the device; nothing is retained on the host. */
int ppcseq_model_create(int32_t G, int32_t S, int32_t C, int32_t K,
                        const int32_t *counts, const double *X, const double *exposure_rate,
                        double lambda_mu_mu, int device, ppcseq_model **out);
void ppcseq_model_free(ppcseq_model *m);

/* Gene-sharded construction (map_rect's shards, negBinomial_MPI.stan:226-240, spread over GPUs):
 * this rank holds genes [g_begin, g_end) of a G_total-gene problem whose first K_total genes are
 * checked.  counts is the local [g_end-g_begin][S] block.  The local unconstrained vector has the
 * layout above with G = g_end-g_begin and K = clamp(K_total - g_begin, 0, G); the 6 scalar
 * hyper-parameters are replicated on every rank. */
int ppcseq_model_create_shard(int32_t G_total, int32_t K_total, int32_t g_begin, int32_t g_end,
                              int32_t S, int32_t C, const int32_t *counts_local, const double *X,
                              const double *exposure_rate, double lambda_mu_mu, int device,
                              ppcseq_model **out);

/* ONE process, SEVERAL GPUs of one box -- what a single R process calling identify_outliers() needs (the reference's
 * model object lives in one process, R/stanmodels.R:10-25 + src/RcppExports.cpp:15-25, and shards the likelihood over
 * map_rect workers inside it, negBinomial_MPI.stan:226-240; chains / shards per process: R/utilities.R:1377-1386).
 * The genes are split into n_devices contiguous blocks, one shard model per device (devices[] = CUDA ordinals, NULL =
 * 0..n_devices-1); direct peer access is enabled between the devices and the shards' mailboxes are wired to each
 * other, so every evaluation is one kernel per device ending in the fused NVLink all-reduce -- no IPC handles, no
 * second process, no NCCL.  The returned handle presents the GLOBAL problem and works with every host-pointer entry
 * point of this header (dims, set_exclusion, log_prob_grad, sample_nuts, advi_meanfield, fit queries, ppc_summary,
 * ppc_draws, flags): theta / gradients / draws use the global layout, lp and hyper-gradients are bitwise those of every
 * shard, PPC runs gene-sharded with Philox streams keyed by the global (gene, sample) pair (same draws as one GPU).
 * The *_device and ppcseq_comm_* entry points return PPCSEQ_ESTATE on such a handle.  Needs peer access between all
 * listed devices (PPCSEQ_ESTATE otherwise) and G >= n_devices; n_devices = 1 is allowed. */
int ppcseq_model_create_multi(int32_t G, int32_t S, int32_t C, int32_t K, const int32_t *counts, const double *X,
                              const double *exposure_rate, double lambda_mu_mu, int32_t n_devices,
                              const int32_t *devices, ppcseq_model **out);

/* Pass-2 exclusion (negBinomial_MPI.stan:105-115; built by R/methods.R:292-300 and
 * R/utilities.R:321-359): `pairs` holds n (g, s) pairs, 0-based, local gene index.  n = 0 clears. */
int ppcseq_model_set_exclusion(ppcseq_model *m, const int32_t *pairs, int64_t n);

/* Likelihood path: 0 = auto; 1 = general (any design matrix, per-element exp); 2 = per-element categorical
 * (X has <= 8 distinct rows: no per-element exp); 3 = Chebyshev-moment categorical (additionally needs a bounded
 * exposure range; the mu-dependent half of the likelihood is evaluated from data-only moments).  Auto picks
 * 3, then 2, then 1.  2 and 3 fail with PPCSEQ_ESTATE when the model is not eligible. */
int ppcseq_model_set_design_path(ppcseq_model *m, int mode);

int ppcseq_model_dims(const ppcseq_model *m, int32_t *G, int32_t *S, int32_t *C, int32_t *K, int64_t *D);

/* ---- log_prob / grad_log_prob (rstan::log_prob, rstan::grad_log_prob on the model) -----------
 * propto / jacobian as in Stan's log_prob<propto, jacobian>.  B thetas at once ([B][D]). */
int ppcseq_log_prob_grad(ppcseq_model *m, int32_t B, const double *theta, int propto, int jacobian,
                         double *lp /*[B]*/, double *grad /*[B][D]*/);
/* device-resident: d_theta [B][D], d_lp [B], d_grad [B][D]; asynchronous on `stream`. */
int ppcseq_log_prob_grad_device(ppcseq_model *m, int32_t B, const double *d_theta, int propto,
                                int jacobian, double *d_lp, double *d_grad, void *stream);
/* sharded variant: gene-block gradients are final; the 8 raw partial sums
 * [lp, d_xi, d_omega, d_skew, d_slope, d_sigma_intercept, d_sigma_sigma, 0] per theta go to
 * d_partials [B][8] for the caller to all-reduce (SUM) across ranks, then ppcseq_finalize_hyper_device
 * applies the hyper-priors, constraints and Jacobians and writes lp and the 6 hyper-gradients. */
int ppcseq_log_prob_grad_partial_device(ppcseq_model *m, int32_t B, const double *d_theta, int propto,
                                        double *d_partials, double *d_grad, void *stream);
int ppcseq_finalize_hyper_device(ppcseq_model *m, int32_t B, const double *d_theta,
                                 const double *d_partials_summed, int propto, int jacobian,
                                 double *d_lp, double *d_grad, void *stream);

/* OPTIONAL per-sample output, NOT part of log_prob's gradient and of no parity claim: out[s] = d log_prob / d
 * exposure_rate[s] = sum_g w_gs dl/d eta_gs (an S-vector; excluded points have weight 0).  exposure_rate is DATA in the
 * reference (inst/stan/negBinomial_MPI.stan:167-168, from TMM, R/methods.R:234-238); BASELINE.json's config 5 names an
 * "exposure-gradient allreduce": this is the quantity, and on gene shards the caller (or the multi-GPU handle) sums the
 * shards' vectors.  Reads the count matrix once (HBM-bound), deterministic two-stage sum. */
int ppcseq_exposure_grad(ppcseq_model *m, const double *theta /*[D]*/, double *out /*[S]*/);
int ppcseq_exposure_grad_device(ppcseq_model *m, const double *d_theta, double *d_out, void *stream);

/* ---- gene shards on several GPUs without a host-visible collective ---------------------------------------------
 * The all-reduce(SUM) of the 8 partial sums is fused INTO the log_prob kernel: the last warp of its grid reduction
 * writes the sums into a mailbox on every peer GPU (stores over NVLink into peer-mapped memory; every value travels as
 * one 16-byte line that carries the launch's sequence number next to the data, so no fence or separate flag is
 * needed), polls its own mailbox for the peers' lines and adds the W contributions in rank order -- bitwise the same
 * result on every rank, one kernel per evaluation, no NCCL call on the data path.  This replaces the gather in sum(map_rect(...))
 * (inst/stan/negBinomial_MPI.stan:226) when the shards live on different GPUs (one process per GPU).
 *   1. every rank: ppcseq_comm_create(model, rank, world, channels, cap, handle)  -> 64-byte IPC handle
 *   2. exchange the handles out of band (e.g. torch.distributed.all_gather), rank order
 *   3. every rank: ppcseq_comm_connect(model, all_handles)
 * After that ppcseq_log_prob_grad[_device] on the shard model returns the GLOBAL lp and hyper-gradients (channel 0).
 * All ranks must issue the same sequence of evaluations; a rank that waits more than ~2 s sets the status flag
 * instead of hanging.  `channels` >= 1 (further channels serve concurrent evaluation streams), `cap` = max batch. */
#define PPCSEQ_COMM_HANDLE_BYTES 64
int ppcseq_comm_create(ppcseq_model *m, int32_t rank, int32_t world, int32_t channels, int32_t cap, uint8_t *handle_out);
int ppcseq_comm_connect(ppcseq_model *m, const uint8_t *all_handles /* [world][PPCSEQ_COMM_HANDLE_BYTES] */);
int ppcseq_comm_status(ppcseq_model *m, int32_t *timed_out);
/* Device-side failure flags of the handle (sticky): bit 0 = fused all-reduce time-out, bit 1 = grid-reduction
 * time-out.  Every host-synchronising entry point (ppcseq_log_prob_grad, ppcseq_stream_sync, the samplers) checks
 * them itself and returns PPCSEQ_ECOMM; this query is for callers of the asynchronous *_device entry points, after
 * they have synchronised their stream.  A flagged evaluation also has lp = NaN. */
int ppcseq_model_status(ppcseq_model *m, int32_t *flags);

/* ---- posterior-predictive summaries and flags -------------------------------------------------
 * Per-pair summary of an explicit draws matrix (what rstan::summary(fit, "counts_rng", prob = c(p, 1-p))
 * returns, R/utilities.R:689-691, or quantile/mean/sd in R/utilities.R:770-776): draws is [n_draws][n_pairs]
 * (draw-major), integer-valued.  lower/upper are R quantile type 7 at p and 1-p, bit-exact. */
int ppcseq_summarise_draws(int device, const double *draws, int32_t n_draws, int64_t n_pairs, double p,
                           double *lower, double *upper, double *mean, double *sd);
/* Outlier flags (R/utilities.R:651-663 check_if_within_posterior, :493-513
 * add_deleterious_if_covariate_exists) and per-gene totals (:597, :604).  lower/upper/mean and the
 * uint8 outputs are [K][S] gene-major; slope [K] is the posterior mean of alpha_sub_1 (:1531).  When
 * C == 1 there is no deleterious column: slope/deleterious/tot_deleterious_outliers may be NULL. */
int ppcseq_flags(ppcseq_model *m, const double *lower, const double *upper, const double *mean,
                 const double *slope, uint8_t *ppc, uint8_t *deleterious, int32_t *ppc_samples_failed,
                 int32_t *tot_deleterious_outliers);

/* ---- fit handle: what rstan::sampling / rstan::vb return (R/utilities.R:1482-1513) ----------------
 * Holds n posterior draws of the unconstrained vector in HBM.  ppcseq_fit_from_draws imports draws the
 * caller already has ([n][D], draw-major) -- the test hook and the way to re-use an external fit. */
int ppcseq_fit_from_draws(ppcseq_model *m, const double *theta_draws, int32_t n, ppcseq_fit **out);
void ppcseq_fit_free(ppcseq_fit *f);
int ppcseq_fit_num_draws(const ppcseq_fit *f, int32_t *n);
/* rstan::extract stand-in (R/utilities.R:738-743): out is [param_count][n_draws], parameter-major. */
int ppcseq_fit_get_draws(const ppcseq_fit *f, int64_t param_begin, int64_t param_count, double *out);
/* posterior means, e.g. slope = mean alpha_sub_1[g] (summary_to_tibble, R/utilities.R:1250-1263, :1531) */
int ppcseq_fit_param_mean(const ppcseq_fit *f, int64_t param_begin, int64_t param_count, double *out);
/* sampler diagnostics written by the sampler that produced the fit.  Slots:
 *   0 algorithm (0 imported draws, 1 NUTS, 2 ADVI)   1 log_prob+grad evaluations   2 wall seconds
 *   NUTS: 3 divergent transitions (post warm-up)  4 transitions that hit max_treedepth  5 mean accept_stat
 *         6 mean adapted step size  7 mean leapfrog steps per post-warm-up iteration
 *   ADVI: 3 iterations run  4 stop reason (1 mean ELBO, 2 median ELBO, 3 both, 0 iteration limit)  5 last ELBO
 *         6 eta used  7 ELBO evaluations
 *   all:  8 gamma draws clamped at 2^30 in the LAST ppcseq_ppc_* call on this fit (Stan's neg_binomial_2_log_rng
 *         raises there; 0 in any sane fit -- callers should treat > 0 as "posterior not usable") */
int ppcseq_fit_info(const ppcseq_fit *f, double *out, int32_t n);

/* ---- inference: the two sampler calls of do_inference() (R/utilities.R:1482-1513) -----------------------
 * Both run entirely on the device (all D-length state in HBM; the host steers from a few scalars per step)
 * and return a fit handle holding the draws of the unconstrained vector.  Defaults are rstan's, with the
 * values the reference overrides noted; *_default_opts fills them in. */
typedef struct ppcseq_nuts_opts {
    int32_t chains;            /* reference: max(3, min(cores, argmin_c draws/c + 150 c)), R/utilities.R:291-303, :1377-1380 */
    int32_t iter;              /* per chain INCLUDING warm-up; reference: ceil(draws/chains) + 150 (:1502) */
    int32_t warmup;            /* 150 (:1503) */
    int32_t max_treedepth;     /* 10 */
    int32_t adapt_init_buffer; /* 75 */
    int32_t adapt_term_buffer; /* 50 */
    int32_t adapt_window;      /* 25 */
    int32_t threads;           /* host threads driving chains concurrently (0 = one per chain); < 0: the batched-chain
                                  driver (one thread, every launch covers all chains; what gene-sharded runs use) */
    double adapt_delta;        /* 0.8 */
    double adapt_gamma;        /* 0.05 */
    double adapt_kappa;        /* 0.75 */
    double adapt_t0;           /* 10 */
    double stepsize;           /* 1 */
    double init_radius;        /* 2: init = "random" draws U(-2,2) on the unconstrained scale (:1506) */
    uint64_t seed;             /* (:1505) */
    const double *init;        /* optional [chains][D] starting points; NULL = random */
} ppcseq_nuts_opts;

typedef struct ppcseq_advi_opts {
    int32_t iter;              /* reference: 50000 (R/utilities.R:1491) */
    int32_t grad_samples;      /* 1 */
    int32_t elbo_samples;      /* 100 */
    int32_t eval_elbo;         /* 100 */
    int32_t output_samples;    /* reference: draws_practical (:1490) */
    int32_t adapt_engaged;     /* 1 */
    int32_t adapt_iter;        /* 50 */
    int32_t reserved;
    double eta;                /* step-size scale when adapt_engaged = 0 */
    double tol_rel_obj;        /* reference: 0.005, hard-coded (:1492) */
    double init_radius;        /* 2 */
    uint64_t seed;
    const double *init;        /* optional [D] starting mean; NULL = random */
} ppcseq_advi_opts;

int ppcseq_nuts_default_opts(ppcseq_nuts_opts *o);
int ppcseq_advi_default_opts(ppcseq_advi_opts *o);
/* rstan::sampling(stanmodels$negBinomial_MPI, chains, iter, warmup, seed, init = "random", save_warmup = FALSE)
 * (R/utilities.R:1497-1512): NUTS with a diagonal metric, Stan's windowed adaptation, multinomial sampling and
 * the generalised U-turn criterion.  The fit holds chains * (iter - warmup) draws, chain-major. */
int ppcseq_sample_nuts(ppcseq_model *m, const ppcseq_nuts_opts *o, ppcseq_fit **out);
/* rstan::vb(algorithm = "meanfield") as driven by vb_iterative (R/utilities.R:246-278): mean-field ADVI with
 * Stan's step-size adaptation and relative-ELBO convergence rule.  Returns PPCSEQ_EDIVERGED where rstan::vb
 * would raise (the caller retries, as vb_iterative does).  The fit holds output_samples draws. */
int ppcseq_advi_meanfield(ppcseq_model *m, const ppcseq_advi_opts *o, ppcseq_fit **out);

/* Fused posterior-predictive draw + summary; outputs [K][S] gene-major.
 *   exact != 0: generated quantities of every saved draw (negBinomial_MPI.stan:259-266) summarised as
 *               rstan::summary does (fit_to_counts_rng, R/utilities.R:685-703); n_draws is ignored.
 *   exact == 0: fit_to_counts_rng_approximated (R/utilities.R:733-784): n_draws NB variates per pair, each
 *               from a posterior draw sampled with replacement.
 * The NB variate is Poisson(Gamma(phi', exp(eta)/phi')), phi' = sigma[g] * truncation_compensation, from a
 * counter-based Philox4x32-10 stream keyed by (seed, pair, draw): draws are never materialised. */
int ppcseq_ppc_summary(ppcseq_fit *f, int exact, int64_t n_draws, double p, double truncation_compensation,
                       uint64_t seed, double *lower, double *upper, double *mean, double *sd);
/* raw counts_rng of every saved draw, [n_draws][K*S] (small problems; save_generated_quantities,
 * R/utilities.R:786-802); same stream as ppcseq_ppc_summary(exact = 1) with the same seed. */
int ppcseq_ppc_draws(ppcseq_fit *f, double truncation_compensation, uint64_t seed, double *counts_rng);

/* ---- input side (host only, no CUDA call): tidy table -> the dense layout ppcseq_model_create takes -----------------
 * Replaces select_to_check_and_house_keeping (R/utilities.R:628-649) + format_input (R/utilities.R:924-959).  The
 * columns are row-aligned, n rows: integer ids of the transcript and of the sample (a caller with string columns
 * passes factor codes), abundance (int32 or int64 by `abundance_itemsize`), significance, do_check (0/1).  Selected
 * rows = rows with do_check, then the rows of the LAST how_many_negative_controls transcripts of
 * distinct(arrange(significance)); G and S are numbered by first appearance in that order, so the K checked genes come
 * first.  Fails (EINVAL) on NaN significance, negative or >= 2^31 abundance, duplicated (transcript, sample) rows and on
 * a table that is not rectangular.  threads == 0: all host cores (fewer for small tables); threads < 0: exactly
 * -threads row chunks whatever the size (the tests use it to run tiny tables through the chunk merge). */
typedef struct ppcseq_prep ppcseq_prep;
int ppcseq_prep_table(int64_t n, const int64_t *transcript, const int64_t *sample, const void *abundance,
                      int32_t abundance_itemsize, const double *significance, const uint8_t *do_check,
                      int64_t how_many_negative_controls, int32_t threads, ppcseq_prep **out);
int ppcseq_prep_dims(const ppcseq_prep *p, int32_t *G, int32_t *S, int32_t *K);
/* gene_ids[G], sample_ids[S] (the ids in G / S order), first_row[S] (a row of the table that belongs to the sample:
 * where its covariates are read), counts[G][S]; any pointer may be NULL */
int ppcseq_prep_fetch(const ppcseq_prep *p, int64_t *gene_ids, int64_t *sample_ids, int64_t *first_row, int32_t *counts);
void ppcseq_prep_free(ppcseq_prep *p);
/* edgeR TMM normalisation factors on dense counts [G][S] (get_scaled_counts_bulk / calcNormFactor, R/tidybulk.R:150-241,
 * :262-323; exposure_rate = -log(multiplier), R/methods.R:222-238).  `order` (NULL = identity) lists the columns in
 * factor(sample) level order; factors[S], lib_size[S] (column sums) and *ref are in that order.  ref_in < 0: the
 * reference is the first level whose median count is the largest; factors are scaled to geometric mean 1. */
int ppcseq_tmm_factors(int32_t G, int32_t S, const int32_t *counts, const int32_t *order, int32_t ref_in, int32_t threads,
                       double *factors, double *lib_size, int32_t *ref);

/* device-side scratch the bench needs */
int ppcseq_device_alloc(int device, int64_t bytes, void **out);
int ppcseq_device_free(int device, void *p);
int ppcseq_memcpy_h2d(void *dst, const void *src, int64_t bytes, void *stream);
int ppcseq_memcpy_d2h(void *dst, const void *src, int64_t bytes, void *stream);
int ppcseq_stream_sync(ppcseq_model *m, void *stream);
/* wall time on the device (CUDA events on `stream`) of `iters` back-to-back evaluations of the
 * B-theta batch; writes milliseconds per launch into ms_each[iters]. */
int ppcseq_time_log_prob_grad_device(ppcseq_model *m, int32_t B, const double *d_theta, int propto,
                                     int jacobian, double *d_lp, double *d_grad, void *stream,
                                     int32_t iters, int flush_l2, float *ms_each);
/* measured DFMA throughput of the device (TFLOP/s, 2 flop per FMA), for the FP64 roofline */
int ppcseq_measure_fp64_peak(int device, double *tflops);
/* kernels launched by this library since load (the bench's gpu_launches counter) */
int64_t ppcseq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
