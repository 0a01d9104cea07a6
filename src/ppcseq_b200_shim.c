/* .Call shim between the ppcseq R package and libppcseq_b200.so.
 *
 * Replaces, in the reference package, src/RcppExports.cpp:15-25 (the registration of the Rcpp module that
 * boots the stanc-generated model class) -- see INTEGRATION.md for the R-side edits (R/stanmodels.R,
 * do_inference in R/utilities.R:1482-1544).  Pure marshalling over the C ABI in include/ppcseq_b200.h:
 * R owns every input and output vector, handles travel as external pointers with finalizers, errors are
 * raised with Rf_error only after the C call has returned (no longjmp across C++/CUDA frames).
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD: the build image has no R headers (R.h / Rinternals.h).  Build
 * inside the R package with
 *     PKG_CPPFLAGS = -I<repo>/include        PKG_LIBS = -L<libdir> -lppcseq_b200 -Wl,-rpath,<libdir>
 * in src/Makevars (replacing the StanHeaders / RcppParallel flags of the reference's src/Makevars:3-9).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "ppcseq_b200.h"

static void check(int rc) {
    if (rc != PPCSEQ_OK) Rf_error("ppcseq_b200 [%d]: %s", rc, ppcseq_last_error());
}

static void model_finalizer(SEXP p) {
    ppcseq_model *m = (ppcseq_model *)R_ExternalPtrAddr(p);
    if (m) { ppcseq_model_free(m); R_ClearExternalPtr(p); }
}
static void fit_finalizer(SEXP p) {
    ppcseq_fit *f = (ppcseq_fit *)R_ExternalPtrAddr(p);
    if (f) { ppcseq_fit_free(f); R_ClearExternalPtr(p); }
}
static ppcseq_model *get_model(SEXP p) {
    ppcseq_model *m = (ppcseq_model *)R_ExternalPtrAddr(p);
    if (!m) Rf_error("ppcseq_b200: model handle is NULL (already freed?)");
    return m;
}
static ppcseq_fit *get_fit(SEXP p) {
    ppcseq_fit *f = (ppcseq_fit *)R_ExternalPtrAddr(p);
    if (!f) Rf_error("ppcseq_b200: fit handle is NULL (already freed?)");
    return f;
}

/* counts: integer matrix S x G (column-major in R == gene-major [G][S] in C); X: numeric matrix S x C;
 * exposure_rate: numeric[S]; K = how_many_to_check; lambda_mu_mu; devices = integer vector of CUDA ordinals: one entry =
 * a single-GPU model, several = ONE handle with the genes split over those GPUs inside this R process
 * (ppcseq_model_create_multi) -- every other routine of this file takes either kind of handle. */
SEXP ppcseqb200_model_create(SEXP counts, SEXP X, SEXP exposure, SEXP K, SEXP lambda_mu_mu, SEXP devices) {
    if (!Rf_isInteger(counts) || !Rf_isMatrix(counts)) Rf_error("counts must be an integer matrix [S, G]");
    const int S = Rf_nrows(counts), G = Rf_ncols(counts), C = Rf_ncols(X);
    if (Rf_nrows(X) != S || LENGTH(exposure) != S) Rf_error("X / exposure_rate do not match counts");
    /* model.matrix is column-major [S][C]; the ABI takes row-major [S][C] */
    double *Xr = (double *)R_alloc((size_t)S * C, sizeof(double));
    const double *Xc = REAL(X);
    for (int s = 0; s < S; ++s)
        for (int c = 0; c < C; ++c) Xr[(size_t)s * C + c] = Xc[(size_t)c * S + s];
    ppcseq_model *m = NULL;
    if (!Rf_isInteger(devices) || LENGTH(devices) < 1) Rf_error("devices must be an integer vector of CUDA ordinals");
    if (LENGTH(devices) == 1)
        check(ppcseq_model_create(G, S, C, Rf_asInteger(K), INTEGER(counts), Xr, REAL(exposure), Rf_asReal(lambda_mu_mu),
                                  INTEGER(devices)[0], &m));
    else
        check(ppcseq_model_create_multi(G, S, C, Rf_asInteger(K), INTEGER(counts), Xr, REAL(exposure),
                                        Rf_asReal(lambda_mu_mu), LENGTH(devices), INTEGER(devices), &m));
    SEXP p = PROTECT(R_MakeExternalPtr(m, Rf_install("ppcseq_model"), R_NilValue));
    R_RegisterCFinalizerEx(p, model_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* pairs: integer matrix 2 x n of 0-based (g, s) */
SEXP ppcseqb200_set_exclusion(SEXP model, SEXP pairs) {
    check(ppcseq_model_set_exclusion(get_model(model), INTEGER(pairs), (int64_t)(LENGTH(pairs) / 2)));
    return R_NilValue;
}

/* rstan::log_prob / grad_log_prob stand-in: theta numeric[D] -> list(lp, grad) */
SEXP ppcseqb200_log_prob_grad(SEXP model, SEXP theta, SEXP propto, SEXP jacobian) {
    int64_t D = 0;
    check(ppcseq_model_dims(get_model(model), NULL, NULL, NULL, NULL, &D));
    if ((int64_t)LENGTH(theta) != D) Rf_error("theta must have length %lld", (long long)D);
    SEXP lp = PROTECT(Rf_allocVector(REALSXP, 1)), grad = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)D));
    int rc = ppcseq_log_prob_grad(get_model(model), 1, REAL(theta), Rf_asLogical(propto), Rf_asLogical(jacobian), REAL(lp),
                                  REAL(grad));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, lp); SET_VECTOR_ELT(out, 1, grad);
    UNPROTECT(3);
    check(rc);
    return out;
}

static SEXP wrap_fit(ppcseq_fit *f) {
    SEXP p = PROTECT(R_MakeExternalPtr(f, Rf_install("ppcseq_fit"), R_NilValue));
    R_RegisterCFinalizerEx(p, fit_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* sampling(stanmodels$negBinomial_MPI, chains, iter, warmup, seed, init = "random")  (R/utilities.R:1497-1512) */
SEXP ppcseqb200_sample_nuts(SEXP model, SEXP chains, SEXP iter, SEXP warmup, SEXP seed) {
    ppcseq_nuts_opts o;
    ppcseq_nuts_default_opts(&o);
    o.chains = Rf_asInteger(chains); o.iter = Rf_asInteger(iter); o.warmup = Rf_asInteger(warmup);
    o.seed = (uint64_t)Rf_asReal(seed);
    ppcseq_fit *f = NULL;
    check(ppcseq_sample_nuts(get_model(model), &o, &f));
    return wrap_fit(f);
}

/* vb(model, output_samples, iter, tol_rel_obj)  (R/utilities.R:256-264); an error here is what vb_iterative retries on */
SEXP ppcseqb200_advi(SEXP model, SEXP output_samples, SEXP iter, SEXP tol_rel_obj, SEXP seed) {
    ppcseq_advi_opts o;
    ppcseq_advi_default_opts(&o);
    o.output_samples = Rf_asInteger(output_samples); o.iter = Rf_asInteger(iter); o.tol_rel_obj = Rf_asReal(tol_rel_obj);
    o.seed = (uint64_t)Rf_asReal(seed);
    ppcseq_fit *f = NULL;
    check(ppcseq_advi_meanfield(get_model(model), &o, &f));
    return wrap_fit(f);
}

/* fit_to_counts_rng / fit_to_counts_rng_approximated (R/utilities.R:685-703, :733-784):
 * returns list(.lower, .upper, mean, sd), each an S x K matrix (R column-major == [K][S] gene-major) */
SEXP ppcseqb200_ppc_summary(SEXP fit, SEXP S_, SEXP K_, SEXP exact, SEXP n_draws, SEXP p, SEXP tc, SEXP seed) {
    const int S = Rf_asInteger(S_), K = Rf_asInteger(K_);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    for (int i = 0; i < 4; ++i) SET_VECTOR_ELT(out, i, Rf_allocMatrix(REALSXP, S, K));
    int rc = ppcseq_ppc_summary(get_fit(fit), Rf_asLogical(exact), (int64_t)Rf_asReal(n_draws), Rf_asReal(p), Rf_asReal(tc),
                                (uint64_t)Rf_asReal(seed), REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)),
                                REAL(VECTOR_ELT(out, 2)), REAL(VECTOR_ELT(out, 3)));
    UNPROTECT(1);
    check(rc);
    return out;
}

/* summary_to_tibble(fit, "alpha_sub_1", "G")$mean (R/utilities.R:1250-1263, :1531): posterior means of `count`
 * consecutive unconstrained parameters starting at 0-based `begin` */
SEXP ppcseqb200_param_mean(SEXP fit, SEXP begin, SEXP count) {
    const R_xlen_t n = (R_xlen_t)Rf_asReal(count);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
    int rc = ppcseq_fit_param_mean(get_fit(fit), (int64_t)Rf_asReal(begin), (int64_t)n, REAL(out));
    UNPROTECT(1);
    check(rc);
    return out;
}

/* rstan::extract stand-in (R/utilities.R:738, :743): n_draws x count matrix */
SEXP ppcseqb200_get_draws(SEXP fit, SEXP begin, SEXP count) {
    int32_t n = 0;
    check(ppcseq_fit_num_draws(get_fit(fit), &n));
    const int cnt = Rf_asInteger(count);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, cnt));      /* column-major [count][n] == the ABI's layout */
    int rc = ppcseq_fit_get_draws(get_fit(fit), (int64_t)Rf_asReal(begin), cnt, REAL(out));
    UNPROTECT(1);
    check(rc);
    return out;
}

/* check_if_within_posterior + add_deleterious_if_covariate_exists + totals (R/utilities.R:651-663, :493-513, :597-606) */
SEXP ppcseqb200_flags(SEXP model, SEXP lower, SEXP upper, SEXP mean, SEXP slope) {
    int32_t S = 0, K = 0, C = 0;
    check(ppcseq_model_dims(get_model(model), NULL, &S, &C, &K, NULL));
    const size_t np = (size_t)S * K;
    uint8_t *ppc = (uint8_t *)R_alloc(np, 1), *del = (uint8_t *)R_alloc(np, 1);
    SEXP failed = PROTECT(Rf_allocVector(INTSXP, K)), tot = PROTECT(Rf_allocVector(INTSXP, K));
    int rc = ppcseq_flags(get_model(model), REAL(lower), REAL(upper), REAL(mean), C > 1 ? REAL(slope) : NULL, ppc,
                          C > 1 ? del : NULL, INTEGER(failed), C > 1 ? INTEGER(tot) : NULL);
    SEXP lppc = PROTECT(Rf_allocMatrix(LGLSXP, S, K)), ldel = PROTECT(Rf_allocMatrix(LGLSXP, S, K));
    for (size_t i = 0; i < np; ++i) { LOGICAL(lppc)[i] = ppc[i]; LOGICAL(ldel)[i] = C > 1 ? del[i] : NA_LOGICAL; }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, lppc); SET_VECTOR_ELT(out, 1, ldel); SET_VECTOR_ELT(out, 2, failed); SET_VECTOR_ELT(out, 3, tot);
    UNPROTECT(5);
    check(rc);
    return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"ppcseqb200_model_create", (DL_FUNC)&ppcseqb200_model_create, 6},
    {"ppcseqb200_set_exclusion", (DL_FUNC)&ppcseqb200_set_exclusion, 2},
    {"ppcseqb200_log_prob_grad", (DL_FUNC)&ppcseqb200_log_prob_grad, 4},
    {"ppcseqb200_sample_nuts", (DL_FUNC)&ppcseqb200_sample_nuts, 5},
    {"ppcseqb200_advi", (DL_FUNC)&ppcseqb200_advi, 5},
    {"ppcseqb200_ppc_summary", (DL_FUNC)&ppcseqb200_ppc_summary, 8},
    {"ppcseqb200_param_mean", (DL_FUNC)&ppcseqb200_param_mean, 3},
    {"ppcseqb200_get_draws", (DL_FUNC)&ppcseqb200_get_draws, 3},
    {"ppcseqb200_flags", (DL_FUNC)&ppcseqb200_flags, 5},
    {NULL, NULL, 0}};

/* same registration pattern as the reference's src/RcppExports.cpp:22-25 */
void R_init_ppcseq(DllInfo *dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
