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

/* select_to_check_and_house_keeping + format_input (R/utilities.R:628-649, :924-959) in one native pass over the tidy
 * table.  transcript / sample: integer codes (as.integer(factor(...)) or any integer ids), abundance: integer, significance:
 * numeric, do_check: logical -- all row-aligned.  Returns list(counts [S, G] integer matrix, gene ids in G order, sample
 * ids in S order, how_many_to_check, first_row = a 1-based row of the table for every sample (its covariates)). */
SEXP ppcseqb200_prep_table(SEXP transcript, SEXP sample, SEXP abundance, SEXP significance, SEXP do_check,
                           SEXP how_many_negative_controls, SEXP threads) {
    const R_xlen_t n = XLENGTH(abundance);
    if (!Rf_isInteger(transcript) || !Rf_isInteger(sample) || !Rf_isInteger(abundance))
        Rf_error("transcript, sample (codes) and abundance must be integer vectors");   /* R/methods.R:139-148 */
    if (XLENGTH(transcript) != n || XLENGTH(sample) != n || XLENGTH(significance) != n || XLENGTH(do_check) != n)
        Rf_error("the columns of the table have different lengths");
    int64_t *t64 = (int64_t *)R_alloc((size_t)n, sizeof(int64_t)), *s64 = (int64_t *)R_alloc((size_t)n, sizeof(int64_t));
    uint8_t *chk = (uint8_t *)R_alloc((size_t)n, 1);
    const int *ti = INTEGER(transcript), *si = INTEGER(sample), *ci = LOGICAL(do_check);
    for (R_xlen_t i = 0; i < n; ++i) {
        if (ti[i] == NA_INTEGER || si[i] == NA_INTEGER || INTEGER(abundance)[i] == NA_INTEGER)
            Rf_error("NA in the transcript, sample or abundance column");                /* check_if_any_NA */
        t64[i] = ti[i]; s64[i] = si[i]; chk[i] = ci[i] == 1;
    }
    ppcseq_prep *h = NULL;
    check(ppcseq_prep_table((int64_t)n, t64, s64, INTEGER(abundance), 4, REAL(significance), chk,
                            (int64_t)Rf_asInteger(how_many_negative_controls), Rf_asInteger(threads), &h));
    int32_t G = 0, S = 0, K = 0;
    ppcseq_prep_dims(h, &G, &S, &K);
    int64_t *gid = (int64_t *)R_alloc((size_t)G, sizeof(int64_t)), *sid = (int64_t *)R_alloc((size_t)S, sizeof(int64_t));
    int64_t *fr = (int64_t *)R_alloc((size_t)S, sizeof(int64_t));
    SEXP counts = PROTECT(Rf_allocMatrix(INTSXP, S, G));          /* column-major [S, G] == gene-major [G][S] */
    SEXP genes = PROTECT(Rf_allocVector(INTSXP, G)), samples = PROTECT(Rf_allocVector(INTSXP, S));
    SEXP first_row = PROTECT(Rf_allocVector(REALSXP, S));
    int rc = ppcseq_prep_fetch(h, gid, sid, fr, INTEGER(counts));
    ppcseq_prep_free(h);
    for (int g = 0; g < G; ++g) INTEGER(genes)[g] = (int)gid[g];
    for (int j = 0; j < S; ++j) { INTEGER(samples)[j] = (int)sid[j]; REAL(first_row)[j] = (double)fr[j] + 1.0; }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SET_VECTOR_ELT(out, 0, counts); SET_VECTOR_ELT(out, 1, genes); SET_VECTOR_ELT(out, 2, samples);
    SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(K)); SET_VECTOR_ELT(out, 4, first_row);
    UNPROTECT(5);
    check(rc);
    return out;
}

/* edgeR::calcNormFactors(method = "TMM") as called by calcNormFactor (R/tidybulk.R:262-323) on the dense counts [S, G];
 * level_order: 1-based column of `counts` for every level of factor(sample) (or NULL); ref: 1-based reference level or NA
 * (= first level with the largest median).  Returns list(nf, lib_size, reference level), all in level order. */
SEXP ppcseqb200_tmm_factors(SEXP counts, SEXP level_order, SEXP ref, SEXP threads) {
    if (!Rf_isInteger(counts) || !Rf_isMatrix(counts)) Rf_error("counts must be an integer matrix [S, G]");
    const int S = Rf_nrows(counts), G = Rf_ncols(counts);
    int32_t *ord = NULL;
    if (!Rf_isNull(level_order)) {
        if (LENGTH(level_order) != S) Rf_error("level_order must have one entry per sample");
        ord = (int32_t *)R_alloc((size_t)S, sizeof(int32_t));
        for (int j = 0; j < S; ++j) ord[j] = INTEGER(level_order)[j] - 1;
    }
    const int r = Rf_asInteger(ref);
    SEXP nf = PROTECT(Rf_allocVector(REALSXP, S)), lib = PROTECT(Rf_allocVector(REALSXP, S));
    int32_t ref_out = 0;
    int rc = ppcseq_tmm_factors(G, S, INTEGER(counts), ord, r == NA_INTEGER ? -1 : r - 1, Rf_asInteger(threads), REAL(nf),
                                REAL(lib), &ref_out);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, nf); SET_VECTOR_ELT(out, 1, lib); SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(ref_out + 1));
    UNPROTECT(3);
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
    {"ppcseqb200_prep_table", (DL_FUNC)&ppcseqb200_prep_table, 7},
    {"ppcseqb200_tmm_factors", (DL_FUNC)&ppcseqb200_tmm_factors, 4},
    {NULL, NULL, 0}};

/* same registration pattern as the reference's src/RcppExports.cpp:22-25 */
void R_init_ppcseq(DllInfo *dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
