/* C restatement of the ppcseq Stan model's hot loop.  ORACLE / CPU BASELINE ONLY -- test
 * infrastructure, never linked into or called by the product (libppcseq_b200.so).
 *
 * PARITY STATUS: parity unpinned.  The reference holds no golden vectors for log_prob/grad and
 * cannot run in this image (no R, no Stan); this file is validated against the 40-digit mpmath
 * evaluation in oracle/model_mp.py (tests/test_oracle.py).
 *
 * What it follows in /root/reference:
 *   inst/stan/negBinomial_MPI.stan:58-120   lp_reduce: neg_binomial_2_log_lpmf over one shard minus
 *                                           the lpmf of the excluded points (:105-115)
 *   inst/stan/negBinomial_MPI.stan:200-206  sigma = exp(-sigma_raw), alpha, lambda_log_param = X*alpha
 *   inst/stan/negBinomial_MPI.stan:210-223  priors
 *   inst/stan/negBinomial_MPI.stan:226-240  sum(map_rect(lp_reduce, ...)) -- one worker per shard
 *   R/utilities.R:125-174 (:133-135)        genes dealt cyclically to shards: gene g -> shard g mod n
 * Stan gets the gradient by reverse-mode AD; here the partials are the closed forms of SURVEY.md 7.4.
 * Without the AD tape this is a *generous* (faster-than-rstan) stand-in for the reference CPU path.
 *
 * Build: make -C oracle   ->  oracle/libppcseq_oracle.so
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HALF_LOG_2PI 0.91893853320467274178
#define SQRT_2_OVER_PI 0.79788456080286535588

static double digamma_pos(double x) {
    /* psi(x), x > 0: upward recurrence to x >= 16, then the asymptotic series */
    double r = 0.0;
    while (x < 16.0) { r -= 1.0 / x; x += 1.0; }
    double w = 1.0 / (x * x);
    double s = w * (1.0 / 12 - w * (1.0 / 120 - w * (1.0 / 252 - w * (1.0 / 240 - w * (1.0 / 132 - w * (691.0 / 32760 - w / 12))))));
    return r + log(x) - 0.5 / x - s;
}

static double stirling_tail(double x) {
    double w = 1.0 / (x * x);
    return (1.0 / x) * (1.0 / 12 + w * (-1.0 / 360 + w * (1.0 / 1260 + w * (-1.0 / 1680 + w * (1.0 / 1188 + w * (-691.0 / 360360))))));
}

/* lgamma(n+phi) - lgamma(n+1) without the large-n cancellation */
static double lgamma_ratio(double n, double phi) {
    if (n < 32.0 || n + phi < 32.0) {
        int sg;
        return lgamma_r(n + phi, &sg) - lgamma_r(n + 1.0, &sg);
    }
    double y = n + 1.0, d = phi - 1.0, x = n + phi;
    return (y - 0.5) * log1p(d / y) + d * log(x) - d + stirling_tail(x) - stirling_tail(y);
}

/* erfcx(t) = exp(t^2) erfc(t) for the skew-normal term; t may be large positive */
static double erfcx_d(double t) {
    if (t < 10.0) return exp(t * t) * erfc(t);      /* t very negative -> +inf, handled by callers */
    double w = 1.0 / (2.0 * t * t), term = 1.0, s = 1.0;
    for (int k = 1; k <= 14; ++k) { term *= -(2.0 * k - 1.0) * w; s += term; }
    return s / (t * 1.7724538509055160273);
}

typedef struct {
    int G, S, C, K;
    const int32_t *counts;      /* [G,S] gene-major */
    const double *X;            /* [S,C] row-major  */
    const double *exposure;     /* [S] */
    const uint8_t *exclude;     /* [G,S] or NULL; 1 = dropped (:105-115) */
    const double *theta;
    double *grad;
    /* constrained hyper-parameters */
    double xi, om, a, sigma_slope, sigma_intercept, sigma_sigma;
    int o_intercept, o_alpha1, o_alpha2, o_sigma_raw;
    int shard, n_shards;
    /* per-shard outputs */
    double lp, g_xi, g_om, g_a, g_slope, g_icpt, g_ss;
} shard_t;

static void *shard_run(void *arg) {
    shard_t *w = (shard_t *)arg;
    const int G = w->G, S = w->S, C = w->C, K = w->K, R = C > 2 ? C - 2 : 0;
    double lp = 0, g_xi = 0, g_om = 0, g_a = 0, g_slope = 0, g_icpt = 0, g_ss = 0;
    double d_alpha[64];
    (void)G;
    for (int g = w->shard; g < w->G; g += w->n_shards) {      /* cyclic: R/utilities.R:133-135 */
        const double ic = w->theta[w->o_intercept + g];
        const double sr = w->theta[w->o_sigma_raw + g];
        const double phi = exp(-sr);
        const double log_phi = -sr;
        int sg;
        const double lg_phi = lgamma_r(phi, &sg), psi_phi = digamma_pos(phi);
        for (int c = 0; c < C; ++c) d_alpha[c] = 0.0;
        double d_phi = 0.0, ll = 0.0;
        const int32_t *row = w->counts + (size_t)g * S;
        const uint8_t *ex = w->exclude ? w->exclude + (size_t)g * S : NULL;
        for (int s = 0; s < S; ++s) {
            if (ex && ex[s]) continue;
            const double *xs = w->X + (size_t)s * C;
            double eta = w->exposure[s] + xs[0] * ic;
            if (g < K && C >= 2) {
                eta += xs[1] * w->theta[w->o_alpha1 + g];
                for (int r = 0; r < R; ++r) eta += xs[2 + r] * w->theta[w->o_alpha2 + (size_t)g * R + r];
            }
            const double n = (double)row[s];
            const double mu = exp(eta), a = mu + phi, log_a = log(a);
            ll += lgamma_ratio(n, phi) - lg_phi + n * eta + phi * log_phi - (n + phi) * log_a;
            const double de = n - (n + phi) * (mu / a);
            d_phi += (mu - n) / a + log_phi - log_a - psi_phi + digamma_pos(n + phi);
            for (int c = 0; c < C; ++c) d_alpha[c] += xs[c] * de;
        }
        lp += ll;
        /* intercept ~ skew_normal(xi, om, a)  (:219) */
        const double z = (ic - w->xi) / w->om;
        const double t = -w->a * z * 0.70710678118654752440;
        const double ecx = erfcx_d(t);
        const double log_erfc = (t < 10.0) ? log(erfc(t)) : log(ecx) - t * t;
        lp += -log(w->om) - 0.5 * z * z + log_erfc;
        const double ratio = isinf(ecx) ? 0.0 : SQRT_2_OVER_PI / ecx;
        const double dz = -z + w->a * ratio;
        double g_ic = d_alpha[0] + dz / w->om;
        g_xi += -dz / w->om;
        g_om += (-1.0 - dz * z) / w->om;
        g_a += ratio * z;
        /* sigma_raw ~ normal(slope*intercept + icpt, ss)  (:223) */
        const double m = w->sigma_slope * ic + w->sigma_intercept;
        const double e = (sr - m) / w->sigma_sigma;
        lp += -log(w->sigma_sigma) - 0.5 * e * e;
        const double g_m = e / w->sigma_sigma;
        g_ic += w->sigma_slope * g_m;
        g_slope += g_m * ic;
        g_icpt += g_m;
        g_ss += -1.0 / w->sigma_sigma + e * e / w->sigma_sigma;
        w->grad[w->o_intercept + g] = g_ic;
        w->grad[w->o_sigma_raw + g] = -phi * d_phi - g_m;
        if (g < K) {
            if (C >= 2) {                                         /* double_exponential(0,1) (:220) */
                const double a1 = w->theta[w->o_alpha1 + g];
                lp += -fabs(a1);
                w->grad[w->o_alpha1 + g] = d_alpha[1] - (a1 > 0 ? 1.0 : (a1 < 0 ? -1.0 : 0.0));
            }
            for (int r = 0; r < R; ++r) {                         /* normal(0,2.5) (:221) */
                const double a2 = w->theta[w->o_alpha2 + (size_t)g * R + r];
                lp += -a2 * a2 / 12.5;
                w->grad[w->o_alpha2 + (size_t)g * R + r] = d_alpha[2 + r] - a2 / 6.25;
            }
        }
    }
    w->lp = lp; w->g_xi = g_xi; w->g_om = g_om; w->g_a = g_a;
    w->g_slope = g_slope; w->g_icpt = g_icpt; w->g_ss = g_ss;
    return NULL;
}

int oracle_dim(int G, int K, int C) { return 6 + 2 * G + K + (C > 2 ? C - 2 : 0) * K; }

/* log_prob<propto,jacobian> and gradient; n_shards worker threads (map_rect with STAN_NUM_THREADS) */
int oracle_log_prob_grad(int G, int S, int C, int K, const int32_t *counts, const double *X,
                         const double *exposure, const uint8_t *exclude, double lambda_mu_mu,
                         const double *theta, int propto, int jacobian, int n_shards,
                         double *lp_out, double *grad) {
    if (C < 1 || C > 64 || K > G || n_shards < 1) return 1;
    const int R = C > 2 ? C - 2 : 0;
    const int o_intercept = 3, o_alpha1 = 3 + G, o_alpha2 = 3 + G + K, o_sigma_raw = 3 + G + K + R * K;
    const int o_tail = o_sigma_raw + G;
    const double L = lambda_mu_mu;
    const double lambda_mu = theta[0] + L, lambda_sigma = exp(theta[1]), lambda_skew = theta[2];
    const double sigma_slope = -exp(theta[o_tail]), sigma_intercept = theta[o_tail + 1];
    const double sigma_sigma = exp(theta[o_tail + 2]);
    if (n_shards > G) n_shards = G > 0 ? G : 1;
    memset(grad, 0, sizeof(double) * (size_t)(o_tail + 3));
    shard_t *w = (shard_t *)calloc((size_t)n_shards, sizeof(shard_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_shards, sizeof(pthread_t));
    for (int i = 0; i < n_shards; ++i) {
        w[i] = (shard_t){G, S, C, K, counts, X, exposure, exclude, theta, grad,
                         lambda_mu + L, lambda_sigma, lambda_skew, sigma_slope, sigma_intercept, sigma_sigma,
                         o_intercept, o_alpha1, o_alpha2, o_sigma_raw, i, n_shards,
                         0, 0, 0, 0, 0, 0, 0};
        if (n_shards > 1) pthread_create(&th[i], NULL, shard_run, &w[i]);
    }
    if (n_shards == 1) shard_run(&w[0]);
    double lp = 0, g_xi = 0, g_om = 0, g_a = 0, g_slope = 0, g_icpt = 0, g_ss = 0;
    for (int i = 0; i < n_shards; ++i) {
        if (n_shards > 1) pthread_join(th[i], NULL);
        lp += w[i].lp; g_xi += w[i].g_xi; g_om += w[i].g_om; g_a += w[i].g_a;
        g_slope += w[i].g_slope; g_icpt += w[i].g_icpt; g_ss += w[i].g_ss;
    }
    free(w); free(th);
    /* hyper-priors (:210-216) */
    lp += -(lambda_mu - L) * (lambda_mu - L) / 8 - lambda_sigma * lambda_sigma / 8 - lambda_skew * lambda_skew / 2
          - sigma_intercept * sigma_intercept / 8 - sigma_slope * sigma_slope / 8 - sigma_sigma * sigma_sigma / 8;
    if (!propto) {
        lp += 5 * (-HALF_LOG_2PI - log(2.0)) - HALF_LOG_2PI - 2.0 * G * HALF_LOG_2PI;
        if (C >= 2) lp += -K * log(2.0);
        if (C >= 3) lp += -(double)R * K * (HALF_LOG_2PI + log(2.5));
    }
    const double jac = jacobian ? 1.0 : 0.0;
    if (jacobian) lp += theta[1] + theta[o_tail] + theta[o_tail + 2];
    grad[0] = g_xi - (lambda_mu - L) / 4;
    grad[1] = (g_om - lambda_sigma / 4) * lambda_sigma + jac;
    grad[2] = g_a - lambda_skew;
    grad[o_tail] = (g_slope - sigma_slope / 4) * sigma_slope + jac;
    grad[o_tail + 1] = g_icpt - sigma_intercept / 4;
    grad[o_tail + 2] = (g_ss - sigma_sigma / 4) * sigma_sigma + jac;
    *lp_out = lp;
    return 0;
}
