// Single-process multi-GPU parent handle (multi.cu): entry points the C ABI dispatches to when a handle owns shards.
#pragma once
#include "model.h"

namespace ppcseq {

struct ShardPool;
void destroy_pool(ShardPool *p);
int multi_create(int G, int S, int C, int K, const int32_t *counts, const double *X, const double *exposure,
                 double lambda_mu_mu, int n_devices, const int32_t *devices, Model **out);
int multi_ensure_comm(Model *P, int channels, int cap);
int multi_set_exclusion(Model *P, const int32_t *pairs, long long n);
int multi_set_design_path(Model *P, int mode);
int multi_status(Model *P, int *flags);
int multi_log_prob_grad(Model *P, int B, const double *theta, int propto, int jacobian, double *lp, double *grad);
int multi_exposure_grad(Model *P, const double *theta, double *out);
int multi_flags(Model *P, const double *lower, const double *upper, const double *mean, const double *slope, uint8_t *ppc,
                uint8_t *deleterious, int32_t *failed, int32_t *tot_del);
int multi_fit_from_draws(Model *P, const double *theta_draws, int n, Fit **out);
int multi_fit_get_draws(const Fit *PF, long long begin, long long count, double *out);
int multi_fit_param_mean(const Fit *PF, long long begin, long long count, double *out);
int multi_sample_nuts(Model *P, const ppcseq_nuts_opts &o, Fit **out);
int multi_advi(Model *P, const ppcseq_advi_opts &o, Fit **out);
int multi_ppc_summary(Fit *PF, int exact, long long n_draws, double p, double tc, uint64_t seed, double *lower, double *upper,
                      double *mean, double *sd);
int multi_ppc_draws(Fit *PF, double tc, uint64_t seed, double *counts_rng);

}  // namespace ppcseq
