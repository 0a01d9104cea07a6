// Host-side launch interface of ppc.cu
#pragma once
#include "common.cuh"

namespace ppcseq {

int launch_summary_matrix(const double *d_draws, int n, int m, double p, double *lower, double *upper, double *mean,
                          double *sd, int *d_bad, cudaStream_t st);
int launch_flags(const int32_t *counts, int counts_stride, int K, int S, const double *lower, const double *upper,
                 const double *mean, const double *slope, const uint8_t *group_right, int has_covariate, uint8_t *ppc,
                 uint8_t *deleterious, int32_t *failed, int32_t *tot_del, cudaStream_t st);

struct PpcArgs;
int ppc_tail_sizes(long long n, double p, int *m_lo, int *m_hi);
int launch_ppc_stream_full(const ModelDev &m, const double *draws_T, int n_post, int ld, int supersample, long long n_draws,
                           double p, double tc, uint64_t seed, int m_lo, int m_hi, double *lower, double *upper,
                           double *mean, double *sd, double *raw, unsigned int *overflow, cudaStream_t st, int skip_summary = 0,
                           long long pair_base = 0);
int launch_transpose_draws(const double *in, int n, long long D, double *out, int ld, cudaStream_t st);
int launch_param_mean(const double *draws_T, int ld, int n, long long begin, long long count, double *out, cudaStream_t st);

}  // namespace ppcseq
