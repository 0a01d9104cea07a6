// Device-resident inference drivers (the stand-ins for rstan::vb and rstan::sampling).
#pragma once
#include "common.cuh"
#include "model.h"

namespace ppcseq {

// all-reduce hook for gene-sharded runs: sums `n` doubles in place at `d_buf` across ranks, enqueued on
// `stream`; must give bitwise-identical results on every rank.  nullptr = single rank.
typedef int (*allreduce_fn)(void *ctx, double *d_buf, int n, void *stream);

struct EvalCtx {
    Model *M;
    cudaStream_t st;
    allreduce_fn allreduce;
    void *ar_ctx;
    double *d_comm;       // [B][8] partials buffer owned by the caller of the hook (may be user memory)
    long long n_evals = 0;
    // log_prob + grad of B thetas; handles the sharded path (partial -> all-reduce -> finalize)
    int eval(int B, const double *d_theta, int propto, int jacobian, double *d_lp, double *d_grad);
};

}  // namespace ppcseq
