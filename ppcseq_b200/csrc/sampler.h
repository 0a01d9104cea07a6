// Device-resident inference drivers: the stand-ins for rstan::sampling (NUTS, diag_e, windowed
// adaptation) and rstan::vb (mean-field ADVI) that the reference calls in do_inference()
// (/root/reference/R/utilities.R:1482-1513, :246-278).  All D-length state lives in HBM; the host
// only steers (tree bookkeeping, dual averaging, convergence checks) from a handful of scalars per step.
#pragma once
#include <cstdint>

#include "common.cuh"
#include "model.h"

namespace ppcseq {

// all-reduce hook for gene-sharded runs: sums `n` doubles in place at `d_buf` across ranks, enqueued on
// `stream`; must give bitwise-identical results on every rank.  nullptr = single rank.
typedef int (*allreduce_fn)(void *ctx, double *d_buf, int n, void *stream);

// One evaluation context = one stream + its own reduction scratch, so that several chains can have
// log_prob kernels in flight at the same time on one model.
struct EvalCtx {
    Model *M = nullptr;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    int channel = 0;                 // comm channel of this context (gene-sharded runs)
    allreduce_fn allreduce = nullptr;
    void *ar_ctx = nullptr;
    int Bcap = 0;
    double *d_block_scratch = nullptr, *d_partials = nullptr;
    unsigned int *d_counters = nullptr;
    long long n_evals = 0;

    int init(Model *model, int B, bool make_stream);
    // log_prob + grad of B thetas ([B][D]); handles the sharded path (partial -> all-reduce -> finalize)
    int eval(int B, const double *d_theta, int propto, int jacobian, double *d_lp, double *d_grad);
    void destroy();
};

// ---- vector kernels (sampler_kernels.cu) --------------------------------------------------------
// scratch: per-context reduction scratch of kRedBlocks*kRedMax doubles + one counter
constexpr int kRedBlocks = 148 * 2;
constexpr int kRedMax = 8;
struct RedScratch {
    double *partials = nullptr;      // [kRedBlocks][kRedMax]
    unsigned int *counter = nullptr;
    // gene-sharded runs (comm.world > 1): every global sum ends with the fused peer all-reduce on (channel, seq);
    // ranks > 0 leave the 6 replicated hyper-parameters out of the sums so that they are counted once
    PeerComm comm;
    int channel = 0;
    unsigned long long seq = 0;
    int skip_hyper = 0;
    long long o_tail = 0;            // local index of the first of the 3 trailing hyper-parameters
    int alloc();
    void free_();
};

// identity of a local parameter across ranks, for counter-based RNG streams: the replicated hyper-parameters get
// the same id on every rank (=> identical momenta / draws), gene-level parameters a rank-unique one
struct ParamIds {
    long long o_tail = 0;
    unsigned long long gene_base = 0;          // (g_begin + 1) << 32
    __host__ __device__ unsigned long long id(long long i) const {
        if (i < 3) return 0x8000000000000000ull | (unsigned long long)i;
        if (i >= o_tail) return 0x8000000000000000ull | (unsigned long long)(3 + i - o_tail);
        return gene_base + (unsigned long long)i;
    }
};

// p = z / sqrt(inv_metric), z ~ N(0,1); out[0] = 1/2 sum z^2 (the kinetic energy)
int launch_sample_p(double *p, const double *inv_metric, long long n, uint64_t seed, uint64_t stream_id,
                    uint64_t counter, ParamIds ids, RedScratch rs, double *out, cudaStream_t st);
// p += eps/2 * grad;  q += eps * inv_metric * p
int launch_leap_a(double *q, double *p, const double *grad, const double *inv_metric, double eps, long long n,
                  cudaStream_t st);
// second half of a leapfrog step fused with the depth-0 case of the NUTS tree: p += eps/2 * grad, then
// rho = p_beg = p_end = p and z_propose = (q, grad) for whichever outputs are non-null; out[0] = 1/2 p' M^-1 p
struct LeapOut {
    double *rho = nullptr, *p_beg = nullptr, *p_end = nullptr, *zq = nullptr, *zg = nullptr;
    const double *q = nullptr;
};
int launch_leap_b(double *p, const double *grad, const double *inv_metric, double eps, LeapOut lo, long long n,
                  RedScratch rs, double *out, cudaStream_t st);
struct BcastDst { double *dst[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; };
int launch_bcast(const double *src, long long n, BcastDst d, cudaStream_t st);
// rho_out = rho_init + rho_final and the six U-turn dot products (Stan base_nuts::compute_criterion x 3), rs = ri + rf:
// out[0..5] = <M^-1 p_beg, rs>, <M^-1 p_end, rs>, <M^-1 p_beg, ri + p_final_beg>, <M^-1 p_final_beg, ri + p_final_beg>,
//             <M^-1 p_init_end, rf + p_init_end>, <M^-1 p_end, rf + p_init_end>
int launch_merge(double *rho_out, const double *rho_init, const double *rho_final, const double *p_beg,
                 const double *p_end, const double *p_init_end, const double *p_final_beg, const double *inv_metric,
                 long long n, RedScratch rs, double *out, cudaStream_t st);
int launch_welford_add(double *mean, double *m2, const double *q, double n_samples_after, long long n, cudaStream_t st);
int launch_welford_finish(const double *m2, double n_samples, double *inv_metric, long long n, cudaStream_t st);
int launch_fill(double *x, double v, long long n, cudaStream_t st);
int launch_store_draw(double *draws_T, int ld, int col, const double *q, long long n, cudaStream_t st);
// ADVI
int launch_advi_draw(const double *mu, const double *omega, double *eta, double *zeta, long long D, int B, uint64_t seed,
                     uint64_t counter, ParamIds ids, cudaStream_t st);
int launch_advi_update(double *mu, double *omega, const double *grad, const double *eta, double *hist_mu,
                       double *hist_omega, long long D, int B, double eta_scaled, int first, int *d_bad, cudaStream_t st);
int launch_advi_output(const double *mu, const double *omega, double *draws_T, int ld, int n, long long D, uint64_t seed,
                       ParamIds ids, cudaStream_t st);
int launch_sum(const double *x, long long n, RedScratch rs, double *out, cudaStream_t st);

// ---- drivers -----------------------------------------------------------------------------------
int run_nuts(Model *M, const ppcseq_nuts_opts &o, Fit **out);
int run_advi(Model *M, const ppcseq_advi_opts &o, Fit **out);

}  // namespace ppcseq
