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

    // make_stream: own stream; else `shared` if given, else the model's stream
    int init(Model *model, int B, bool make_stream, cudaStream_t shared = nullptr);
    // log_prob + grad of B thetas ([B][D]); handles the sharded path (partial -> all-reduce -> finalize)
    // skip: optional device flag, non-zero => the launch is a no-op (see LpGradArgs::skip)
    int eval(int B, const double *d_theta, int propto, int jacobian, double *d_lp, double *d_grad,
             const double *skip = nullptr);
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
    int channel = 0, entry = 0;      // mailbox cell of the exchange (entry = chain in the batched-chain driver)
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

// ---- device-side bookkeeping of one NUTS transition --------------------------------------------------------------
// Everything Stan's base_nuts::transition / build_tree decide from scalars (energy error, divergence, multinomial
// weights and acceptances, the U-turn criteria) is decided ON THE DEVICE by the last block of the kernel that produces
// the scalars, in a small state vector per chain.  The host enqueues a whole subtree (2^depth leapfrogs and their
// merges) without reading anything back: once the state's STOP flag is up (divergence, or a U-turn inside the
// subtree) every later kernel of the subtree returns at once, and the host looks at the state once per tree doubling.
enum : int {
    TS_H0 = 0,          // Hamiltonian at the start of the transition
    TS_LSW,             // log sum of weights of the trajectory (Stan: log_sum_weight)
    TS_METRO,           // sum of min(1, exp(H0 - h)) over the leapfrogs (accept_stat numerator)
    TS_NLEAP,           // leapfrog steps actually taken
    TS_DIV,             // divergent transition
    TS_STOP,            // the subtree being built is invalid: remaining kernels skip themselves
    TS_PERSIST,         // top-level U-turn criterion after the last completed doubling (1 = keep going)
    TS_SPARE,
    TS_VPROP = 8,       // potential of proposal k: 0 = z_sample, 1 = z_propose, 2 + d = the level-d right proposal
    TS_ACC = 40,        // log-sum-weight accumulators: 0 = the subtree of the current doubling, 2d-1 / 2d = left / right
    TS_SIZE = 104       //                              half of a level-d subtree (d <= 20)
};
struct LeapBook {       // depth-0 case of build_tree, folded into the second half-step kernel
    double *ts = nullptr;
    const double *lp = nullptr;            // log_prob written by the evaluation that precedes the kernel on the stream
    int acc_id = 0, prop_id = -1;
    unsigned long long reset_mask = 0;     // accumulators that start a new subtree with this leapfrog
};
struct MergeBook {      // subtree merge (top = 1: the trajectory-level merge of base_nuts::transition)
    double *ts = nullptr;
    int acc_init = 0, acc_final = 0, acc_parent = 0, prop_dst = 0, prop_src = 0, top = 0;
    uint64_t seed = 0, tctr = 0;
    uint32_t chain = 0, node = 0;
    double *zq_dst = nullptr, *zg_dst = nullptr;
    const double *zq_src = nullptr, *zg_src = nullptr;
};
// H0 = V + kinetic, every accumulator back to -inf / 0, V of the current sample
int launch_tree_init(double *ts, const double *kinetic, double V, cudaStream_t st);

// p = z / sqrt(inv_metric), z ~ N(0,1); out[0] = 1/2 sum z^2 (the kinetic energy)
int launch_sample_p(double *p, const double *inv_metric, long long n, uint64_t seed, uint64_t stream_id,
                    uint64_t counter, ParamIds ids, RedScratch rs, double *out, cudaStream_t st);
// p += eps/2 * grad;  q += eps * inv_metric * p
int launch_leap_a(double *q, double *p, const double *grad, const double *inv_metric, double eps, long long n,
                  cudaStream_t st, const double *skip = nullptr);
// second half of a leapfrog step fused with the depth-0 case of the NUTS tree: p += eps/2 * grad, then
// rho = p_beg = p_end = p and z_propose = (q, grad) for whichever outputs are non-null; out[0] = 1/2 p' M^-1 p
struct LeapOut {
    double *rho = nullptr, *p_beg = nullptr, *p_end = nullptr, *zq = nullptr, *zg = nullptr;
    const double *q = nullptr;
    // fuse the FIRST half of the next leapfrog of the same trajectory end into this kernel (p += eps/2 grad once more,
    // q += eps M^-1 p): saves a launch and a pass over p, q, grad, M^-1 per leaf.  q_next = the position vector to advance.
    double *q_next = nullptr;
};
int launch_leap_b(double *p, const double *grad, const double *inv_metric, double eps, LeapOut lo, long long n,
                  RedScratch rs, double *out, cudaStream_t st, LeapBook book = LeapBook());
struct BcastDst { double *dst[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; };
int launch_bcast(const double *src, long long n, BcastDst d, cudaStream_t st);
// rho_out = rho_init + rho_final and the six U-turn dot products (Stan base_nuts::compute_criterion x 3), rs = ri + rf:
// out[0..5] = <M^-1 p_beg, rs>, <M^-1 p_end, rs>, <M^-1 p_beg, ri + p_final_beg>, <M^-1 p_final_beg, ri + p_final_beg>,
//             <M^-1 p_init_end, rf + p_init_end>, <M^-1 p_end, rf + p_init_end>
int launch_merge(double *rho_out, const double *rho_init, const double *rho_final, const double *p_beg,
                 const double *p_end, const double *p_init_end, const double *p_final_beg, const double *inv_metric,
                 long long n, RedScratch rs, double *out, cudaStream_t st, MergeBook book = MergeBook());
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

// ---- batched chains (gene-sharded NUTS): one launch covers the same tree step of every chain (grid.y = chain) -----
// Every chain owns its buffers, its step size, its tree state and its reduction scratch; the tree STRUCTURE of a
// doubling (which leaf, which merge, which accumulators) is common, so the kernel parameters below are per-chain
// pointer tables plus common indices.  The cross-GPU exchange of a reduction uses mailbox entry = chain.
constexpr int kMaxBatch = 8;
struct PtrTab { double *p[kMaxBatch] = {}; };
struct CPtrTab { const double *p[kMaxBatch] = {}; };
struct BatchRed {
    double *partials[kMaxBatch] = {};
    unsigned int *counter[kMaxBatch] = {};
    PeerComm comm;
    int channel = 0, skip_hyper = 0;
    long long o_tail = 0;
};
struct LeapOutB { PtrTab rho, p_beg, p_end, zq, zg, q_next; CPtrTab q; int has_prop = 0, fuse_next = 0; };
struct LeapBookB { PtrTab ts; CPtrTab lp; int acc_id = 0, prop_id = -1; unsigned long long reset_mask = 0; };
struct MergeBookB {
    PtrTab ts, zq_dst, zg_dst;
    CPtrTab zq_src, zg_src;
    int acc_init = 0, acc_final = 0, acc_parent = 0, prop_dst = 0, prop_src = 0, top = 0;
    uint64_t seed = 0, tctr = 0;
    uint32_t node = 0;
};
struct EpsTab { double e[kMaxBatch] = {}; };
int launch_leap_a_batched(int B, PtrTab q, PtrTab p, CPtrTab grad, CPtrTab inv_metric, EpsTab eps, long long n, CPtrTab skip,
                          cudaStream_t st);
int launch_leap_b_batched(int B, PtrTab p, CPtrTab grad, CPtrTab inv_metric, EpsTab eps, LeapOutB lo, long long n, BatchRed rs,
                          PtrTab out, LeapBookB book, cudaStream_t st);
int launch_merge_batched(int B, PtrTab rho_out, CPtrTab rho_init, CPtrTab rho_final, CPtrTab p_beg, CPtrTab p_end,
                         CPtrTab p_init_end, CPtrTab p_final_beg, CPtrTab inv_metric, long long n, BatchRed rs, PtrTab out,
                         MergeBookB book, cudaStream_t st);

// ---- drivers -----------------------------------------------------------------------------------
int run_nuts(Model *M, const ppcseq_nuts_opts &o, Fit **out);
int run_nuts_batched(Model *M, const ppcseq_nuts_opts &o, Fit **out);       // nuts_batched.cu
int run_advi(Model *M, const ppcseq_advi_opts &o, Fit **out);

}  // namespace ppcseq
