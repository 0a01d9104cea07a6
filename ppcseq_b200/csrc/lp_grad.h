// Host-side launch interface of lp_grad.cu
#pragma once
#include "common.cuh"

namespace ppcseq {

struct LpGradArgs;
int lp_grad_num_blocks(const ModelDev &m);
// per-theta scratch of the two-level grid reduction: one 8-double slot per CTA plus one per group of 32 CTAs,
// and one arrival counter per group plus the top-level one
// counters per theta: [top-level arrivals, epoch (sequence number of the last launch), one per group]; scratch per theta:
// one cell per CTA then one per group, a cell = 8 lines of 16 bytes = 16 doubles (2 x kNumPartials)
inline size_t lp_grad_counter_slots(const ModelDev &m) { return (size_t)lp_grad_num_blocks(m) / 32 + 3; }
inline size_t lp_grad_scratch_slots(const ModelDev &m) { return 2 * ((size_t)lp_grad_num_blocks(m) + lp_grad_counter_slots(m)); }
// single-rank (finalize=1: lp[B] and complete gradient) or shard mode (finalize=0: partials[B][8])
// comm (optional): fused peer all-reduce of the partial sums inside the kernel (gene shards on several GPUs)
struct CommCall {
    const PeerComm *comm = nullptr;
    int channel = 0;
    unsigned long long seq = 0;
};
// per-theta buffers of a batched launch (the chains of a sampler), B <= 8
struct LpGradTab {
    const double *theta[8] = {};
    double *grad[8] = {};
    double *lp[8] = {};
    const double *skip[8] = {};
};
int launch_lp_grad_full(const ModelDev &m, int B, const double *theta, double *grad, double *lp, double *partials,
                        unsigned int *counters, double *block_scratch, int propto, int jacobian, int finalize,
                        cudaStream_t st, CommCall cc = CommCall(), const double *skip = nullptr,
                        const LpGradTab *tab = nullptr);
int launch_finalize_hyper(const ModelDev &m, int B, const double *theta, const double *partials, int propto,
                          int jacobian, double *lp, double *grad, cudaStream_t st);
int launch_scatter_sentinel(const ModelDev &m, int32_t *counts_p, const int *perm_pos, const int32_t *pairs,
                            long long n, int restore, cudaStream_t st);
// Chebyshev-moment path (lp_grad_mom.cu)
int launch_lp_grad_mom(const LpGradArgs &a, int B, cudaStream_t st);
int mom_record_slots(int n_groups, int J, int xm = 0);
int mom_tile_genes();
int launch_moments(const ModelDev &m, const double *Tz, double *rec, uint8_t *mflags, double *mconst, cudaStream_t st);
int launch_gene_consts(const ModelDev &m, double *gconst, uint8_t *gflags, cudaStream_t st);
// load every kernel an evaluation / a sampler step of a C-column model can launch onto the current device (no lazy
// module loading left once peer-waiting kernels are in flight)
int preload_lp_grad_kernels(int C);
int preload_mom_kernels(int C);
int preload_sampler_kernels();

}  // namespace ppcseq
