// Shared declarations for the ppcseq_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/ppcseq_b200.h"

namespace ppcseq {

void set_error(const std::string &msg);
extern std::atomic<long long> g_launches;

#define PPCSEQ_CUDA(call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ::ppcseq::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));             \
            return PPCSEQ_ECUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define PPCSEQ_CHECK_LAUNCH()                                                                     \
    do {                                                                                          \
        ::ppcseq::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
        PPCSEQ_CUDA(cudaGetLastError());                                                          \
    } while (0)

constexpr int kMaxC = 8;          // design-matrix columns supported by the fused kernel
constexpr int kNumPartials = 8;   // [lp, d_xi, d_omega, d_skew, d_slope, d_sig_icpt, d_sig_sigma, pad]

// Everything the kernels need about one (shard of a) model; passed by value.
struct ModelDev {
    int G, S, C, K;               // local genes, samples, design columns, local checked genes
    int R;                        // max(0, C-2)
    int W;                        // 32-bit mask words per gene row = ceil(S/32)
    long long D;                  // local unconstrained dimension
    int o_intercept, o_alpha1, o_alpha2, o_sigma_raw, o_tail;
    double lambda_mu_mu;
    const int32_t *counts;        // [G][S]
    const double *Xt;             // [C][S]   (column-major copy of the model.matrix)
    const double *exposure;       // [S]
    const uint32_t *mask;         // [G][W] bit s%32 of word s/32 set = excluded; nullptr in pass 1
    const double *gconst;         // [(3+C)][G]: S_eff, sum n*exposure, sum lgamma(n+1), sum n*X[:,c]
    const void *log_tab;          // LogTabEntry[128] (nb_math.cuh)
    const uint8_t *gflags;        // [G] bit0: the gene has a count < 32 (needs the small-count table)
    // categorical-design fast path (<= 8 distinct rows of X): samples sorted by design row, every
    // group padded to a multiple of 32 so that one warp iteration never straddles two groups.
    int n_groups;                 // 0 = general path
    int S_pad;                    // padded row length = 32 * grp_chunk_begin[n_groups]
    int grp_chunk_begin[9];       // first 32-sample chunk of each group (prefix sums)
    int grp_size[8];              // true number of samples per group
    const int32_t *counts_p;      // [G][S_pad] permuted + padded counts; -1 = padding or pass-2 excluded
    const double *exp_exposure_p; // [S_pad] exp(exposure_rate) in permuted order (pad = 1)
    const double *Xg;             // [8][C] the distinct design rows
};

}  // namespace ppcseq
