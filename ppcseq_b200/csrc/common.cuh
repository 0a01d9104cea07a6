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

// One-shot all-reduce over peer-mapped memory (NVLink / NVSwitch), used INSIDE the kernels that produce the
// values: every rank owns a mailbox; the last CTA of a reduction writes its partial sums into slot [rank] of every
// peer's mailbox (remote stores), publishes a sequence number, waits for the W sequence numbers in its own mailbox
// and adds the W slots in rank order -- a fixed order, so every rank obtains bitwise the same sum.  Two parities
// alternate so that a fast rank never overwrites a slot a slow rank is still reading.
constexpr int kCommMaxWorld = 8;
constexpr int kCommSlot = 8;                       // doubles per slot
struct PeerComm {
    int world = 1, rank = 0;                       // world == 1: no exchange
    int channels = 0, cap = 0;                     // mailbox geometry: [2 parities][channels][cap entries][world][kCommSlot]
    double *slots[kCommMaxWorld] = {};             // slot arrays of every rank (own one included), peer-mapped
    unsigned long long *flags[kCommMaxWorld] = {}; // [2][channels][cap][world] sequence numbers
    uint4 *ll[kCommMaxWorld] = {};                 // same cells as 16-byte {lo, seq, hi, seq} lines, kCommSlot per cell:
                                                   // data and flag travel in one store (lp_grad kernels)
    unsigned long long *exec_seq = nullptr;        // [channels][cap], THIS rank's memory: all-reduces executed so far per
                                                   // (channel, entry).  The sequence number of an exchange is taken from
                                                   // here on the device, not from the host: kernels that skip themselves
                                                   // (NUTS subtrees cut short) leave no gaps, so the two-parity argument
                                                   // holds whatever the host enqueued
    int *error = nullptr;                          // the model's status word (ModelDev::status): kStatusPeerTimeout is
                                                   // OR-ed in when a wait times out (ranks out of step)
};
// Device-side failure flags of a model (ModelDev::status, a word in mapped pinned host memory: kernels write it only
// when something went wrong, the host reads it without a copy at every synchronisation point).  A flagged evaluation
// also carries NaN in lp, so it can never be mistaken for a result.
constexpr int kStatusPeerTimeout = 1;              // fused cross-GPU all-reduce: a peer's line did not arrive within ~2 s
constexpr int kStatusReduceTimeout = 2;            // grid reduction: a CTA's partial sums did not arrive within ~1 s

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
    const double *gconst;         // [(5+C)][G]: S_eff, sum n*exposure, sum lgamma(n+1), sum n*X[:,c], #(n>=32), sum_{n>=32} n
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
    // Chebyshev-moment path (lp_grad_mom.cu): the mu-dependent part of the likelihood from per-(gene, design row)
    // moments of the counts in T_j(z_s), z_s = (exp(exposure_s) - E_c) / E_hw.  mom_J = 0 disables the path.
    int mom_J;                    // series length (j = 0..mom_J)
    // Moment groups = (design row, exposure bin): the samples of a design row are sorted by exposure in the permuted
    // layout and cut into bins of bounded exposure ratio, each with its own centre / half-width, so that the series
    // stays short (J <= ~31) for ANY exposure range (piecewise Chebyshev).  One bin per row when the range is narrow.
    int mom_ng;                   // number of moment groups (<= kMomMaxGroups)
    int mom_xm;                   // 1: pass-2 exclusions through per-gene T_j moments of the excluded points (heavy lists),
                                  // stored next to the count moments (row 2j = count moment j, row 2j+1 = excluded T_j
                                  // moment j); 0: through the per-point correction list (excl_off / excl_E / excl_r)
    int mom_begin[17];            // group r covers permuted sample positions [mom_begin[r], mom_end[r])
    int mom_end[16];
    const double *mom_Eg;         // [kMomMaxGroups][4]: E_c, E_hw, E_min, E_max of every group
    const double *mom_Xg;         // [kMomMaxGroups][C]: design row of every group
    // data-only record of the moment kernel: [tile = g / 16][rec_slots rows][32 lanes] doubles, zero padded; a row
    // carries one double per lane: lanes 0-15 the "half 0" data of the tile's 16 genes, lanes 16-31 the "half 1" data.
    // ("n_groups" below = mom_ng, "design row r" = moment group r)
    // rows: 8 x small-count tail counts (4 x u16: #{s not excluded: k < n_s < 64} for k = s, s+16, s+32, s+48; half h
    // holds s = row + 8 h), ceil(n_groups / 2) x mom_J1p x count moments sum_{s in r, not excluded} n_s T_j(z_s) / max(j, 1)
    // of design row r = 2 pair + h in descending order j, 16 x Taylor coefficients P_k = sum_{n_s >= 64} psi^(k-1)(n_s) / k!
    // in descending order (half 0 the even powers, half 1 the odd ones).
    const double *rec;
    int rec_slots;                // 8 + ceil(n_groups / 2) * mom_J1p + 16 (a multiple of 8: the kernel streams 8-row batches)
    int mom_J1p;                  // mom_J + 1 rounded up to a multiple of 8
    const double *mom_1;          // [8][mom_J1p]: sum_{s in r} T_j(z_s) / max(j, 1)  (every sample of the design row; zero padded)
    const int *excl_off;          // [G + 1]: first entry of gene g in excl_E / excl_r; nullptr in pass 1
    const double *excl_E;         // exp(exposure) of the excluded points, sorted by gene
    const uint8_t *excl_r;        // their design rows
    const uint8_t *mflags;        // [G] bit0: some count < 64, bit1: no count >= 64
    const double *mconst;         // [4][G]: #(n >= 64), sum_{n >= 64} n, min_{n >= 64} n, sum lgamma(n+1) - sum_{n >= 64} lgamma(n)
    const void *log_tab_mom;      // LogTabEntry[kMomLogTab] for the moment kernel
    int *status;                  // device-side failure flags, int[2] in mapped pinned host memory: [0] peer time-out, [1] reduction time-out
};
constexpr int kSerK = 26;         // Taylor terms of sum_s lgamma(n_s + phi) about phi = 0, valid for phi <= kSerRatio * min n
constexpr double kSerRatio = 0.2; // (0.2^27 / 27 < 1e-20)
constexpr int kMomLogTab = 256;   // c_i = 1 + (i + 1/2)/256
constexpr int kMomJCap = 48;      // longest supported series (per exposure bin)
constexpr int kMomJTarget = 31;   // bins are added until the series is at most this long (J1p <= 32)
constexpr int kMomMaxGroups = 16; // design rows x exposure bins
constexpr int kMomXmPerGene = 12; // more excluded points per gene AND group pair than this on average: exclusion by per-gene T_j moments

}  // namespace ppcseq
