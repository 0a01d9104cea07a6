// Host-side model handle (opaque `ppcseq_model` of the C ABI).
#pragma once
#include <functional>
#include <memory>
#include <vector>

#include "common.cuh"
#include "lp_grad.h"

namespace ppcseq {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct Model {
    int device = 0;
    ModelDev m{};
    cudaStream_t stream = nullptr;
    int G_total = 0, K_total = 0, g_begin = 0;   // shard placement (single rank: 0..G)
    int n_groups_detected = 0;                    // distinct design rows (0 = more than 8)
    // data (HBM-resident for the life of the handle)
    int32_t *d_counts = nullptr;
    double *d_Xt = nullptr, *d_exposure = nullptr, *d_gconst = nullptr, *d_Xg = nullptr;
    uint8_t *d_gflags = nullptr;
    uint32_t *d_mask = nullptr;
    int *d_perm_pos = nullptr;
    int32_t *d_excl_pairs = nullptr;            // device copy of the current exclusion list
    long long n_excl = 0;
    int32_t *d_counts_p = nullptr;               // group-sorted, padded copy for the categorical path
    double *d_exp_exposure_p = nullptr;
    void *d_log_tab = nullptr, *d_log_tab512 = nullptr;
    // Chebyshev-moment path
    double *d_rec = nullptr, *d_mom_1 = nullptr, *d_Tz = nullptr;
    size_t rec_doubles = 0;
    int *d_excl_off = nullptr;                   // exclusion list by (gene, design row): offsets and exp(exposure)
    double *d_excl_E = nullptr;
    uint8_t *d_excl_r = nullptr;
    std::vector<double> h_exp_exposure;          // exp(exposure_rate[s]), original sample order
    std::vector<int> h_grp;                      // design row of sample s (categorical designs)
    std::vector<double> h_Xg;                    // the distinct design rows [8][C] (host copy)
    std::vector<int> h_mgrp;                     // moment group (design row x exposure bin) of sample s
    double *d_mom_Eg = nullptr, *d_mom_Xg = nullptr;
    double *d_xg_partial = nullptr;              // scratch of the optional exposure-gradient output (exposure_grad.cu)
    size_t xg_cap = 0;
    uint8_t *d_mflags = nullptr;
    double *d_mconst = nullptr;
    int mom_J_detected = 0;                      // 0 = not eligible (design not categorical or exposure range too wide)
    int design_mode = 0;                         // ppcseq_model_set_design_path: 0 auto, 1 general, 2 per-element, 3 moments
    std::vector<double> hX;                      // host copy of the model.matrix (S x C), for flags
    std::vector<int> perm_pos;                   // original sample s -> position in the padded row
    // per-evaluation scratch, sized for Bcap simultaneous thetas
    int Bcap = 0;
    double *d_block_scratch = nullptr, *d_lp = nullptr, *d_theta = nullptr, *d_grad = nullptr, *d_partials = nullptr;
    unsigned int *d_counters = nullptr;

    // fused peer all-reduce (ppcseq_comm_create / _connect): mailbox owned by this rank + peers' mapped mailboxes
    PeerComm comm;
    void *d_mailbox = nullptr;                   // [slots | flags | error]
    size_t mailbox_bytes = 0;
    std::vector<void *> peer_mailboxes;          // cudaIpcOpenMemHandle results (nullptr for self)
    std::vector<unsigned long long> chan_seq;    // next sequence number per channel
    CommCall next_comm_call(int channel);

    // host-pointer batches (ppcseq_log_prob_grad, B > 1): copy streams + per-theta events of the three-stage pipeline
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done;
    int ensure_pipeline(int B);

    // device-side failure flags (ModelDev::status): int[2] in mapped pinned host memory, [0] peer all-reduce time-out,
    // [1] grid-reduction time-out.  check_status() is called at every host synchronisation point of the library
    // (host-pointer log_prob, the samplers' scalar fetches): a raised flag is fatal (PPCSEQ_ECOMM) and sticky.
    int *h_status = nullptr;
    int check_status() const;

    int ensure_batch(int B);

    // ---- single-process multi-GPU parent (ppcseq_model_create_multi, multi.cu) -------------------------------------
    // A parent owns one shard Model per device (contiguous gene blocks, peer mailboxes wired by direct peer access) and
    // presents the GLOBAL problem: m.G/K/D and the o_* offsets are the global ones, theta / gradients / fit queries use
    // the global layout.  It holds no device data of its own.
    std::vector<double> h_stage_theta, h_stage_grad, h_stage_lp;   // a shard's host staging for parent-level calls
    std::vector<Model *> shards;
    std::vector<int> shard_g0;                   // first global gene of every shard (+ G at the end)
    struct ShardPool *pool = nullptr;            // one persistent host thread per shard
    // set on a shard by its parent around a sampler run: a rendez-vous of all shard threads between "every buffer is
    // allocated" and "the first peer-waiting kernel is launched" (allocations of pinned host memory or peer-mapped
    // device memory may synchronise OTHER devices of the process; a kernel spinning there for this shard's next launch
    // would close a wait cycle)
    std::function<void()> pre_run_barrier;
    bool is_multi() const { return !shards.empty(); }
    ~Model();
};

// hits Model::pre_run_barrier exactly once on every path out of a sampler driver (an early error return included, so
// that the other shards' threads are never left waiting)
struct PreRunBarrier {
    Model *M;
    bool done = false;
    explicit PreRunBarrier(Model *m) : M(m) {}
    void hit() {
        if (done) return;
        done = true;
        if (M->pre_run_barrier) M->pre_run_barrier();
    }
    ~PreRunBarrier() { hit(); }
};

int comm_alloc(Model *M, int rank, int world, int channels, int cap);
int comm_attach(Model *M, void *const *bases);
void comm_release(Model *M);
// creation of a (shard of a) model on one device (capi.cu)
int create_impl(int G_total, int K_total, int g_begin, int g_end, int S, int C, const int32_t *counts, const double *X,
                const double *exposure, double lambda_mu_mu, int device, Model **out);

// Posterior draws of the unconstrained vector, device-resident, parameter-major [D][ld]
// (the stand-in for the stanfit object that rstan::sampling / rstan::vb return).
struct Fit {
    Model *model = nullptr;
    std::vector<Fit *> shard_fits;               // multi-GPU parent fit: one fit per shard model (owned)
    int n_draws = 0, ld = 0;
    double *d_draws_T = nullptr;
    // sampler diagnostics (host)
    std::vector<double> info;
    ~Fit();
};

}  // namespace ppcseq
