// Single-process multi-GPU: one handle, gene shards on several GPUs of one box.
//
// The reference's compiled model lives in ONE R process (R/stanmodels.R:10-25, src/RcppExports.cpp:15-25) and
// shards the likelihood over map_rect workers inside it (inst/stan/negBinomial_MPI.stan:226-240).  This file gives the
// C ABI the same shape: ppcseq_model_create_multi builds one shard Model per device (contiguous gene blocks), turns
// on direct peer access between the devices and wires the shards' mailboxes to each other by plain device pointers
// (no cudaIpc handles, no torch.distributed, no second process), so that every shard's log_prob kernel ends in the
// fused NVLink all-reduce of lp_grad_common.cuh.  The parent handle presents the GLOBAL problem: theta, gradients
// and every fit query use the global unconstrained layout; one persistent host thread per device drives its shard.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>

#include "host_util.h"
#include "model.h"
#include "multi.h"
#include "ppc.h"
#include "sampler.h"

namespace ppcseq {

// ---- one persistent host thread per shard ---------------------------------------------------------------------------
struct ShardPool {
    int n;
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;
    unsigned long long gen = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::vector<std::string> err;

    explicit ShardPool(int n_) : n(n_), rc(n_, 0), err(n_) {
        for (int q = 0; q < n; ++q) th.emplace_back([this, q] { loop(q); });
    }
    ~ShardPool() {
        { std::lock_guard<std::mutex> l(mu); stop = true; ++gen; }
        cv_go.notify_all();
        for (auto &t : th) t.join();
    }
    void loop(int q) {
        unsigned long long seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> l(mu);
                cv_go.wait(l, [&] { return gen != seen; });
                seen = gen;
                if (stop) return;
                f = job;
            }
            const int r = f(q);
            std::string e = r ? ppcseq_last_error() : "";
            {
                std::lock_guard<std::mutex> l(mu);
                rc[q] = r; err[q] = e;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs fn(q) on every shard thread; returns the first non-zero code (its message becomes the caller's last error)
    int run(std::function<int(int)> fn) {
        std::unique_lock<std::mutex> l(mu);
        job = std::move(fn);
        pending = n;
        ++gen;
        cv_go.notify_all();
        cv_done.wait(l, [&] { return pending == 0; });
        // a peer time-out on one shard is usually the echo of a real error on another: report the real one first
        for (int q = 0; q < n; ++q)
            if (rc[q] && rc[q] != PPCSEQ_ECOMM) { set_error("shard " + std::to_string(q) + ": " + err[q]); return rc[q]; }
        for (int q = 0; q < n; ++q)
            if (rc[q]) { set_error("shard " + std::to_string(q) + ": " + err[q]); return rc[q]; }
        return PPCSEQ_OK;
    }
};

void destroy_pool(ShardPool *p) { delete p; }

// rendez-vous of the shard threads (Model::pre_run_barrier)
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0;
    unsigned long long gen = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> l(mu);
        const unsigned long long g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};
struct BarrierScope {                                  // installs the barrier on every shard for the life of the scope
    Model *P;
    std::shared_ptr<HostBarrier> b;
    explicit BarrierScope(Model *p) : P(p), b(std::make_shared<HostBarrier>((int)p->shards.size())) {
        for (Model *s : P->shards) { auto bb = b; s->pre_run_barrier = [bb] { bb->wait(); }; }
    }
    ~BarrierScope() { for (Model *s : P->shards) s->pre_run_barrier = nullptr; }
};

// ---- global <-> local index algebra -----------------------------------------------------------------------------------
// global vector: [3 | intercept G | alpha_sub_1 K | alpha_2 (K x R, gene-major) | sigma_raw G | 3]; a shard holds the same
// blocks for its genes [g0, g1) (checked ones: the first Kl) plus its own copy of the 6 hyper-parameters.
struct Piece { long long glob, loc, len; };       // global offset, local offset, length
static void shard_pieces(const Model *P, int q, std::vector<Piece> &out, bool with_hyper) {
    const ModelDev &g = P->m;
    const ModelDev &l = P->shards[q]->m;
    const long long g0 = P->shard_g0[q], Gl = l.G, Kl = l.K, R = g.R;
    out.clear();
    if (with_hyper) out.push_back({0, 0, 3});
    out.push_back({g.o_intercept + g0, l.o_intercept, Gl});
    if (Kl > 0) {
        out.push_back({g.o_alpha1 + g0, l.o_alpha1, Kl});
        if (R > 0) out.push_back({g.o_alpha2 + g0 * R, l.o_alpha2, Kl * R});
    }
    out.push_back({g.o_sigma_raw + g0, l.o_sigma_raw, Gl});
    if (with_hyper) out.push_back({g.o_tail, l.o_tail, 3});
}

static void gather_local(const Model *P, int q, const double *glob, double *loc) {
    std::vector<Piece> pc;
    shard_pieces(P, q, pc, true);
    for (const Piece &p : pc) memcpy(loc + p.loc, glob + p.glob, sizeof(double) * (size_t)p.len);
}
static void scatter_global(const Model *P, int q, const double *loc, double *glob) {
    std::vector<Piece> pc;
    shard_pieces(P, q, pc, q == 0);                 // the replicated hyper entries are taken from shard 0
    for (const Piece &p : pc) memcpy(glob + p.glob, loc + p.loc, sizeof(double) * (size_t)p.len);
}

// ---- creation -------------------------------------------------------------------------------------------------------------
static int wire_mailboxes(Model *P, int channels, int cap) {
    const int W = (int)P->shards.size();
    int rc;
    for (int q = 0; q < W; ++q) {
        if (P->shards[q]->d_mailbox) comm_release(P->shards[q]);
        if ((rc = comm_alloc(P->shards[q], q, W, channels, cap))) return rc;
    }
    std::vector<void *> bases(W);
    for (int q = 0; q < W; ++q) bases[q] = P->shards[q]->d_mailbox;
    for (int q = 0; q < W; ++q)
        if ((rc = comm_attach(P->shards[q], bases.data()))) return rc;
    return PPCSEQ_OK;
}

int multi_ensure_comm(Model *P, int channels, int cap) {
    if (P->shards.size() < 2) return PPCSEQ_OK;
    const PeerComm &c = P->shards[0]->comm;
    if (c.channels >= channels && c.cap >= cap) return PPCSEQ_OK;
    for (Model *s : P->shards) {                    // nothing may be in flight while the mailboxes are replaced
        DeviceGuard g(s->device);
        PPCSEQ_CUDA(cudaDeviceSynchronize());
    }
    return wire_mailboxes(P, std::max(channels, c.channels), std::max(cap, c.cap));
}

int multi_create(int G, int S, int C, int K, const int32_t *counts, const double *X, const double *exposure,
                 double lambda_mu_mu, int n_devices, const int32_t *devices, Model **out) {
    if (!out) { set_error("out is NULL"); return PPCSEQ_EINVAL; }
    *out = nullptr;
    if (n_devices < 1 || n_devices > kCommMaxWorld) { set_error("need 1 <= n_devices <= 8"); return PPCSEQ_EINVAL; }
    if (G < n_devices) { set_error("fewer genes than devices"); return PPCSEQ_EINVAL; }
    if (!counts || !X || !exposure) { set_error("NULL data pointer"); return PPCSEQ_EINVAL; }
    if (K < 0 || K > G || S < 1 || C < 1 || C > kMaxC) { set_error("bad dimensions"); return PPCSEQ_EINVAL; }
    const bool trace = getenv("PPCSEQ_TRACE") != nullptr;       // set-up wall-clock split on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double>(b - a).count();
    };
    const auto t_begin = now();
    int ndev = 0;
    PPCSEQ_CUDA(cudaGetDeviceCount(&ndev));
    std::vector<int> dev(n_devices);
    for (int q = 0; q < n_devices; ++q) {
        dev[q] = devices ? devices[q] : q;
        if (dev[q] < 0 || dev[q] >= ndev) { set_error("no such CUDA device"); return PPCSEQ_EINVAL; }
        for (int r = 0; r < q; ++r)
            if (dev[r] == dev[q]) { set_error("a device is listed twice"); return PPCSEQ_EINVAL; }
    }
    // every pair of devices must be able to address each other's memory (NVLink / NVSwitch on a B200 box)
    for (int a = 0; a < n_devices; ++a) {
        DeviceGuard g(dev[a]);
        for (int b = 0; b < n_devices; ++b) {
            if (a == b) continue;
            int can = 0;
            PPCSEQ_CUDA(cudaDeviceCanAccessPeer(&can, dev[a], dev[b]));
            if (!can) { set_error("devices cannot access each other's memory (no peer access)"); return PPCSEQ_ESTATE; }
            cudaError_t e = cudaDeviceEnablePeerAccess(dev[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return PPCSEQ_ECUDA; }
        }
    }
    const auto t_peer = now();
    std::unique_ptr<Model> P(new (std::nothrow) Model());
    if (!P) return PPCSEQ_ENOMEM;
    P->device = dev[0];
    P->G_total = G; P->K_total = K; P->g_begin = 0;
    ModelDev &m = P->m;
    memset(&m, 0, sizeof(m));
    m.G = G; m.S = S; m.C = C; m.K = K; m.R = std::max(0, C - 2);
    m.o_intercept = 3; m.o_alpha1 = 3 + G; m.o_alpha2 = 3 + G + K; m.o_sigma_raw = m.o_alpha2 + m.R * K;
    m.o_tail = m.o_sigma_raw + G; m.D = (long long)m.o_tail + 3; m.lambda_mu_mu = lambda_mu_mu;
    P->hX.assign(X, X + (size_t)S * C);
    P->shards.assign(n_devices, nullptr);
    P->shard_g0.assign(n_devices + 1, 0);
    const int base = G / n_devices, extra = G % n_devices;
    for (int q = 0; q < n_devices; ++q) P->shard_g0[q + 1] = P->shard_g0[q] + base + (q < extra ? 1 : 0);
    P->pool = new ShardPool(n_devices);
    Model *Pp = P.get();
    int rc = P->pool->run([&](int q) {               // uploads and set-up kernels of all shards run concurrently
        const int g0 = Pp->shard_g0[q], g1 = Pp->shard_g0[q + 1];
        return create_impl(G, K, g0, g1, S, C, counts + (size_t)g0 * S, X, exposure, lambda_mu_mu, dev[q], &Pp->shards[q]);
    });
    if (rc) return rc;
    const auto t_shards = now();
    // default mailbox geometry: channel 0 for plain evaluations, 8 sampler channels, batches up to 128 thetas
    // (an ELBO estimate is 100); the samplers grow it on demand (multi_ensure_comm)
    if (n_devices > 1 && (rc = wire_mailboxes(Pp, 9, 128))) return rc;
    if (trace)
        fprintf(stderr, "[ppcseq] create_multi on %d devices: contexts + peer access %.2f s, shard upload + set-up %.2f s, "
                "mailboxes %.2f s\n", n_devices, secs(t_begin, t_peer), secs(t_peer, t_shards), secs(t_shards, now()));
    *out = P.release();
    return PPCSEQ_OK;
}

// ---- model entry points on a parent ---------------------------------------------------------------------------------------
int multi_set_exclusion(Model *P, const int32_t *pairs, long long n) {
    const int W = (int)P->shards.size();
    if (n < 0 || (n > 0 && !pairs)) { set_error("bad exclusion list"); return PPCSEQ_EINVAL; }
    std::vector<std::vector<int32_t>> loc(W);
    for (long long i = 0; i < n; ++i) {
        const int g = pairs[2 * i], s = pairs[2 * i + 1];
        if (g < 0 || g >= P->m.G || s < 0 || s >= P->m.S) { set_error("exclusion pair out of range"); return PPCSEQ_EINVAL; }
        const int q = (int)(std::upper_bound(P->shard_g0.begin(), P->shard_g0.end(), g) - P->shard_g0.begin()) - 1;
        loc[q].push_back(g - P->shard_g0[q]);
        loc[q].push_back(s);
    }
    return P->pool->run([&](int q) {
        return ppcseq_model_set_exclusion((ppcseq_model *)P->shards[q], loc[q].data(), (int64_t)(loc[q].size() / 2));
    });
}

int multi_set_design_path(Model *P, int mode) {
    for (Model *s : P->shards) {
        const int rc = ppcseq_model_set_design_path((ppcseq_model *)s, mode);
        if (rc) return rc;
    }
    return PPCSEQ_OK;
}

int multi_status(Model *P, int *flags) {
    *flags = 0;
    for (Model *s : P->shards) {
        int32_t f = 0;
        ppcseq_model_status((ppcseq_model *)s, &f);
        *flags |= f;
    }
    return PPCSEQ_OK;
}

// log_prob + gradient of B global thetas (host pointers).  Every shard thread slices its local thetas, runs the
// shard's own host-pointer entry point (copy/compute pipeline + fused all-reduce) and scatters its gene block into
// the global gradient; lp and the 6 hyper-gradients are bitwise identical on all shards and taken from shard 0.
int multi_log_prob_grad(Model *P, int B, const double *theta, int propto, int jacobian, double *lp, double *grad) {
    const long long D = P->m.D;
    int rc = multi_ensure_comm(P, 1, B);
    if (rc) return rc;
    // Phase 1 (only when a shard's scratch has to grow): every allocation of every shard, with the pool's rendez-vous
    // before the first peer-waiting kernel is launched -- a cudaMalloc / cudaFree on one device may have to synchronise
    // its peers, and a kernel spinning there for this shard's launch would close a wait cycle.
    bool grow = false;
    for (Model *s : P->shards) grow = grow || s->Bcap < B || (B > 1 && (int)s->ev_in.size() < B);
    if (grow) {
        rc = P->pool->run([&](int q) {
            Model *s = P->shards[q];
            DeviceGuard g(s->device);
            int r = s->ensure_batch(B);
            if (r == PPCSEQ_OK && B > 1) r = s->ensure_pipeline(B);
            return r;
        });
        if (rc) return rc;
    }
    return P->pool->run([&](int q) {
        Model *s = P->shards[q];
        const long long Dl = s->m.D;
        std::vector<double> &th = s->h_stage_theta, &gr = s->h_stage_grad, &l = s->h_stage_lp;
        th.resize((size_t)B * Dl); gr.resize((size_t)B * Dl); l.resize(B);
        for (int b = 0; b < B; ++b) gather_local(P, q, theta + (size_t)b * D, th.data() + (size_t)b * Dl);
        const int r = ppcseq_log_prob_grad((ppcseq_model *)s, B, th.data(), propto, jacobian, l.data(), gr.data());
        if (r) return r;
        for (int b = 0; b < B; ++b) scatter_global(P, q, gr.data() + (size_t)b * Dl, grad + (size_t)b * D);
        if (q == 0) memcpy(lp, l.data(), sizeof(double) * B);
        return (int)PPCSEQ_OK;
    });
}

int multi_flags(Model *P, const double *lower, const double *upper, const double *mean, const double *slope, uint8_t *ppc,
                uint8_t *deleterious, int32_t *failed, int32_t *tot_del) {
    const size_t S = (size_t)P->m.S;
    return P->pool->run([&](int q) {
        Model *s = P->shards[q];
        if (s->m.K == 0) return (int)PPCSEQ_OK;
        const size_t g0 = (size_t)P->shard_g0[q];
        return ppcseq_flags((ppcseq_model *)s, lower + g0 * S, upper + g0 * S, mean + g0 * S, slope ? slope + g0 : nullptr,
                            ppc + g0 * S, deleterious ? deleterious + g0 * S : nullptr, failed + g0,
                            tot_del ? tot_del + g0 : nullptr);
    });
}

// per-sample exposure gradient: every shard's S-vector, added in shard order (deterministic)
int multi_exposure_grad(Model *P, const double *theta, double *out) {
    const int W = (int)P->shards.size(), S = P->m.S;
    std::vector<std::vector<double>> part(W, std::vector<double>((size_t)S));
    int rc = P->pool->run([&](int q) {
        Model *s = P->shards[q];
        std::vector<double> th((size_t)s->m.D);
        gather_local(P, q, theta, th.data());
        return ppcseq_exposure_grad((ppcseq_model *)s, th.data(), part[q].data());
    });
    if (rc) return rc;
    for (int i = 0; i < S; ++i) {
        double t = 0.0;
        for (int q = 0; q < W; ++q) t += part[q][i];
        out[i] = t;
    }
    return PPCSEQ_OK;
}

// ---- fits ---------------------------------------------------------------------------------------------------------------
static Fit *new_parent_fit(Model *P, std::vector<Fit *> &parts) {
    Fit *F = new (std::nothrow) Fit();
    if (!F) { for (Fit *f : parts) delete f; return nullptr; }
    F->model = P;
    F->shard_fits = parts;
    F->n_draws = parts[0]->n_draws; F->ld = parts[0]->ld;
    F->info = parts[0]->info;                        // identical on every shard (same decisions everywhere)
    return F;
}

int multi_fit_from_draws(Model *P, const double *theta_draws, int n, Fit **out) {
    const int W = (int)P->shards.size();
    const long long D = P->m.D;
    std::vector<Fit *> parts(W, nullptr);
    int rc = P->pool->run([&](int q) {
        Model *s = P->shards[q];
        const long long Dl = s->m.D;
        std::vector<double> loc((size_t)n * Dl);
        for (int i = 0; i < n; ++i) gather_local(P, q, theta_draws + (size_t)i * D, loc.data() + (size_t)i * Dl);
        return ppcseq_fit_from_draws((ppcseq_model *)s, loc.data(), n, (ppcseq_fit **)&parts[q]);
    });
    if (rc) { for (Fit *f : parts) delete f; return rc; }
    *out = new_parent_fit(P, parts);
    return *out ? PPCSEQ_OK : PPCSEQ_ENOMEM;
}

// per-parameter query over the global range [begin, begin + count): fn(shard fit, local begin, count, out offset)
template <typename F>
static int for_global_range(const Fit *PF, long long begin, long long count, F fn) {
    const Model *P = PF->model;
    std::vector<Piece> pc;
    for (int q = 0; q < (int)P->shards.size(); ++q) {
        shard_pieces(P, q, pc, q == 0);
        for (const Piece &p : pc) {
            const long long lo = std::max(begin, p.glob), hi = std::min(begin + count, p.glob + p.len);
            if (lo >= hi) continue;
            const int rc = fn(PF->shard_fits[q], p.loc + (lo - p.glob), hi - lo, lo - begin);
            if (rc) return rc;
        }
    }
    return PPCSEQ_OK;
}

int multi_fit_get_draws(const Fit *PF, long long begin, long long count, double *out) {
    const size_t n = (size_t)PF->n_draws;
    return for_global_range(PF, begin, count, [&](Fit *f, long long lb, long long c, long long off) {
        return ppcseq_fit_get_draws((const ppcseq_fit *)f, lb, c, out + (size_t)off * n);
    });
}

int multi_fit_param_mean(const Fit *PF, long long begin, long long count, double *out) {
    return for_global_range(PF, begin, count, [&](Fit *f, long long lb, long long c, long long off) {
        return ppcseq_fit_param_mean((const ppcseq_fit *)f, lb, c, out + off);
    });
}

int multi_sample_nuts(Model *P, const ppcseq_nuts_opts &o, Fit **out) {
    const int W = (int)P->shards.size();
    const long long D = P->m.D;
    int rc = multi_ensure_comm(P, 1 + o.chains, 1);
    if (rc) return rc;
    std::vector<Fit *> parts(W, nullptr);
    BarrierScope barrier(P);
    rc = P->pool->run([&](int q) {
        Model *s = P->shards[q];
        ppcseq_nuts_opts lo = o;
        std::vector<double> init;
        if (o.init) {                               // [chains][D] global -> [chains][D_local]
            init.resize((size_t)o.chains * s->m.D);
            for (int c = 0; c < o.chains; ++c) gather_local(P, q, o.init + (size_t)c * D, init.data() + (size_t)c * s->m.D);
            lo.init = init.data();
        }
        return lo.chains <= kMaxBatch ? run_nuts_batched(s, lo, &parts[q]) : run_nuts(s, lo, &parts[q]);
    });
    if (rc) { for (Fit *f : parts) delete f; return rc; }
    *out = new_parent_fit(P, parts);
    return *out ? PPCSEQ_OK : PPCSEQ_ENOMEM;
}

int multi_advi(Model *P, const ppcseq_advi_opts &o, Fit **out) {
    const int W = (int)P->shards.size();
    int rc = multi_ensure_comm(P, 2, std::max(o.grad_samples, o.elbo_samples));
    if (rc) return rc;
    std::vector<Fit *> parts(W, nullptr);
    BarrierScope barrier(P);
    rc = P->pool->run([&](int q) {
        Model *s = P->shards[q];
        ppcseq_advi_opts lo = o;
        std::vector<double> init;
        if (o.init) {
            init.resize((size_t)s->m.D);
            gather_local(P, q, o.init, init.data());
            lo.init = init.data();
        }
        return run_advi(s, lo, &parts[q]);
    });
    if (rc) { for (Fit *f : parts) delete f; return rc; }
    *out = new_parent_fit(P, parts);
    return *out ? PPCSEQ_OK : PPCSEQ_ENOMEM;
}

// posterior-predictive summaries: genes are independent, every shard fills the rows of its own checked genes (no collective)
int multi_ppc_summary(Fit *PF, int exact, long long n_draws, double p, double tc, uint64_t seed, double *lower, double *upper,
                      double *mean, double *sd) {
    Model *P = PF->model;
    const size_t S = (size_t)P->m.S;
    int rc = P->pool->run([&](int q) {
        if (P->shards[q]->m.K == 0) return (int)PPCSEQ_OK;
        const size_t off = (size_t)P->shard_g0[q] * S;
        // the Philox streams are keyed by the GLOBAL (gene, sample) pair: same seed everywhere, same draws as one GPU
        return ppcseq_ppc_summary((ppcseq_fit *)PF->shard_fits[q], exact, n_draws, p, tc, seed, lower + off, upper + off,
                                  mean + off, sd + off);
    });
    if (rc) return rc;
    double ovf = 0.0;
    for (Fit *f : PF->shard_fits) ovf += f->info.size() > 8 ? f->info[8] : 0.0;
    if (PF->info.size() < 9) PF->info.resize(9, 0.0);
    PF->info[8] = ovf;
    return PPCSEQ_OK;
}

int multi_ppc_draws(Fit *PF, double tc, uint64_t seed, double *counts_rng) {
    Model *P = PF->model;
    const size_t S = (size_t)P->m.S, K = (size_t)P->m.K, n = (size_t)PF->n_draws;
    return P->pool->run([&](int q) {
        const size_t Kl = (size_t)P->shards[q]->m.K;
        if (Kl == 0) return (int)PPCSEQ_OK;
        std::vector<double> loc(n * Kl * S);
        const int rc = ppcseq_ppc_draws((ppcseq_fit *)PF->shard_fits[q], tc, seed, loc.data());
        if (rc) return rc;
        const size_t off = (size_t)P->shard_g0[q] * S;
        for (size_t d = 0; d < n; ++d) memcpy(counts_rng + d * K * S + off, loc.data() + d * Kl * S, sizeof(double) * Kl * S);
        return (int)PPCSEQ_OK;
    });
}

}  // namespace ppcseq
