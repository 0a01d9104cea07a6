// Mean-field ADVI, device-resident: the stand-in for rstan::vb(algorithm = "meanfield") that the reference
// calls through vb_iterative (/root/reference/R/utilities.R:246-278, options at :1487-1494).
//
// Algorithm (Stan's published ADVI, Kucukelbir et al. 2017, as rstan runs it -- restated, the Stan sources
// are not in the reference tree): variational family N(mu, diag(exp(omega))^2) on the unconstrained scale;
// stochastic gradient ascent on the ELBO with `grad_samples` reparameterised draws per step and the
// adaptive step-size sequence  s_k = 0.1 g_k^2 + 0.9 s_{k-1},  rho_k = eta k^{-1/2} / (1 + sqrt(s_k));
// eta picked from {100, 10, 1, 0.1, 0.01} by `adapt_iter` trial steps each; every `eval_elbo` iterations the
// ELBO is estimated from `elbo_samples` draws of log_prob<propto = false, jacobian = true> and the run stops
// when the mean or the median of the relative ELBO changes in a circular buffer drops below tol_rel_obj.
//
// B200 mapping: mu, omega, the step-size history and the draws live in HBM; one iteration is three
// launches (Philox draw, fused log_prob+grad, update) with no host synchronisation; the ELBO estimate is ONE
// batched log_prob launch over all `elbo_samples` draws (grid.y = batch).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <deque>
#include <memory>
#include <vector>

#include "host_util.h"
#include "ppc.h"
#include "sampler.h"

namespace ppcseq {

namespace {

struct Advi {
    Model *M;
    const ppcseq_advi_opts &o;
    EvalCtx ctx;
    RedScratch rs;
    ParamIds ids;
    DevBuf buf;
    long long D;
    int Bmax;
    double *mu = nullptr, *omega = nullptr, *hist_mu = nullptr, *hist_om = nullptr, *init = nullptr;
    double *eta = nullptr, *zeta = nullptr, *grad = nullptr, *lp = nullptr, *scal = nullptr;
    int *d_bad = nullptr;
    std::vector<double> h_lp;
    uint64_t ctr = 1;
    long long elbo_evals = 0;
    double D_global = 0.0;             // dimension of the whole model (sum over gene shards, hyper-parameters once)
    PreRunBarrier *barrier = nullptr;

    Advi(Model *m, const ppcseq_advi_opts &opts) : M(m), o(opts) {}
    ~Advi() { ctx.destroy(); rs.free_(); }

    int setup() {
        D = M->m.D;
        Bmax = std::max(o.grad_samples, o.elbo_samples);
        int rc;
        if ((rc = ctx.init(M, Bmax, false))) return rc;
        if ((rc = rs.alloc())) return rc;
        ctx.channel = 1;                   // gene-sharded run: comm channel 1 (channel 0 is the model's own)
        rs.comm = M->comm; rs.channel = 1; rs.o_tail = M->m.o_tail;
        rs.skip_hyper = (M->comm.world > 1 && M->comm.rank != 0) ? 1 : 0;
        ids.o_tail = M->m.o_tail;
        ids.gene_base = ((unsigned long long)(M->g_begin + 1)) << 32;
        if ((rc = buf.get(&mu, D)) || (rc = buf.get(&omega, D)) || (rc = buf.get(&hist_mu, D)) ||
            (rc = buf.get(&hist_om, D)) || (rc = buf.get(&init, D)) || (rc = buf.get(&eta, (size_t)Bmax * D)) ||
            (rc = buf.get(&zeta, (size_t)Bmax * D)) || (rc = buf.get(&grad, (size_t)Bmax * D)) ||
            (rc = buf.get(&lp, Bmax)) || (rc = buf.get(&scal, 8)) || (rc = buf.get(&d_bad, 1)))
            return rc;
        h_lp.resize(Bmax);
        PPCSEQ_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx.st));
        D_global = (double)D;
        PPCSEQ_CUDA(cudaStreamSynchronize(ctx.st));
        if (barrier) barrier->hit();                     // single-process multi-GPU: all shards allocated before any runs
        if (rs.comm.world > 1) {
            if ((rc = launch_fill(hist_mu, 1.0, D, ctx.st))) return rc;
            if ((rc = launch_sum(hist_mu, D, R(), scal, ctx.st))) return rc;
            PPCSEQ_CUDA(cudaMemcpyAsync(&D_global, scal, sizeof(double), cudaMemcpyDeviceToHost, ctx.st));
            PPCSEQ_CUDA(cudaStreamSynchronize(ctx.st));
        }
        return PPCSEQ_OK;
    }

    RedScratch R(bool count_all = false) {
        RedScratch r = rs;
        if (count_all) r.skip_hyper = 0;
        if (r.comm.world > 1) r.seq = ++M->chan_seq[r.channel];
        return r;
    }

    int reset_variational() {          // Q(cont_params): mu = init, omega = 0
        PPCSEQ_CUDA(cudaMemcpyAsync(mu, init, sizeof(double) * D, cudaMemcpyDeviceToDevice, ctx.st));
        PPCSEQ_CUDA(cudaMemsetAsync(omega, 0, sizeof(double) * D, ctx.st));
        return PPCSEQ_OK;
    }
    int reset_history() {
        PPCSEQ_CUDA(cudaMemsetAsync(hist_mu, 0, sizeof(double) * D, ctx.st));
        PPCSEQ_CUDA(cudaMemsetAsync(hist_om, 0, sizeof(double) * D, ctx.st));
        return PPCSEQ_OK;
    }

    // one stochastic-gradient step; asynchronous
    int step(double eta_scale, int iter) {
        int rc;
        const int B = o.grad_samples;
        if ((rc = launch_advi_draw(mu, omega, eta, zeta, D, B, o.seed, ctr, ids, ctx.st))) return rc;
        ctr += (uint64_t)B;
        if ((rc = ctx.eval(B, zeta, 1, 1, lp, grad))) return rc;
        return launch_advi_update(mu, omega, grad, eta, hist_mu, hist_om, D, B, eta_scale / std::sqrt((double)iter),
                                  iter == 1, d_bad, ctx.st);
    }

    // non-finite gradient seen since the last call?  (Stan throws from calc_grad; here the flag is polled)
    int poll_bad(bool *bad) {
        int h = 0;
        PPCSEQ_CUDA(cudaMemcpyAsync(&h, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx.st));
        PPCSEQ_CUDA(cudaStreamSynchronize(ctx.st));
        { const int src = M->check_status(); if (src) return src; }
        if (h) PPCSEQ_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx.st));
        if (rs.comm.world > 1) {           // every rank must reach the same verdict: all-reduce the flag
            const double hv = (double)h;
            double tot = 0.0;
            PPCSEQ_CUDA(cudaMemcpyAsync(scal + 4, &hv, sizeof(double), cudaMemcpyHostToDevice, ctx.st));
            int rc2 = launch_sum(scal + 4, 1, R(true), scal + 5, ctx.st);
            if (rc2) return rc2;
            PPCSEQ_CUDA(cudaMemcpyAsync(&tot, scal + 5, sizeof(double), cudaMemcpyDeviceToHost, ctx.st));
            PPCSEQ_CUDA(cudaStreamSynchronize(ctx.st));
            h = tot != 0.0;
        }
        *bad = h != 0;
        return PPCSEQ_OK;
    }

    // ELBO estimate; *ok = false when too many evaluations were dropped (Stan throws there)
    int calc_elbo(double *elbo, bool *ok) {
        const int n = o.elbo_samples;
        double sum = 0.0;
        int got = 0, dropped = 0, rc;
        *ok = true;
        while (got < n) {
            if ((rc = launch_advi_draw(mu, omega, eta, zeta, D, n, o.seed, ctr, ids, ctx.st))) return rc;
            ctr += (uint64_t)n;
            if ((rc = ctx.eval(n, zeta, 0, 1, lp, grad))) return rc;
            PPCSEQ_CUDA(cudaMemcpyAsync(h_lp.data(), lp, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx.st));
            PPCSEQ_CUDA(cudaStreamSynchronize(ctx.st));
            if ((rc = M->check_status())) return rc;
            for (int i = 0; i < n && got < n; ++i) {
                if (std::isfinite(h_lp[i])) { sum += h_lp[i]; ++got; }
                else if (++dropped >= n) { *ok = false; return PPCSEQ_OK; }
            }
        }
        ++elbo_evals;
        double h_sum = 0.0;
        if ((rc = launch_sum(omega, D, R(), scal, ctx.st))) return rc;
        PPCSEQ_CUDA(cudaMemcpyAsync(&h_sum, scal, sizeof(double), cudaMemcpyDeviceToHost, ctx.st));
        PPCSEQ_CUDA(cudaStreamSynchronize(ctx.st));
        *elbo = sum / n + 0.5 * D_global * (1.0 + 1.8378770664093454836) + h_sum;       // + entropy
        return PPCSEQ_OK;
    }

    int adapt_eta(double *eta_out) {
        static const double seq[5] = {100, 10, 1, 0.1, 0.01};
        double elbo = -INFINITY, elbo_best = -INFINITY, elbo_init, eta_best = 0.0;
        bool ok;
        int rc;
        if ((rc = calc_elbo(&elbo_init, &ok))) return rc;
        if (!ok) { set_error("ADVI: cannot compute the ELBO at the initial variational distribution"); return PPCSEQ_EDIVERGED; }
        for (int idx = 0;; ++idx) {
            const double eta_try = seq[idx];
            if ((rc = reset_history())) return rc;
            for (int it = 1; it <= o.adapt_iter; ++it)
                if ((rc = step(eta_try, it))) return rc;
            bool bad;
            if ((rc = poll_bad(&bad))) return rc;
            if (bad) elbo = -INFINITY;          // Stan zeroes the offending gradient and lets the ELBO decide
            else {
                if ((rc = calc_elbo(&elbo, &ok))) return rc;
                if (!ok || !std::isfinite(elbo)) elbo = -INFINITY;
            }
            bool done = false;
            if (elbo < elbo_best && elbo_best > elbo_init) {
                done = true;                      // the previous eta was the best
            } else {
                if (idx < 4) { elbo_best = elbo; eta_best = eta_try; }
                else if (elbo > elbo_init) { eta_best = eta_try; done = true; }
                else { set_error("ADVI: all proposed step-sizes failed"); return PPCSEQ_EDIVERGED; }
            }
            if ((rc = reset_variational())) return rc;
            if (done) break;
        }
        *eta_out = eta_best;
        return PPCSEQ_OK;
    }
};

}  // namespace

int run_advi(Model *M, const ppcseq_advi_opts &o, Fit **out) {
    *out = nullptr;
    PreRunBarrier barrier(M);          // hit once on every path out (other shards' threads must never be left waiting)
    if (o.iter < 1 || o.grad_samples < 1 || o.elbo_samples < 1 || o.eval_elbo < 1 || o.output_samples < 1 ||
        o.adapt_iter < 1 || !(o.tol_rel_obj > 0.0) || !(o.init_radius >= 0.0)) {
        set_error("bad ADVI options"); return PPCSEQ_EINVAL;
    }
    if (M->comm.world > 1 && (M->comm.channels < 2 || M->comm.cap < std::max(o.grad_samples, o.elbo_samples))) {
        set_error("gene-sharded ADVI needs ppcseq_comm_create(channels >= 2, cap >= max(grad_samples, elbo_samples))");
        return PPCSEQ_ESTATE;
    }
    DeviceGuard guard(M->device);
    const auto t0 = std::chrono::steady_clock::now();
    Advi A(M, o);
    A.barrier = &barrier;
    int rc;
    if ((rc = A.setup())) return rc;
    const long long D = A.D;
    // initial point: user-supplied, or U(-r, r) retried until log_prob and gradient are finite (Stan: 100 attempts)
    {
        std::vector<double> h(D);
        bool ok = false;
        for (int attempt = 0; attempt < 100 && !ok; ++attempt) {
            if (o.init) std::copy(o.init, o.init + D, h.begin());
            else for (long long i = 0; i < D; ++i) {
                uint32_t w[4];
                const unsigned long long pid = A.ids.id(i);
                philox4x32_10((uint32_t)pid, (uint32_t)(pid >> 32), 0x76626979u, 0x696e6974u + (uint32_t)attempt,
                              (uint32_t)o.seed, (uint32_t)(o.seed >> 32), w);
                const double u = ((double)(((uint64_t)w[0] << 21) | (w[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
                h[i] = (2.0 * u - 1.0) * o.init_radius;
            }
            PPCSEQ_CUDA(cudaMemcpyAsync(A.init, h.data(), sizeof(double) * D, cudaMemcpyHostToDevice, A.ctx.st));
            if ((rc = A.ctx.eval(1, A.init, 1, 1, A.lp, A.grad))) return rc;
            if ((rc = launch_sum(A.grad, D, A.R(), A.scal, A.ctx.st))) return rc;
            double v[2];
            PPCSEQ_CUDA(cudaMemcpyAsync(&v[0], A.lp, sizeof(double), cudaMemcpyDeviceToHost, A.ctx.st));
            PPCSEQ_CUDA(cudaMemcpyAsync(&v[1], A.scal, sizeof(double), cudaMemcpyDeviceToHost, A.ctx.st));
            PPCSEQ_CUDA(cudaStreamSynchronize(A.ctx.st));
            ok = std::isfinite(v[0]) && std::isfinite(v[1]);
            if (o.init) break;
        }
        if (!ok) { set_error("ADVI: could not find a finite starting point"); return PPCSEQ_EDIVERGED; }
    }
    if ((rc = A.reset_variational())) return rc;
    double eta = o.eta;
    if (o.adapt_engaged) { if ((rc = A.adapt_eta(&eta))) return rc; }
    // ---- stochastic gradient ascent -------------------------------------------------------------------
    if ((rc = A.reset_history())) return rc;
    const int cb_size = (int)std::max(0.1 * o.iter / o.eval_elbo, 2.0);
    std::deque<double> cb;
    double elbo = 0.0, elbo_prev;
    int stop_reason = 0, iters = 0;
    for (int it = 1;; ++it) {
        if ((rc = A.step(eta, it))) return rc;
        iters = it;
        bool stop = false;
        if (it % o.eval_elbo == 0) {
            bool bad, ok;
            if ((rc = A.poll_bad(&bad))) return rc;
            if (bad) { set_error("ADVI: non-finite gradient during stochastic gradient ascent"); return PPCSEQ_EDIVERGED; }
            elbo_prev = elbo;
            if ((rc = A.calc_elbo(&elbo, &ok))) return rc;
            if (!ok) { set_error("ADVI: too many dropped ELBO evaluations"); return PPCSEQ_EDIVERGED; }
            const double delta = std::fabs((elbo - elbo_prev) / elbo_prev);
            if ((int)cb.size() == cb_size) cb.pop_front();
            cb.push_back(delta);
            double ave = 0.0;
            for (double d : cb) ave += d;
            ave /= (double)cb.size();
            std::vector<double> v(cb.begin(), cb.end());
            const size_t mid = v.size() / 2;
            std::nth_element(v.begin(), v.begin() + mid, v.end());
            const double med = v[mid];
            if (ave < o.tol_rel_obj) { stop_reason |= 1; stop = true; }
            if (med < o.tol_rel_obj) { stop_reason |= 2; stop = true; }
        }
        if (it == o.iter) stop = true;
        if (stop) break;
    }
    {
        bool bad;
        if ((rc = A.poll_bad(&bad))) return rc;
        if (bad) { set_error("ADVI: non-finite gradient during stochastic gradient ascent"); return PPCSEQ_EDIVERGED; }
    }
    // ---- output_samples draws from the fitted Gaussian (what rstan keeps after dropping the mean row) ----
    std::unique_ptr<Fit> F(new (std::nothrow) Fit());
    if (!F) return PPCSEQ_ENOMEM;
    F->model = M; F->n_draws = o.output_samples; F->ld = (o.output_samples + 31) & ~31;
    PPCSEQ_CUDA(cudaMalloc((void **)&F->d_draws_T, (size_t)F->ld * D * sizeof(double)));
    PPCSEQ_CUDA(cudaMemsetAsync(F->d_draws_T, 0, (size_t)F->ld * D * sizeof(double), A.ctx.st));
    if ((rc = launch_advi_output(A.mu, A.omega, F->d_draws_T, F->ld, F->n_draws, D, o.seed ^ 0x9e3779b97f4a7c15ull, A.ids, A.ctx.st)))
        return rc;
    PPCSEQ_CUDA(cudaStreamSynchronize(A.ctx.st));
    if ((rc = M->check_status())) return rc;
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    F->info = {2.0, (double)A.ctx.n_evals, secs, (double)iters, (double)stop_reason, elbo, eta, (double)A.elbo_evals};
    *out = F.release();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
