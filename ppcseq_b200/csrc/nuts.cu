// NUTS (diagonal metric, windowed adaptation), device-resident: the stand-in for
// rstan::sampling(stanmodels$negBinomial_MPI, ...) as the reference calls it
// (/root/reference/R/utilities.R:1497-1512: chains, iter = ceil(draws/chains) + 150, warmup = 150,
// init = "random", save_warmup = FALSE; every other control is rstan's default).
//
// Algorithm (Stan's published sampler -- Hoffman & Gelman 2014, Betancourt 2017; restated, the Stan sources
// are not in the reference tree): multinomial NUTS with biased progressive sampling, the generalised
// U-turn criterion on rho and p# = M^-1 p including the two extra cross-subtree checks, divergence when
// H - H0 > 1000, dual-averaging step size (delta 0.8, gamma 0.05, kappa 0.75, t0 10) and the
// 75 / 25.. / 50 windowed diagonal variance estimate regularised as (n/(n+5)) var + 1e-3 (5/(n+5)),
// with the step size re-initialised by the doubling/halving heuristic after every metric update.
//
// B200 mapping: every D-length vector (position, momentum, gradient, the rho / p boundary vectors of
// every tree level, the metric, Welford moments) lives in HBM.  One leapfrog = 3 launches (half-step +
// drift, fused log_prob+grad, half-step fused with the tree's depth-0 bookkeeping + kinetic energy);
// one subtree merge = 1 launch giving rho and the six U-turn dot products.  The tree bookkeeping (energy error,
// divergence, multinomial weights and acceptances, U-turn verdicts) is done ON THE DEVICE by the last block of those
// kernels (sampler.h: TS_*, LeapBook, MergeBook): the host enqueues a whole subtree of 2^depth leapfrogs without
// reading anything back -- kernels that follow a divergence or an internal U-turn skip themselves -- and synchronises
// ONCE PER TREE DOUBLING (depth + 1 round trips per transition instead of ~2 per leapfrog).  Proposals are copied by
// the merge kernel instead of swapped by pointer, because the host no longer sees the acceptance.  Chains run
// concurrently, one host thread + one stream each.
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "host_util.h"
#include "sampler.h"

namespace ppcseq {

namespace {

inline double log_sum_exp(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + std::log1p(std::exp(b - a)) : b + std::log1p(std::exp(a - b));
}

struct ZFull { double *q = nullptr, *p = nullptr, *g = nullptr; double V = 0.0; };   // a phase-space point
struct ZProp { double *q = nullptr, *g = nullptr; double V = 0.0; };                 // a proposal (position only)

struct Level {                       // scratch of one recursion level of build_tree
    double *p_init_end, *rho_init, *p_final_beg, *rho_final;
    ZProp zpf;
};

struct ChainStats {
    long long n_leapfrog_post = 0, n_div_post = 0, n_maxdepth_post = 0, n_post = 0;
    double sum_accept_post = 0.0, eps_final = 0.0;
};

// Gene-sharded runs: every kernel of a chain ends in a cross-GPU exchange and waits there for the same chain's kernel
// on the peers.  If two chains could reach the device in different orders on two GPUs (separate streams that alias
// onto one hardware queue, or simply two host threads racing), A's chain 0 would wait for B's chain 0, queued behind
// B's chain 1, which waits for A's chain 1, queued behind A's chain 0.  So in that mode all chains of a device share
// ONE stream and take turns, in strict rotation, to enqueue the work up to their next host read: a chain holds the
// turn while it enqueues and passes it on while it waits for its scalars.  Every rank runs the same chains with
// bitwise the same decisions, hence the same rotation: the kernels of all ranks line up in the same order, whatever the
// hardware does with streams.  The device stays busy: while one chain's host thread reads and decides, the segments the
// other chains enqueued are running.
struct Turnstile {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> alive;
    int cur = 0;
    explicit Turnstile(int n) : alive(n, 1) {}
    void acquire(int c) {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return cur == c; });
    }
    void pass_from(int c) {                            // mu held
        const int n = (int)alive.size();
        for (int k = 1; k <= n; ++k) {
            const int nxt = (c + k) % n;
            if (alive[nxt]) { cur = nxt; cv.notify_all(); return; }
        }
        cur = -1;                                      // nobody left
        cv.notify_all();
    }
    void release(int c) { std::lock_guard<std::mutex> l(mu); pass_from(c); }
    void finish(int c) {                               // leaves the rotation (end of the chain, or an error)
        std::lock_guard<std::mutex> l(mu);
        alive[c] = 0;
        if (cur == c) pass_from(c);
    }
};

struct Chain {
    int id;
    Model *M;
    const ppcseq_nuts_opts &o;
    Fit *F;
    int col0;                                  // first draw column of this chain in the fit
    EvalCtx ctx;
    RedScratch rs;
    ParamIds ids;
    DevBuf buf;
    HostRng rng;
    long long D = 0;
    cudaStream_t st = nullptr;
    ZFull z, z_fwd, z_bck;
    ZProp z_sample, z_propose;
    double *p_ff, *p_fb, *p_bf, *p_bb, *rho, *rho_fwd, *rho_bck;
    double *inv_metric, *w_mean, *w_m2;
    double *d_scal;                            // [0] lp, [1] kinetic, [2..7] merge dot products
    double *h_scal = nullptr;                  // pinned mirror
    double *d_ts = nullptr, *h_ts = nullptr;   // device-side tree state (TS_*) and its pinned mirror
    uint64_t t_ctr = 0;                        // transitions started (keys the device-side acceptance draws)
    uint32_t node_ctr = 0;                     // merges enqueued in the current transition
    unsigned long long pending_reset = 0;      // accumulators the next leapfrog must reset
    long long leaf_idx = 0, n_leaves = 0;      // position of the next leaf in the subtree being enqueued
    Turnstile *turn = nullptr;                 // gene-sharded runs: the rotation of the chains on the shared stream
    cudaEvent_t ev = nullptr;                  // ... and the event this chain waits on instead of the whole stream
    std::vector<Level> lv;
    double eps = 1.0;
    uint64_t p_ctr = 0;
    bool divergent = false;
    int depth = 0;
    // dual averaging
    double da_mu = 0, da_sbar = 0, da_xbar = 0; int da_counter = 0;
    // windowed adaptation
    int w_counter = 0, w_size = 0, w_next = 0; double w_n = 0;
    ChainStats stats;
    int rc = PPCSEQ_OK;
    std::string err;

    Chain(int id_, Model *m, const ppcseq_nuts_opts &opts, Fit *f, int col)
        : id(id_), M(m), o(opts), F(f), col0(col), rng(opts.seed, 0x4e550000u + (uint32_t)id_) {}
    ~Chain() {
        ctx.destroy(); rs.free_();
        if (h_scal) cudaFreeHost(h_scal);
        if (h_ts) cudaFreeHost(h_ts);
        if (ev) cudaEventDestroy(ev);
    }

    // everything enqueued so far has to finish before the host reads: own stream -> synchronise it; shared stream ->
    // wait for this chain's event only, and let the other chains enqueue meanwhile
    int sync_point() {
        if (!turn) {
            PPCSEQ_CUDA(cudaStreamSynchronize(st));
        } else {
            PPCSEQ_CUDA(cudaEventRecord(ev, st));
            turn->release(id);
            const cudaError_t e = cudaEventSynchronize(ev);
            turn->acquire(id);
            if (e != cudaSuccess) { set_error(std::string("cudaEventSynchronize: ") + cudaGetErrorString(e)); return PPCSEQ_ECUDA; }
        }
        return M->check_status();               // a timed-out device-side wait is fatal, not a rejected proposal
    }

    // shared: the stream of the host thread that will run this chain (gene-sharded runs, see run_nuts); nullptr = a
    // stream of its own
    int setup(cudaStream_t shared) {
        D = M->m.D;
        int r;
        if ((r = ctx.init(M, 1, shared == nullptr, shared))) return r;
        if (shared) PPCSEQ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        st = ctx.st;
        if ((r = rs.alloc())) return r;
        // gene-sharded run: this chain owns comm channel 1 + id; ranks > 0 leave the replicated hyper-parameters out
        // of global sums; RNG streams are keyed by cross-rank parameter ids
        ctx.channel = 1 + id;
        rs.comm = M->comm; rs.channel = 1 + id; rs.o_tail = M->m.o_tail;
        rs.skip_hyper = (M->comm.world > 1 && M->comm.rank != 0) ? 1 : 0;
        ids.o_tail = M->m.o_tail;
        ids.gene_base = ((unsigned long long)(M->g_begin + 1)) << 32;
        auto vec = [&](double **p) { return buf.get(p, (size_t)D); };
        for (ZFull *zz : {&z, &z_fwd, &z_bck})
            if ((r = vec(&zz->q)) || (r = vec(&zz->p)) || (r = vec(&zz->g))) return r;
        for (ZProp *zp : {&z_sample, &z_propose})
            if ((r = vec(&zp->q)) || (r = vec(&zp->g))) return r;
        for (double **p : {&p_ff, &p_fb, &p_bf, &p_bb, &rho, &rho_fwd, &rho_bck, &inv_metric, &w_mean, &w_m2})
            if ((r = vec(p))) return r;
        lv.resize(o.max_treedepth + 1);
        for (int d = 1; d <= o.max_treedepth; ++d) {
            Level &L = lv[d];
            if ((r = vec(&L.p_init_end)) || (r = vec(&L.rho_init)) || (r = vec(&L.p_final_beg)) ||
                (r = vec(&L.rho_final)) || (r = vec(&L.zpf.q)) || (r = vec(&L.zpf.g)))
                return r;
        }
        if ((r = buf.get(&d_scal, 16))) return r;
        if ((r = buf.get(&d_ts, TS_SIZE))) return r;
        PPCSEQ_CUDA(cudaMallocHost((void **)&h_scal, 16 * sizeof(double)));
        PPCSEQ_CUDA(cudaMallocHost((void **)&h_ts, TS_SIZE * sizeof(double)));
        return launch_fill(inv_metric, 1.0, D, st);
    }

    RedScratch R() {                           // reduction scratch stamped with this call's comm sequence number
        RedScratch r = rs;
        if (r.comm.world > 1) r.seq = ++M->chan_seq[r.channel];
        return r;
    }

    int fetch(int first, int count) {          // device scalars -> host, synchronising the chain's stream
        PPCSEQ_CUDA(cudaMemcpyAsync(h_scal + first, d_scal + first, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
        return sync_point();
    }

    // potential and gradient at z.q  (hamiltonian.init / update_potential_gradient)
    int init_point(ZFull &zz) {
        int r;
        if ((r = ctx.eval(1, zz.q, 1, 1, d_scal, zz.g))) return r;
        if ((r = fetch(0, 1))) return r;
        zz.V = -h_scal[0];
        return PPCSEQ_OK;
    }

    // p ~ N(0, M); returns the kinetic energy
    int sample_p(ZFull &zz, double *kinetic) {
        int r;
        if ((r = launch_sample_p(zz.p, inv_metric, D, o.seed, 0x100u + (uint32_t)id, ++p_ctr, ids, R(), d_scal + 1, st))) return r;
        if ((r = fetch(1, 1))) return r;
        *kinetic = h_scal[1];
        return PPCSEQ_OK;
    }

    // one leapfrog step of z (expl_leapfrog::evolve); h = H(z) afterwards
    int leapfrog(double e, const LeapOut &lo, double *h) {
        int r;
        if ((r = launch_leap_a(z.q, z.p, z.g, inv_metric, e, D, st))) return r;
        if ((r = ctx.eval(1, z.q, 1, 1, d_scal, z.g))) return r;
        if ((r = launch_leap_b(z.p, z.g, inv_metric, e, lo, D, R(), d_scal + 1, st))) return r;
        if ((r = fetch(0, 2))) return r;
        z.V = -h_scal[0];
        double hh = z.V + h_scal[1];
        if (std::isnan(hh)) hh = INFINITY;
        *h = hh;
        return PPCSEQ_OK;
    }

    ZProp &prop(int id) { return id == 0 ? z_sample : (id == 1 ? z_propose : lv[id - 2].zpf); }

    // one leapfrog of z with the depth-0 bookkeeping on the device; nothing is read back
    // Consecutive leaves of a subtree continue from the same trajectory end: the second half-step kernel of a leaf also
    // takes the first half-step (and the position update) of the next one, except after the last leaf, whose end state
    // must stay exact for the next doubling.
    int leapfrog_async(double e, LeapOut lo, int acc_id, int prop_id) {
        int r;
        const double *skip = d_ts + TS_STOP;
        const bool first = leaf_idx == 0, last = leaf_idx == n_leaves - 1;
        ++leaf_idx;
        if (first && (r = launch_leap_a(z.q, z.p, z.g, inv_metric, e, D, st, skip))) return r;
        if (!last) lo.q_next = z.q;
        if ((r = ctx.eval(1, z.q, 1, 1, d_scal, z.g, skip))) return r;
        LeapBook bk;
        bk.ts = d_ts; bk.lp = d_scal; bk.acc_id = acc_id; bk.prop_id = prop_id; bk.reset_mask = pending_reset;
        pending_reset = 0;
        return launch_leap_b(z.p, z.g, inv_metric, e, lo, D, R(), d_scal + 1, st, bk);
    }

    // Stan base_nuts::build_tree, enqueued without host round trips.  The subtree's proposal ends up in proposal
    // `prop_id`, the momenta at its two ends in p_beg / p_end, the sum of its momenta in rho_out, its log sum of
    // weights is added to accumulator `acc_id`.  A divergence or a U-turn inside raises TS_STOP on the device.
    int build_tree(int dep, int prop_id, double *p_beg, double *p_end, double *rho_out, int acc_id, double sign) {
        int r;
        if (dep == 0) {
            LeapOut lo;
            ZProp &zp = prop(prop_id);
            lo.rho = rho_out; lo.p_beg = p_beg; lo.p_end = p_end; lo.zq = zp.q; lo.zg = zp.g; lo.q = z.q;
            return leapfrog_async(sign * eps, lo, acc_id, prop_id);
        }
        Level &L = lv[dep];
        const int a_init = 2 * dep - 1, a_final = 2 * dep, right = 2 + dep;
        pending_reset |= (1ull << a_init) | (1ull << a_final);
        if ((r = build_tree(dep - 1, prop_id, p_beg, L.p_init_end, L.rho_init, a_init, sign))) return r;
        if ((r = build_tree(dep - 1, right, L.p_final_beg, p_end, L.rho_final, a_final, sign))) return r;
        // multinomial sample from the right subtree, rho of the merged subtree, the three U-turn checks (around, and
        // across the two halves): one launch, verdicts on the device
        MergeBook bk;
        bk.ts = d_ts; bk.acc_init = a_init; bk.acc_final = a_final; bk.acc_parent = acc_id; bk.prop_dst = prop_id;
        bk.prop_src = right; bk.top = 0; bk.seed = o.seed; bk.tctr = t_ctr; bk.chain = (uint32_t)id; bk.node = ++node_ctr;
        bk.zq_dst = prop(prop_id).q; bk.zg_dst = prop(prop_id).g; bk.zq_src = L.zpf.q; bk.zg_src = L.zpf.g;
        return launch_merge(rho_out, L.rho_init, L.rho_final, p_beg, p_end, L.p_init_end, L.p_final_beg, inv_metric, D, R(),
                            d_scal + 2, st, bk);
    }

    int fetch_tree() {                         // the tree state -> host: the one round trip of a tree doubling
        PPCSEQ_CUDA(cudaMemcpyAsync(h_ts, d_ts, sizeof(double) * (TS_VPROP + 2), cudaMemcpyDeviceToHost, st));
        return sync_point();
    }

    // Stan base_nuts::transition.  On entry z.q / z.g / z.V hold the current state; on exit the new one.
    int transition(double *accept_stat, long long *n_leap, int *depth_out, bool *div_out) {
        int r;
        if ((r = launch_sample_p(z.p, inv_metric, D, o.seed, 0x100u + (uint32_t)id, ++p_ctr, ids, R(), d_scal + 1, st))) return r;
        if ((r = launch_tree_init(d_ts, d_scal + 1, z.V, st))) return r;      // H0 = V + kinetic, on the device
        ++t_ctr; node_ctr = 0; pending_reset = 0;
        // both ends of the trajectory, the sample and every boundary momentum start at z
        {
            BcastDst dq; dq.dst[0] = z_fwd.q; dq.dst[1] = z_sample.q;
            BcastDst dg; dg.dst[0] = z_fwd.g; dg.dst[1] = z_sample.g;
            BcastDst dp; dp.dst[0] = z_fwd.p; dp.dst[1] = p_ff; dp.dst[2] = p_fb; dp.dst[3] = p_bf; dp.dst[4] = p_bb; dp.dst[5] = rho;
            if ((r = launch_bcast(z.q, D, dq, st)) || (r = launch_bcast(z.g, D, dg, st)) || (r = launch_bcast(z.p, D, dp, st))) return r;
        }
        std::swap(z, z_bck);                          // z_bck = initial point; z becomes scratch
        depth = 0; divergent = false;
        while (depth < o.max_treedepth) {
            pending_reset |= 1ull;                    // accumulator 0: the subtree of this doubling
            MergeBook bk;
            bk.ts = d_ts; bk.acc_parent = 0; bk.prop_dst = 0; bk.prop_src = 1; bk.top = 1; bk.seed = o.seed; bk.tctr = t_ctr;
            bk.chain = (uint32_t)id;
            bk.zq_dst = z_sample.q; bk.zg_dst = z_sample.g; bk.zq_src = z_propose.q; bk.zg_src = z_propose.g;
            if (rng.uniform() > 0.5) {               // extend forward
                std::swap(z, z_fwd);
                std::swap(rho, rho_bck);             // rho_bck = rho of the old trajectory
                std::swap(p_bf, p_ff);               // p_bck_fwd = old forward end
                leaf_idx = 0; n_leaves = 1ll << depth;
                if ((r = build_tree(depth, 1, p_fb, p_ff, rho_fwd, 0, 1.0))) return r;
                std::swap(z, z_fwd);
            } else {                                  // extend backwards
                std::swap(z, z_bck);
                std::swap(rho, rho_fwd);
                std::swap(p_fb, p_bb);               // p_fwd_bck = old backward end
                leaf_idx = 0; n_leaves = 1ll << depth;
                if ((r = build_tree(depth, 1, p_bf, p_bb, rho_bck, 0, -1.0))) return r;
                std::swap(z, z_bck);
            }
            // valid subtree: the sample moves to the new proposal with probability min(1, w_new / w_old), the weights
            // merge, rho = rho_bck + rho_fwd and the three U-turn checks over the whole trajectory -- one launch that
            // skips itself when the subtree was cut short
            bk.node = ++node_ctr;
            if ((r = launch_merge(rho, rho_bck, rho_fwd, p_bb, p_ff, p_bf, p_fb, inv_metric, D, R(), d_scal + 2, st, bk))) return r;
            if ((r = fetch_tree())) return r;
            if (h_ts[TS_STOP] != 0.0) break;          // divergence or U-turn inside the new subtree
            ++depth;
            if (h_ts[TS_PERSIST] == 0.0) break;
        }
        const double nl = h_ts[TS_NLEAP];
        divergent = h_ts[TS_DIV] != 0.0;
        *accept_stat = h_ts[TS_METRO] / nl;
        *n_leap = (long long)nl; *depth_out = depth; *div_out = divergent;
        // z = z_sample
        std::swap(z.q, z_sample.q); std::swap(z.g, z_sample.g); z.V = h_ts[TS_VPROP];
        return PPCSEQ_OK;
    }

    // Stan base_hmc::init_stepsize: double / halve eps until the one-step acceptance crosses 0.8
    int init_stepsize() {
        if (eps == 0 || eps > 1e7 || std::isnan(eps)) return PPCSEQ_OK;
        int r;
        // back up the current point in z_propose
        BcastDst dq; dq.dst[0] = z_propose.q;
        BcastDst dg; dg.dst[0] = z_propose.g;
        if ((r = launch_bcast(z.q, D, dq, st)) || (r = launch_bcast(z.g, D, dg, st))) return r;
        z_propose.V = z.V;
        auto restore = [&]() -> int {
            BcastDst bq; bq.dst[0] = z.q;
            BcastDst bg; bg.dst[0] = z.g;
            int rr;
            if ((rr = launch_bcast(z_propose.q, D, bq, st)) || (rr = launch_bcast(z_propose.g, D, bg, st))) return rr;
            z.V = z_propose.V;
            return PPCSEQ_OK;
        };
        auto one_step = [&](double *delta) -> int {
            double kin, h;
            int rr;
            if ((rr = sample_p(z, &kin))) return rr;
            const double H0 = z.V + kin;
            if ((rr = leapfrog(eps, LeapOut(), &h))) return rr;
            *delta = H0 - h;
            return PPCSEQ_OK;
        };
        double delta;
        if ((r = one_step(&delta))) return r;
        const int direction = delta > std::log(0.8) ? 1 : -1;
        for (;;) {
            if ((r = restore())) return r;
            if ((r = one_step(&delta))) return r;
            if (direction == 1 && !(delta > std::log(0.8))) break;
            if (direction == -1 && !(delta < std::log(0.8))) break;
            eps = direction == 1 ? 2.0 * eps : 0.5 * eps;
            if (eps > 1e7) { set_error("NUTS: posterior is improper (step size diverged)"); return PPCSEQ_EDIVERGED; }
            if (eps == 0) { set_error("NUTS: no acceptably small step size could be found"); return PPCSEQ_EDIVERGED; }
        }
        return restore();
    }

    void learn_stepsize(double adapt_stat) {
        ++da_counter;
        adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
        const double eta = 1.0 / (da_counter + o.adapt_t0);
        da_sbar = (1.0 - eta) * da_sbar + eta * (o.adapt_delta - adapt_stat);
        const double x = da_mu - da_sbar * std::sqrt((double)da_counter) / o.adapt_gamma;
        const double x_eta = std::pow((double)da_counter, -o.adapt_kappa);
        da_xbar = (1.0 - x_eta) * da_xbar + x_eta * x;
        eps = std::exp(x);
    }

    // Stan windowed_adaptation + var_adaptation::learn_variance; returns true when the metric was updated
    int learn_variance(bool *updated) {
        *updated = false;
        int r;
        const int nw = o.warmup, tb = o.adapt_term_buffer, ib = o.adapt_init_buffer;
        const bool in_window = (w_counter >= ib) && (w_counter < nw - tb) && (w_counter != nw);
        if (in_window) {
            w_n += 1.0;
            if ((r = launch_welford_add(w_mean, w_m2, z.q, w_n, D, st))) return r;
        }
        const bool end_window = (w_counter == w_next) && (w_counter != nw);
        if (end_window) {
            // compute_next_window
            if (w_next != nw - tb - 1) {
                w_size *= 2;
                w_next = w_counter + w_size;
                if (w_next != nw - tb - 1) {
                    const int boundary = w_next + 2 * w_size;
                    if (boundary >= nw - tb) w_next = nw - tb - 1;
                }
            }
            if ((r = launch_welford_finish(w_m2, w_n, inv_metric, D, st))) return r;
            PPCSEQ_CUDA(cudaMemsetAsync(w_mean, 0, sizeof(double) * D, st));
            PPCSEQ_CUDA(cudaMemsetAsync(w_m2, 0, sizeof(double) * D, st));
            w_n = 0.0;
            *updated = true;
        }
        ++w_counter;
        return PPCSEQ_OK;
    }

    int run() {
        DeviceGuard guard(M->device);
        int r;
        // ---- initial point -----------------------------------------------------------------------
        {
            std::vector<double> h(D);
            bool ok = false;
            for (int attempt = 0; attempt < 100 && !ok; ++attempt) {
                if (o.init) std::copy(o.init + (size_t)id * D, o.init + (size_t)(id + 1) * D, h.begin());
                else for (long long i = 0; i < D; ++i) {       // keyed by the cross-rank parameter id: hyper-parameters agree on every rank
                    uint32_t w[4];
                    const unsigned long long pid = ids.id(i);
                    philox4x32_10((uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)id, 0x696e6974u + (uint32_t)attempt,
                                  (uint32_t)o.seed, (uint32_t)(o.seed >> 32), w);
                    const double u = ((double)(((uint64_t)w[0] << 21) | (w[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
                    h[i] = (2.0 * u - 1.0) * o.init_radius;
                }
                PPCSEQ_CUDA(cudaMemcpyAsync(z.q, h.data(), sizeof(double) * D, cudaMemcpyHostToDevice, st));
                if ((r = init_point(z))) return r;
                if ((r = launch_sum(z.g, D, R(), d_scal + 1, st))) return r;
                if ((r = fetch(1, 1))) return r;
                ok = std::isfinite(z.V) && std::isfinite(h_scal[1]);
                if (o.init) break;
            }
            if (!ok) { set_error("NUTS: could not find a finite starting point"); return PPCSEQ_EDIVERGED; }
        }
        PPCSEQ_CUDA(cudaMemsetAsync(w_mean, 0, sizeof(double) * D, st));
        PPCSEQ_CUDA(cudaMemsetAsync(w_m2, 0, sizeof(double) * D, st));
        eps = o.stepsize;
        const bool adapt = o.warmup > 0;
        bool windows = adapt;
        w_counter = 0; w_size = o.adapt_window; w_next = o.adapt_init_buffer + o.adapt_window - 1; w_n = 0;
        if (adapt && o.adapt_init_buffer + o.adapt_window + o.adapt_term_buffer > o.warmup) {
            // Stan: with fewer than 20 warm-up iterations no windows; otherwise 15% / 75% / 10%
            if (o.warmup < 20) windows = false;
            else {
                const int ib = (int)(0.15 * o.warmup), tb = (int)(0.1 * o.warmup);
                const_cast<ppcseq_nuts_opts &>(o).adapt_init_buffer = ib;
                const_cast<ppcseq_nuts_opts &>(o).adapt_term_buffer = tb;
                const_cast<ppcseq_nuts_opts &>(o).adapt_window = o.warmup - ib - tb;
                w_size = o.adapt_window; w_next = ib + w_size - 1;
            }
        }
        if ((r = init_stepsize())) return r;
        da_mu = std::log(10.0 * eps); da_sbar = 0; da_xbar = 0; da_counter = 0;
        // ---- iterations ------------------------------------------------------------------------------
        const int n_keep = o.iter - o.warmup;
        for (int it = 0; it < o.iter; ++it) {
            double accept; long long nl; int dep; bool div;
            if ((r = transition(&accept, &nl, &dep, &div))) return r;
            if (it < o.warmup) {
                learn_stepsize(accept);
                if (windows) {
                    bool upd;
                    if ((r = learn_variance(&upd))) return r;
                    if (upd) {
                        if ((r = init_stepsize())) return r;
                        da_mu = std::log(10.0 * eps); da_sbar = 0; da_xbar = 0; da_counter = 0;
                    }
                }
                if (it == o.warmup - 1) eps = std::exp(da_xbar);          // complete_adaptation
            } else {
                if ((r = launch_store_draw(F->d_draws_T, F->ld, col0 + (it - o.warmup), z.q, D, st))) return r;
                stats.n_leapfrog_post += nl; stats.n_post += 1; stats.sum_accept_post += accept;
                stats.n_div_post += div ? 1 : 0;
                stats.n_maxdepth_post += (dep >= o.max_treedepth) ? 1 : 0;
            }
        }
        (void)n_keep;
        stats.eps_final = eps;
        return sync_point();
    }
};

}  // namespace

int run_nuts(Model *M, const ppcseq_nuts_opts &o_in, Fit **out) {
    *out = nullptr;
    PreRunBarrier barrier(M);
    ppcseq_nuts_opts o = o_in;
    if (o.chains < 1 || o.iter < 1 || o.warmup < 0 || o.warmup >= o.iter || o.max_treedepth < 1 || o.max_treedepth > 20 ||
        !(o.adapt_delta > 0 && o.adapt_delta < 1) || !(o.stepsize > 0) || !(o.init_radius >= 0)) {
        set_error("bad NUTS options"); return PPCSEQ_EINVAL;
    }
    if (M->comm.world > 1 && (M->comm.channels < 1 + o.chains || M->comm.cap < 1)) {
        set_error("gene-sharded NUTS needs ppcseq_comm_create(channels >= 1 + chains)"); return PPCSEQ_ESTATE;
    }
    DeviceGuard guard(M->device);
    const auto t0 = std::chrono::steady_clock::now();
    const long long D = M->m.D;
    const int n_keep = o.iter - o.warmup;
    std::unique_ptr<Fit> F(new (std::nothrow) Fit());
    if (!F) return PPCSEQ_ENOMEM;
    F->model = M; F->n_draws = o.chains * n_keep; F->ld = (F->n_draws + 31) & ~31;
    PPCSEQ_CUDA(cudaMalloc((void **)&F->d_draws_T, (size_t)F->ld * D * sizeof(double)));
    PPCSEQ_CUDA(cudaMemset(F->d_draws_T, 0, (size_t)F->ld * D * sizeof(double)));
    PPCSEQ_CUDA(cudaDeviceSynchronize());
    // Gene-sharded runs: one stream for all chains of this device and a strict rotation of the chains (Turnstile)
    int n_threads = o.threads > 0 ? std::min(o.threads, o.chains) : o.chains;
    std::vector<cudaStream_t> shared_streams;
    std::unique_ptr<Turnstile> turnstile;
    if (M->comm.world > 1) {
        n_threads = o.chains;                            // every chain needs its own host thread to take its turns
        shared_streams.assign(1, nullptr);
        PPCSEQ_CUDA(cudaStreamCreateWithFlags(&shared_streams[0], cudaStreamNonBlocking));
        turnstile.reset(new Turnstile(o.chains));
    }
    struct StreamsGuard {
        std::vector<cudaStream_t> &v;
        ~StreamsGuard() { for (auto s : v) if (s) cudaStreamDestroy(s); }
    } streams_guard{shared_streams};
    std::vector<std::unique_ptr<Chain>> chains;
    std::vector<ppcseq_nuts_opts> copts(o.chains, o);       // per-chain copy (window sizes may be adjusted)
    for (int c = 0; c < o.chains; ++c) chains.emplace_back(new Chain(c, M, copts[c], F.get(), c * n_keep));
    // Every allocation happens HERE, before any chain runs: cudaMalloc / cudaMallocHost synchronise the device, and a
    // chain thread stuck in one while another chain's kernel spins on a peer GPU (gene-sharded runs) can close a
    // cross-rank wait cycle (rank A: chain 0 kernel waits for rank B; rank B: chain 0 thread waits in cudaMalloc for
    // chain 1's kernel, which waits for rank A's chain 1, whose thread waits in cudaMalloc for chain 0's kernel).
    for (auto &ch : chains) {
        ch->turn = turnstile.get();
        const int r = ch->setup(shared_streams.empty() ? nullptr : shared_streams[0]);
        if (r) { set_error("chain " + std::to_string(ch->id) + ": " + ppcseq_last_error()); return r; }
    }
    PPCSEQ_CUDA(cudaDeviceSynchronize());
    barrier.hit();                                       // single-process multi-GPU: all shards allocated before any runs
    auto worker = [&](int t) {
        for (int c = t; c < o.chains; c += n_threads) {
            Chain &ch = *chains[c];
            if (ch.turn) ch.turn->acquire(ch.id);        // a chain only ever enqueues while it holds the turn
            ch.rc = ch.run();
            if (ch.rc) ch.err = ppcseq_last_error();
            if (ch.turn) ch.turn->finish(ch.id);
        }
    };
    if (n_threads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
        for (auto &t : th) t.join();
    }
    long long evals = 0, nl = 0, ndiv = 0, nmax = 0, npost = 0;
    double acc = 0, eps = 0;
    for (auto &ch : chains) {
        if (ch->rc) { set_error("chain " + std::to_string(ch->id) + ": " + ch->err); return ch->rc; }
        evals += ch->ctx.n_evals; nl += ch->stats.n_leapfrog_post; ndiv += ch->stats.n_div_post;
        nmax += ch->stats.n_maxdepth_post; npost += ch->stats.n_post; acc += ch->stats.sum_accept_post;
        eps += ch->stats.eps_final;
    }
    chains.clear();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    F->info = {1.0, (double)evals, secs, (double)ndiv, (double)nmax, acc / std::max<long long>(npost, 1),
               eps / o.chains, (double)nl / std::max<long long>(npost, 1)};
    *out = F.release();
    return PPCSEQ_OK;
}

}  // namespace ppcseq
