"""Plain NumPy restatement of Stan's NUTS (diag_e metric, windowed adaptation).  ORACLE ONLY: the CPU
reference the GPU sampler's posterior means are compared with (north_star: "posterior means concordant
within Monte Carlo error").  Follows the algorithm rstan::sampling runs for the reference
(/root/reference/R/utilities.R:1497-1512 passes chains / iter / warmup = 150 / seed / init = "random" and
leaves every control at rstan's default): multinomial NUTS, generalised U-turn criterion with the two
cross-subtree checks, divergence at H - H0 > 1000, dual averaging (delta .8, gamma .05, kappa .75, t0 10),
75 / 25 / 50 windowed diagonal variance with the (n/(n+5)) var + 1e-3 (5/(n+5)) shrinkage.
Stan's sources are not in the reference tree (external dependency, DESCRIPTION:32,59-60): parity unpinned.
"""
from __future__ import annotations

import math

import numpy as np


def _lse(a, b):
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    return max(a, b) + math.log1p(math.exp(-abs(a - b)))


class Nuts:
    def __init__(self, lp_grad, D, rng, max_depth=10, delta=0.8):
        self.f, self.D, self.rng, self.max_depth, self.delta = lp_grad, D, rng, max_depth, delta
        self.inv_m = np.ones(D)
        self.eps = 1.0
        self.n_evals = 0

    def grad(self, q):
        lp, g = self.f(q)
        self.n_evals += 1
        return -lp, g

    def leap(self, q, p, g, e):
        p = p + 0.5 * e * g
        q = q + e * self.inv_m * p
        V, g = self.grad(q)
        p = p + 0.5 * e * g
        return q, p, g, V

    def H(self, V, p):
        h = V + 0.5 * np.dot(p, self.inv_m * p)
        return math.inf if math.isnan(h) else h

    def crit(self, ps_a, ps_b, rho):
        return np.dot(ps_b, rho) > 0 and np.dot(ps_a, rho) > 0

    def build(self, depth, st, H0, sign, acc):
        """st: dict with q,p,g,V (the integrator state, advanced in place).  Returns
        (valid, zprop(q,g,V), p_beg, p_end, rho, log_sum_weight)."""
        if depth == 0:
            st["q"], st["p"], st["g"], st["V"] = self.leap(st["q"], st["p"], st["g"], sign * self.eps)
            acc["n"] += 1
            h = self.H(st["V"], st["p"])
            if h - H0 > 1000:
                acc["div"] = True
            acc["metro"] += 1.0 if H0 - h > 0 else math.exp(H0 - h)
            return (not acc["div"]), (st["q"].copy(), st["g"].copy(), st["V"]), st["p"].copy(), st["p"].copy(), st["p"].copy(), H0 - h
        ok, zp, p_beg, p_ie, rho_i, lw_i = self.build(depth - 1, st, H0, sign, acc)
        if not ok:
            return False, zp, p_beg, p_ie, rho_i, lw_i
        ok, zpf, p_fb, p_end, rho_f, lw_f = self.build(depth - 1, st, H0, sign, acc)
        if not ok:
            return False, zp, p_beg, p_end, rho_i, lw_i
        lw = _lse(lw_i, lw_f)
        if lw_f > lw or self.rng.uniform() < math.exp(lw_f - lw):
            zp = zpf
        rho = rho_i + rho_f
        im = self.inv_m
        ok = self.crit(im * p_beg, im * p_end, rho)
        ok = ok and self.crit(im * p_beg, im * p_fb, rho_i + p_fb)
        ok = ok and self.crit(im * p_ie, im * p_end, rho_f + p_ie)
        return ok, zp, p_beg, p_end, rho, lw

    def transition(self, q, g, V):
        p = self.rng.standard_normal(self.D) / np.sqrt(self.inv_m)
        H0 = self.H(V, p)
        fwd = dict(q=q.copy(), p=p.copy(), g=g.copy(), V=V)
        bck = dict(q=q.copy(), p=p.copy(), g=g.copy(), V=V)
        zs = (q.copy(), g.copy(), V)
        p_ff = p_fb = p_bf = p_bb = p.copy()
        rho = p.copy()
        lsw = 0.0
        acc = dict(n=0, metro=0.0, div=False)
        depth = 0
        while depth < self.max_depth:
            if self.rng.uniform() > 0.5:
                rho_bck = rho
                p_bf = p_ff
                ok, zp, p_fb, p_ff, rho_fwd, lw = self.build(depth, fwd, H0, 1.0, acc)
            else:
                rho_fwd = rho
                p_fb = p_bb
                ok, zp, p_bf, p_bb, rho_bck, lw = self.build(depth, bck, H0, -1.0, acc)
            if not ok:
                break
            depth += 1
            if lw > lsw or self.rng.uniform() < math.exp(lw - lsw):
                zs = zp
            lsw = _lse(lsw, lw)
            rho = rho_bck + rho_fwd
            im = self.inv_m
            ok = self.crit(im * p_bb, im * p_ff, rho)
            ok = ok and self.crit(im * p_bb, im * p_fb, rho_bck + p_fb)
            ok = ok and self.crit(im * p_bf, im * p_ff, rho_fwd + p_bf)
            if not ok:
                break
        return zs[0], zs[1], zs[2], acc["metro"] / acc["n"], acc["n"], acc["div"]

    def init_stepsize(self, q, g, V):
        def one():
            p = self.rng.standard_normal(self.D) / np.sqrt(self.inv_m)
            H0 = self.H(V, p)
            _, p1, _, V1 = self.leap(q, p, g, self.eps)
            return H0 - self.H(V1, p1)
        d = one()
        direction = 1 if d > math.log(0.8) else -1
        while True:
            d = one()
            if direction == 1 and not d > math.log(0.8):
                break
            if direction == -1 and not d < math.log(0.8):
                break
            self.eps = self.eps * 2 if direction == 1 else self.eps / 2
            if self.eps > 1e7 or self.eps == 0:
                raise RuntimeError("step size search failed")


def sample(lp_grad, D, n_iter, warmup, seed, init=None, init_buffer=75, term_buffer=50, window=25):
    """One chain.  Returns draws [n_iter - warmup, D] and a stats dict."""
    rng = np.random.default_rng(seed)
    s = Nuts(lp_grad, D, rng)
    q = rng.uniform(-2, 2, D) if init is None else np.array(init, dtype=float)
    V, g = s.grad(q)
    s.init_stepsize(q, g, V)
    mu, sbar, xbar, cnt = math.log(10 * s.eps), 0.0, 0.0, 0
    wc, wsize, wnext = 0, window, init_buffer + window - 1
    wn, wmean, wm2 = 0, np.zeros(D), np.zeros(D)
    draws = np.empty((n_iter - warmup, D))
    ndiv = 0
    nleap = 0
    for it in range(n_iter):
        q, g, V, a, n, div = s.transition(q, g, V)
        if it < warmup:
            cnt += 1
            a = min(a, 1.0)
            eta = 1.0 / (cnt + 10.0)
            sbar = (1 - eta) * sbar + eta * (s.delta - a)
            x = mu - sbar * math.sqrt(cnt) / 0.05
            xe = cnt ** -0.75
            xbar = (1 - xe) * xbar + xe * x
            s.eps = math.exp(x)
            if init_buffer <= wc < warmup - term_buffer:
                wn += 1
                d = q - wmean
                wmean += d / wn
                wm2 += (q - wmean) * d
            if wc == wnext and wc != warmup:
                if wnext != warmup - term_buffer - 1:
                    wsize *= 2
                    wnext = wc + wsize
                    if wnext != warmup - term_buffer - 1 and wnext + 2 * wsize >= warmup - term_buffer:
                        wnext = warmup - term_buffer - 1
                var = wm2 / (wn - 1)
                s.inv_m = (wn / (wn + 5.0)) * var + 1e-3 * (5.0 / (wn + 5.0))
                wn, wmean, wm2 = 0, np.zeros(D), np.zeros(D)
                s.init_stepsize(q, g, V)
                mu, sbar, xbar, cnt = math.log(10 * s.eps), 0.0, 0.0, 0
            wc += 1
            if it == warmup - 1:
                s.eps = math.exp(xbar)
        else:
            draws[it - warmup] = q
            ndiv += int(div)
            nleap += n
    return draws, dict(eps=s.eps, divergent=ndiv, n_evals=s.n_evals, leapfrogs=nleap)
