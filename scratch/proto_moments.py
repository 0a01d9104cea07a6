"""NumPy prototype of the Chebyshev-moment formulation of the mu-dependent part of the NB likelihood."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy import special as sp
from oracle import model_np as M
from ppcseq_b200 import synthetic

def choose_J(Emin, Emax, Jmax=48):
    Ec, hw = 0.5*(Emin+Emax), 0.5*(Emax-Emin)
    if hw == 0: return 1
    q0 = hw/(Ec + math.sqrt(Emin*Emax))
    for J in range(1, Jmax+1):
        if 2*q0**(J+1)/((J+1)*(1-q0)) < 2e-17: return J
    return None

def moments_likelihood(d, alpha, phi):
    """returns ll (scalar), d_alpha [G,C], d_phi [G] using group Chebyshev moments"""
    G,S,C = d.G, d.S, d.C
    n = d.counts.astype(np.float64)
    w = np.ones((G,S)) if d.exclude is None else (~d.exclude).astype(float)
    E = np.exp(d.exposure)
    Emin, Emax = E.min(), E.max()
    Ec, hw = 0.5*(Emin+Emax), 0.5*(Emax-Emin)
    J = choose_J(Emin, Emax)
    z = (E-Ec)/hw if hw>0 else np.zeros(S)
    T = np.empty((J+1, S)); T[0]=1; T[1]=z
    for j in range(2,J+1): T[j] = 2*z*T[j-1]-T[j-2]
    rows, grp = np.unique(d.X, axis=0, return_inverse=True)
    ng = len(rows)
    # data-only moments
    Mn = np.zeros((G, ng, J+1)); M1 = np.zeros((G, ng, J+1))
    for r in range(ng):
        sel = grp==r
        Mn[:,r,:] = (w[:,sel]*n[:,sel]) @ T[:,sel].T
        M1[:,r,:] = w[:,sel] @ T[:,sel].T
    big = (n>=32)&(w>0); small=(n<32)&(w>0)
    n_big = big.sum(1); Sn_big=(n*big).sum(1)
    S_eff = w.sum(1)
    A = (w*n*d.exposure[None,:]).sum(1); LG1=(w*sp.gammaln(n+1)).sum(1); Bc=(w*n)@d.X
    cum = np.zeros((G,32))
    for k in range(32): cum[:,k] = ((n>k)&small).sum(1)
    # ---- evaluation -----
    Mr = np.exp(rows @ alpha)   # [ng, G]
    ll = np.zeros(G); dphi=np.zeros(G); dal=np.zeros((G,C))
    invj = np.concatenate([[0.0], 1.0/np.arange(1,J+1)])
    for g in range(G):
        ph = phi[g]
        lp_g = A[g]-LG1[g]+alpha[:,g]@Bc[g] + S_eff[g]*ph*math.log(ph)
        lgphi, psphi = sp.gammaln(ph), sp.digamma(ph)
        # small part
        k = np.arange(32)
        lp_g += (cum[g]*np.log(ph+k)).sum()
        dp = (cum[g]/(ph+k)).sum()
        # big part (Stirling per element)
        xb = n[g][big[g]]+ph
        lx = np.log(xb); rx=1/xb; ww=rx*rx
        P = 1/12 + ww*(-1/360 + ww*(1/1260))
        Q = 1/12 + ww*(-1/120 + ww*(1/252))
        lp_g += ((xb-0.5)*lx + rx*P).sum() + n_big[g]*M.HALF_LOG_2PI - Sn_big[g] - n_big[g]*ph - n_big[g]*lgphi
        dp += (lx - 0.5*rx - ww*Q).sum() - n_big[g]*psphi
        dp += S_eff[g]*math.log(ph)
        da = np.zeros(C)
        for r in range(ng):
            m = Mr[r,g]
            Dm = (ph + m*Ec) + math.sqrt((ph+m*Emin)*(ph+m*Emax))
            q = m*hw/Dm; t=-q
            W = Mn[g,r] + ph*M1[g,r]
            An=0.0; A2=0.0; B=0.0
            for j in range(J,0,-1):
                An = An*t + W[j]*invj[j]
                A2 = A2*t + M1[g,r,j]*invj[j]
                B = B*t + W[j]
            An*=t; A2*=t; B*=t
            lD = math.log(Dm/2)
            Rs = 2/(Dm*(1-q*q))*(W[0]+2*B)
            Nr = M1[g,r,0]
            lp_g += -(W[0]*lD - 2*An)
            dp += Nr - Rs - (Nr*lD - 2*A2)
            da += ph*rows[r]*(Rs-Nr)
        ll[g]=lp_g; dphi[g]=dp; dal[g]=da
    return ll.sum(), dal, dphi, J

w = synthetic.make(G=300, S=500, C=3, mask=True, seed=5)
excl = np.zeros((w.G,w.S),bool); excl[w.exclude_pairs[:,0], w.exclude_pairs[:,1]]=True
d = M.ModelData(w.counts, w.X, w.exposure, w.K, exclude=excl)
for th in (w.theta_true, synthetic.random_thetas(w,1)[0]):
    p = M.unpack(th, d.G, d.K, d.C)
    phi = np.exp(-p["sigma_raw"]); alpha = M.alpha_matrix(p, d.G, d.K, d.C)
    eta = (d.X@alpha).T + d.exposure[None,:]
    ll0, d_eta, d_phi0 = M._likelihood(d, eta, phi)
    dal0 = d_eta @ d.X
    ll1, dal1, dphi1, J = moments_likelihood(d, alpha, phi)
    sc = np.maximum(np.abs(dal0), 1e-3*np.abs(dal0).max())
    print("J", J, "ll rel", abs(ll1-ll0)/abs(ll0), "dal", (np.abs(dal1-dal0)/sc).max(), "dphi", (np.abs(dphi1-d_phi0)/np.maximum(np.abs(d_phi0),1e-3*np.abs(d_phi0).max())).max())

# diagnose worst component at theta_true against mpmath
import mpmath as mp
mp.mp.dps = 40
th = w.theta_true
p = M.unpack(th, d.G, d.K, d.C)
phi = np.exp(-p["sigma_raw"]); alpha = M.alpha_matrix(p, d.G, d.K, d.C)
eta = (d.X@alpha).T + d.exposure[None,:]
ll0, d_eta, d_phi0 = M._likelihood(d, eta, phi)
dal0 = d_eta @ d.X
ll1, dal1, dphi1, J = moments_likelihood(d, alpha, phi)
sc = np.maximum(np.abs(dal0), 1e-3*np.abs(dal0).max())
err = np.abs(dal1-dal0)/sc
g, c = np.unravel_index(err.argmax(), err.shape)
print("worst", g, c, dal0[g], dal1[g], "scale", sc[g,c], "max count", d.counts[g].max(), "phi", phi[g])
# mp truth for gene g
wts = (~d.exclude[g])
tot = [mp.mpf(0)]*d.C
for s in range(d.S):
    if not wts[s]: continue
    n = mp.mpf(int(d.counts[g,s])); e = mp.mpf(d.exposure[s]) + sum(mp.mpf(d.X[s,cc])*mp.mpf(alpha[cc,g]) for cc in range(d.C))
    mu = mp.e**e; ph = mp.mpf(phi[g])
    de = n - (n+ph)*mu/(mu+ph)
    for cc in range(d.C): tot[cc] += de*mp.mpf(d.X[s,cc])
print("mp", [float(t) for t in tot])
