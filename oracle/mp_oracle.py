"""Second, independent restatement of the target density in mpmath (50 digits), written from the
mathematics ("one-screen statement of the target") rather than from
oracle.hpp, to catch shared misreadings.  Pure-Python loops: small trees only.

*** TEST INFRASTRUCTURE ONLY (see oracle.hpp). ***

ln pi(x) = A + BD + C + lnL + lnJ with
  A   soft-bound calibrations (bounds / H), node-order constraints, braces
      (lib/Mcmc/Tree/Prior/Node/{Calibration,Constraint,Brace,Combined}.hs)
  BD  exponential(1) priors on lambda, mu + Stadler (2011) birth-death density conditioned on the
      MRCA, by the D/E recursion (lib/Mcmc/Tree/Prior/BirthDeath.hs)
  C   exponential(ht) on m, gamma(3/2, 1/6) on v, per-branch clock density
      (app/Probability.hs:96-124, lib/Mcmc/Tree/Prior/Branch/RelaxedClock.hs)
  lnL multivariate normal on d_k = H m sum t_i r_i (app/Probability.hs:166-207)
  lnJ -ln d_0 (app/Probability.hs:393-410)
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50
NINF = mp.mpf("-inf")


def _mp(v):
    return v if isinstance(v, mp.mpf) else mp.mpf(float(v))


def _soft(delta, p):
    s = mp.sqrt(mp.mpf(2) / mp.pi) * mp.mpf(p)
    return -(delta ** 2) / (2 * s ** 2)


def ln_posterior_parts(md, x):
    """md: ModelDesc-like; x: sequence of floats (one state).  Returns dict of mp values for valid
    states (all branches > 0, rates > 0, hyper-parameters in support, |lambda-mu| >= 1e-6)."""
    N = len(md.parent)
    x = [_mp(v) for v in x]
    la, mu_, H = x[0], x[1], x[2]
    h = x[3:3 + N]
    m, v = x[3 + N], x[4 + N]
    r = x[5 + N:5 + 2 * N]
    parent = [int(p) for p in md.parent]
    kids = [[] for _ in range(N)]
    for i in range(1, N):
        kids[parent[i]].append(i)
    t = [mp.mpf(0)] + [h[parent[i]] - h[i] for i in range(1, N)]

    # ---- A
    A = mp.mpf(0)
    for c in range(len(md.cal_node)):
        hc = h[int(md.cal_node[c])]
        lo, hi = float(md.cal_lo[c]), float(md.cal_hi[c])
        if lo > 0 and hc < mp.mpf(lo) / H:
            A += _soft(mp.mpf(lo) / H - hc, md.cal_lo_p[c])
        if hi != float("inf") and hc > mp.mpf(hi) / H:
            A += _soft(hc - mp.mpf(hi) / H, md.cal_hi_p[c])
    for c in range(len(md.con_young)):
        hy, ho = h[int(md.con_young[c])], h[int(md.con_old[c])]
        if not hy < ho:
            A += _soft(hy - ho, md.con_p[c])
    for b in range(len(md.brace_sd)):
        nodes = [int(z) for z in md.brace_node[md.brace_off[b]:md.brace_off[b + 1]]]
        hs = [h[z] for z in nodes]
        mean = sum(hs) / len(hs)
        sd = mp.mpf(float(md.brace_sd[b]))
        A += sum(-(z - mean) ** 2 / (2 * sd ** 2) for z in hs)

    # ---- BD: D/E recursion, rho = 1, E from the left child
    d = la - mu_

    def de(dt, e0, rho):
        xx = mp.e ** (-d * dt)
        c = (1 - rho) + rho * e0
        y = (mu_ - c * la) * xx
        den = la * (c - 1) + y
        return d * d * xx / den ** 2, (mu_ * (c - 1) + y) / den

    def rec(i):
        if not kids[i]:
            D, E = de(t[i], mp.mpf(0), mp.mpf(1))
            return mp.log(D), E
        (lnDl, El), (lnDr, _) = rec(kids[i][0]), rec(kids[i][1])
        D, E = de(t[i], El, mp.mpf(1))
        return mp.log(D * la) + lnDl + lnDr, E

    BD = -la - mu_ + rec(kids[0][0])[0] + rec(kids[0][1])[0]

    # ---- C
    C = mp.log(mp.mpf(float(md.ht))) - mp.mpf(float(md.ht)) * m
    C += mp.mpf("0.5") * mp.log(v) - 6 * v - mp.loggamma(mp.mpf("1.5")) + mp.mpf("1.5") * mp.log(6)
    for i in range(1, N):
        ri, ti = r[i], t[i]
        if md.clock_model in (0, 2):  # gamma with mean 1 and variance v (uncorrelated) or v / t (white noise)
            var = v if md.clock_model == 0 else v / ti
            k, th = 1 / var, var
            C += (k - 1) * mp.log(ri) - ri / th - mp.loggamma(k) - k * mp.log(th)
        else:  # log-normal with mean 1, variance parameter v or v t
            w = v if md.clock_model == 1 else v * ti
            C += -mp.log(mp.sqrt(2 * mp.pi)) - mp.log(ri * mp.sqrt(w)) - (mp.log(ri) + w / 2) ** 2 / (2 * w)

    # ---- likelihood + Jacobian
    K = N - 2
    rr = kids[0][1]

    def kidx(i):
        return 0 if i in (1, rr) else (i - 1 if i < rr else i - 2)

    dvec = [mp.mpf(0)] * K
    for i in range(1, N):
        dvec[kidx(i)] += t[i] * r[i]
    dvec = [z * H * m for z in dvec]
    if md.likelihood == 0:
        dx = [dvec[k] - mp.mpf(float(md.mean[k])) for k in range(K)]
        P = md.precision.reshape(K, K)
        quad = mp.mpf(0)
        for i in range(K):
            quad += dx[i] * sum(mp.mpf(float(P[i, j])) * dx[j] for j in range(K))
        lnL = -K * mp.log(mp.sqrt(2 * mp.pi)) - (mp.mpf(float(md.logdet_sigma)) + quad) / 2
    elif md.likelihood == 3:
        dx = [dvec[k] - mp.mpf(float(md.mean[k])) for k in range(K)]
        quad = mp.mpf(0)
        for i, j, val in zip(md.sparse_row, md.sparse_col, md.sparse_val):
            quad += dx[int(i)] * mp.mpf(float(val)) * dx[int(j)]
        lnL = -K * mp.log(mp.sqrt(2 * mp.pi)) - (mp.mpf(float(md.logdet_sigma)) + quad) / 2
    elif md.likelihood == 1:
        es = sum((dvec[k] - mp.mpf(float(md.mean[k]))) ** 2 / mp.mpf(float(md.precision[k])) for k in range(K))
        lnL = -K * mp.log(mp.sqrt(2 * mp.pi)) - (mp.mpf(float(md.logdet_sigma)) + es) / 2
    else:
        lnL = mp.mpf(0)
    lnJ = -mp.log(dvec[0])
    return {"A": A, "B": BD, "C": C, "prior": A + BD + C, "lik": lnL, "jac": lnJ, "post": A + BD + C + lnL + lnJ}


def grad_fd(md, x, idx, rel_step=mp.mpf("1e-25")):
    """central finite difference of ln post w.r.t. x[idx], evaluated at 50 digits"""
    x = [_mp(v) for v in x]
    hstep = rel_step * max(mp.mpf(1), abs(x[idx]))
    xp, xm = list(x), list(x)
    xp[idx] += hstep
    xm[idx] -= hstep
    return (ln_posterior_parts(md, xp)["post"] - ln_posterior_parts(md, xm)["post"]) / (2 * hstep)
