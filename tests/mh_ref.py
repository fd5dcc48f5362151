"""Host restatement of the device-resident Metropolis-Hastings-Green proposals, heated chains and MC3 swaps
(mcmc-date_b200/csrc/mh_kernels.cuh) for the tests.  Each proposal follows the reference's Haskell literally
(lib/Mcmc/Tree/Proposal/{Ultrametric,Unconstrained,Contrary,Brace,Internal}.hs, the truncated normal of
lib/Statistics/Distribution/TruncatedNormal.hs:61-131; `genericContinuous`, `scaleUnbiased`, `scaleContrarily` and the
MHG ratio of the un-vendored `mcmc` package from their published definitions) -- tree recursions as recursions, sums in
the reference's order -- with the ORACLE's value path and the same Philox uniforms as the device.  Test infrastructure only."""
import math

import numpy as np
from scipy.special import erf, erfinv

from nuts_ref import MASK, philox4x32_10

(SLIDE_NODE, SCALE_SUBTREE, PULLEY, SLIDE_BRACE, SCALE_BRANCH, SCALE_RATE_SUBTREE, SCALE_NORM_TREE_CONTRA_M,
 SCALE_NORM_TREE_CONTRA_H, SCALE_VAR_TREE, SCALE_VAR_TREE_AUTO, SLIDE_NODE_CONTRA, SCALE_SUBTREE_CONTRA, SLIDE_BRACE_CONTRA,
 SLIDE_ROOT_CONTRA, SCALE_RATES_TREE_CONTRA, SCALE_SCALAR, SCALE_H_M_CONTRA) = range(17)
NODE_KINDS = (SLIDE_NODE, SCALE_SUBTREE, SCALE_RATE_SUBTREE, SLIDE_NODE_CONTRA, SCALE_SUBTREE_CONTRA)
MULT_KINDS = (SCALE_BRANCH, SCALE_RATE_SUBTREE, SCALE_NORM_TREE_CONTRA_M, SCALE_NORM_TREE_CONTRA_H, SCALE_VAR_TREE,
              SCALE_VAR_TREE_AUTO, SCALE_SCALAR, SCALE_H_M_CONTRA)


def uniform2(seed, chain, iteration, draw, stream=2):
    c = philox4x32_10((chain, iteration, draw, stream), (seed & MASK, (seed >> 32) & MASK))
    return ((((c[0] >> 5) << 26) | (c[1] >> 6)) + 0.5) * 2.0 ** -53, ((((c[2] >> 5) << 26) | (c[3] >> 6)) + 0.5) * 2.0 ** -53


def uniform(seed, chain, iteration, draw, stream=2):
    return uniform2(seed, chain, iteration, draw, stream)[0]


def topology(parent):
    N = len(parent)
    child = [[] for _ in range(N)]
    for i in range(1, N):
        child[parent[i]].append(i)
    size, inner = [1] * N, [0] * N
    for i in range(N - 1, -1, -1):
        if child[i]:
            inner[i] += 1
        if i > 0:
            size[parent[i]] += size[i]
            inner[parent[i]] += inner[i]
    inner_list = [i for i in range(1, N) if child[i]]
    return child, size, inner, inner_list


def phi2(x):
    return 0.5 * (1.0 + erf(x * 0.70710678118654752440))


LAST = (0.0, 0.0)  # (ln q, ln |J|) of the last proposal, for the property tests
FORCED = None      # tests only: the sampled value (new height / shift / multiplier) instead of a draw


def truncated_normal_sample(m, s, a, b, p):
    """truncatedNormalSample (Internal.hs:107-138) -> (value, ln(qYX / qXY)) or None where the reference calls `error`"""
    if not s > 0 or not a < b or a > m or b < m or m != m:
        return None
    phiA = phi2((a - m) / s)
    z = phi2((b - m) / s) - phiA
    u = erfinv(2.0 * (p * z + phiA) - 1.0) * 1.41421356237309504880 * s + m
    if FORCED is not None:
        u = FORCED
    if a > u or b < u or u != u or not z > 0:
        return None
    z2 = phi2((b - u) / s) - phi2((a - u) / s)
    return float(u), math.log(z) - math.log(z2)


def gamma_sample(shape, scale, seed, chain, iteration):
    """Marsaglia & Tsang with the device's draw numbering"""
    if not shape > 0 or not scale > 0:
        return None
    if FORCED is not None:
        return FORCED
    boost, a = 1.0, shape
    if a < 1.0:
        boost = uniform(seed, chain, iteration, 7) ** (1.0 / a)
        a += 1.0
    d = a - 1.0 / 3.0
    c = 1.0 / math.sqrt(9.0 * d)
    for i in range(64):
        u0, u1 = uniform2(seed, chain, iteration, 8 + 2 * i)
        x = math.sqrt(-2.0 * math.log(u0)) * math.cos(2.0 * math.pi * u1)
        t = 1.0 + c * x
        if t <= 0.0:
            continue
        v = t * t * t
        uu = uniform(seed, chain, iteration, 9 + 2 * i)
        if math.log(uu) < 0.5 * x * x + d - d * v + d * math.log(v):
            return d * v * scale * boost
    return None


def propose(x, parent, topo, braces, kind, node, param, tune, seed, chain, iteration):
    """-> (proposed state, ln(q |J|), node) or (None, 0, node) where the reference would call `error`"""
    child, size, inner, inner_list = topo
    N = len(parent)
    OH, OR, OM, OV = 3, 5 + N, 3 + N, 4 + N
    y = x.copy()
    h, r = x[OH:OH + N], x[OR:OR + N]
    s = param * tune
    p = uniform(seed, chain, iteration, 0)
    if node < 0:
        un = uniform(seed, chain, iteration, 2)
        if kind in NODE_KINDS:
            node = inner_list[min(int(un * len(inner_list)), len(inner_list) - 1)]
        elif kind == SCALE_BRANCH:
            node = 1 + min(int(un * (N - 1)), N - 2)
        elif kind in (SLIDE_BRACE, SLIDE_BRACE_CONTRA):
            node = min(int(un * len(braces)), len(braces) - 1)
    j = node
    u, lnq, lnj = 1.0, 0.0, 0.0
    if kind in MULT_KINDS:
        kk, th = param / tune, tune / param
        u = gamma_sample(kk, th, seed, chain, iteration)
        if u is None:
            return None, 0.0, node
        # genericContinuous: qYX / qXY = pdf(1/u) / pdf(u) of gammaDistr kk th
        lnq = ((kk - 1.0) * math.log(1.0 / u) - (1.0 / u) / th) - ((kk - 1.0) * math.log(u) - u / th)
    root_l, root_r = child[0]
    if kind in (SLIDE_NODE, SLIDE_NODE_CONTRA):
        hj, hP, hcs = h[j], h[parent[j]], [h[c] for c in child[j]]
        res = truncated_normal_sample(hj, s, max(hcs), hP, p)
        if res is None:
            return None, 0.0, node
        hn, lnq = res
        y[OH + j] = hn
        if kind == SLIDE_NODE_CONTRA:
            xiS = (hP - hj) / (hP - hn)
            xis = [(hj - hc) / (hn - hc) for hc in hcs]
            y[OR + j] *= xiS
            for c, xi in zip(child[j], xis):
                y[OR + c] *= xi
            lnj = sum(math.log(xi) for xi in xis) + math.log(xiS)
    elif kind in (SCALE_SUBTREE, SCALE_SUBTREE_CONTRA):
        hj, hP = h[j], h[parent[j]]
        res = truncated_normal_sample(hj, s, 0.0, hP, p)
        if res is None:
            return None, 0.0, node
        hn, lnq = res
        xi = hn / hj
        y[OH + j + 1:OH + j + size[j]] *= xi
        y[OH + j] = hn
        if kind == SCALE_SUBTREE:
            lnj = (inner[j] - 1) * math.log(xi)
        else:
            xiR, xiS = 1.0 / xi, (hP - hj) / (hP - hn)
            y[OR + j + 1:OR + j + size[j]] *= xiR
            y[OR + j] *= xiS
            lnj = (inner[j] - size[j]) * math.log(xi) + math.log(xiS)
    elif kind == PULLEY:
        ht, hL, hR = h[0], h[root_l], h[root_r]
        brL, brR = ht - hL, ht - hR
        if brL <= 0 or brR <= 0:
            return None, 0.0, node
        a, b = -min(brL, ht - brR), min(brR, ht - brL)
        res = truncated_normal_sample(0.0, s, a, b, p)
        if res is None:
            return None, 0.0, node
        uu, lnq = res
        hLn, hRn = hL - uu, hR + uu
        xiL, xiR = hLn / hL, hRn / hR
        y[OH + root_l + 1:OH + root_l + size[root_l]] *= xiL
        y[OH + root_r + 1:OH + root_r + size[root_r]] *= xiR
        y[OH + root_l], y[OH + root_r] = hLn, hRn
        lnj = (inner[root_l] - 1) * math.log(xiL) + (inner[root_r] - 1) * math.log(xiR)
    elif kind in (SLIDE_BRACE, SLIDE_BRACE_CONTRA):
        nodes = braces[j]
        lo = max(max(h[c] for c in child[n]) - h[n] for n in nodes)
        hi = min(h[parent[n]] - h[n] for n in nodes)
        res = truncated_normal_sample(0.0, s, lo, hi, p)
        if res is None:
            return None, 0.0, node
        dl, lnq = res
        for n in nodes:
            y[OH + n] += dl
        if kind == SLIDE_BRACE_CONTRA:
            for n in reversed(nodes):          # foldr
                hNo, hPa = h[n], h[parent[n]]
                xiS = (hPa - hNo) / (hPa - hNo - dl)
                xis = [(hNo - h[c]) / (hNo + dl - h[c]) for c in child[n]]
                y[OR + n] *= xiS
                for c, xi in zip(child[n], xis):
                    y[OR + c] *= xi
                lnj += sum(math.log(xi) for xi in xis) + math.log(xiS)
    elif kind == SCALE_BRANCH:
        y[OR + j] *= u
        lnj = math.log(1.0 / u)
    elif kind == SCALE_RATE_SUBTREE:
        y[OR + j:OR + j + size[j]] *= u
        lnj = (size[j] - 2) * math.log(u)
    elif kind in (SCALE_NORM_TREE_CONTRA_M, SCALE_NORM_TREE_CONTRA_H):
        o = OM if kind == SCALE_NORM_TREE_CONTRA_M else 2
        y[o] = x[o] / u
        y[OR + 1:OR + N] *= u
        lnj = ((N - 1) - 2 - 1) * math.log(u)
    elif kind == SCALE_VAR_TREE:
        n = N - 1
        ssum = 0.0
        for i in range(1, N):                  # sum $ concatMap branches $ forest tr
            ssum += r[i]
        mu = ssum / n
        y[OV] = x[OV] * u * u
        for i in range(1, N):
            b2 = (r[i] - mu) * u + mu
            y[OR + i] = b2 if b2 > 0 else math.nan
        n1 = 1.0 / n
        lnj = n * math.log(u - n1 * u + n1)
    elif kind == SCALE_VAR_TREE_AUTO:
        muR = x[OM]
        y[OV] = x[OV] * u * u

        def scale_f(i, mu, mu2):               # scaleF oldParentRate newParentRate tree
            stack = [(i, mu, mu2)]
            while stack:
                i, mu, mu2 = stack.pop()
                d2 = u * (r[i] - mu)
                yy = mu2 + d2
                y[OR + i] = yy if yy > 0 else math.nan
                for c in child[i]:
                    stack.append((c, r[i], yy))

        for c in child[0]:
            scale_f(c, muR, muR)
        lnj = (N - 1) * math.log(u)
    elif kind == SLIDE_ROOT_CONTRA:
        H = x[2]
        if abs(h[0] - 1.0) > 1e-14:
            return None, 0.0, node
        hcs = [h[c] for c in child[0]]
        res = truncated_normal_sample(H, s, H * max(hcs), math.inf, p)
        if res is None:
            return None, 0.0, node
        Hn, lnq = res
        uu = Hn / H
        xis = [(1 - hc) / (uu - hc) for hc in hcs]
        y[2] = Hn
        y[OH + 1:OH + N] = h[1:] / uu
        for c, xi in zip(child[0], xis):
            y[OR + c] *= xi
        lnj = -inner[0] * math.log(uu) + sum(math.log(xi) for xi in xis)
    elif kind == SCALE_RATES_TREE_CONTRA:
        nn = inner[0] - 1
        if nn < 1:
            return None, 0.0, node
        hc = max(h[c] for c in child[0])
        res = truncated_normal_sample(hc, s, 0.0, h[0], p)
        if res is None:
            return None, 0.0, node
        hn, lnq = res
        xi = hn / hc
        y[OH + 1:OH + N] *= xi
        y[0], y[OM] = x[0] / xi, x[OM] / xi   # timeBirthRate, rateMean (ratesTimeTreeL, app/Definitions.hs:236-237)
        lnj = (nn - 1 - 2) * math.log(xi)
    elif kind == SCALE_SCALAR:
        o = [0, 1, 2, OM, OV][j]
        y[o] = x[o] * u
        lnj = math.log(1.0 / u)
    elif kind == SCALE_H_M_CONTRA:
        y[2] = x[2] * u
        y[OM] = x[OM] / u
        lnj = math.log(1.0 / (u * u))
    else:
        raise ValueError(kind)
    global LAST
    LAST = (lnq, lnj)
    return y, lnq + lnj, node


def mh_step(orc, parent, X, out_cur, st_cur, kind, node, sd, tune, use_root_jacobian, seed, iteration, braces=(),
            beta_prior=None, beta_lik=None, return_lr=False):
    """one proposal on every chain; returns the accepted flags (1 / 0 / -1) and updates X, out_cur, st_cur in place;
    beta_*[b]: heats of the chains (None: cold)"""
    topo = topology(parent)
    acc = np.zeros(len(X), np.int32)
    lrs = np.full(len(X), np.nan)
    for b in range(len(X)):
        y, lqj, _ = propose(X[b], parent, topo, braces, kind, node, sd, tune, seed, b, iteration)
        if y is None:
            acc[b] = -1
            continue
        o1, s1 = orc.eval(y[None, :])
        o1, s1 = o1[0], int(s1[0])
        bp = 1.0 if beta_prior is None else beta_prior[b]
        bl = 1.0 if beta_lik is None else beta_lik[b]
        with np.errstate(invalid="ignore"):
            lr = bp * (o1[3] - out_cur[b, 3]) + bl * (o1[4] - out_cur[b, 4]) + lqj
            if use_root_jacobian:
                lr += o1[5] - out_cur[b, 5]
        lrs[b] = lr
        u = uniform(seed, b, iteration, 1)
        if math.log(u) < lr:          # False for NaN
            acc[b] = 1
            X[b] = y
            out_cur[b, :7] = o1[:7]
            st_cur[b] = s1
    return (acc, lrs) if return_lr else acc


def mc3_swap(stats, slot, chain_of_slot, ladder_prior, ladder_lik, C, pair, seed, iteration):
    """one swap attempt per group; stats[c] = (ln prior, ln lik); updates slot / chain_of_slot in place -> accepted[g]"""
    G = len(slot) // C
    acc = np.zeros(G, np.int32)
    for g in range(G):
        u0, u1 = uniform2(seed, g, iteration, 0, 3)
        p = pair if pair >= 0 else min(int(u1 * (C - 1)), C - 2)
        i, j = chain_of_slot[g * C + p], chain_of_slot[g * C + p + 1]
        lr = (ladder_prior[p] - ladder_prior[p + 1]) * (stats[j, 0] - stats[i, 0]) + \
             (ladder_lik[p] - ladder_lik[p + 1]) * (stats[j, 1] - stats[i, 1])
        if math.log(u0) < lr:
            acc[g] = 1
            slot[i], slot[j] = p + 1, p
            chain_of_slot[g * C + p], chain_of_slot[g * C + p + 1] = j, i
    return acc
