"""Host restatement of the device-resident Metropolis-Hastings proposals (mcmc-date_b200/csrc/mh_kernels.cuh) for the
tests: slideNodeAtUltrametric / scaleSubTreeAtUltrametric (lib/Mcmc/Tree/Proposal/Ultrametric.hs:50-62,126-147) with the
reference's truncated normal (lib/Statistics/Distribution/TruncatedNormal.hs:61-131, lib/Mcmc/Tree/Proposal/Internal.hs:
100-137), the ORACLE's value path and the same Philox uniforms.  Test infrastructure only."""
import math

import numpy as np
from scipy.special import erf, erfinv

from nuts_ref import uniform

SLIDE_NODE, SCALE_SUBTREE = 0, 1


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


def mh_step(orc, parent, X, out_cur, st_cur, kind, node, sd, tune, use_root_jacobian, seed, iteration):
    """one proposal on every chain; returns the accepted flags (1 / 0 / -1) and updates X, out_cur, st_cur in place"""
    N = len(parent)
    child, size, inner, inner_list = topology(parent)
    s = sd * tune
    acc = np.zeros(len(X), np.int32)
    for b in range(len(X)):
        h = X[b, 3:3 + N]
        j = node
        if j < 0:
            pick = min(int(uniform(seed, b, iteration, 2, 2) * len(inner_list)), len(inner_list) - 1)
            j = inner_list[pick]
        hj, hP = h[j], h[parent[j]]
        a = max(h[c] for c in child[j]) if kind == SLIDE_NODE else 0.0
        bb = hP
        if not (s > 0 and a < bb and not (a > hj) and not (bb < hj) and hj == hj):
            acc[b] = -1
            continue
        p = uniform(seed, b, iteration, 0, 2)
        phiA = phi2((a - hj) / s)
        z = phi2((bb - hj) / s) - phiA
        hnew = erfinv(2.0 * (p * z + phiA) - 1.0) * 1.41421356237309504880 * s + hj
        if a > hnew or bb < hnew or hnew != hnew or not z > 0:
            acc[b] = -1
            continue
        z2 = phi2((bb - hnew) / s) - phi2((a - hnew) / s)
        lnq = math.log(z) - math.log(z2)
        y = X[b].copy()
        if kind == SCALE_SUBTREE:
            xi = hnew / hj
            lnq += (inner[j] - 1) * math.log(xi)
            y[3 + j:3 + j + size[j]] *= xi
        y[3 + j] = hnew
        o1, s1 = orc.eval(y[None, :])
        o1, s1 = o1[0], int(s1[0])
        lr = (o1[3] + o1[4]) - (out_cur[b, 3] + out_cur[b, 4]) + lnq
        if use_root_jacobian:
            lr += o1[5] - out_cur[b, 5]
        u = uniform(seed, b, iteration, 1, 2)
        if math.log(u) < lr:          # False for NaN
            acc[b] = 1
            X[b] = y
            out_cur[b, :7] = o1[:7]
            st_cur[b] = s1
    return acc
