"""Seeded synthetic trees, models and chain states of the shapes BASELINE.json names
(SURVEY.md section 8d).  Pure numpy; used by tests and bench.py on both the CUDA path and the oracle.
"""
from __future__ import annotations

import numpy as np

from . import model as _m
from . import tree as _tree

BASE_SEED = 0xD47E


def synthetic_precision(mu: np.ndarray, rng: np.random.Generator, band: int = 64, fill: float = 0.01):
    """Full-rank, well-conditioned Sigma^-1 = L L^T with L lower triangular: L_kk = 1/sigma_k,
    sigma_k = 0.1 mu_k + 1e-3, off-diagonals N(0, (0.05 L_kk)^2) inside a band plus random fill.
    Returns (precision [K,K] exactly symmetric, logdet Sigma)."""
    K = len(mu)
    sigma = 0.1 * mu + 1e-3
    diag = 1.0 / sigma
    L = np.zeros((K, K))
    i, j = np.tril_indices(K, -1)
    keep = (i - j <= band) | (rng.random(len(i)) < fill)
    i, j = i[keep], j[keep]
    L[i, j] = rng.normal(0.0, 1.0, len(i)) * 0.05 * diag[i] / np.sqrt(1.0 + np.minimum(i - j, band))
    L[np.arange(K), np.arange(K)] = diag
    P = L @ L.T
    P = 0.5 * (P + P.T)
    logdet_sigma = -2.0 * float(np.sum(np.log(diag)))
    return P, logdet_sigma


def synthetic_model(n_leaves: int, seed: int = BASE_SEED, clock_model: int = _m.UNCORRELATED_LOGNORMAL,
                    n_cal: int = 0, n_con: int = 0, n_brace: int = 0, likelihood: int = _m.LIK_FULL,
                    root_age: float = 100.0, parent: np.ndarray | None = None):
    """A random bifurcating tree with n_leaves leaves and a model around it.  Returns
    (ModelDesc, true_heights).  With n_cal > 0 the first calibration is on the root
    (bounds = root_age x [0.8, 1.2]) and ht = its mean (getMeanRootHeight, Calibration.hs:324-339)."""
    rng = np.random.default_rng(seed)
    if parent is None:
        parent = _tree.random_topology(n_leaves, rng)
    child0, child1 = _tree.children_from_parent(parent)
    N = len(parent)
    h = _tree.random_ultrametric_heights(parent, rng)
    t = np.zeros(N)
    t[1:] = h[parent[1:]] - h[1:]
    rho = rng.lognormal(-0.02, 0.2, N)
    bidx = _tree.branch_index(parent)
    K = N - 2
    mu = np.zeros(K)
    np.add.at(mu, bidx[1:], t[1:] * rho[1:])
    if likelihood == _m.LIK_FULL:
        prec, logdet = synthetic_precision(mu, rng)
    elif likelihood == _m.LIK_UNIVARIATE:
        prec = (0.1 * mu + 1e-3) ** 2
        logdet = float(np.sum(np.log(prec)))
    elif likelihood == _m.LIK_SPARSE:
        # banded Cholesky factor without fill -> an exactly banded (sparse), positive definite precision
        dense, logdet = synthetic_precision(mu, rng, band=6, fill=0.0)
        si, sj = np.nonzero(dense)
        sparse = (si.astype(np.int32), sj.astype(np.int32), dense[si, sj])
        prec = np.zeros(0)
    else:
        prec, logdet = np.zeros(0), 0.0

    inner = np.array([i for i in range(1, N) if child0[i] >= 0], dtype=np.int32)
    ht = 1.0
    cal_node, cal_lo, cal_hi = [], [], []
    if n_cal > 0:
        ht = root_age
        nodes = [0] + list(rng.choice(inner, size=min(n_cal - 1, len(inner)), replace=False))
        for nd in nodes:
            age = h[nd] * root_age
            cal_node.append(int(nd))
            cal_lo.append(0.8 * age)
            cal_hi.append(1.2 * age)
    con_y, con_o = [], []
    tries = 0
    while len(con_y) < n_con and tries < 1000 and len(inner) >= 2:
        tries += 1
        a, b = rng.choice(inner, size=2, replace=False)
        y, o = (a, b) if h[a] < h[b] else (b, a)
        # unrelated nodes only (ancestor/descendant pairs are redundant, Constraint.hs:230-241)
        x, anc = int(y), False
        while x > 0:
            x = int(parent[x])
            anc = anc or x == int(o)
        if not anc:
            con_y.append(int(y))
            con_o.append(int(o))
    br_off, br_node, br_sd = [0], [], []
    # braces: pairs of inner nodes that are neighbours in height (braced nodes are meant to be coeval)
    by_h = inner[np.argsort(h[inner])] if len(inner) else inner
    used = set()
    for _ in range(n_brace):
        cand = [j for j in range(len(by_h) - 1) if int(by_h[j]) not in used and int(by_h[j + 1]) not in used]
        if not cand:
            break
        j = int(rng.choice(cand))
        a, b = sorted((int(by_h[j]), int(by_h[j + 1])))
        used.update((a, b))
        br_node += [a, b]
        br_off.append(len(br_node))
        br_sd.append(1e-4)
    md = _m.ModelDesc(
        parent=parent, mean=mu, precision=prec, logdet_sigma=logdet, clock_model=clock_model,
        likelihood=likelihood, ht=ht,
        cal_node=cal_node, cal_lo=cal_lo, cal_lo_p=[0.025] * len(cal_node), cal_hi=cal_hi,
        cal_hi_p=[0.025] * len(cal_node),
        con_young=con_y, con_old=con_o, con_p=[0.025] * len(con_y),
        brace_off=br_off, brace_node=br_node, brace_sd=br_sd)
    if likelihood == _m.LIK_SPARSE:
        md.sparse_row, md.sparse_col, md.sparse_val = sparse
    return md, h


def synthetic_states(md: _m.ModelDesc, heights: np.ndarray, B: int, seed: int = BASE_SEED + 1,
                     jitter: float = 0.3) -> np.ndarray:
    """[B][S] chain-major states around `heights`: each inner non-root height moved by
    U(-jitter, jitter) x (gap to the nearest of parent / children), rates ~ LogNormal(0, 0.3),
    lambda, mu ~ LogNormal(0, 0.3), H ~ ht LogNormal(0, 0.05), m ~ LogNormal(0, 0.2)/ht,
    v ~ Gamma(1.5, 1/6).  All states are valid (every branch positive)."""
    rng = np.random.default_rng(seed)
    N, S = md.n_nodes, md.state_len
    parent, c0, c1 = md.parent, md.child0, md.child1
    X = np.zeros((B, S))
    X[:, 0] = rng.lognormal(0.0, 0.3, B)
    X[:, 1] = rng.lognormal(0.0, 0.3, B)
    near = np.abs(X[:, 0] - X[:, 1]) < 1e-3
    X[near, 1] += 0.01
    X[:, 2] = md.ht * rng.lognormal(0.0, 0.05, B) if md.calibrations_available else 1.0
    inner = np.array([i for i in range(1, N) if c0[i] >= 0], dtype=np.int64)
    h = np.tile(heights, (B, 1))
    if len(inner):
        up = heights[parent[inner]] - heights[inner]
        dn = heights[inner] - np.maximum(heights[c0[inner]], heights[c1[inner]])
        gap = np.minimum(up, dn)
        h[:, inner] += rng.uniform(-jitter, jitter, (B, len(inner))) * gap[None, :]
    # braced nodes stay nearly coeval: the later node follows the first one within its own safe interval
    for b in range(md.n_brace):
        nodes = md.brace_node[md.brace_off[b]:md.brace_off[b + 1]]
        a0 = int(nodes[0])
        for nd in nodes[1:]:
            nd = int(nd)
            up = heights[parent[nd]] - heights[nd]
            dn = heights[nd] - max(heights[c0[nd]], heights[c1[nd]])
            gp = jitter * min(up, dn)
            target = h[:, a0] + rng.normal(0.0, 2.0 * md.brace_sd[b], B)
            h[:, nd] = np.clip(target, heights[nd] - gp, heights[nd] + gp)
    X[:, 3:3 + N] = h
    X[:, 3 + N] = rng.lognormal(0.0, 0.2, B) / md.ht
    X[:, 4 + N] = rng.gamma(1.5, 1.0 / 6.0, B) + 1e-3
    X[:, 5 + N:5 + 2 * N] = rng.lognormal(0.0, 0.3, (B, N))
    X[:, 5 + N] = 0.0  # rate stem (app/Definitions.hs:96-123: stem 0)
    return X
