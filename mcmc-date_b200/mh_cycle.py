"""Host-side mirror of the reference's proposal cycle (`proposals`, app/Definitions.hs:125-279): which proposals exist for a
given tree, on which nodes, with which standard deviations / shapes, weights and root-branch Jacobian lifts.  The device
executes the list (binding.Evaluator.mh_cycle -> mcd_mh_cycle); order, tuning and monitors stay with the host."""
from __future__ import annotations

import math

import numpy as np

from . import binding as _b
from . import model as _m


def _topology(parent):
    N = len(parent)
    child = [[] for _ in range(N)]
    for i in range(1, N):
        child[int(parent[i])].append(i)
    depth = [1] * N           # elynx `depth`: leaves 1
    plen = [0] * N            # length of the path from the root
    for i in range(1, N):
        plen[i] = plen[int(parent[i])] + 1
    for i in range(N - 1, 0, -1):
        p = int(parent[i])
        depth[p] = max(depth[p], depth[i] + 1)
    return child, depth, plen


def weight_n_branches(n: int) -> int:
    """weightNBranches (app/Definitions.hs:127-130)"""
    return int(math.floor(math.log(float(n)) / math.log(1.3)))


def reference_cycle(md: _m.ModelDesc, calibrations_available: bool | None = None, tune: float = 1.0):
    """-> list of (kind, node, param, tune, use_root_jacobian, weight) in the order of `proposals`
    (app/Definitions.hs:256-279; the Hamiltonian proposal is mcd_nuts and not part of this list)"""
    parent = np.asarray(md.parent)
    N = len(parent)
    child, depth, plen = _topology(parent)
    cal = md.calibrations_available if calibrations_available is None else calibrations_available
    w = weight_n_branches(N)
    braces = range(md.n_brace)
    inner = [i for i in range(1, N) if child[i]]
    out = []

    def add(kind, node, param, jac, weight):
        out.append((kind, node, param, tune, int(jac), int(weight)))

    def w_depth(i):           # min (wMin + depth - 2) wMax with wMin 3, wMax 8
        return min(3 + depth[i] - 2, 8)

    # hyper-parameters (:259-263)
    for scalar in (0, 1, 3, 4):
        add(_b.MH_SCALE_SCALAR, scalar, 10.0, False, w)
    if len(inner) >= 1:
        add(_b.MH_SCALE_RATES_TREE_CONTRA, 0, 0.1, True, w)
    # proposalsTimeTree (:145-166): [R] children of the root, [O] other nodes, [B] braces
    rl, rr = child[0]
    if child[rl] and child[rr]:
        add(_b.MH_PULLEY, 0, 0.01, True, 6)
    for at_root in (True, False):
        nodes = [i for i in inner if (plen[i] == 1) == at_root]
        for i in nodes:
            add(_b.MH_SLIDE_NODE, i, 0.01, at_root, 5)
        for i in nodes:
            add(_b.MH_SCALE_SUBTREE, i, 0.01, at_root, w_depth(i))
    for b in braces:
        add(_b.MH_SLIDE_BRACE, b, 0.01, False, 5)
    # proposalsRateTree (:180-201)
    add(_b.MH_SCALE_NORM_TREE_CONTRA_M, 0, 100.0, True, w)
    add(_b.MH_SCALE_VAR_TREE, 0, 100.0, True, w)
    add(_b.MH_SCALE_VAR_TREE_AUTO, 0, 100.0, True, w)
    for at_root in (True, False):
        allnodes = [i for i in range(1, N) if (plen[i] == 1) == at_root]
        for i in allnodes:
            add(_b.MH_SCALE_BRANCH, i, 100.0, at_root, 3)
        for i in allnodes:
            if child[i]:
                add(_b.MH_SCALE_RATE_SUBTREE, i, 100.0, at_root, w_depth(i))
    # proposalsTimeRateTreeContra (:204-224)
    for at_root in (True, False):
        nodes = [i for i in inner if (plen[i] == 1) == at_root]
        for i in nodes:
            add(_b.MH_SLIDE_NODE_CONTRA, i, 0.1, at_root, w_depth(i))
        for i in nodes:
            add(_b.MH_SCALE_SUBTREE_CONTRA, i, 0.1, at_root, w_depth(i))
    for b in braces:
        add(_b.MH_SLIDE_BRACE_CONTRA, b, 0.1, False, 5)
    # proposalsChangingTimeHeight (:241-253), only with calibrations
    if cal:
        add(_b.MH_SCALE_SCALAR, 2, 3000.0, False, w)
        add(_b.MH_SCALE_H_M_CONTRA, 0, 10.0, False, w)
        add(_b.MH_SCALE_NORM_TREE_CONTRA_H, 0, 100.0, True, w)
        add(_b.MH_SLIDE_ROOT_CONTRA, 0, 10.0, True, w)
    return out


# ----------------------------------------------------------------------------- order, dimensions, auto tuning (host side)
def proposal_dimension(md: _m.ModelDesc, entry) -> int:
    """PDimension of one cycle entry, as the reference's proposal constructors state it (the auto tuner's target acceptance
    rate depends on it): lib/Mcmc/Tree/Proposal/Ultrametric.hs:97,185,311; Unconstrained.hs:50,115,280,366,435;
    Contrary.hs:133,262,382,482; Brace.hs:90,200; `scaleUnbiased` 1 and `scaleContrarily` 2 (`mcmc` package)."""
    kind, node = entry[0], entry[1]
    parent = np.asarray(md.parent)
    N = len(parent)
    child, _, plen = _topology(parent)
    size = [1] * N
    inner = [1 if child[i] else 0 for i in range(N)]
    for i in range(N - 1, 0, -1):
        size[int(parent[i])] += size[i]
        inner[int(parent[i])] += inner[i]
    brace_nodes = lambda b: [int(x) for x in md.brace_node[md.brace_off[b]:md.brace_off[b + 1]]]
    if kind in (_b.MH_SLIDE_NODE, _b.MH_SCALE_BRANCH, _b.MH_SCALE_SCALAR):
        return 1
    if kind == _b.MH_SCALE_H_M_CONTRA:
        return 2
    if kind == _b.MH_SCALE_SUBTREE:
        return inner[node]
    if kind == _b.MH_PULLEY:
        return inner[child[0][0]] + inner[child[0][1]]
    if kind == _b.MH_SCALE_RATE_SUBTREE:
        return size[node]                                   # scaleTree: `length tr`, stem included
    if kind in (_b.MH_SCALE_NORM_TREE_CONTRA_M, _b.MH_SCALE_NORM_TREE_CONTRA_H, _b.MH_SCALE_VAR_TREE, _b.MH_SCALE_VAR_TREE_AUTO):
        return (N - 1) + 1
    if kind == _b.MH_SLIDE_NODE_CONTRA:
        return 1 + 1 + len(child[node])                      # never the root here
    if kind == _b.MH_SCALE_SUBTREE_CONTRA:
        return inner[node] + size[node]
    if kind == _b.MH_SLIDE_ROOT_CONTRA:
        return 1 + inner[0] + len(child[0])
    if kind == _b.MH_SCALE_RATES_TREE_CONTRA:
        return (inner[0] - 1) + 2
    if kind == _b.MH_SLIDE_BRACE:
        return len(brace_nodes(node))
    if kind == _b.MH_SLIDE_BRACE_CONTRA:
        ns = brace_nodes(node)
        return len(ns) + sum(1 for i in ns if plen[i] > 0) + sum(len(child[i]) for i in ns)
    raise ValueError(f"unknown proposal kind {kind}")


def optimal_rate(dim: int) -> float:
    """target acceptance rate of the auto tuner by proposal dimension: 0.44 in one dimension falling linearly (step 0.0515) to
    0.234 from five dimensions on (Roberts & Rosenthal optimal scaling; the `mcmc` package's `getOptimalRate`, restated from
    its published definition -- the package is not vendored)."""
    if dim <= 0:
        raise ValueError("getOptimalRate: Proposal dimension is zero or negative.")
    return 0.44 - 0.0515 * (min(dim, 5) - 1)


TUNE_MIN, TUNE_MAX = 1e-5, 1e3


def auto_tune(md: _m.ModelDesc, cycle, accepted, proposed):
    """one auto-tuning step after a tuning period: t <- t * exp(2 (rate - optimal rate)) for every entry of the cycle
    (`mcmc` package: tuningFunction / autoTuneCycle, restated from its published definition); entries that were never
    proposed keep their parameter.  -> new cycle list"""
    out = []
    for e, a, n in zip(cycle, accepted, proposed):
        kind, node, param, tune, jac, w = e
        if n > 0:
            rate = float(a) / float(n)
            tune = min(TUNE_MAX, max(TUNE_MIN, tune * math.exp(2.0 * (rate - optimal_rate(proposal_dimension(md, e))))))
        out.append((kind, node, param, tune, jac, w))
    return out


def random_order(cycle, rng: np.random.Generator):
    """one iteration in the `mcmc` package's default order RandomO: every proposal replicated by its weight, the whole list
    shuffled anew for each iteration.  -> (list of entries with repeat 1, index of each into `cycle`)"""
    idx = np.repeat(np.arange(len(cycle)), [e[5] for e in cycle])
    rng.shuffle(idx)
    return [cycle[i][:5] + (1,) for i in idx], idx


BURN_IN_FAST = [10, 10] + list(range(10, 131, 10))      # burnIn (app/Definitions.hs:417-422)
BURN_IN_SLOW = list(range(100, 401, 20))


def run_iterations(ev, md, cycle, n_iterations: int, rng, seed: int, k0: int):
    """n_iterations of the cycle in random order on the evaluator's resident chains -> (accepted, proposed per entry, next k).
    All chains share the order (each chain is still a realisation of the same Markov chain)."""
    acc = np.zeros(len(cycle))
    prop = np.zeros(len(cycle))
    n_chains = ev.n_resident()
    k = k0
    for _ in range(n_iterations):
        lst, idx = random_order(cycle, rng)
        a, inv, k = ev.mh_cycle(lst, 1, seed=seed, iteration0=k)
        np.add.at(acc, idx, a.astype(float))
        np.add.at(prop, idx, float(n_chains))
    return acc, prop, k


def burn_in(ev, md, cycle, rng, seed: int = 0, k0: int = 0, periods=None):
    """BurnInWithCustomAutoTuning (app/Definitions.hs:417-422): tuning periods of the given lengths, an auto-tuning step after
    each.  The acceptance rate of an entry is the mean over the resident chains (they share the tuning parameters).
    -> (tuned cycle, next k)"""
    for n in (BURN_IN_FAST + BURN_IN_SLOW) if periods is None else periods:
        acc, prop, k0 = run_iterations(ev, md, cycle, n, rng, seed, k0)
        cycle = auto_tune(md, cycle, acc, prop)
    return cycle, k0
