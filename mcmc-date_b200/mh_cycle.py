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
