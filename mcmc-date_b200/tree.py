"""Flattened rooted trees in pre-order -- the host-side layout the CUDA path consumes.

The reference keeps `Tree e a` values (elynx-tree) and folds them in pre-order
(lib/Mcmc/Tree/Types.hs:91-95, 146-150); calibrations/constraints/braces address nodes by their
pre-order index (`identify`, lib/Mcmc/Tree/Prior/Node/Calibration.hs:173).  Here a tree is three
int32 arrays over pre-order node ids (root = 0): parent, child0, child1 (-1 = none).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- Newick
def parse_newick(text: str):
    """Parse one Newick tree -> nested dict {name, length, children}.  Handles plain and quoted labels,
    branch lengths and [comments]; enough for the reference's tests/*/data/*.tree files."""
    s = text.strip()
    if s.endswith(";"):
        s = s[:-1]
    pos = 0

    def skip_comment():   # white space and [comments]
        nonlocal pos
        while pos < len(s) and (s[pos] == "[" or s[pos].isspace()):
            pos = s.index("]", pos) + 1 if s[pos] == "[" else pos + 1

    def node():
        nonlocal pos
        children = []
        skip_comment()
        if pos < len(s) and s[pos] == "(":
            pos += 1
            while True:
                children.append(node())
                skip_comment()
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"newick: unexpected {s[pos]!r} at {pos}")
        skip_comment()
        if pos < len(s) and s[pos] in "'\"":   # quoted label (MCMCtree calibrations: 'B(6,8,2.5e-2,2.5e-2)')
            q = s[pos]
            end = s.index(q, pos + 1)
            name = s[pos + 1:end]
            pos = end + 1
        else:
            start = pos
            while pos < len(s) and s[pos] not in ",():;[":
                pos += 1
            name = s[start:pos].strip()
        skip_comment()
        length = None
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",();[":
                pos += 1
            length = float(s[start:pos])
        skip_comment()
        return {"name": name, "length": length, "children": children}

    return node()


def flatten_preorder(tree):
    """nested dict -> (parent, child0, child1, names, lengths) in pre-order (children in file order,
    which is what the reference relies on, app/Main.hs:184-190).  Bifurcating trees only."""
    parent, names, lengths, kids = [], [], [], []
    stack = [(tree, -1)]
    while stack:
        nd, p = stack.pop()
        i = len(parent)
        parent.append(p)
        names.append(nd["name"])
        lengths.append(nd["length"] if nd["length"] is not None else 0.0)
        kids.append([])
        if p >= 0:
            kids[p].append(i)
        ch = nd["children"]
        if len(ch) not in (0, 2):
            raise ValueError("flatten_preorder: tree is not bifurcating")
        for c in reversed(ch):
            stack.append((c, i))
    n = len(parent)
    child0 = np.full(n, -1, np.int32)
    child1 = np.full(n, -1, np.int32)
    for i, k in enumerate(kids):
        if k:
            child0[i], child1[i] = k
    return np.asarray(parent, np.int32), child0, child1, names, np.asarray(lengths, np.float64)


def children_from_parent(parent: np.ndarray):
    """child0/child1 from a pre-order parent array (first child = lower index)."""
    n = len(parent)
    child0 = np.full(n, -1, np.int32)
    child1 = np.full(n, -1, np.int32)
    for i in range(1, n):
        p = parent[i]
        if child0[p] < 0:
            child0[p] = i
        elif child1[p] < 0:
            child1[p] = i
        else:
            raise ValueError("children_from_parent: node with more than two children")
    return child0, child1


def node_heights_from_lengths(parent, lengths):
    """Relative node heights (root = 1, leaves 0 for an ultrametric tree) from branch lengths --
    what `toHeightTreeUltrametric . normalizeHeight` produce (app/Definitions.hs:96-123)."""
    n = len(parent)
    depth = np.zeros(n)
    for i in range(1, n):
        depth[i] = depth[parent[i]] + lengths[i]
    total = depth.max()
    h = (total - depth) / total
    h[0] = 1.0
    return h


# ----------------------------------------------------------------------------- index maps
def branch_index(parent: np.ndarray) -> np.ndarray:
    """node -> MVN dimension k(i) after getBranches + sumFirstTwo (app/Tools.hs:36-48), closed form
    (SURVEY.md R2): with s_l = #nodes in the left root subtree, k(1) = k(1+s_l) = 0,
    k(i) = i-1 for 2 <= i <= s_l, k(i) = i-2 for i >= s_l+2; root -> -1."""
    n = len(parent)
    roots = np.nonzero(parent == 0)[0]
    if len(roots) != 2:
        raise ValueError("getBranches: Root node is not bifurcating.")
    r = int(roots[1])  # = 1 + s_l
    k = np.empty(n, np.int32)
    k[0] = -1
    idx = np.arange(n)
    k[1:r] = idx[1:r] - 1
    k[r:] = idx[r:] - 2
    k[1] = 0
    k[r] = 0
    return k


# ----------------------------------------------------------------------------- synthetic trees
def random_topology(n_leaves: int, rng: np.random.Generator) -> np.ndarray:
    """Random bifurcating topology by recursive random splits; returns the pre-order parent array."""
    parent = []
    stack = [(n_leaves, -1)]
    while stack:
        n, p = stack.pop()
        i = len(parent)
        parent.append(p)
        if n > 1:
            k = int(rng.integers(1, n))
            stack.append((n - k, i))  # right subtree (popped after the whole left subtree)
            stack.append((k, i))
    return np.asarray(parent, np.int32)


def random_ultrametric_heights(parent: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Root 1, leaves 0, inner heights = sorted U(0,1) handed out along a random topological order,
    so every parent is older than its children."""
    child0, child1 = children_from_parent(parent)
    n = len(parent)
    inner = [i for i in range(1, n) if child0[i] >= 0]
    vals = np.sort(rng.uniform(0.02, 0.98, size=len(inner)))[::-1]
    h = np.zeros(n)
    h[0] = 1.0
    avail = [c for c in (child0[0], child1[0]) if child0[c] >= 0]
    for v in vals:
        j = int(rng.integers(0, len(avail)))
        avail[j], avail[-1] = avail[-1], avail[j]
        node = avail.pop()
        h[node] = v
        for c in (child0[node], child1[node]):
            if child0[c] >= 0:
                avail.append(c)
    return h
