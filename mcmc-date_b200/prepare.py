"""Host-side restatement of the reference's `prepare` step and of its auxiliary-data loaders --
the data formats either side of the hot path.  numpy only.

  prepare_from_treelist   app/Main.hs:159-307   tree list -> mean vector, Sigma^-1, ln det Sigma, mean tree
  load_calibrations       lib/Mcmc/Tree/Prior/Node/Calibration.hs:186-319   CSV -> node index table
  load_constraints        lib/Mcmc/Tree/Prior/Node/Constraint.hs:275-374    CSV -> node index table
  load_braces             lib/Mcmc/Tree/Prior/Node/Brace.hs:152-192         JSON -> node index table
  mean_root_height        lib/Mcmc/Tree/Prior/Node/Calibration.hs:324-339   getMeanRootHeight
  initial_state           app/Definitions.hs:96-123                         initWith

  load_calibrations_from_tree  lib/Mcmc/Tree/Prior/Node/CalibrationFromTree.hs:27-130  MCMCtree-labelled Newick tree -> table
  outgroup                app/Main.hs:179-180 (elynx-tree `outgroup`, un-vendored: restated from its published algorithm)

Constraint validation, conflict detection and the pruning of duplicate / redundant constraints
(Constraint.hs:117-146, 211-253, 306-374) are restated in load_constraints.  Not restated: glasso (`SparseMultivariateNormal`,
app/Main.hs:257-277, a Fortran dependency) -- a sparse `.data` file written by the reference is read as is.
"""
from __future__ import annotations

import csv
import io
import json

import numpy as np

from . import model as _m
from . import tree as _tree


def _branches_row(parent, lengths):
    """sumFirstTwo . getBranches (app/Tools.hs:36-48) for one tree."""
    k = _tree.branch_index(parent)
    row = np.zeros(len(parent) - 2)
    np.add.at(row, k[1:], lengths[1:])
    return row


def _leaf_set(nd):
    if not nd["children"]:
        return frozenset([nd["name"]])
    return frozenset().union(*[_leaf_set(c) for c in nd["children"]])


def _with_length(nd, length):
    return {"name": nd["name"], "length": length, "children": nd["children"]}


def _half(x):
    return None if x is None else 0.5 * x       # elynx `split` on a branch length


def _plus(a, b):
    return None if a is None and b is None else (a or 0.0) + (b or 0.0)   # `<>` on branch lengths


def outgroup(og, tree):
    """Root `tree` (nested dict from tree.parse_newick) at the branch that separates the leaf set `og` from the rest:
    elynx-tree `outgroup` (ELynx.Tree.Rooted, rev 34fc58b0; the package is not vendored in the reference checkout, so this
    follows its published algorithm and documentation, as called at app/Main.hs:179-180).

    * a leaf root or a root of degree one is an error;
    * a multifurcating root first gets a bifurcating root above it by splitting the leftmost branch in two halves:
      `[leftmost child, (the other children)]`, the old root label moving to the second node;
    * a tree that is already rooted at the bipartition is returned unchanged (`roots t` starts with `t` itself);
    * otherwise the root moves down to the branch above the sub tree X whose leaves are `og` or its complement; that branch is
      split in halves.  `descend` puts the part that hangs "upside down" FIRST and X SECOND; the upside-down node keeps the
      label of X's former parent and lists [the former up-going part (branch = the two halves / root branches joined),
      then X's former siblings in their old order].
    The sub-tree order is observable: the mean tree takes the order of the re-rooted tree list (app/Main.hs:289-291), and
    every node index follows from it."""
    og = frozenset(og)
    if not og:
        raise ValueError("outgroup: Outgroup is empty.")
    ch = tree["children"]
    if len(ch) == 0:
        raise ValueError("outgroup: Root node is a leaf.")
    if len(ch) == 1:
        raise ValueError("outgroup: Root node has degree two.")
    if len(ch) > 2:
        o = ch[0]
        h = _half(o["length"])
        tree = {"name": "", "length": tree["length"],
                "children": [_with_length(o, h), {"name": tree["name"], "length": h, "children": ch[1:]}]}
        ch = tree["children"]
    allv = _leaf_set(tree)
    if not og < allv:
        raise ValueError("outgroup: Outgroup is not a proper subset of the leaves.")
    target = (og, allv - og)
    tl, tr = ch
    if _leaf_set(tl) in target:
        return tree

    def descend(tc, d):
        """the root sits above d, tc is everything else (its branch = the other root branch); -> re-rooted tree or None"""
        if not d["children"]:
            return None
        tc2 = _with_length(tc, _plus(tc["length"], d["length"]))
        kids = d["children"]
        cfs = [[tc2] + kids[:i] + kids[i + 1:] for i in range(len(kids))]
        for dd, f in zip(kids, cfs):
            if _leaf_set(dd) in target:
                h = _half(dd["length"])
                return {"name": tree["name"], "length": tree["length"],
                        "children": [{"name": d["name"], "length": h, "children": f}, _with_length(dd, h)]}
        for dd, f in zip(kids, cfs):
            if og <= _leaf_set(dd) or (allv - og) <= _leaf_set(dd):
                h = _half(dd["length"])
                return descend({"name": d["name"], "length": h, "children": f}, _with_length(dd, h))
        return None

    out = descend(tr, tl) or descend(tl, tr)
    if out is None:
        raise ValueError("outgroup: No rooted tree has the required bipartition (the outgroup is not monophyletic).")
    return out


def _topology_key(nd, ordered):
    """T.fromBranchLabelTree: the topology with leaf labels only; ordered = sub-tree order matters (`nub`, app/Main.hs:184),
    unordered = `T.equal'` (:198)"""
    if not nd["children"]:
        return nd["name"]
    ks = [_topology_key(c, ordered) for c in nd["children"]]
    return tuple(ks) if ordered else frozenset(ks)


def prepare_from_treelist(text: str, rooted_tree_text: str | None = None):
    """-> dict(parent, names, mean, cov, precision, logdet_sigma, mean_lengths).
    `prepare` (app/Main.hs:159-307): drop the first len/6 trees (:166-168); no duplicate leaves (:171-173); re-root every tree
    at the outgroup of the given rooted tree (:176-180; skipped when no rooted tree is given -- the list must then already be
    rooted); identical topology AND sub-tree order (:184-190); same topology as the rooted tree up to sub-tree order
    (:195-205); mean / covariance like hmatrix meanCov (unbiased, 1/(n-1)); invlndet (:230) -- the inverse is NOT symmetrised
    here, as in the reference's `.data` file (mcd_create takes the symmetric part: same quadratic form, and its gradient)."""
    lines = [ln for ln in text.replace(";", ";\n").splitlines() if ln.strip()]
    nested_all = [_tree.parse_newick(ln) for ln in lines]
    for t in nested_all:
        lv = [n for n in _all_leaves(t)]
        if len(set(lv)) != len(lv):
            raise ValueError("prepare: Trees have duplicate leaves.")
    nested = nested_all[len(nested_all) // 6:]
    rooted = None
    if rooted_tree_text is not None:
        rooted = _tree.parse_newick(rooted_tree_text)
        if len(rooted["children"]) != 2:
            raise ValueError("bipartition: Root node is not bifurcating.")
        og = min((_leaf_set(c) for c in rooted["children"]), key=lambda x: sorted(x))   # fst . fromBipartition
        nested = [outgroup(og, t) for t in nested]
    keys = {_topology_key(t, True) for t in nested}
    if len(keys) != 1:
        raise ValueError("prepare: A single topology and equal sub tree orders are required.")
    if rooted is not None and _topology_key(rooted, False) != _topology_key(nested[0], False):
        raise ValueError("prepare: A single topology is required.")
    trees = [_tree.flatten_preorder(t) for t in nested]
    parent, c0, c1, names, _ = trees[0]
    rows = np.array([_branches_row(p, ln) for p, _, _, _, ln in trees])
    mean = rows.mean(axis=0)
    cov = np.cov(rows, rowvar=False, ddof=1).reshape(len(mean), len(mean))
    if np.min(np.diag(cov)) <= 0:
        raise ValueError("prepare: Minimum variance is zero or negative.")
    sign, logdet = np.linalg.slogdet(cov)
    if sign != 1.0:
        raise ValueError("prepare: Determinant of covariance matrix is negative?")
    prec = np.linalg.inv(cov)
    all_len = np.array([ln for _, _, _, _, ln in trees]).mean(axis=0)  # mean tree incl. both root branches
    return {"parent": parent, "names": names, "mean": mean, "cov": cov, "precision": prec,
            "logdet_sigma": float(logdet), "mean_lengths": all_len}


def _all_leaves(nd):
    if not nd["children"]:
        yield nd["name"]
    for c in nd["children"]:
        yield from _all_leaves(c)


def _mrca(parent, names, leaf_a: str, leaf_b: str) -> int:
    ia, ib = names.index(leaf_a), names.index(leaf_b)
    anc = set()
    x = ia
    while x >= 0:
        anc.add(x)
        x = int(parent[x])
    x = ib
    while x not in anc:
        x = int(parent[x])
    return x


def _opt(s):
    s = s.strip()
    return float(s) if s else None


def load_calibrations(text: str, parent, names, on_problem: str = "warn", log=None):
    """CSV `Name,LeafA,LeafB,YoungAge,YoungProbabilityMass,OldAge,OldProbabilityMass` -> dict of arrays
    (loadCalibrations, Calibration.hs:283-319)."""
    rows = []
    for row in csv.reader(io.StringIO(text)):
        if not row or row[0].strip() == "Name":
            continue
        rows.append((row[0], row[1].strip(), row[2].strip()) + tuple(_opt(x) for x in row[3:7]))
    if not rows:
        raise ValueError("loadCalibrations: No calibrations found in file.")
    return _calibration_table(rows, parent, names, on_problem, log)


_MCMCTREE_LABEL = None


def _parse_mcmctree_label(label: str):
    """pABounded (CalibrationFromTree.hs:36-88): `L(l[,p,c[,pL]])`, `U(u[,pU])`, `B(l,u[,pL[,pU]])` at the START of a node label
    (attoparsec `parseOnly` does not demand the end of the input) -> (lo, lo_p, hi, hi_p) with None for absent bounds, or
    None when the label is no calibration.  Missing probability masses default to 0.01 (defPM, :93-95); the Cauchy
    parameters of L are read and ignored."""
    import re
    global _MCMCTREE_LABEL
    if _MCMCTREE_LABEL is None:
        num = r"[+-]?\d+(?:\.\d+)?(?:[eE][+-]?\d+)?"   # attoparsec `double`: at least one leading digit (".06" does not parse)
        _MCMCTREE_LABEL = {
            "L": re.compile(rf"L\(({num})(?:,({num}))?(?:,({num}))?(?:,({num}))?\)"),
            "U": re.compile(rf"U\(({num})(?:,({num}))?\)"),
            "B": re.compile(rf"B\(({num}),({num})(?:,({num}))?(?:,({num}))?\)"),
        }
    f = lambda x: None if x is None else float(x)
    d = lambda x: 0.01 if x is None else float(x)
    m = _MCMCTREE_LABEL["L"].match(label)
    if m:
        return float(m.group(1)), d(m.group(4)), None, None
    m = _MCMCTREE_LABEL["U"].match(label)
    if m:
        return None, None, float(m.group(1)), d(m.group(2))
    m = _MCMCTREE_LABEL["B"].match(label)
    if m:
        return float(m.group(1)), d(m.group(3)), float(m.group(2)), d(m.group(4))
    return None


def load_calibrations_from_tree(text: str, parent, names, on_problem: str = "warn", log=None):
    """loadCalibrationsFromTree (CalibrationFromTree.hs:97-130): every node of the Newick tree whose label parses as an MCMCtree
    bound becomes a calibration on the MRCA of the node's leftmost and rightmost leaf (filterBoundedNodes, :107-116), named
    `<leftmost>-<rightmost>`, in pre-order of the labelled tree; then the same checks as the CSV loader
    (checkAndConvertCalibrationData, Calibration.hs:254-281)."""
    t = _tree.parse_newick(text)
    rows = []

    def left(nd):
        return nd["name"] if not nd["children"] else left(nd["children"][0])

    def right(nd):
        return nd["name"] if not nd["children"] else right(nd["children"][-1])

    def walk(nd):
        b = _parse_mcmctree_label(nd["name"])
        if b is not None:
            rows.append((left(nd) + "-" + right(nd), left(nd), right(nd)) + b)
        for c in nd["children"]:
            walk(c)

    walk(t)
    if not rows:
        raise ValueError("loadCalibrationsFromTree: no calibrations found")
    return _calibration_table(rows, parent, names, on_problem, log)


def _calibration_table(rows, parent, names, on_problem="warn", log=None):
    """calibrationDataToCalibration + checkAndConvertCalibrationData (Calibration.hs:210-281) on rows
    (name, leafA, leafB, lo, lo_p, hi, hi_p) with None for absent fields"""
    node, lo, lop, hi, hip, nm = [], [], [], [], [], []
    for name, la, lb, a, pa, b, pb in rows:
        err = lambda m: ValueError(f"calibrationDataToCalibration: {name}: {m}")
        if a is None and pa is not None:
            raise err("Lower probability mass given but no lower boundary.")
        if b is None and pb is not None:
            raise err("Upper probability mass given but no upper boundary.")
        if a is not None and pa is None:
            raise err("Lower boundary given but no lower probability mass.")
        if b is not None and pb is None:
            raise err("Upper boundary given but no upper probability mass.")
        if a is None and b is None:
            raise err("No boundaries provided.")
        if a is not None and b is not None and a >= b:
            raise err("Lower boundary larger equal upper boundary.")
        for pm in (pa, pb):
            if pm is not None and not 0.0 < pm < 1.0:
                raise err("probabilityMass: out of (0, 1)")
        if a is not None and a <= 0:
            raise err("positiveLowerBoundary: Zero or negative value.")
        if b is not None and b <= 0:
            raise err("positiveUpperBoundary: Zero or negative value.")
        nm.append(name)
        node.append(_mrca(parent, names, la, lb))
        lo.append(a if a is not None else 0.0)
        lop.append(pa if pa is not None else 0.5)
        hi.append(b if b is not None else np.inf)
        hip.append(pb if pb is not None else 0.5)
    dups = sorted({n for n in node if node.count(n) > 1})
    if dups:
        msg = "loadCalibrations: Duplicate/conflicting/redundant calibrations have been detected."
        if on_problem == "error":
            raise ValueError(msg)
        (log or (lambda m: None))("WARNING: " + msg + f" (nodes {dups})")
    return {"names": nm, "node": np.array(node, np.int32), "lo": np.array(lo, float), "lo_p": np.array(lop, float),
            "hi": np.array(hi, float), "hi_p": np.array(hip, float)}


def _is_ancestor(parent, x: int, y: int) -> bool:
    """x is an ancestor of y or y itself (the reference compares node PATHS with isPrefixOf, Internal.hs:69-75)"""
    while y >= 0:
        if y == x:
            return True
        y = int(parent[y])
    return False


def load_constraints(text: str, parent, names, on_problem: str = "drop", log=None):
    """CSV `Name,YoungerLeafA,YoungerLeafB,OlderLeafA,OlderLeafB,ProbabilityMass` -> validated, pruned table
    (loadConstraints, Constraint.hs:275-374):

    * a constraint whose two nodes coincide, or whose younger node is an ancestor of the older one, is an error
      (validateConstraint, :117-146); one whose older node is an ancestor of the younger one is vacuous and is
      dropped with a warning (`on_problem="drop"`, WarnAboutAndDropProblematicConstraints) or an error (`"error"`);
    * conflicting pairs are an error: given a < b, the constraint c < d conflicts iff c is an ancestor of b and d a
      descendant of a or of b (isConflictingWith, :233-235);
    * of constraints on the same pair of nodes the later ones are removed (duplicate, :101-102, :333-349);
    * given a < b, the constraint c < d is redundant iff c is a descendant of a and d an ancestor of b
      (isRedundantWith, :222-224); redundant ones are removed (:351-369).

    "ancestor" / "descendant" include the node itself, as path-prefix tests do."""
    say = log if log is not None else (lambda msg: None)
    rows = []
    for row in csv.reader(io.StringIO(text)):
        if not row or row[0].strip() == "Name":
            continue
        name, pm = row[0], float(row[5])
        if not 0.0 < pm < 1.0:
            raise ValueError(f"constraintDataToConstraint: {name}: probabilityMass: out of (0, 1)")
        y = _mrca(parent, names, row[1].strip(), row[2].strip())
        o = _mrca(parent, names, row[3].strip(), row[4].strip())
        if y == o:
            raise ValueError(f"validateConstraint: {name!r}: Bogus constraint; both nodes are equal (?).")
        if _is_ancestor(parent, y, o):
            raise ValueError(f"validateConstraint: {name!r}: Bogus constraint; younger node is direct ancestor of older node (?).")
        if _is_ancestor(parent, o, y):
            msg = f"validateConstraint: {name!r}: Redundant constraint; old node is direct ancestor of young node."
            if on_problem == "error":
                raise ValueError(msg)
            say("WARNING: Dropping constraint: " + msg)
            continue
        rows.append((name, y, o, pm))
    if not rows and not text.strip():
        raise ValueError("loadConstraints: No constraints found.")
    anc = lambda x, y: _is_ancestor(parent, x, y)          # A(x, y)
    desc = lambda x, y: _is_ancestor(parent, y, x)         # D(x, y)
    # conflicts (validateWith: ordered pairs of different constraints)
    for l in rows:
        for r in rows:
            if l != r and anc(r[1], l[2]) and (desc(r[2], l[1]) or desc(r[2], l[2])):
                raise ValueError(f"loadConstraints: Constraint {r[0]} is conflicting given constraint {l[0]}.")
    # duplicates (validateWithCommutative: the later one of each pair goes)
    dup = []
    for i, l in enumerate(rows):
        for r in rows[i + 1:]:
            if l != r and l[1] == r[1] and l[2] == r[2] and r not in dup:
                dup.append(r)
                say(f"Constraints {l[0]} and {r[0]} affect the same nodes.")
    for r in dup:
        rows.remove(r)
    # redundancies
    red = []
    for l in rows:
        for r in rows:
            if l != r and desc(r[1], l[1]) and anc(r[2], l[2]) and r not in red:
                red.append(r)
                say(f"Constraint {r[0]} is redundant given constraint {l[0]}.")
    for r in red:
        rows.remove(r)
    return {"names": [r[0] for r in rows], "young": np.array([r[1] for r in rows], np.int32),
            "old": np.array([r[2] for r in rows], np.int32), "p": np.array([r[3] for r in rows], dtype=float)}


def load_braces(text: str, parent, names):
    """JSON list of {braceDataName, braceDataNodes: [[leafA, leafB], ...], braceDataStandardDeviation};
    nodes sorted by index (Brace.hs:112)."""
    off, idx, sd, nm = [0], [], [], []
    for b in json.loads(text):
        nodes = sorted(_mrca(parent, names, a, c) for a, c in b["braceDataNodes"])
        if len(nodes) < 2:
            raise ValueError("brace: need at least two nodes")
        idx += nodes
        off.append(len(idx))
        sd.append(float(b["braceDataStandardDeviation"]))
        nm.append(b["braceDataName"])
    return {"names": nm, "off": np.array(off, np.int32), "node": np.array(idx, np.int32), "sd": np.array(sd)}


def mean_root_height(cal) -> float:
    """getMeanRootHeight: exactly one root calibration with a finite upper bound -> its mean, else 1.0."""
    roots = [i for i, n in enumerate(cal["node"]) if n == 0]
    if len(roots) != 1 or not np.isfinite(cal["hi"][roots[0]]):
        return 1.0
    i = roots[0]
    return float((cal["lo"][i] + cal["hi"][i]) / 2.0) if cal["lo"][i] > 0 else float(cal["hi"][i] / 2.0)


def ultrametric_heights(parent, lengths):
    """makeUltrametric (extend terminal branches) + normalizeHeight + toHeightTreeUltrametric:
    node height = longest path to a leaf below, divided by the root's; leaves 0."""
    n = len(parent)
    h = np.zeros(n)
    for i in range(n - 1, 0, -1):
        h[parent[i]] = max(h[parent[i]], h[i] + lengths[i])
    c0, _ = _tree.children_from_parent(parent)
    h[c0 < 0] = 0.0
    return h / h[0]


def initial_state(parent, mean_lengths):
    """initWith (app/Definitions.hs:96-123): lambda = mu = H = m = v = 1, rates 1 with stem 0, heights
    from the mean tree (zero branches replaced by the average)."""
    ln = np.array(mean_lengths, float)
    nz = ln[1:][ln[1:] > 0]
    ln[1:][ln[1:] <= 0] = nz.mean() if len(nz) else 1.0
    N = len(parent)
    x = np.ones(5 + 2 * N)
    x[3:3 + N] = ultrametric_heights(parent, ln)
    x[5 + N] = 0.0
    return x


# ----------------------------------------------------------------------------- on-disk formats
def write_data_file(path, kind, mean=None, precision=None, logdet_sigma=None, variances=None, sparse=None):
    """`<name>.data` as `prepare` writes it (app/Main.hs:75-81,286): aeson `deriveJSON defaultOptions` of
    LikelihoodDataStore, i.e. {"tag": "FullS", "contents": [mu, rows, logdet]} etc."""
    if kind == "FullS":
        obj = {"tag": kind, "contents": [list(map(float, mean)), [list(map(float, r)) for r in precision], float(logdet_sigma)]}
    elif kind == "SparseS":
        r, c, v = sparse
        obj = {"tag": kind, "contents": [list(map(float, mean)), [[[int(i), int(j)], float(x)] for i, j, x in zip(r, c, v)],
                                         float(logdet_sigma)]}
    elif kind == "UnivariateS":
        obj = {"tag": kind, "contents": [list(map(float, mean)), list(map(float, variances))]}
    elif kind == "NoLikelihoodS":
        obj = {"tag": kind}
    else:
        raise ValueError("unknown LikelihoodDataStore constructor " + kind)
    with open(path, "w") as f:
        json.dump(obj, f)


def read_data_file(path):
    """getData (app/Main.hs:85-99) -> dict(likelihood, mean, precision | variances | sparse, logdet_sigma)"""
    with open(path) as f:
        obj = json.load(f)
    tag, c = obj["tag"], obj.get("contents")
    if tag == "FullS":
        return {"likelihood": _m.LIK_FULL, "mean": np.array(c[0]), "precision": np.array(c[1]), "logdet_sigma": float(c[2])}
    if tag == "SparseS":
        r = np.array([e[0][0] for e in c[1]], np.int32)
        col = np.array([e[0][1] for e in c[1]], np.int32)
        v = np.array([e[1] for e in c[1]], float)
        return {"likelihood": _m.LIK_SPARSE, "mean": np.array(c[0]), "sparse": (r, col, v), "logdet_sigma": float(c[2])}
    if tag == "UnivariateS":
        var = np.array(c[1])
        return {"likelihood": _m.LIK_UNIVARIATE, "mean": np.array(c[0]), "variances": var,
                "logdet_sigma": float(np.sum(np.log(var)))}   # logSigmaSquaredProduct (app/Probability.hs:274)
    if tag == "NoLikelihoodS":
        return {"likelihood": _m.LIK_NONE}
    raise ValueError("getData: Could not decode data file: " + path)


def mean_tree_newick(parent, names, mean_lengths) -> str:
    """`<name>.meantree` (app/Main.hs:289-307): the first tree's topology with mean branch lengths; unnamed
    (inner) nodes get their running pre-order index as label (assignIndices, app/Tools.hs:73-81)."""
    c0, c1 = _tree.children_from_parent(np.asarray(parent))

    def label(i):
        nm = names[i]
        return nm if nm and not nm.isdigit() else str(i)

    def rec(i):
        s = label(i) + ":" + repr(float(mean_lengths[i]))
        if c0[i] < 0:
            return s
        return "(" + rec(int(c0[i])) + "," + rec(int(c1[i])) + ")" + s

    return rec(0) + ";"


def model_from_data_file(data_path, meantree_text, calibrations_text=None, constraints_text=None, braces_text=None,
                         clock_model=_m.UNCORRELATED_LOGNORMAL):
    """run / continue mode inputs (app/Main.hs:370-417): `<name>.data` + `<name>.meantree` (+ auxiliary
    files) -> (ModelDesc, initial state)."""
    dat = read_data_file(data_path)
    parent, c0, c1, names, lengths = _tree.flatten_preorder(_tree.parse_newick(meantree_text))
    cal = load_calibrations(calibrations_text, parent, names) if calibrations_text else None
    con = load_constraints(constraints_text, parent, names) if constraints_text else None
    br = load_braces(braces_text, parent, names) if braces_text else None
    kw = {}
    if cal:
        kw.update(cal_node=cal["node"], cal_lo=cal["lo"], cal_lo_p=cal["lo_p"], cal_hi=cal["hi"], cal_hi_p=cal["hi_p"])
    if con:
        kw.update(con_young=con["young"], con_old=con["old"], con_p=con["p"])
    if br:
        kw.update(brace_off=br["off"], brace_node=br["node"], brace_sd=br["sd"])
    lik = dat["likelihood"]
    K = len(parent) - 2
    md = _m.ModelDesc(parent=parent, mean=dat.get("mean", np.zeros(K)),
                      precision=dat.get("precision", dat.get("variances", np.zeros(0))),
                      logdet_sigma=dat.get("logdet_sigma", 0.0), clock_model=clock_model, likelihood=lik,
                      ht=mean_root_height(cal) if cal else 1.0, **kw)
    if lik == _m.LIK_SPARSE:
        md.sparse_row, md.sparse_col, md.sparse_val = dat["sparse"]
    return md, initial_state(parent, lengths)


def model_from_files(treelist_text, calibrations_text=None, constraints_text=None, braces_text=None,
                     clock_model=_m.UNCORRELATED_LOGNORMAL, likelihood=_m.LIK_FULL):
    """Everything `getMcmcProps` assembles (app/Main.hs:370-457) -> (ModelDesc, prepared dict)."""
    pr = prepare_from_treelist(treelist_text)
    parent, names = pr["parent"], pr["names"]
    cal = load_calibrations(calibrations_text, parent, names) if calibrations_text else None
    con = load_constraints(constraints_text, parent, names) if constraints_text else None
    br = load_braces(braces_text, parent, names) if braces_text else None
    prec = pr["precision"] if likelihood == _m.LIK_FULL else np.diag(pr["cov"]).copy()
    logdet = pr["logdet_sigma"] if likelihood == _m.LIK_FULL else float(np.sum(np.log(np.diag(pr["cov"]))))
    kw = {}
    if cal:
        kw.update(cal_node=cal["node"], cal_lo=cal["lo"], cal_lo_p=cal["lo_p"], cal_hi=cal["hi"], cal_hi_p=cal["hi_p"])
    if con:
        kw.update(con_young=con["young"], con_old=con["old"], con_p=con["p"])
    if br:
        kw.update(brace_off=br["off"], brace_node=br["node"], brace_sd=br["sd"])
    md = _m.ModelDesc(parent=parent, mean=pr["mean"], precision=prec, logdet_sigma=logdet, clock_model=clock_model,
                      likelihood=likelihood, ht=mean_root_height(cal) if cal else 1.0, **kw)
    pr.update(calibrations=cal, constraints=con, braces=br)
    return md, pr
