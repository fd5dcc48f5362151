"""Host-side restatement of the reference's `prepare` step and of its auxiliary-data loaders --
the data formats either side of the hot path.  numpy only.

  prepare_from_treelist   app/Main.hs:159-307   tree list -> mean vector, Sigma^-1, ln det Sigma, mean tree
  load_calibrations       lib/Mcmc/Tree/Prior/Node/Calibration.hs:186-319   CSV -> node index table
  load_constraints        lib/Mcmc/Tree/Prior/Node/Constraint.hs:275-374    CSV -> node index table
  load_braces             lib/Mcmc/Tree/Prior/Node/Brace.hs:152-192         JSON -> node index table
  mean_root_height        lib/Mcmc/Tree/Prior/Node/Calibration.hs:324-339   getMeanRootHeight
  initial_state           app/Definitions.hs:96-123                         initWith

Restrictions (documented): the trees of the list must already be rooted at the outgroup
and bifurcating (the reference re-roots with elynx `outgroup`, app/Main.hs:179-180; all of its own
tests/*/data/test.treelist files already are).  Constraint validation, conflict detection and the pruning of
duplicate / redundant constraints (Constraint.hs:117-146, 211-253, 306-374) are restated in load_constraints.
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


def prepare_from_treelist(text: str):
    """-> dict(parent, names, mean, cov, precision, logdet_sigma, mean_lengths).
    Drops the first len/6 trees (app/Main.hs:166-168), requires identical topology AND sub-tree
    order (:184-190), mean/covariance like hmatrix meanCov (unbiased, 1/(n-1)), then invlndet (:230)."""
    lines = [ln for ln in text.splitlines() if ln.strip()]
    trees = [_tree.flatten_preorder(_tree.parse_newick(ln)) for ln in lines]
    burn = len(trees) // 6
    trees = trees[burn:]
    parent, c0, c1, names, _ = trees[0]
    for p, _, _, nm, _ in trees:
        if not np.array_equal(p, parent) or nm != names:
            raise ValueError("prepare: A single topology and equal sub tree orders are required.")
    rows = np.array([_branches_row(p, ln) for p, _, _, _, ln in trees])
    mean = rows.mean(axis=0)
    cov = np.cov(rows, rowvar=False, ddof=1).reshape(len(mean), len(mean))
    if np.min(np.diag(cov)) <= 0:
        raise ValueError("prepare: Minimum variance is zero or negative.")
    sign, logdet = np.linalg.slogdet(cov)
    if sign != 1.0:
        raise ValueError("prepare: Determinant of covariance matrix is negative?")
    prec = np.linalg.inv(cov)
    prec = 0.5 * (prec + prec.T)
    all_len = np.array([ln for _, _, _, _, ln in trees]).mean(axis=0)  # mean tree incl. both root branches
    return {"parent": parent, "names": names, "mean": mean, "cov": cov, "precision": prec,
            "logdet_sigma": float(logdet), "mean_lengths": all_len}


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


def load_calibrations(text: str, parent, names):
    """CSV `Name,LeafA,LeafB,YoungAge,YoungProbabilityMass,OldAge,OldProbabilityMass` -> dict of arrays."""
    node, lo, lop, hi, hip, nm = [], [], [], [], [], []
    for row in csv.reader(io.StringIO(text)):
        if not row or row[0].strip() == "Name":
            continue
        name, la, lb = row[0], row[1].strip(), row[2].strip()
        a, pa, b, pb = (_opt(x) for x in row[3:7])
        if (a is None) != (pa is None) or (b is None) != (pb is None) or (a is None and b is None):
            raise ValueError(f"calibrationDataToCalibration: {name}: inconsistent boundaries")
        if a is not None and b is not None and a >= b:
            raise ValueError(f"calibrationDataToCalibration: {name}: Lower boundary larger equal upper boundary.")
        nm.append(name)
        node.append(_mrca(parent, names, la, lb))
        lo.append(a if a is not None else 0.0)
        lop.append(pa if pa is not None else 0.5)
        hi.append(b if b is not None else np.inf)
        hip.append(pb if pb is not None else 0.5)
    return {"names": nm, "node": np.array(node, np.int32), "lo": np.array(lo), "lo_p": np.array(lop),
            "hi": np.array(hi), "hi_p": np.array(hip)}


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
