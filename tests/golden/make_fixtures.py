"""Generate tests/golden/*.npz from the reference's own datasets (run HERE, where /root/reference
exists; the fixtures travel to the GPU box, the reference does not).

    python tests/golden/make_fixtures.py

For each of the reference's test configurations (BASELINE.json configs[0..3]) this writes
  <name>.npz : the flattened model (topology from the first tree of the tree list, mean vector /
               precision / ln det from a restatement of `prepare`, calibration / constraint / brace
               node tables, ht), a batch of seeded states (valid states, the reference's initial state,
               and edge states: non-positive branch, zero rate, v <= 0, H <= 0, lambda < 0, NaN rate,
               near-critical birth/death rates), and for each of the four clock models the ORACLE's
               outputs, status words, analytic gradient and (first four states) dual-number gradient.
Integer fixtures (parent arrays, node indices) are additionally pinned against the values derived in
SURVEY.md section 8(c) by tests/test_fixtures.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mcmc_date_b200 import model, prepare, synth, tree  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def edge_states(md, X):
    """append edge-case states derived from X[0]"""
    N = md.n_nodes
    base = X[0].copy()
    inner = [i for i in range(1, N) if md.child0[i] >= 0]
    out = []
    e = base.copy(); e[3 + inner[0]] = 1.5; out.append(e)                 # child older than the root: branch <= 0
    e = base.copy(); e[3 + inner[-1]] = -0.1; out.append(e)               # negative height
    e = base.copy(); e[5 + N + 2] = 0.0; out.append(e)                    # zero rate
    e = base.copy(); e[5 + N + 3] = -1.0; out.append(e)                   # negative rate
    e = base.copy(); e[4 + N] = 0.0; out.append(e)                        # v = 0
    e = base.copy(); e[4 + N] = -0.5; out.append(e)                       # v < 0
    e = base.copy(); e[2] = 0.0; out.append(e)                            # H = 0
    e = base.copy(); e[2] = -3.0; out.append(e)                           # H < 0
    e = base.copy(); e[0] = -0.2; out.append(e)                           # lambda < 0
    e = base.copy(); e[1] = -0.2; out.append(e)                           # mu < 0
    e = base.copy(); e[3 + N] = -1e-3; out.append(e)                      # m < 0
    e = base.copy(); e[5 + N + 4] = np.nan; out.append(e)                 # NaN rate (proposals emit these)
    e = base.copy(); e[1] = e[0] + 3e-7; out.append(e)                    # near-critical
    e = base.copy(); e[1] = e[0]; out.append(e)                           # exactly critical
    e = base.copy(); e[2] = 1.0; out.append(e)                            # H == 1: transformCalibration shortcut
    e = base.copy(); e[0] = 0.0; out.append(e)                            # lambda = 0
    return np.vstack([X, np.array(out)])


def make(name, md, pr, init_state, n_valid=24, seed=0):
    heights = init_state[3:3 + md.n_nodes]
    X = synth.synthetic_states(md, heights, n_valid, seed=synth.BASE_SEED + 100 + seed)
    X = np.vstack([init_state[None, :], X])
    if md.calibrations_available:
        X[0, 2] = md.ht  # a start inside the calibrations (the reference itself starts at H = 1, app/Definitions.hs:101, and lets the burn-in climb)
    X = edge_states(md, X)
    arrs = dict(
        parent=md.parent, mean=md.mean, precision=md.precision, logdet_sigma=md.logdet_sigma, ht=md.ht,
        likelihood=md.likelihood, cal_node=md.cal_node, cal_lo=md.cal_lo, cal_lo_p=md.cal_lo_p, cal_hi=md.cal_hi,
        cal_hi_p=md.cal_hi_p, con_young=md.con_young, con_old=md.con_old, con_p=md.con_p, brace_off=md.brace_off,
        brace_node=md.brace_node, brace_sd=md.brace_sd, states=X, n_valid=n_valid + 1,
        leaf_names=np.array(pr["names"]) if pr else np.array([]))
    for clock in range(4):
        md.clock_model = clock
        orc = O.Oracle(md)
        out, grad, st = orc.eval_grad(X)
        gd = np.array([orc.grad_dual(x) for x in X[:4]])
        arrs[f"out_{clock}"] = out
        arrs[f"grad_{clock}"] = grad
        arrs[f"status_{clock}"] = st
        arrs[f"graddual_{clock}"] = gd
        if clock == 1:
            arrs["branch_index"] = orc.branch_index()
            arrs["mask"] = orc.mask
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, "N", md.n_nodes, "K", md.dim, "states", X.shape, "ht", md.ht,
          "cal", md.cal_node.tolist(), "con", md.con_young.tolist(), md.con_old.tolist(), "brace", md.brace_node.tolist())


def dataset(name, d, cal=True, con=False, br=False, seed=0, trees="test.treelist"):
    md, pr = prepare.model_from_files(
        read(f"tests/{d}/data/{trees}"),
        read(f"tests/{d}/data/calibrations.csv") if cal else None,
        read(f"tests/{d}/data/constraints.csv") if con else None,
        read(f"tests/{d}/data/braces.json") if br else None)
    init = prepare.initial_state(md.parent, pr["mean_lengths"])
    make(name, md, pr, init, seed=seed)


def mtcdnapri():
    """7-taxon primate set of bench/comparison_with_mcmctree (BASELINE.json configs[3]).  Topology: the tree list re-rooted
    at the outgroup of the rooted tree (prepare.outgroup; the list is already rooted there, sub-tree order (apes, gibbon)).
    Calibrations: parsed from the MCMCtree-labelled tree the reference run used (`calibrations="data/mtCDNApri_MD.trees"`,
    02_McmcDate/01_McmcDate/analysis.conf) with prepare.load_calibrations_from_tree: U(100,.025) root, B(12,16,.025,.025),
    B(6,8,.025,.025).  This checkout keeps only 10 trees of the tree list (too few for a non-singular 11x11 covariance), so
    the mean vector comes from those trees (after the len/6 burn-in) and the precision is synthetic."""
    base = "bench/comparison_with_mcmctree/02_McmcDate/01_McmcDate/data/"
    rooted = tree.parse_newick(read(base + "pb_rooted_mitCDNApri.tree"))
    og = min((prepare._leaf_set(c) for c in rooted["children"]), key=lambda x: sorted(x))
    lines = [ln for ln in read(base + "unr_lg_g5_ncat1.treelist").splitlines() if ln.strip()]
    lines = lines[len(lines) // 6:]
    trees = [tree.flatten_preorder(prepare.outgroup(og, tree.parse_newick(ln))) for ln in lines]
    parent, c0, c1, names, _ = trees[0]
    rows = np.array([prepare._branches_row(p, ln) for p, _, _, _, ln in trees])
    mu = rows.mean(axis=0)
    rng = np.random.default_rng(7)
    prec, logdet = synth.synthetic_precision(mu, rng, band=4)
    cal = prepare.load_calibrations_from_tree(read(base + "mtCDNApri_MD.trees"), parent, names)
    assert cal["node"].tolist() == [0, 1, 3] and cal["names"] == ["human-gibbon", "human-sumatran", "human-bonobo"]
    md = model.ModelDesc(parent=parent, mean=mu, precision=prec, logdet_sigma=logdet, ht=prepare.mean_root_height(cal),
                         cal_node=cal["node"], cal_lo=cal["lo"], cal_lo_p=cal["lo_p"], cal_hi=cal["hi"], cal_hi_p=cal["hi_p"])
    assert md.ht == 50.0
    mean_len = np.array([ln for _, _, _, _, ln in trees]).mean(axis=0)
    init = prepare.initial_state(parent, mean_len)
    make("mtcdnapri-7-leaves", md, {"names": names}, init, seed=3)


def abi_case():
    """tests/golden/abi_case_12_leaves.txt: the 12-leaf data set as plain numbers for the C program tests/c_abi/abi_check.c
    (model arrays, 8 states incl. two edge states, the oracle's outputs / gradient / status for the log-normal clock)"""
    z = np.load(os.path.join(HERE, "12-leaves-variable-rate.npz"))
    clock = 1
    pick = [0, 1, 2, 3, 4, 5, 25, 27]                      # valid states + a non-positive branch + a zero rate
    X, out, grad, st = z["states"][pick], z[f"out_{clock}"][pick], z[f"grad_{clock}"][pick], z[f"status_{clock}"][pick]
    fmt = lambda a: " ".join("inf" if v == np.inf else "-inf" if v == -np.inf else "nan" if v != v else repr(float(v)) for v in np.ravel(a))
    ints = lambda a: " ".join(str(int(v)) for v in np.ravel(a))
    N = len(z["parent"])
    with open(os.path.join(HERE, "abi_case_12_leaves.txt"), "w") as f:
        f.write(f"{N} {clock} {int(z['likelihood'])} {len(z['cal_node'])} {len(z['con_young'])} {len(pick)}\n")
        for a, kind in ((z["parent"], ints), (z["mean"], fmt), (z["precision"], fmt), ([float(z["logdet_sigma"]), float(z["ht"])], fmt),
                        (z["cal_node"], ints), (z["cal_lo"], fmt), (z["cal_lo_p"], fmt), (z["cal_hi"], fmt), (z["cal_hi_p"], fmt),
                        (z["con_young"], ints), (z["con_old"], ints), (z["con_p"], fmt), (X, fmt), (out[:, :7], fmt),
                        (np.where(np.isfinite(grad), grad, 0.0), fmt), (st, ints)):
            f.write(kind(a) + "\n")
    print("abi_case_12_leaves.txt", X.shape, st.tolist())


def more_datasets():
    """the reference's three remaining tests/ directories: a calibration pinned to a 1e-6-wide interval with probability mass 1e-6
    (very steep soft bounds), the 10-leaf autocorrelated-rate simulation (ages in the hundreds, ht = 1000), and the 25-leaf
    empirical set with calibrations and constraints (3000 trees in the list)"""
    dataset("06-leaves-pinned-node", "06-leaves-pinned-node", seed=4)
    dataset("10-leaves-autocorrelated-rate", "10-leaves-autocorrelated-rate", seed=5)
    dataset("25-leaves-bastien", "25-leaves-bastien", con=True, seed=6, trees="alignment.fasta.trees.only")


if __name__ == "__main__":
    if sys.argv[1:] == ["more"]:           # only the fixtures added later; the earlier files stay byte-identical
        more_datasets()
        sys.exit(0)
    dataset("06-leaves-constant-rate", "06-leaves-constant-rate", seed=0)
    dataset("12-leaves-variable-rate", "12-leaves-variable-rate", con=True, seed=1)
    dataset("24-leaves-braces", "24-leaves-braces", con=True, br=True, seed=2)
    mtcdnapri()
    abi_case()
    more_datasets()
