"""The reference's prior-only run of the 7-taxon primate set (bench/comparison_with_mcmctree, README.md:617-622:
`./run -c ul n r`) on device-resident chains, compared with the statistics of the reference's own samples
(tests/golden/mtcdnapri-prior-samples.npz).  The `-m gpu` test tests/test_reference_samples.py runs the same function.

    python tools/prior_samples.py [n_chains] [burn-in scale] [sampling iterations]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mcmc_date_b200 import binding, mh_cycle, model  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_model(clock=model.UNCORRELATED_LOGNORMAL, root_upper=None):
    """the model of the reference's `./run -c ul n r`: calibrations of data/mtCDNApri_MD.trees (parsed into the fixture by
    prepare.load_calibrations_from_tree), NoLikelihood; root_upper replaces the root's upper bound (U(100,.025) in the file)"""
    z = np.load(os.path.join(GOLDEN, "mtcdnapri-7-leaves.npz"))
    K = len(z["mean"])
    cal_hi = z["cal_hi"].copy()
    if root_upper is not None:
        assert z["cal_node"][0] == 0
        cal_hi[0] = root_upper
    md = model.ModelDesc(parent=z["parent"], mean=np.zeros(K), precision=np.zeros(0), logdet_sigma=0.0, clock_model=clock,
                         likelihood=model.LIK_NONE, ht=float(z["ht"]), cal_node=z["cal_node"], cal_lo=z["cal_lo"],
                         cal_lo_p=z["cal_lo_p"], cal_hi=cal_hi, cal_hi_p=z["cal_hi_p"])
    x0 = z["states"][0].copy()     # initWith (app/Definitions.hs:96-123) but H = ht instead of 1: a shorter climb for the burn-in
    return md, x0


def exact_cycle(md):
    """proposals whose stated Jacobian is the determinant of the move, none lifted with the root-branch Jacobian: the chain
    then targets the prior itself (used to tell the prior's own marginals from the stationary distribution of the
    reference's cycle, which mixes kernels with and without the root-branch Jacobian)"""
    B = binding
    cyc = [(B.MH_SCALE_SCALAR, s, 10.0, 1.0, 0, 3) for s in (0, 1, 2, 3, 4)]
    inner = [i for i in range(1, md.n_nodes) if md.child0[i] >= 0]
    cyc += [(B.MH_SLIDE_NODE, i, 0.01, 1.0, 0, 5) for i in inner]
    cyc += [(B.MH_SCALE_BRANCH, i, 100.0, 1.0, 0, 1) for i in range(1, md.n_nodes)]
    return cyc


def sample(n_chains=4096, periods=None, n_sampling=400, thin=50, seed=11, cycle=None, log=None, root_upper=None):
    """burn-in with auto tuning, then `n_sampling` iterations with the node ages H h_i of all chains recorded every `thin`-th
    iteration -> (ages [n_records * n_chains][N], tuned cycle, acceptance rates per cycle entry)"""
    md, x0 = load_model(root_upper=root_upper)
    N = md.n_nodes
    ev = binding.Evaluator(md)
    ev.chains_set(np.tile(x0, (n_chains, 1)))
    rng = np.random.default_rng(seed)
    cycle = mh_cycle.reference_cycle(md) if cycle is None else cycle(md)
    cycle, k = mh_cycle.burn_in(ev, md, cycle, rng, seed=seed, k0=0, periods=periods)
    ages, acc, prop = [], np.zeros(len(cycle)), np.zeros(len(cycle))
    for _ in range(n_sampling // thin):
        a, p, k = mh_cycle.run_iterations(ev, md, cycle, thin, rng, seed, k)
        acc += a
        prop += p
        X, out, st = ev.chains_get()
        assert (st == 0).all() and np.isfinite(out[:, 6]).all()
        ages.append(X[:, 2:3] * X[:, 3:3 + N])
        if log:
            log(f"  record {len(ages)}: mean root age {ages[-1][:, 0].mean():.3f}")
    ev.close()
    return np.concatenate(ages), cycle, acc / np.maximum(prop, 1)


def compare(ages, g, prefix=""):
    """our statistics beside the reference's: per node (mean, sd, 2.5 %, 50 %, 97.5 %) of ours, of the pooled reference samples
    and the range over the six reference runs; prefix "below_" = the statistics conditional on a root age below g["below"]"""
    if prefix:
        ages = ages[ages[:, 0] < float(g["below"])]
        g = {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)} | {"nodes": g["nodes"], "quantile_grid": g["quantile_grid"]}
    nodes = g["nodes"]
    q = g["quantile_grid"]
    iq = [int(np.argmin(np.abs(q - x))) for x in (0.025, 0.5, 0.975)]
    rows = []
    for j, nd in enumerate(nodes):
        a = ages[:, nd]
        ours = np.array([a.mean(), a.std(ddof=1)] + list(np.quantile(a, q[iq])))
        runs = np.column_stack([g["run_mean"][:, j], g["run_sd"][:, j]] + [g["run_quantiles"][:, i, j] for i in iq])
        pooled = np.array([g["pooled_mean"][j], g["pooled_sd"][j]] + [g["pooled_quantiles"][i, j] for i in iq])
        rows.append(dict(node=int(nd), ours=ours, pooled=pooled, spread=runs.max(0) - runs.min(0), run_min=runs.min(0), run_max=runs.max(0)))
    return rows


if __name__ == "__main__":
    nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    periods = [max(5, int(round(n * scale))) for n in mh_cycle.BURN_IN_FAST + mh_cycle.BURN_IN_SLOW]
    t0 = time.time()
    mode = sys.argv[4] if len(sys.argv) > 4 else "reference"
    ru = float(sys.argv[6]) if len(sys.argv) > 6 else None
    ages, cycle, rates = sample(nch, periods, ns, cycle=exact_cycle if mode == "exact" else None, log=print, root_upper=ru)
    print(f"{time.time() - t0:.1f} s, {len(ages)} samples")
    if len(sys.argv) > 5 and sys.argv[5] != "-":
        np.save(sys.argv[5], ages.astype(np.float32))
    g = np.load(os.path.join(GOLDEN, "mtcdnapri-prior-samples.npz"))
    print("node  stat: ours  pooled-reference  [run min, run max]")
    for r in compare(ages, g):
        for i, nm in enumerate(("mean", "sd", "q2.5", "q50", "q97.5")):
            flag = "" if r["run_min"][i] - r["spread"][i] <= r["ours"][i] <= r["run_max"][i] + r["spread"][i] else "  <-- outside"
            print(f"{r['node']:3d} {nm:6s} {r['ours'][i]:8.3f} {r['pooled'][i]:8.3f}  [{r['run_min'][i]:.3f}, {r['run_max'][i]:.3f}]{flag}")
    md, _ = load_model()
    for e, rt in zip(cycle, rates):
        print(f"kind {e[0]:2d} node {e[1]:2d} param {e[2]:7.2f} tune {e[3]:9.4f} jac {e[4]} w {e[5]} dim {mh_cycle.proposal_dimension(md, e)} rate {rt:.3f}")
