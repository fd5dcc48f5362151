"""Golden statistics of the reference's OWN McmcDate output for the 7-taxon primate set (run HERE, where /root/reference
exists; the small .npz travels to the GPU box, the reference does not).

    python tests/golden/make_reference_samples.py

bench/comparison_with_mcmctree/03_compare_estimates/prior_samples_run{1..6}.tsv hold the node ages (time height x relative
node height) of McmcDate nodes 0, 1, 2, 3, 5, 9 sampled by six real `./run -c ul n r` runs (README.md:617-622: calibrations
from data/mtCDNApri_MD.trees, uncorrelated log-normal clock, NoLikelihood, i.e. the prior plus the proposal cycle): 12 930
iterations (burn-in 4 930 + 8 000, app/Definitions.hs:417-441), the time-tree monitor every 2nd iteration
(app/Definitions.hs:376), the first 25 % discarded by scripts-analyze-stack/analyze -> 4 850 samples per run.

Written: per run and pooled, for each node: mean, standard deviation and a 201-point quantile grid (0, 0.5 %, ..., 100 %);
and the figures the reference's own summary table reports (03_compare_estimates/out/compare_divtimes.tsv:2-4, MD_CLK columns).
The same statistics restricted to the samples whose root age is below 28 (`below_*`): the committed samples stop at a root
age of about 30-31.5 in all six runs although the committed calibration file says U(100) -- see
tests/test_reference_samples.py -- so the part of the distribution that does not depend on the root bound is kept as well.
"""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BELOW = 28.0
BASE = "/root/reference/bench/comparison_with_mcmctree/03_compare_estimates/"


def main():
    runs = []
    for i in range(1, 7):
        with open(BASE + f"prior_samples_run{i}.tsv") as f:
            header = f.readline().split()
        assert header == ["Iter", "0", "1", "2", "3", "5", "9"], header
        a = np.loadtxt(BASE + f"prior_samples_run{i}.tsv", skiprows=1)[:, 1:]
        assert a.shape == (4850, 6)
        runs.append(a)
    q = np.linspace(0.0, 1.0, 201)
    pooled = np.vstack(runs)
    table = {}
    with open(BASE + "out/compare_divtimes.tsv") as f:
        cols = f.readline().rstrip("\n").split("\t")
        for ln in f:
            v = ln.rstrip("\n").split("\t")
            table[int(v[cols.index("McmcDate")])] = [float(v[cols.index(c)]) for c in ("MD_CLK-mean_t", "MD_CLK-q2.5%", "MD_CLK-q97.5%")]
    np.savez_compressed(
        os.path.join(HERE, "mtcdnapri-prior-samples.npz"),
        nodes=np.array([0, 1, 2, 3, 5, 9], np.int32), quantile_grid=q,
        run_mean=np.array([r.mean(0) for r in runs]), run_sd=np.array([r.std(0, ddof=1) for r in runs]),
        run_quantiles=np.array([np.quantile(r, q, axis=0) for r in runs]),          # [6 runs][201][6 nodes]
        pooled_mean=pooled.mean(0), pooled_sd=pooled.std(0, ddof=1), pooled_quantiles=np.quantile(pooled, q, axis=0),
        table_nodes=np.array(sorted(table), np.int32), table_mean_q025_q975=np.array([table[k] for k in sorted(table)]),
        samples_per_run=4850, below=BELOW,
        below_fraction=np.array([(r[:, 0] < BELOW).mean() for r in runs]),
        below_run_mean=np.array([r[r[:, 0] < BELOW].mean(0) for r in runs]),
        below_run_sd=np.array([r[r[:, 0] < BELOW].std(0, ddof=1) for r in runs]),
        below_run_quantiles=np.array([np.quantile(r[r[:, 0] < BELOW], q, axis=0) for r in runs]),
        below_pooled_mean=pooled[pooled[:, 0] < BELOW].mean(0), below_pooled_sd=pooled[pooled[:, 0] < BELOW].std(0, ddof=1),
        below_pooled_quantiles=np.quantile(pooled[pooled[:, 0] < BELOW], q, axis=0),
        root_age_max_per_run=np.array([r[:, 0].max() for r in runs]))
    print("pooled mean", pooled.mean(0).round(3), "table", table)


if __name__ == "__main__":
    main()
