#!/usr/bin/env python
"""Accuracy of the contraction pipes on the library's own models (run on a GPU box):
prints, per mode, the largest relative error of ln likelihood / ln posterior and of the gradient
(relative to the chain's largest gradient component) against the CPU oracle, and against an mpmath
evaluation of the quadratic form for a few chains.   usage: contraction_accuracy.py [n_leaves=1000] [B=256]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcmc_date_b200 import binding, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    n_leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    cond = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    md, h = synth.synthetic_model(n_leaves, seed=synth.BASE_SEED + 4, n_cal=16 if n_leaves > 100 else 3, n_con=2, n_brace=1)
    if cond > 0:   # make the precision matrix ill-conditioned: P' = D P D with a wide spread of row scales
        rng = np.random.default_rng(7)
        d = np.exp(rng.uniform(-0.5 * np.log(cond), 0.5 * np.log(cond), md.dim) * 0.5)
        md.precision = (md.precision * d[:, None]) * d[None, :]
        md.precision = 0.5 * (md.precision + md.precision.T)
        md.logdet_sigma = float(md.logdet_sigma - 2 * np.log(d).sum())
    X = synth.synthetic_states(md, h, B)
    orc = O.Oracle(md)
    oo, og, ost = orc.eval_grad(X, nthreads=16)
    # exact ln likelihood for a few chains: long-double quadratic form
    P = md.precision.astype(np.longdouble)
    mu = md.mean.astype(np.longdouble)
    ev = binding.Evaluator(md)
    ok = ost == 0
    print(f"n_leaves {n_leaves}  K {md.dim}  B {B}  finite chains {ok.sum()}  spread {cond:g}")
    gsc = np.maximum(1.0, np.abs(og).max(axis=1, keepdims=True))
    res = {}
    for mode in ("dmma", "i8s7", "i8s6"):
        ev.set_contraction(mode)
        out, grad, st = ev.eval_grad(X)
        e_lik = np.abs(out[ok, 4] - oo[ok, 4]) / np.maximum(1.0, np.abs(oo[ok, 4]))
        e_post = np.abs(out[ok, 6] - oo[ok, 6]) / np.maximum(1.0, np.abs(oo[ok, 6]))
        e_g = (np.abs(grad - og) / gsc)[ok]
        res[mode] = out
        print(f"  {mode:5s}: lnL rel err max {e_lik.max():.2e}  ln post {e_post.max():.2e}  gradient max {e_g.max():.2e} "
              f"rms {np.sqrt((e_g ** 2).mean()):.2e}")
    print(f"  |lnL| ~ {np.abs(oo[ok, 4]).mean():.3e}")
    ev.close()


if __name__ == "__main__":
    main()
