#!/usr/bin/env python
"""Throughput of device-resident Metropolis-Hastings moves on the benchmark model (run on a GPU box).
usage: mh_bench.py [n_leaves=1000] [B=8192] [steps=200]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_date_b200 import binding, mh_cycle, synth  # noqa: E402
from mcmc_date_b200.binding import (MH_SCALE_BRANCH, MH_SCALE_NORM_TREE_CONTRA_M, MH_SCALE_SUBTREE, MH_SLIDE_BRACE_CONTRA,  # noqa: E402
                                    MH_SLIDE_NODE, MH_SLIDE_NODE_CONTRA)


def main():
    n_leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    md, h = synth.synthetic_model(n_leaves, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, B)
    kinds = ((MH_SLIDE_NODE, "slide node", 0.002), (MH_SLIDE_NODE_CONTRA, "slide node contrarily", 0.002),
             (MH_SCALE_BRANCH, "scale branch", 100.0), (MH_SLIDE_BRACE_CONTRA, "slide braced nodes contrarily", 0.0002),
             (MH_SCALE_SUBTREE, "scale sub tree (random node: full evaluation)", 0.002),
             (MH_SCALE_NORM_TREE_CONTRA_M, "scale rate mean and tree contrarily (full evaluation)", 3000.0))
    for inc in (True, False):
        ev = binding.Evaluator(md, max_batch=B)
        ev.mh_set_incremental(inc)
        ev.chains_set(X)
        print(f"--- incremental evaluation {'on' if ev.mh_incremental_active() else 'off'}")
        for kind, name, par in kinds:
            ev.mh_cycle([(kind, -1, par, 1.0, 0, 3)], 1, seed=3, iteration0=0)     # warm-up
            ev.synchronize()
            t0 = time.perf_counter()
            acc, inv, _ = ev.mh_cycle([(kind, -1, par, 1.0, 0, steps)], 1, seed=3, iteration0=10)
            dt = time.perf_counter() - t0
            print(f"{name}: {steps} steps x {B} chains in {dt * 1e3:.1f} ms = {steps * B / dt / 1e6:.2f} M proposals/s "
                  f"({dt * 1e6 / steps:.1f} us per step); acceptance {acc[0] / (steps * B):.2f}, invalid {inv[0] / (steps * B):.4f}")
        # one sweep of the reference's whole cycle (app/Definitions.hs:256-279)
        props = mh_cycle.reference_cycle(md)
        nsteps = sum(p[5] for p in props)
        ev.synchronize()
        t0 = time.perf_counter()
        acc, inv, _ = ev.mh_cycle(props, 1, seed=5, iteration0=100000)
        dt = time.perf_counter() - t0
        print(f"one iteration of the reference cycle: {len(props)} proposals, {nsteps} steps x {B} chains in {dt:.2f} s = "
              f"{nsteps * B / dt / 1e6:.2f} M proposals/s ({dt * 1e6 / nsteps:.1f} us per step); acceptance "
              f"{acc.sum() / (nsteps * B):.2f}")
        Xd, out, st = ev.chains_get()
        o2, s2 = ev.eval(Xd[:256])
        err = np.abs(out[:256, :7] - o2[:, :7]) / np.maximum(1.0, np.abs(o2[:, :7]))
        print(f"finite posteriors: {np.isfinite(out[:, 6]).mean():.3f}; drift of the resident values vs a fresh evaluation: {err.max():.2e}")
        ev.close()


if __name__ == "__main__":
    main()
