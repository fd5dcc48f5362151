#!/usr/bin/env python
"""Throughput of device-resident Metropolis-Hastings moves on the benchmark model (run on a GPU box).
usage: mh_bench.py [n_leaves=1000] [B=8192] [steps=200]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_date_b200 import binding, synth  # noqa: E402


def main():
    n_leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    md, h = synth.synthetic_model(n_leaves, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md, max_batch=B)
    ev.chains_set(X)
    for kind, name, sd in ((0, "slide node", 0.002), (1, "scale sub tree", 0.002)):
        ev.mh_step(kind, -1, sd, seed=3, iteration=0)     # warm-up
        acc = 0
        t0 = time.perf_counter()
        for it in range(steps):
            a = ev.mh_step(kind, -1, sd, seed=3, iteration=1 + it, want_accepted=(it == steps - 1))
        ev.synchronize()
        dt = time.perf_counter() - t0
        print(f"{name}: {steps} steps x {B} chains in {dt * 1e3:.1f} ms = {steps * B / dt / 1e6:.2f} M proposals/s "
              f"({dt * 1e3 / steps:.3f} ms per step); acceptance of the last step {np.mean(a == 1):.2f}, invalid {np.mean(a < 0):.3f}")
    Xd, out, st = ev.chains_get()
    print("finite posteriors:", np.isfinite(out[:, 6]).mean())
    ev.close()


if __name__ == "__main__":
    main()
