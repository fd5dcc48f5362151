"""Quick GPU-vs-oracle parity sweep (development helper; the real tests live in tests/)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_date_b200 import binding, model, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def relerr(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def check(n_leaves, B, clock, lik=model.LIK_FULL, n_cal=3, n_con=2, n_brace=1, dual=False):
    md, h = synth.synthetic_model(n_leaves, seed=1234 + n_leaves, clock_model=clock, n_cal=n_cal, n_con=n_con,
                                  n_brace=n_brace, likelihood=lik)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    assert np.array_equal(ev.branch_index(), orc.branch_index())
    assert np.array_equal(ev.mask(), orc.mask)
    t0 = time.time()
    out, grad, st = ev.eval_grad(X)
    t1 = time.time()
    oo, og, ost = orc.eval_grad(X, nthreads=8)
    t2 = time.time()
    ev_val = relerr(out[:, :7], oo).max(axis=0)
    gscale = np.maximum(1.0, np.abs(og).max(axis=1, keepdims=True))
    gerr = (np.abs(grad - og) / gscale).max()
    msg = f"n={n_leaves} B={B} clock={clock} lik={lik}: val relerr {ev_val.max():.2e} grad err {gerr:.2e} status eq {np.array_equal(st, ost)} gpu {t1 - t0:.3f}s cpu {t2 - t1:.3f}s"
    if dual:
        gd = orc.grad_dual(X[0])
        msg += f" | dual-vs-gpu {np.abs(gd - grad[0]).max() / max(1, np.abs(gd).max()):.2e} dual-vs-port {np.abs(gd - og[0]).max() / max(1, np.abs(gd).max()):.2e}"
    out2, st2 = ev.eval(X)
    msg += f" | eval-vs-evalgrad {np.abs(out2[:, :7] - out[:, :7]).max():.1e}"
    print(msg, flush=True)
    ev.close()


if __name__ == "__main__":
    for clock in range(4):
        check(12, 64, clock, dual=True)
    check(6, 8, 1, n_cal=1, n_con=0, n_brace=0, dual=True)
    check(24, 1024, 1, dual=True)
    check(12, 64, 1, lik=model.LIK_UNIVARIATE, dual=True)
    check(12, 64, 1, lik=model.LIK_NONE, dual=True)
    check(60, 300, 2, n_cal=5, n_con=4, n_brace=2)
    check(1000, 300, 1, n_cal=16, n_con=8, n_brace=4)
    check(1000, 130, 3, n_cal=16, n_con=8, n_brace=4)
