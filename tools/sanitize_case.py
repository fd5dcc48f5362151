"""Small driver for compute-sanitizer runs: touches every kernel once on small inputs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcmc_date_b200 import binding, model, synth  # noqa: E402


def run(n_leaves, B, clock, lik):
    md, h = synth.synthetic_model(n_leaves, seed=9 + n_leaves, clock_model=clock, n_cal=3, n_con=2, n_brace=1, likelihood=lik)
    X = synth.synthetic_states(md, h, B)
    X[1, 1] = X[1, 0] + 2e-7          # one near-critical chain
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    o2, s2 = ev.eval(X)
    mask = ev.mask().astype(bool)
    theta = np.ascontiguousarray(X[:, mask][:, ::-1])
    o3, g3, s3 = ev.eval_grad_theta(theta, X[0])
    assert np.isfinite(out[:, 6]).all() and np.isfinite(grad).all()
    ev.close()
    print("ok", n_leaves, B, clock, lik, flush=True)


if __name__ == "__main__":
    run(12, 20, 1, model.LIK_FULL)        # fused small-tree kernel
    run(24, 9, 2, model.LIK_FULL)
    run(150, 130, 1, model.LIK_FULL)      # residual + DMMA contraction + posterior kernels (+ Cholesky path)
    run(150, 40, 3, model.LIK_SPARSE)     # CSR contraction
    run(150, 40, 0, model.LIK_UNIVARIATE)
