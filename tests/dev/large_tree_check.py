"""Maximum sizes: a 4000-leaf tree (N = 7999 nodes, K = 7997, 512 MB precision matrix) through every large-tree kernel,
checked against the CPU oracle on a few chains; MH steps (incremental and full) on the same model.  Run by hand on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcmc_date_b200 import binding, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
import mh_ref as R  # noqa: E402

n_leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
t0 = time.time()
md, h = synth.synthetic_model(n_leaves, seed=4000, n_cal=16, n_con=8, n_brace=4)
B = 300
X = synth.synthetic_states(md, h, B)
print(f"model built in {time.time() - t0:.1f} s: N = {md.n_nodes}, K = {md.dim}")
ev = binding.Evaluator(md)
orc = O.Oracle(md)
out, grad, st = ev.eval_grad(X)
oo, og, ost = orc.eval_grad(X[:6], nthreads=6)
rel = np.abs(out[:6, :7] - oo) / np.maximum(1.0, np.abs(oo))
gsc = np.maximum(1.0, np.abs(og).max(axis=1, keepdims=True))
print(f"value relerr {rel.max():.2e}, gradient relerr {(np.abs(grad[:6] - og) / gsc).max():.2e}, status equal {np.array_equal(st[:6], ost)}")
assert rel.max() < 1e-10 and (np.abs(grad[:6] - og) / gsc).max() < 1e-10
o2, s2 = ev.eval(X)
assert (np.abs(o2[:, :7] - out[:, :7]) / np.maximum(1.0, np.abs(out[:, :7]))).max() < 1e-10
ev.chains_set(X)
print("incremental:", ev.mh_incremental_active())
k = 0
for kind, par in ((R.SLIDE_NODE, 0.001), (R.SCALE_BRANCH, 100.0), (R.SLIDE_NODE_CONTRA, 0.001), (R.SLIDE_BRACE_CONTRA, 0.0001),
                  (R.SCALE_SUBTREE, 0.0005), (R.SCALE_NORM_TREE_CONTRA_M, 5000.0)):
    acc, inv, k = ev.mh_cycle([(kind, -1, par, 1.0, 0, 5)], 1, seed=1, iteration0=k)
    print(kind, "accepted", int(acc[0]), "of", 5 * B, "invalid", int(inv[0]))
Xd, od, sd = ev.chains_get()
o3, s3 = ev.eval(Xd)
err = (np.abs(od[:, :7] - o3[:, :7]) / np.maximum(1.0, np.abs(o3[:, :7]))).max()
print(f"resident values vs fresh evaluation after the moves: {err:.2e}; statuses equal {np.array_equal(sd, s3)}")
assert err < 1e-10
print("ok")
