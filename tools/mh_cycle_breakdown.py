import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_date_b200 import binding, synth, mh_cycle
md, h = synth.synthetic_model(1000, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
X = synth.synthetic_states(md, h, 8192)
ev = binding.Evaluator(md, max_batch=8192)
ev.chains_set(X)
props = mh_cycle.reference_cycle(md)
import collections
def run(sel, name):
    ps = [p for p in props if sel(p)]
    n = sum(p[5] for p in ps)
    ev.mh_cycle(ps[:50], 1, seed=1, iteration0=0); ev.synchronize()
    t0 = time.perf_counter(); ev.mh_cycle(ps, 1, seed=2, iteration0=1000); dt = time.perf_counter() - t0
    print(f"{name}: {len(ps)} proposals, {n} steps, {dt*1e3:.0f} ms, {dt*1e6/n:.1f} us/step")
B = binding
parent = np.asarray(md.parent); N = len(parent)
size = np.ones(N, int)
for i in range(N - 1, 0, -1): size[parent[i]] += size[i]
run(lambda p: p[0] == B.MH_SLIDE_NODE, "slide node (fixed nodes)")
run(lambda p: p[0] == B.MH_SCALE_BRANCH, "scale branch (fixed)")
run(lambda p: p[0] == B.MH_SLIDE_NODE_CONTRA, "slide node contra (fixed)")
run(lambda p: p[0] in (B.MH_SCALE_SUBTREE, B.MH_SCALE_SUBTREE_CONTRA) and size[p[1]] <= 32, "scale sub tree <= 32 (incremental)")
run(lambda p: p[0] == B.MH_SCALE_RATE_SUBTREE and size[p[1]] <= 64, "scale rate sub tree <= 64 (incremental)")
run(lambda p: p[0] in (B.MH_SCALE_SUBTREE, B.MH_SCALE_SUBTREE_CONTRA) and size[p[1]] > 32, "scale sub tree > 32 (full)")
run(lambda p: p[0] == B.MH_SCALE_RATE_SUBTREE and size[p[1]] > 64, "scale rate sub tree > 64 (full)")
run(lambda p: p[0] in (B.MH_SCALE_SCALAR, B.MH_SCALE_NORM_TREE_CONTRA_M, B.MH_SCALE_NORM_TREE_CONTRA_H, B.MH_SCALE_VAR_TREE, B.MH_SCALE_VAR_TREE_AUTO, B.MH_SLIDE_ROOT_CONTRA, B.MH_SCALE_RATES_TREE_CONTRA, B.MH_SCALE_H_M_CONTRA, B.MH_PULLEY), "global (full)")
run(lambda p: p[0] in (B.MH_SLIDE_BRACE, B.MH_SLIDE_BRACE_CONTRA), "braces")
