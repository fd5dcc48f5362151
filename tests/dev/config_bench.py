"""Throughput of every BASELINE.json configuration (device-resident and host-buffer paths), with the
CPU port beside it.  Development / reporting helper: bench.py remains the contract."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from mcmc_date_b200 import binding, model, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import load_fixture  # noqa: E402


def bench_model(name, md, X, iters=50, cpu=True):
    B, S = X.shape
    ev = binding.Evaluator(md, max_batch=B)
    dev = torch.device("cuda", 0)
    d_x = torch.from_numpy(X).to(dev)
    d_out = torch.empty((B, model.OUT_COLS), dtype=torch.float64, device=dev)
    d_grad = torch.empty((B, S), dtype=torch.float64, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        ev.eval_grad_device(B, d_x.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_st.data_ptr(), st)

    def step_v():
        ev.eval_device(B, d_x.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), st)

    res = {"config": name, "n_leaves": md.n_leaves, "K": md.dim, "chains": B}
    for key, fn in (("grad", step), ("value", step_v)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[f"device_{key}_us_per_batch"] = 1e3 * ms
        res[f"device_{key}_evals_per_s"] = B / (ms * 1e-3)
    hx = torch.from_numpy(X).pin_memory()
    ho = torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory()
    hg = torch.empty((B, S), dtype=torch.float64).pin_memory()
    hs = torch.empty(B, dtype=torch.int32).pin_memory()
    for _ in range(3):
        ev.eval_grad_ptr(B, hx.data_ptr(), ho.data_ptr(), hg.data_ptr(), hs.data_ptr())
    t0 = time.perf_counter()
    for _ in range(iters):
        ev.eval_grad_ptr(B, hx.data_ptr(), ho.data_ptr(), hg.data_ptr(), hs.data_ptr())
    dt = (time.perf_counter() - t0) / iters
    res["host_grad_us_per_batch"] = 1e6 * dt
    res["host_grad_evals_per_s"] = B / dt
    if cpu:
        orc = O.Oracle(md)
        n = min(B, 2048)
        thr = O.max_threads()
        orc.eval_grad(X[:n], nthreads=thr)
        t0 = time.perf_counter()
        orc.eval_grad(X[:n], nthreads=thr)
        res["cpu_port_evals_per_s"] = n / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        orc.eval_grad(X[: min(n, 256)], nthreads=1)
        res["cpu_port_1thread_evals_per_s"] = min(n, 256) / (time.perf_counter() - t0)
        res["cpu_threads"] = thr
    ev.close()
    return res


def main():
    out = []
    rng = np.random.default_rng(0)
    for name, B in (("06-leaves-constant-rate", 1), ("06-leaves-constant-rate", 1024), ("12-leaves-variable-rate", 1024),
                    ("24-leaves-braces", 1024), ("mtcdnapri-7-leaves", 64), ("mtcdnapri-7-leaves", 8)):
        md, z = load_fixture(name, 3 if name.startswith("mtcdna") else 1)
        nv = int(z["n_valid"])
        X = z["states"][rng.integers(1, nv, size=B)]
        out.append(bench_model(f"{name} x{B}", md, np.ascontiguousarray(X)))
        print(json.dumps(out[-1]), flush=True)
    md, h = synth.synthetic_model(1000, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    for B in (1024, 8192):
        X = synth.synthetic_states(md, h, B, seed=synth.BASE_SEED + 5)
        out.append(bench_model(f"synthetic-1000-leaves x{B}", md, X, iters=10))
        print(json.dumps(out[-1]), flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
