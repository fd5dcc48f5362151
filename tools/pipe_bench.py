#!/usr/bin/env python
"""A/B timing of the device-path step (mcd_eval_grad_device, 1000 leaves) under the experiment switches of the chunk pipeline:
    python tools/pipe_bench.py            -> runs every (library build, MCD_PIPE) combination in a subprocess
    python tools/pipe_bench.py one        -> one measurement in this process (env: MCD_LIB_PATH, MCD_PIPE, PB_CHAINS)
Prints ms per step (CUDA events, 30 steps after 5 warm-up steps) and a checksum of the outputs (must agree across variants)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch
    import bench
    from mcmc_date_b200 import binding, model
    B = int(os.environ.get("PB_CHAINS", "8192"))
    md, X = bench.build_workload(B)
    ev = binding.Evaluator(md, device=0, max_batch=B)
    dev = torch.device("cuda", 0)
    d_states = torch.from_numpy(X).to(dev)
    d_out = torch.empty((B, model.OUT_COLS), dtype=torch.float64, device=dev)
    d_grad = torch.empty((B, md.state_len), dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def step():
        ev.eval_grad_device(B, d_states.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_status.data_ptr(), st.cuda_stream)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    best = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / 30)
    chk = float(d_out[:, 6].sum().item()), float(d_grad.abs().sum().item())
    print(json.dumps({"lib": os.path.basename(os.environ.get("MCD_LIB_PATH", "default")), "pipe": os.environ.get("MCD_PIPE", "0"),
                      "chains": B, "ms_per_step": min(best), "all": best, "evals_per_s": B / min(best) * 1e3, "checksum": chk}))
    ev.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        libs = [None] + [os.path.join(ROOT, "mcmc-date_b200", f) for f in sorted(os.listdir(os.path.join(ROOT, "mcmc-date_b200")))
                         if f.startswith("libmcd_") and f.endswith(".so")]
        for chains in os.environ.get("PB_CHAINS_LIST", "8192").split(","):
            for lib in libs:
                for pipe in os.environ.get("PB_PIPES", "0,2,3,4").split(","):
                    env = dict(os.environ, MCD_PIPE=pipe, PB_CHAINS=chains)
                    if lib:
                        env["MCD_LIB_PATH"] = lib
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
                    print(r.stdout.strip() or ("FAILED " + r.stderr[-400:]), flush=True)
