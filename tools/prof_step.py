#!/usr/bin/env python
"""Smallest program that runs the benchmark step (mcd_eval_grad_device: residual_split_kernel, gemm_i8_ozaki_kernel,
posterior_kernel; 1000 leaves) a few times -- the command ncu captures are taken of.  usage: prof_step.py [chains=8192] [steps=4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from mcmc_date_b200 import binding, model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
md, X = bench.build_workload(B)
ev = binding.Evaluator(md, device=0, max_batch=B)
dev = torch.device("cuda", 0)
d_states = torch.from_numpy(X).to(dev)
d_out = torch.empty((B, model.OUT_COLS), dtype=torch.float64, device=dev)
d_grad = torch.empty((B, md.state_len), dtype=torch.float64, device=dev)
d_status = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(steps):
    ev.eval_grad_device(B, d_states.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_status.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(d_out[:, 6].sum().item()))
ev.close()
