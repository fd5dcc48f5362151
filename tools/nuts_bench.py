#!/usr/bin/env python
"""Throughput of the device-resident NUTS transition on the benchmark model (run on a GPU box).
usage: nuts_bench.py [n_leaves=1000] [B=8192] [max_depth=6] [eps=2e-4]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_date_b200 import binding, synth  # noqa: E402


def main():
    n_leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    max_depth = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    eps = float(sys.argv[4]) if len(sys.argv) > 4 else 2e-4
    md, h = synth.synthetic_model(n_leaves, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md, max_batch=B)
    mask = ev.mask().astype(bool)
    theta = np.ascontiguousarray(X[:, mask][:, ::-1])
    out, grad, st = ev.eval_grad(X[:256])
    g = np.abs(grad[:, mask][:, ::-1]).max(axis=0)
    inv_mass = 1.0 / np.maximum(1.0, g) ** 2
    D = ev.D
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_theta, h_base, h_im, h_eps = pin(theta), pin(X[0]), pin(inv_mass), pin(np.full(B, eps))
    h_th, h_out, h_acc = torch.empty((B, D), dtype=torch.float64).pin_memory(), torch.empty((B, 8), dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory()
    h_info, h_st = torch.empty((B, 4), dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()

    def call(th_in, it, depth):
        ev.nuts_ptr(B, th_in.data_ptr(), h_base.data_ptr(), h_im.data_ptr(), h_eps.data_ptr(), 0, depth, 1, it, h_th.data_ptr(),
                    h_out.data_ptr(), h_acc.data_ptr(), h_info.data_ptr(), h_st.data_ptr())

    call(h_theta, 0, max_depth)   # warm-up / allocation
    for it in range(2):
        l0 = ev.kernel_launches()
        ev.set_kernel_timing(it == 1)
        t0 = time.perf_counter()
        call(h_theta, it + 1, max_depth)
        dt = time.perf_counter() - t0
        info, acc = h_info.numpy(), h_acc.numpy()
        nl = info[:, 1].astype(np.int64)
        ticks = int(nl.max())
        print(f"iteration {it}: {dt * 1e3:.1f} ms, {ticks} ticks ({dt * 1e3 / max(ticks, 1):.3f} ms/tick), leapfrog steps "
              f"{nl.sum()} ({nl.sum() / dt / 1e6:.2f} M useful gradient evals/s, lockstep efficiency {nl.sum() / (ticks * B):.2f}), "
              f"depth histogram {np.bincount(info[:, 0], minlength=max_depth + 1).tolist()}, diverged {int(info[:, 2].sum())}, "
              f"mean accept {acc.mean():.3f}, launches {ev.kernel_launches() - l0}")
        if it == 1:
            kms, ncalls = ev.kernel_times()
            print(f"   evaluation kernels inside the call: {ncalls} evaluations, residual {kms[0] / ncalls:.3f} + contraction "
                  f"{kms[1] / ncalls:.3f} + posterior {kms[2] / ncalls:.3f} ms each = {sum(kms):.1f} ms of the {dt * 1e3:.1f} ms")
            ev.set_kernel_timing(False)
        h_theta.copy_(h_th)
    ev.close()


if __name__ == "__main__":
    main()
