"""Chains partitioned across the GPUs of one box (one process per GPU, torch.distributed).

The evaluation itself needs no collective: chains are independent and the model is replicated.
The only exchange is MC3's swap statistics -- (ln prior, ln likelihood) per chain -- which every
rank needs in order to take the same swap decisions.  In the reference MC3 lives in the third-party
`mcmc` package (called at app/Main.hs:476-479 with `MC3Settings (NChains 4) (SwapPeriod 2) (NSwaps 3)`);
heated chain i targets prior^beta_i * likelihood^beta_i.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_chains: int, world: int, rank: int):
    """Contiguous block of chains owned by `rank` (first `n_chains % world` ranks get one extra)."""
    base, rem = divmod(n_chains, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allgather_swap_stats(local_stats, world: int, dist=None):
    """local_stats: torch tensor [B_local, 2] = (ln prior, ln likelihood), same B_local on every rank.
    Returns [world * B_local, 2] on every rank (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    if world == 1 or dist is None:
        return local_stats
    out = torch.empty((world * local_stats.shape[0], local_stats.shape[1]), dtype=local_stats.dtype,
                      device=local_stats.device)
    dist.all_gather_into_tensor(out, local_stats.contiguous())
    return out


def mc3_swap_decisions(stats: np.ndarray, betas: np.ndarray, n_swaps: int, seed: int):
    """Propose `n_swaps` swaps between neighbouring heated chains and accept with the MC3 ratio
    ((prior*lik)_j / (prior*lik)_i)^(beta_i - beta_j).  Deterministic in (stats, betas, seed), so
    every rank that holds the gathered statistics reaches the same permutation.
    Returns the list of accepted (i, j) pairs."""
    rng = np.random.default_rng(seed)
    lp = stats[:, 0] + stats[:, 1]
    n = len(betas)
    accepted = []
    for _ in range(n_swaps):
        i = int(rng.integers(0, n - 1))
        j = i + 1
        log_ratio = (betas[i] - betas[j]) * (lp[j] - lp[i])
        u = rng.random()
        if np.isfinite(log_ratio) and np.log(u) < log_ratio:
            accepted.append((i, j))
            lp[[i, j]] = lp[[j, i]]
    return accepted
