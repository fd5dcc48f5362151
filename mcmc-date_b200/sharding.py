"""Chains partitioned across the GPUs of one box (one process per GPU, torch.distributed).

The evaluation itself needs no collective: chains are independent and the model is replicated.
The only exchange is MC3's swap statistics -- (ln prior, ln likelihood) per chain -- which every
rank needs in order to take the same swap decisions.  In the reference MC3 lives in the third-party
`mcmc` package (called at app/Main.hs:476-479 with `MC3Settings (NChains 4) (SwapPeriod 2) (NSwaps 3)`);
heated chain i targets prior^beta_i * likelihood^beta_i.
"""
from __future__ import annotations


def shard_range(n_chains: int, world: int, rank: int):
    """Contiguous block of chains owned by `rank`.  The chains must divide evenly over the ranks: the all-gather of the swap
    statistics and the slot tables of mcd_mc3_configure assume the same number of chains on every rank."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("shard_range: rank out of range")
    if n_chains % world != 0:
        raise ValueError(f"shard_range: {n_chains} chains do not divide evenly over {world} ranks")
    per = n_chains // world
    return rank * per, (rank + 1) * per


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
