#!/usr/bin/env python
"""BASELINE.json configs[3]: mtCDNApri-shaped data set (7 leaves), autocorrelated rates, MC3 with 64 heated chains spread
over the GPUs of one box (one process per GPU; torchrun).  Every iteration = one sweep of the reference's proposal cycle
(mcmc-date_b200/mh_cycle.py -> mcd_mh_cycle) on the resident chains, then -- every SWAP_PERIOD iterations -- the all-gather of
(ln prior, ln likelihood) over NCCL (mcd_allgather_stats: the library's own communicator, set up from an id that rank 0 creates and
torch.distributed broadcasts; MC3_TORCH_ALLGATHER=1 uses torch.distributed's all-gather instead) and N_SWAPS slot swaps
(mcd_mc3_swap), decided identically on every rank
(`MC3Settings (NChains ..) (SwapPeriod 2) (NSwaps 3)`, app/Main.hs:477).
usage: [torchrun --nproc-per-node N] tools/mc3_bench.py [groups=1] [iterations=200]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_date_b200 import binding, mh_cycle, model  # noqa: E402

N_CHAINS, SWAP_PERIOD, N_SWAPS = 64, 2, 3


def main():
    groups = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    z = np.load(os.path.join(ROOT, "tests", "golden", "mtcdnapri-7-leaves.npz"))    # model + states of the 7-taxon data set
    md = model.ModelDesc(parent=z["parent"], mean=z["mean"], precision=z["precision"], logdet_sigma=float(z["logdet_sigma"]),
                         clock_model=model.AUTOCORRELATED_LOGNORMAL, likelihood=int(z["likelihood"]), ht=float(z["ht"]),
                         cal_node=z["cal_node"], cal_lo=z["cal_lo"], cal_lo_p=z["cal_lo_p"], cal_hi=z["cal_hi"], cal_hi_p=z["cal_hi_p"])
    n_global = groups * N_CHAINS
    assert n_global % world == 0, "chains must divide over the ranks"
    B = n_global // world
    valid = np.asarray(z["states"], float)[:int(z["n_valid"])]
    X0 = np.tile(valid[0], (n_global, 1))                     # all chains start from the same valid state
    ev = binding.Evaluator(md, device=local, max_batch=B)
    ev.chains_set(X0[rank * B:(rank + 1) * B])
    ladder = 1.0 / (1.0 + 0.05 * np.arange(N_CHAINS))         # incremental heating, beta_i = 1 / (1 + i dT)
    ev.mc3_configure(n_global, rank * B, N_CHAINS, ladder, ladder)
    lib_comm = world > 1 and not os.environ.get("MC3_TORCH_ALLGATHER")
    if lib_comm:   # the library's communicator: rank 0 creates the NCCL id, everybody joins
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(binding.Evaluator.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ev.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    props = mh_cycle.reference_cycle(md)
    steps_per_sweep = sum(p[5] for p in props)
    local_stats = torch.empty((B, 2), dtype=torch.float64, device=dev)
    gathered = torch.empty((n_global, 2), dtype=torch.float64, device=dev)

    def iteration(it, k):
        acc, inv, k = ev.mh_cycle(props, 1, seed=17, iteration0=k)
        if it % SWAP_PERIOD == SWAP_PERIOD - 1:
            if lib_comm:
                ev.allgather_stats(gathered.data_ptr())        # (ln prior, ln lik) of every rank's resident chains
            else:
                ev.chains_stats_device(local_stats.data_ptr())
                if world > 1:
                    dist.all_gather_into_tensor(gathered, local_stats)
                else:
                    gathered.copy_(local_stats)
                torch.cuda.synchronize()
            for s in range(N_SWAPS):
                ev.mc3_swap(-1, seed=23, iteration=it * N_SWAPS + s, d_stats_global=gathered.data_ptr(), want_accepted=False)
        return acc, k

    k = 0
    for it in range(5):
        _, k = iteration(it, k)
    ev.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tot_acc = 0
    for it in range(iters):
        acc, k = iteration(it, k)
        tot_acc += int(acc.sum())
    ev.synchronize()
    dt = time.perf_counter() - t0
    slots = ev.mc3_slots()
    ok = bool((np.sort(slots.reshape(groups, N_CHAINS), axis=1) == np.arange(N_CHAINS)).all())
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        sl = torch.from_numpy(slots.astype(np.int64)).to(dev)
        sl0 = sl.clone()
        dist.broadcast(sl0, 0)
        same = torch.tensor([int(torch.equal(sl, sl0))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        ok = ok and bool(same.item())
    X, out, st = ev.chains_get()
    if rank == 0:
        moved = int((slots != np.arange(n_global) % N_CHAINS).sum())
        print(f"MC3: {groups} group(s) x {N_CHAINS} heated chains on {world} GPU(s), {len(props)} proposals in the cycle: {iters} iterations x "
              f"{steps_per_sweep} proposal steps in {dt * 1e3:.1f} ms = {iters / dt:.1f} iterations/s, "
              f"{iters * steps_per_sweep * n_global / dt / 1e6:.2f} M proposals/s; acceptance "
              f"{tot_acc / (iters * steps_per_sweep * B):.2f}; chains off their initial slot {moved}/{n_global}; "
              f"slots a permutation per group and identical on all ranks: {ok}; finite posteriors {np.isfinite(out[:, 6]).mean():.2f}")
        import json
        print(json.dumps({"config": "mtCDNApri 7 leaves, autocorrelated log-normal clock, MC3", "groups": groups, "heated_chains_per_group": N_CHAINS,
                          "n_gpus": world, "chains_per_gpu": B, "iterations": iters, "proposal_steps_per_iteration": steps_per_sweep,
                          "swap_period": SWAP_PERIOD, "n_swaps": N_SWAPS, "iterations_per_s": iters / dt,
                          "proposals_per_s": iters * steps_per_sweep * n_global / dt, "acceptance": tot_acc / (iters * steps_per_sweep * B),
                          "chains_off_initial_slot": moved, "slot_tables_identical_on_all_ranks": ok,
                          "allgather": "mcd_allgather_stats (library NCCL communicator)" if lib_comm else "torch.distributed / local copy"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ev.close()


if __name__ == "__main__":
    main()
