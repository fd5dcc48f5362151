#!/usr/bin/env python
"""Host <-> device copy bandwidth of N concurrent processes, one per GPU (the traffic pattern of bench.py's host-buffer `e2e`
leg at N GPUs): every rank copies pinned host memory to its GPU and back at the same time as all the others.  With PCIE_NUMA=1 in the environment
each rank first pins itself to the CPUs of its GPU's NUMA node (bench.bind_to_gpu_numa_node), so that the pinned buffers are
allocated there.  Rank 0 prints one JSON line: per-rank and aggregate GB/s for H2D alone, D2H alone and both together.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_multi.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = bench.bind_to_gpu_numa_node(local) if bool(os.environ.get("PCIE_NUMA")) else {"bound": False}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(0)           # first touch on this rank's (possibly pinned) CPUs
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = 8

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(h2d, d2h):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        cur = torch.cuda.current_stream()
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9     # GB/s per direction

    for _ in range(2):
        run(True, True)
    res = torch.tensor([run(True, False), run(False, True), run(True, True)], dtype=torch.float64, device=dev)
    allr = [torch.empty_like(res) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, res)
    else:
        allr = [res]
    if rank == 0:
        t = torch.stack(allr).cpu()
        print(json.dumps({"n_processes": world, "numa_pinned": bool(os.environ.get("PCIE_NUMA")), "rank0_placement": placement,
                          "per_rank_GBps": {"h2d_alone": t[:, 0].tolist(), "d2h_alone": t[:, 1].tolist(), "both_each_direction": t[:, 2].tolist()},
                          "aggregate_GBps": {"h2d_alone": float(t[:, 0].sum()), "d2h_alone": float(t[:, 1].sum()),
                                             "both_each_direction": float(t[:, 2].sum()), "both_total": 2 * float(t[:, 2].sum())}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
