#!/usr/bin/env python
"""bench.py -- log-posterior + gradient evaluations / second (batched chains).

Workload (BASELINE.json configs[4], the one the north-star target is quoted on): synthetic 1000-leaf
tree (N = 1999 nodes, K = 1997 MVN dimensions, dense precision matrix), uncorrelated log-normal clock,
16 calibrations / 8 constraints / 4 braces.  One "step" = one batched evaluation of ln prior (three parts),
ln likelihood, ln Jacobian and the full HMC gradient for every chain.  Chains are independent and the model is
replicated, so there is no data-path collective.
  --scaling weak   (default) 8192 chains PER GPU;
  --scaling strong 8192 chains in total, 8192 / N per GPU (the split SURVEY 8e names).
With N > 1 the default (weak) line also carries a `strong_scaling` object measured in the same run.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scaling weak|strong] [--impl reference]

  value : chains/s with states resident in HBM (mcd_eval_grad_device), CUDA events, max over ranks
  e2e   : chains/s through the host-buffer C-ABI call mcd_eval_grad (pinned host states in, ln-posterior
          parts + gradient out; H2D and D2H copies inside the timed region)
  roofline     : the contraction kernel (INT8 tensor cores on digit planes by default, FP64 DMMA with
                 --contraction dmma), algorithmic 2 K^2 flops per chain
  cpu_baseline : the oracle's CPU port of the same evaluation on the box's host cores (bounded sample)

--impl reference times the CPU implementation (oracle port; the Haskell reference cannot be built in
this image) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_LEAVES = 1000
CHAINS_PER_GPU = 8192
METRIC = "log-posterior+gradient evals/sec (batched chains)"
UNIT = "evals/s"
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SMs x 64 FP64 FMA/clk/SM x 1965 MHz = 37.2
INT8_NOMINAL_TOPS = 4500.0                           # dense int8 tensor peak (2x the 2.25 PFLOP/s bf16 figure)
INT8_LIBRARY_GEMM_TOPS = 3079.7                      # cuBLASLt int8 GEMM measured on this pool (profiles/r01_int8_peak_measured.json)
GLOBAL_CHAINS_STRONG = 8192


def ncu_dram_bytes(oz_s):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the contraction kernel at the benchmark shape, from the
    tracked summary of the last `ncu --set full` capture of the shipped kernel (profiles/dram_traffic.json)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        e = t["gemm_f64_dmma_kernel" if oz_s == 0 else f"gemm_i8_ozaki_kernel<{oz_s}>"]
        return float(e["dram_read_bytes"]) + float(e["dram_write_bytes"])
    except Exception:
        return None


def workload_config(n_gpus, chains):
    mb = lambda x: f"{x * chains / 8192:.0f}"
    return {
        "workload": "synthetic 1000-leaf tree (N=1999 nodes, K=1997, dense precision), "
                    f"{chains} chains per GPU, value+gradient, uncorrelated log-normal clock, 16 cal / 8 con / 4 braces",
        "n_leaves": N_LEAVES, "chains_per_gpu": chains, "global_chains": chains * n_gpus,
        "parallelism": f"chains sharded over {n_gpus} GPU(s), model replicated",
        "l2": f"inputs larger than L2 (states {mb(262)} MB + residual planes {mb(117)} MB + gradient {mb(262)} MB per step vs 126 MB L2)"
              if chains >= 4096 else
              f"states {mb(262)} MB + residual planes {mb(117)} MB + contraction result {mb(134)} MB + gradient {mb(262)} MB per step "
              "exceed the 126 MB L2 together; every step streams all of them",
    }


def chains_per_gpu(args, world):
    if args.chains is not None:
        return args.chains
    if args.scaling == "strong":
        if GLOBAL_CHAINS_STRONG % world:
            raise SystemExit(f"--scaling strong: {GLOBAL_CHAINS_STRONG} chains do not divide over {world} GPUs")
        return GLOBAL_CHAINS_STRONG // world
    return CHAINS_PER_GPU


def bind_to_gpu_numa_node(local):
    """Rank placement for the host-buffer path: pin this process (and therefore the pinned buffers it allocates afterwards,
    first touch) to the CPUs of the NUMA node its GPU hangs off.  Best effort: returns a description for the JSON line."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                      text=True, timeout=20).strip().lower()
        dom, rest = bus.split(":", 1)
        dev = f"{dom[-4:]}:{rest}"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read())
        cpus = open(f"/sys/bus/pci/devices/{dev}/local_cpulist").read().strip()
        if node < 0 or not cpus:
            return {"numa_node": node, "bound": False}
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, ids)
        return {"numa_node": node, "bound": True, "cpus": cpus}
    except Exception as exc:  # no sysfs / nvidia-smi: run unpinned
        return {"bound": False, "why": type(exc).__name__}


def build_workload(chains, seed_offset=0):
    from mcmc_date_b200 import synth
    md, h = synth.synthetic_model(N_LEAVES, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, chains, seed=synth.BASE_SEED + 5 + seed_offset)
    return md, X


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            txt, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            txt = ""
        sm, smax, reasons = [], [], set()
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [x for x in sm if x > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------- CPU arm
def cpu_port_rate(md, X, threads, target_seconds=12.0):
    """Time the oracle's CPU port (value + analytic gradient) on a bounded sample of X."""
    from oracle import oracle as O
    orc = O.Oracle(md)
    probe = min(len(X), 4 * threads)
    t0 = time.perf_counter()
    orc.eval_grad(X[:probe], nthreads=threads)
    dt = time.perf_counter() - t0
    n = int(max(probe, min(len(X), probe * target_seconds / max(dt, 1e-6))))
    reps = max(1, int(round(1.5 * probe / max(dt, 1e-6) / n)))   # ~1.5 s wall on all threads (20-30 core-seconds)
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.eval_grad(X[:n], nthreads=threads)
    dt = time.perf_counter() - t0
    return n * reps / dt, n * reps, dt


def cpu_blas_contraction_rate(md, n=2048, reps=3):
    """Upper bound for ANY CPU implementation of the step: the contraction alone, Y = DX . Sigma^-1, as one batched DGEMM through
    the BLAS numpy links (OpenBLAS, all host threads) -- chains/s if everything else were free."""
    try:
        P = np.ascontiguousarray(md.precision, dtype=np.float64).reshape(md.dim, md.dim)
        DX = np.random.default_rng(0).normal(size=(n, md.dim))
        DX @ P
        t0 = time.perf_counter()
        for _ in range(reps):
            DX @ P
        dt = (time.perf_counter() - t0) / reps
        return {"value": n / dt, "unit": UNIT, "gflops": 2.0 * n * md.dim * md.dim / dt / 1e9,
                "note": f"contraction only ({n} chains x K x K DGEMM, numpy/OpenBLAS, all host threads): an upper bound for a CPU "
                        "implementation, not the reference's algorithm (it does one gemv per state, app/Probability.hs:169)"}
    except Exception as exc:
        return {"unavailable": type(exc).__name__}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    threads = O.max_threads()
    chains = chains_per_gpu(args, max(1, args.gpus))
    md, X = build_workload(max(chains, 4 * threads))
    orc = O.Oracle(md)
    probe = 4 * threads
    t0 = time.perf_counter()
    orc.eval_grad(X[:probe], nthreads=threads)
    rate0 = probe / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    n = int(max(threads, min(len(X), rate0 * min(budget, 15.0))))
    for _ in range(args.warmup):
        orc.eval_grad(X[:n], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.eval_grad(X[:n], nthreads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cfg = workload_config(args.gpus, chains)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} of {chains} chains per step (bounded sample of the same workload); "
                                   "oracle CPU port (C++ -O3 -march=native, std::thread over chains); the Haskell "
                                   "reference cannot be built here (no GHC)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------- roofline of the dominant kernel
def roofline(oz_s, K, B, gemm_ms, kms, ncalls, args):
    """The precision-matrix contraction Y = DX . Sigma^-1 dominates the step.  Algorithmic work: 2 K^2 FP64 flops per
    chain.  On the INT8 tensor pipe (default) the same product costs S(S+1)/2 int8 products of that
    size (S digit planes per operand, gemm_i8_ozaki.cuh); the roofline fraction is quoted on the pipe the kernel runs
    on, and the FP64-equivalent rate against the FP64 peak is given beside it."""
    flops_alg = 2.0 * K * K * B
    share = {"residual_ms": kms[0] / max(1, ncalls), "contraction_ms": gemm_ms, "posterior_ms": kms[2] / max(1, ncalls)}
    traffic = args.traffic if args.traffic is not None else ncu_dram_bytes(oz_s)
    if B != CHAINS_PER_GPU:
        traffic = None   # the capture is of the 8192-chain launch
    fp64_equiv = flops_alg / (gemm_ms * 1e-3) / 1e12
    if oz_s == 0:
        return {"kernel": "gemm_f64_dmma_kernel", "bound": "tensor", "achieved": fp64_equiv, "peak": FP64_NOMINAL_TFLOPS,
                "unit": "TFLOP/s", "frac": fp64_equiv / FP64_NOMINAL_TFLOPS, "traffic": traffic,
                "peak_source": "nominal FP64 (148 SMs x 64 FMA/clk x 1965 MHz); MEASURED_PEAKS.json has no FP64 entry; "
                               "cublasDgemm on this shape measured 35.6 TFLOP/s executed (profiles/)",
                "kernel_ms": gemm_ms, "algorithmic_flops_per_launch": flops_alg, "step_share": share,
                "hbm_kernels": hbm_kernels(oz_s, K, B, kms, ncalls)}
    hbm = hbm_kernels(oz_s, K, B, kms, ncalls)
    pairs = oz_s * (oz_s + 1) // 2
    ops = pairs * flops_alg                                # int8 multiply-adds x 2 the split needs, unpadded K
    achieved = ops / (gemm_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return {"kernel": f"gemm_i8_ozaki_v2_kernel<{oz_s}>", "bound": "tensor", "achieved": achieved, "peak": INT8_NOMINAL_TOPS,
            "unit": "TOP/s (int8)", "frac": achieved / INT8_NOMINAL_TOPS, "traffic": traffic,
            "peak_source": "nominal dense INT8 tensor peak (4.5 POP/s); MEASURED_PEAKS.json has no int8 entry -- the figure derived "
                           f"from it is twice its measured bf16 burst, {2 * peaks.get('bf16_tflops', 1662.5):.0f} TOP/s "
                           "(frac_derived_from_measured)",
            "frac_derived_from_measured": achieved / (2 * peaks.get("bf16_tflops", 1662.5)),
            "frac_of_measured": {
                "vs_2x_bf16_burst_of_MEASURED_PEAKS": achieved / (2 * peaks.get("bf16_tflops", 1662.5)),
                "vs_library_int8_gemm_measured": achieved / INT8_LIBRARY_GEMM_TOPS,
                "library_int8_gemm_tops": INT8_LIBRARY_GEMM_TOPS,
                "note": "cuBLASLt int8 GEMM (torch._int_mm, 8192^3, best of 10) measured on this pool's B200 with "
                        "tools/int8_peak.py -> profiles/r01_int8_peak_measured.json"},
            "kernel_ms": gemm_ms, "algorithmic_int8_ops_per_launch": ops, "int8_products": pairs,
            "algorithmic_flops_per_launch": flops_alg,
            "fp64_equivalent": {"achieved_tflops": fp64_equiv, "fp64_nominal_peak_tflops": FP64_NOMINAL_TFLOPS,
                                "ratio_to_fp64_peak": fp64_equiv / FP64_NOMINAL_TFLOPS,
                                "note": "2 K^2 FP64 flops per chain delivered by exact int8 products (Ozaki split); "
                                        "the FP64 DMMA kernel it replaces reaches 0.92 of that peak"},
            "step_share": share, "hbm_kernels": hbm}


def hbm_kernels(oz_s, K, B, kms, ncalls):
    """The two HBM-side kernels of the step against the measured copy bandwidth (north_star: "achieved HBM GB/s for the
    gather and prior kernels"): algorithmic bytes per chain x chains / CUDA-event time."""
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        peak, src = 6650.0, "fallback 6.65 TB/s (of fallback)"
    N = K + 2
    S = 5 + 2 * N
    ld8 = (K + 63) // 64 * 64
    k1_bytes = B * (8 * S + (oz_s * ld8 if oz_s else 8 * K))     # state row in, digit planes (or FP64 residuals) out
    k3_bytes = B * (8 * S + 8 * K + 8 * S + 64)                  # state row + y in, gradient + ln-posterior parts out
    out = []
    for name, nbytes, ms in (("residual_split_kernel (tree gather + residuals + digit planes)", k1_bytes, kms[0] / max(1, ncalls)),
                             ("posterior_kernel (priors, Jacobian, gradient)", k3_bytes, kms[2] / max(1, ncalls))):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        out.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                    "algorithmic_bytes_per_launch": nbytes, "kernel_ms": ms, "peak_source": src})
    out[1]["note"] = ("ncu (profiles/r02_ncu_summary.md): 199 M warp instructions, 66 % of the issue slots, FP64 pipe 36 % -- the FP64 "
                      "arithmetic per node (logarithm, exponential, reciprocals) and the CTA-wide barriers between the passes bind "
                      "this kernel before HBM does")
    return out


# --------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from mcmc_date_b200 import binding, model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): whatever libraries print while the run lasts (NCCL announces its
    # version on stdout when the first communicator is created) goes to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = bind_to_gpu_numa_node(local) if not args.no_numa else {"bound": False, "why": "--no-numa"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = chains_per_gpu(args, world)
    md, X = build_workload(B, seed_offset=rank)
    S, K = md.state_len, md.dim
    ev = binding.Evaluator(md, device=local, max_batch=B)
    if args.contraction:
        ev.set_contraction(args.contraction)
    oz_s = ev.get_contraction()      # 0: FP64 DMMA, 6..8: INT8 tensor cores with that many digit planes

    d_states = torch.from_numpy(X).to(dev)
    d_out = torch.empty((B, model.OUT_COLS), dtype=torch.float64, device=dev)
    d_grad = torch.empty((B, S), dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    # MC3 swap statistics: (ln prior, ln likelihood) of every chain to every rank, every step.  The exchange is not on the
    # evaluation's critical path (swaps are decided every SwapPeriod iterations), so it runs on a side stream from a
    # double-buffered copy of the two columns and overlaps the next step's kernels.
    if world > 1:
        side = torch.cuda.Stream(device=dev)
        stats = [torch.empty((B, 2), dtype=torch.float64, device=dev) for _ in range(2)]
        gathered = [torch.empty((world * B, 2), dtype=torch.float64, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        for e in done:
            e.record(stream)
    step_no = [0]

    def step():
        ev.eval_grad_device(B, d_states.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_status.data_ptr(),
                            stream.cuda_stream)
        if world > 1:
            k = step_no[0] & 1
            step_no[0] += 1
            stream.wait_event(done[k])               # the exchange that last used this buffer pair has finished
            stats[k].copy_(d_out[:, 3:5])
            ready[k].record(stream)
            side.wait_event(ready[k])
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(gathered[k], stats[k])
                done[k].record(side)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ev.kernel_launches()
    ev.set_kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    if world > 1:  # the last exchanges belong to the timed region
        stream.wait_event(done[0])
        stream.wait_event(done[1])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kms, ncalls = ev.kernel_times()
    ev.set_kernel_timing(False)
    launches = ev.kernel_launches() - launches0
    # the same step on the strong-scaling split (8192 chains in total, 8192 / N on this GPU), measured in the same run
    strong = None
    if world > 1 and args.scaling == "weak" and args.chains is None and GLOBAL_CHAINS_STRONG % world == 0:
        Bs = GLOBAL_CHAINS_STRONG // world

        def strong_step():
            ev.eval_grad_device(Bs, d_states.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_status.data_ptr(), stream.cuda_stream)

        for _ in range(3):
            strong_step()
        barrier()
        ev.set_kernel_timing(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            strong_step()
        s1.record()
        barrier()
        skms, sncalls = ev.kernel_times()
        ev.set_kernel_timing(False)
        strong = {"ms": s0.elapsed_time(s1), "chains_per_gpu": Bs, "kernel_ms": [x / max(1, sncalls) for x in skms]}
    # value-only evaluation (what the Metropolis-Hastings proposals need): triangular contraction on the Cholesky factor
    ev.eval_device(B, d_states.data_ptr(), d_out.data_ptr(), d_status.data_ptr(), stream.cuda_stream)  # factorises once
    for _ in range(2):
        ev.eval_device(B, d_states.data_ptr(), d_out.data_ptr(), d_status.data_ptr(), stream.cuda_stream)
    barrier()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    for _ in range(args.steps):
        ev.eval_device(B, d_states.data_ptr(), d_out.data_ptr(), d_status.data_ptr(), stream.cuda_stream)
    v1.record()
    barrier()
    value_only_ms = v0.elapsed_time(v1) / args.steps
    # leave d_out as the value+gradient call wrote it (the parity guard below compares it with the host entry points)
    ev.eval_grad_device(B, d_states.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_status.data_ptr(), stream.cuda_stream)
    barrier()
    # --- end-to-end through the host-buffer C ABI (pinned host memory, copies inside) ---------
    h_states = torch.from_numpy(X).pin_memory()
    h_out = torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory()
    h_grad = torch.empty((B, S), dtype=torch.float64).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()

    def e2e_state_step():
        ev.eval_grad_ptr(B, h_states.data_ptr(), h_out.data_ptr(), h_grad.data_ptr(), h_status.data_ptr())

    # HMC form (what NUTS exchanges, app/Hamiltonian.hs:49-60): packed position vectors in, packed gradient out
    D = ev.D
    mask = ev.mask().astype(bool)
    h_theta = torch.from_numpy(np.ascontiguousarray(X[:, mask][:, ::-1])).pin_memory()
    h_base = torch.from_numpy(X[0].copy()).pin_memory()
    h_gtheta = torch.empty((B, D), dtype=torch.float64).pin_memory()
    h_out2 = torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory()

    def e2e_step():
        ev.eval_grad_theta_ptr(B, h_theta.data_ptr(), h_base.data_ptr(), h_out2.data_ptr(), h_gtheta.data_ptr(),
                               h_status.data_ptr())

    # the same call, double buffered: step i + 1 is queued before the host waits for step i (mcd_eval_grad_theta_async /
    # mcd_wait), so the PCIe fill of one step overlaps the drain of the previous one.  Every step still moves its own
    # inputs to the device and its own results back to (a second set of) pinned host buffers.
    h_theta_b = h_theta.clone().pin_memory()
    h_gtheta_b = torch.empty((B, D), dtype=torch.float64).pin_memory()
    h_out2_b = torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory()
    h_status_b = torch.empty(B, dtype=torch.int32).pin_memory()
    bufs = ((h_theta, h_out2, h_gtheta, h_status), (h_theta_b, h_out2_b, h_gtheta_b, h_status_b))

    def e2e_pipelined(n_steps):
        prev = None
        for i in range(n_steps):
            th, oo, gg, ss = bufs[i & 1]
            t = ev.eval_grad_theta_async_ptr(B, th.data_ptr(), h_base.data_ptr(), oo.data_ptr(), gg.data_ptr(), ss.data_ptr())
            if prev is not None:
                ev.wait(prev)          # step i - 1 is complete: its buffers are the host's again
            prev = t
        ev.wait(prev)

    # HMC trajectory form (mcd_leapfrog): positions + momenta in, end point + energies out, TRAJ_L leapfrog steps
    # (= TRAJ_L + 1 value+gradient evaluations per chain) resident on the device in between
    TRAJ_L = 10
    h_mom = torch.from_numpy(np.random.default_rng(11 + rank).normal(size=(B, D)) * 1e-3).pin_memory()
    h_invm = torch.ones(D, dtype=torch.float64).pin_memory()
    h_eps = torch.full((B,), 1e-6, dtype=torch.float64).pin_memory()
    h_th_out = torch.empty((B, D), dtype=torch.float64).pin_memory()
    h_mom_out = torch.empty((B, D), dtype=torch.float64).pin_memory()
    h_out3 = torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory()
    h_energy = torch.empty((B, 2), dtype=torch.float64).pin_memory()
    h_status3 = torch.empty(B, dtype=torch.int32).pin_memory()

    def traj_step():
        ev.leapfrog_ptr(B, TRAJ_L, h_theta.data_ptr(), h_mom.data_ptr(), h_base.data_ptr(), h_invm.data_ptr(),
                        h_eps.data_ptr(), h_th_out.data_ptr(), h_mom_out.data_ptr(), h_out3.data_ptr(),
                        h_energy.data_ptr(), h_status3.data_ptr())

    for _ in range(2):
        e2e_state_step()
        e2e_step()
    traj_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 4)):
        traj_step()
    torch.cuda.synchronize()
    traj_s = (time.perf_counter() - t0) / max(1, args.steps // 4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_state_step()
    torch.cuda.synchronize()
    e2e_state_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_pipelined(2)
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_pipe_s = time.perf_counter() - t0
    pipe_ok = bool(torch.equal(h_gtheta_b, h_gtheta)) and bool(torch.equal(h_out2_b, h_out2))
    # Metropolis-Hastings steps on chains resident in HBM: one proposal + evaluation + accept per step
    ev.chains_set(X)
    mh = {}
    for name, kind, par in (("slide_node_incremental", binding.MH_SLIDE_NODE, 0.002),
                            ("scale_rate_mean_and_tree_contrarily_no_contraction", binding.MH_SCALE_NORM_TREE_CONTRA_M, 3000.0),
                            ("scale_variance_and_tree_from_scratch", binding.MH_SCALE_VAR_TREE, 3000.0)):
        ev.mh_cycle([(kind, -1, par, 1.0, 0, 3)], 1, seed=3, iteration0=0)
        ev.synchronize()
        n_mh = 10 * args.steps
        t0 = time.perf_counter()
        acc, _, _ = ev.mh_cycle([(kind, -1, par, 1.0, 0, n_mh)], 1, seed=3, iteration0=10)
        dt = time.perf_counter() - t0
        mh[name] = {"value": B * world * n_mh / dt, "unit": "proposals/s", "us_per_step": 1e6 * dt / n_mh,
                    "acceptance": float(acc[0]) / (n_mh * B)}
    mh["note"] = ("mcd_mh_cycle on this rank's chains x world (per-rank wall time incl. the final synchronisation); small moves "
                  "are scored from the cached contraction result, moves that leave the distances unchanged by the posterior "
                  "kernel against it, the remaining global moves by the full value-only evaluation")
    clocks = sampler.stop() if rank == 0 else None
    # parity guard: the timed outputs are the real thing (finite, and equal through both entry points)
    ok = bool(torch.isfinite(d_out[:, 6]).all().item()) and bool(
        torch.allclose(d_out.cpu()[:, :7], h_out[:, :7], rtol=0, atol=0, equal_nan=True)) and bool(
        torch.equal(h_out2, h_out)) and bool(torch.equal(h_gtheta, torch.from_numpy(
            np.ascontiguousarray(h_grad.numpy()[:, mask][:, ::-1]))))

    t = torch.tensor([ms, e2e_s * 1e3, e2e_state_s * 1e3, traj_s * 1e3, e2e_pipe_s * 1e3, strong["ms"] if strong else 0.0],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e_state_ms_max, traj_ms_max = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    e2e_pipe_ms_max, strong_ms_max = float(t[4]), float(t[5])
    if rank == 0:
        total = B * world
        value = total * args.steps / (ms_max * 1e-3)
        e2e_value = total * args.steps / (e2e_pipe_ms_max * 1e-3)
        e2e_sync_value = total * args.steps / (e2e_ms_max * 1e-3)
        gemm_ms = kms[1] / max(1, ncalls)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.chains is not None else args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, B), "global_chains": total,
            "contraction": "fp64 dmma" if oz_s == 0 else f"int8 tensor cores, {oz_s} base-256 digit planes per operand (error-free split)",
            "rank_placement": placement,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * 8 + S * 8,
                    "d2h_bytes_per_step": B * (D + model.OUT_COLS) * 8 + B * 4,
                    "note": "mcd_eval_grad_theta_async / mcd_wait on two sets of pinned host buffers (HMC position vectors in, packed "
                            "gradient + ln-posterior parts out; step i + 1 is queued before the host waits for step i, every "
                            "step moves its own inputs and results over PCIe); bytes per rank",
                    "synchronous_call": {"value": e2e_sync_value, "unit": UNIT,
                                         "note": "mcd_eval_grad_theta: one call at a time, the host waits for every call"},
                    "full_state_api": {"value": total * args.steps / (e2e_state_ms_max * 1e-3), "unit": UNIT,
                                       "h2d_bytes_per_step": B * S * 8,
                                       "d2h_bytes_per_step": B * (S + model.OUT_COLS) * 8 + B * 4,
                                       "note": "mcd_eval_grad (full canonical states in, full-layout gradient out)"},
                    "hmc_trajectory_api": {"value": total * (TRAJ_L + 1) / (traj_ms_max * 1e-3), "unit": UNIT,
                                           "leapfrog_steps": TRAJ_L, "ms_per_call": traj_ms_max,
                                           "h2d_bytes_per_step": 2 * B * D * 8 + (S + D + B) * 8,
                                           "d2h_bytes_per_step": B * (2 * D + model.OUT_COLS + 2) * 8 + B * 4,
                                           "note": "mcd_leapfrog: one call = positions + momenta in, TRAJ_L leapfrog steps "
                                                   "(TRAJ_L + 1 value+gradient evaluations per chain) resident in HBM, end "
                                                   "point + energies out; what the reference's Hamiltonian proposal asks for"}},
            "gpu_launches": int(launches),
            "value_only": {"value": B * world / (value_only_ms * 1e-3), "unit": "evals/s", "ms_per_step": value_only_ms,
                           "note": "mcd_eval_device (no gradient), per-rank time of this rank x world; quadratic form from the "
                                   "triangular Cholesky-factor contraction"},
            "mh": mh,
            "roofline": roofline(oz_s, K, B, gemm_ms, kms, ncalls, args),
            "clocks": clocks, "outputs_ok": ok and pipe_ok,
        }
        if strong:
            line["strong_scaling"] = {
                "value": GLOBAL_CHAINS_STRONG * args.steps / (strong_ms_max * 1e-3), "unit": UNIT, "scaling": "strong",
                "global_chains": GLOBAL_CHAINS_STRONG, "chains_per_gpu": strong["chains_per_gpu"],
                "ms_per_step": strong_ms_max / args.steps,
                "kernel_ms_rank0": {"residual": strong["kernel_ms"][0], "contraction": strong["kernel_ms"][1], "posterior": strong["kernel_ms"][2]},
                "contraction_tiles": f"{(strong['chains_per_gpu'] + 127) // 128} x {(K + 63) // 64} tiles of 128 chains x 64 rows on {torch.cuda.get_device_properties(local).multi_processor_count} persistent CTAs",
                "note": "same step, 8192 chains in total split evenly over the GPUs (SURVEY 8e); device-resident, max over ranks; "
                        "speed-up over one GPU = value / the one-GPU value of this bench"}
        if world == 1 and not args.no_cpu:
            from oracle import oracle as O
            threads = O.max_threads()
            rate, n, dt = cpu_port_rate(md, X, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n} chains of the same workload in {dt:.1f} s wall on {threads} threads "
                                              f"= {dt * threads:.0f} core-seconds (oracle CPU port, std::thread over chains)",
                                    "best_case_blas_contraction": cpu_blas_contraction_rate(md)}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ev.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (overrides --scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 8192 chains per GPU; strong: 8192 chains in total, split evenly over the GPUs")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the ranks to their GPU's NUMA node")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch of the contraction kernel from ncu (default: the committed capture)")
    ap.add_argument("--contraction", default=None, choices=["dmma", "i8s6", "i8s7"],
                    help="arithmetic pipe of the contraction (default: the library's default, i8s7)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
