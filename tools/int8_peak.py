#!/usr/bin/env python
"""Measured dense INT8 tensor throughput of this GPU (library GEMM): the denominator MEASURED_PEAKS.json lacks for the
roofline of gemm_i8_ozaki_kernel.  torch._int_mm (cuBLASLt int8 x int8 -> int32), same recipe as the driver's bf16 figure:
best of 10 (burst) and back to back for 4 s (sustained).  Prints one JSON line."""
import json
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    res = {}
    for n in (8192, 16384):
        a = torch.randint(-127, 127, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-127, 127, (n, n), dtype=torch.int8, device=dev).t().contiguous().t()
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k = 0
        while time.perf_counter() - t0 < 4.0:
            for _ in range(10):
                torch._int_mm(a, b)
            k += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / k
        ops = 2.0 * n ** 3
        res[str(n)] = {"burst_tops": ops / (best * 1e-3) / 1e12, "sustained_tops": ops / (sus * 1e-3) / 1e12}
    print(json.dumps({"int8_tops": res, "how": "torch._int_mm n^3, best of 10 / 4 s back to back, CUDA events",
                      "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
