#!/bin/bash
# Multi-GPU records of one box (run under `gpurun --gpus 8`): PCIe bandwidth with 1/2/4/8 concurrent processes, the MC3
# configuration of BASELINE.json configs[3] on 1/2/4/8 GPUs, bench.py at 2/4/8 GPUs (weak line + strong_scaling object).
# Everything goes to gpurun_out/r02_multi/.
set -u
OUT=gpurun_out/r02_multi
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | head -30 > $OUT/lscpu.txt 2>&1
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29600 + n)) tools/pcie_multi.py > $OUT/pcie_n$n.json 2> $OUT/pcie_n$n.err
  PCIE_NUMA=1 $TR --nproc-per-node $n --master-port $((29610 + n)) tools/pcie_multi.py > $OUT/pcie_numa_n$n.json 2> $OUT/pcie_numa_n$n.err
done
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29620 + n)) tools/mc3_bench.py 1 300 > $OUT/mc3_n$n.log 2> $OUT/mc3_n$n.err
done
$TR --nproc-per-node 8 --master-port 29641 tools/mc3_bench.py 1024 100 > $OUT/mc3_65536chains_n8.log 2> $OUT/mc3_65536chains_n8.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29630 + n)) bench.py --gpus $n --steps 20 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -n 2 $OUT/*.json $OUT/mc3_*.log | cut -c1-400
