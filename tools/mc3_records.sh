#!/bin/bash
# MC3 records of BASELINE.json configs[3] on 1/2/4/8 GPUs of one box (run under `gpurun --gpus 8`) -> gpurun_out/r02_mc3/
set -u
OUT=gpurun_out/r02_mc3
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29620 + n)) tools/mc3_bench.py 1 300 > $OUT/mc3_n$n.log 2> $OUT/mc3_n$n.err
done
$TR --nproc-per-node 8 --master-port 29641 tools/mc3_bench.py 1024 100 > $OUT/mc3_65536chains_n8.log 2> $OUT/mc3_65536chains_n8.err
MCD_MH_PER_STEP=1 $TR --nproc-per-node 8 --master-port 29642 tools/mc3_bench.py 1 300 > $OUT/mc3_perstep_n8.log 2> $OUT/mc3_perstep_n8.err
tail -qn 1 $OUT/mc3_*.log | cut -c1-600
