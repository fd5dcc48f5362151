#!/bin/bash
# development helper: posterior kernel occupancy cap sweep (rebuilds the library on the GPU box)
for mb in 3 4 2; do
  touch mcmc-date_b200/csrc/mcd_api.cu
  make -C mcmc-date_b200/csrc -s EXTRA=-DPOST_MINB=$mb 2>&1 | grep -E "error|posterior_kernelILi256ELi1ELb1" -A2 | grep -E "error|registers|spill"
  python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minb', $mb, round(d['value']), d['ms_per_step'], d['roofline']['step_share'], d['outputs_ok'])"
done
touch mcmc-date_b200/csrc/mcd_api.cu; make -C mcmc-date_b200/csrc -s 2>&1 | grep error
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
