#!/bin/bash
# development helper: e2e throughput vs pipeline chunk size
for c in ${@:-256 512 1024 2048}; do
  MCD_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk', $c, 'theta', round(d['e2e']['value']), 'state', round(d['e2e']['full_state_api']['value']))"
done
