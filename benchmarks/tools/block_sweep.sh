#!/usr/bin/env bash
# CTA size sweep for the run kernel (B200LP_BLOCK), under gpurun.
cd "$(dirname "$0")/../.."
for S in 1024 2048; do
  for b in 32 64 128; do
    Y=50; [ $S = 2048 ] && Y=10
    B200LP_BLOCK=$b python bench.py --scenarios $S --years $Y --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S', $S, 'block', $b, round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms')"
  done
done
