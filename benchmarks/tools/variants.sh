for v in 3 8 9; do
  for S in 1024 32768; do
    Y=50; if [ $S = 32768 ]; then Y=5; fi
    B200LP_VARIANT=$v python bench.py --scenarios $S --years $Y --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant', d['kernel_variant'], 'S', d['config']['scenarios_per_gpu'], 'value %.1fM' % (d['value']/1e6), 'ms', round(d['ms_per_step'],2), 'fail', d['failed_scenarios'])"
  done
done
