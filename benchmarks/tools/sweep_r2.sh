for S in 512 1024 2048 4096; do for v in 1 2; do for b in 128 64; do
B200LP_VARIANT=$v B200LP_BLOCK=$b python bench.py --workload thames --scenarios $S --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /tmp/o.json 2>/tmp/o.err; python -c "
import json,sys; d=json.load(open('/tmp/o.json')); print('thames S=$S variant=$v block=$b', round(d['value']/1e6,1), round(d['ms_per_step'],2), d.get('failed_scenarios'))"; done; done; done
for S in 512 1024 2048; do for b in 128 64 32; do
B200LP_BLOCK=$b python bench.py --workload two_reservoir --scenarios $S --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /tmp/o.json 2>/tmp/o.err; python -c "
import json,sys; d=json.load(open('/tmp/o.json')); print('two_reservoir S=$S block=$b', round(d['value']/1e6,1), round(d['ms_per_step'],2), d.get('failed_scenarios'))"; done; done
