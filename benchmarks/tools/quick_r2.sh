for w in "thames 512" "thames 1024" "thames 4096" "two_reservoir 1024"; do set -- $w
python bench.py --workload $1 --scenarios $2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > /tmp/o.json 2>/tmp/o.err; python -c "
import json,sys; d=json.load(open('/tmp/o.json')); print('$1 S=$2', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],3), 'ms', d.get('failed_scenarios'), d['pivots_per_scenario_timestep'])"; done
