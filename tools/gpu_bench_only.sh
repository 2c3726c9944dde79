#!/bin/bash
# the default bench line alone (N = 1), with its e2e diagnostics printed
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-b}
timeout 900 python bench.py --gpus 1 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(e['value']), e['ms_per_step_of_each_repeat'], 'warm', e['warm_runs'], e['host_ms_per_step']['enqueue'], e['host_ms_per_step']['waiting_in_collect'])
print('ceiling', e['h2d_copy_bound']['utt_s'], 'clocks', d['clocks'], 'cpu1', d['cpu_baseline'].get('one_thread_value'))
PY
