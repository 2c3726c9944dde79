#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2}
timeout 900 python -m pytest -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench.err
for w in config3 config5; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; echo "$w rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), (d.get('frontend_roofline') or {}).get('frac'), {k:v['ms_per_step'] for k,v in d.get('stages',{}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
