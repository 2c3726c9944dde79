#!/bin/bash
# Round-2 visit A: GPU tests (incl. the new mirror tests) + every bench workload at N=1.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=r2a
python -m pytest -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
for w in config3 config5 config4; do
  python bench.py --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; echo "$w rc=$?"; tail -3 $OUT/${TAG}_bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), (d.get('frontend_roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
