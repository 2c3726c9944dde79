#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-tr}
timeout 600 python -m pytest -m gpu -x -q tests/test_gpu_train.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --workload config4 > $OUT/${TAG}_bench4.json 2> $OUT/${TAG}_bench4.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench4.err
timeout 600 python bench.py --workload config4 --no-graph --no-cpu-baseline > $OUT/${TAG}_bench4_eager.json 2>> $OUT/${TAG}_bench4.err; echo "bench eager rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench4.json','gpurun_out/${TAG}_bench4_eager.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['train']['cuda_graph'][:40], d['train']['gpu_launches_per_step'], d.get('cpu_baseline',{}))
    except Exception as e: print(f,'ERR',e)
PY
