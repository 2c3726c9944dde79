#!/bin/bash
# what the driver runs at round end, in its order: the GPU tests, smoke(), the reference arm, the bench (N = 1)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-final}
timeout 600 python -m pytest tests/ -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $OUT/${TAG}_smoke.log)"
timeout 600 python bench.py --impl reference --gpus 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "reference rc=$?"
timeout 900 python bench.py --gpus 1 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/${TAG}_bench_ref.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'ref', round(r['value'],1), 'ratio e2e', round(d['e2e']['value']/r['value'],1))
print('frontend', d['frontend_roofline']['frac'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['cpu_baseline'].items() if k!='sample'})
print('train', d['train']['ms_per_step'])
PY
