#!/bin/bash
# quick check after a kernel change: the tensor-core contraction tests, the pipeline tests and the default bench line
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-q}
timeout 600 python -m pytest -m gpu -x -q tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_train.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("host_ms_per_step",{}).get("enqueue"), {k:round(v['ms_per_step'],4) for k,v in d.get('stages',{}).items()})
PY
