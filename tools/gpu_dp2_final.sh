#!/bin/bash
# 2 GPUs, last build: the training tests (with the NCCL data-parallel parity check) and the default bench line at N = 2
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-dp2f}
timeout 300 python -m pytest -m gpu -x -q tests/test_gpu_train.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --train-steps 20 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; tail -2 $OUT/${TAG}_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n2.json').read().strip().splitlines()[-1])
t=d['train']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'train ms', round(t['ms_per_step'],4), 'allreduce wait', t.get('all_reduce_ms'))
PY
