#!/bin/bash
# 2-GPU round: training tests (incl. the NCCL data-parallel check) and the config-4 bench at N = 1 and N = 2
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-dp2}
timeout 900 python -m pytest -m gpu -x -q tests/test_gpu_train.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --workload config4 --no-cpu-baseline > $OUT/${TAG}_bench4_n1.json 2> $OUT/${TAG}_bench4_n1.err; echo "bench n1 rc=$?"; tail -3 $OUT/${TAG}_bench4_n1.err
N=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload config4 > $OUT/${TAG}_bench4_n$n.json 2> $OUT/${TAG}_bench4_n$n.err; echo "bench n$n rc=$?"; tail -3 $OUT/${TAG}_bench4_n$n.err
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench4_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        t=d['train']
        print(f, 'value',round(d['value']), 'ms',round(d['ms_per_step'],4), 'allreduce_ms',t['all_reduce_ms'], t['all_reduce_share'], t['cuda_graph'][:30])
    except Exception as e: print(f,'ERR',e)
PY
