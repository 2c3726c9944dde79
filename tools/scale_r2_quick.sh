#!/bin/bash
# cheap multi-rank check of the bench's collective logic: default workload + config3 at N ranks, short runs
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2t}; n=${2:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $n --steps 20 --train-steps 20 > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err; echo "bench rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29653 bench.py --gpus $n --workload config3 --steps 4 > $OUT/${TAG}_bench3_n$n.json 2> $OUT/${TAG}_bench3_n$n.err; echo "config3 rc=$?"
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_n$n.json", "$OUT/${TAG}_bench3_n$n.json"):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("warm_runs"))
    except Exception as e: print(f, "ERR", e)
PY
