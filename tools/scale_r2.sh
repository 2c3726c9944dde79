#!/bin/bash
# Round 2: the driver's scaling commands at N GPUs of one box - the default bench (config 2 incl. its `train` object: the
# bucketed, overlapped gradient all-reduce) and the config-4 training bench.  usage: tools/scale_r2.sh <tag> <N>
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2s}; n=${2:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $n > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $n --workload config4 > $OUT/${TAG}_bench4_n$n.json 2> $OUT/${TAG}_bench4_n$n.err; echo "config4 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus $n --workload config3 > $OUT/${TAG}_bench3_n$n.json 2> $OUT/${TAG}_bench3_n$n.err; echo "config3 rc=$?"
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_n$n.json", "$OUT/${TAG}_bench4_n$n.json", "$OUT/${TAG}_bench3_n$n.json"):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        t=d.get("train") or {}
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "train", t.get("ms_per_step"), t.get("all_reduce_ms"), str(t.get("cuda_graph"))[:24])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $OUT/${TAG}_bench_n$n.err
