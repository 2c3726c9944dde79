#!/bin/bash
# 1 -> 8 GPU scaling of the bench (inference, weak scaling) and of the fused training step (config 4: batch 16 per GPU).
# usage: tools/scale_round.sh <tag> [list of N]     (outputs under gpurun_out/<tag>_*)
OUT=gpurun_out
TAG=${1:-scale}
NS=${2:-"1 2 4 8"}
mkdir -p $OUT
for n in $NS; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2>$OUT/${TAG}_err.log
    python tools/train_time.py 16 100 > $OUT/${TAG}_train_n1.log 2>>$OUT/${TAG}_err.log
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_n$n.json 2>>$OUT/${TAG}_err.log
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 tools/train_time.py 16 100 > $OUT/${TAG}_train_n$n.log 2>>$OUT/${TAG}_err.log
  fi
  python - <<PY
import json
lines=[l for l in open("$OUT/${TAG}_bench_n$n.json").read().splitlines() if l.startswith("{")]
print("stdout lines:", len(open("$OUT/${TAG}_bench_n$n.json").read().splitlines()))
d=json.loads(lines[-1]); print("N=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "bound", d["e2e"]["h2d_copy_bound"]["utt_s"], "pcm16", round(d["e2e_pcm16"]["value"]), "ms/step", round(d["ms_per_step"],4))
PY
  grep "train step" $OUT/${TAG}_train_n$n.log
done
python -m pytest tests/test_gpu_train.py -m gpu -q -k two_gpus 2>&1 | tail -1
