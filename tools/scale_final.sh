OUT=gpurun_out
for n in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --e2e-repeats 3 > $OUT/r1h_bench_n$n.json 2>$OUT/r1h_err_n$n.log
  python - <<PY
import json
txt=open("$OUT/r1h_bench_n$n.json").read().splitlines()
d=json.loads(txt[-1]); print("lines", len(txt), "N=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "bound", d["e2e"]["h2d_copy_bound"]["utt_s"], "pcm16", round(d["e2e_pcm16"]["value"]), "ms/step", round(d["ms_per_step"],4))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 tools/train_time.py 16 100 2>/dev/null | grep "train step"
