#!/bin/bash
# One GPU-box visit: parity tests, the bench (own arm + reference arm), launch list and ncu captures.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
  # full capture of the tensor-core contractions of one warm step (conv2, conv3, gemm l0, gemm l1) and the frontend
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_persistent|conv3x3_stream|gemm_persistent" --launch-skip 16 -c 4 \
      -o $OUT/${TAG}_tc -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_tc.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"logmel_frontend|gru_layer|conv1_bn" --launch-skip 16 -c 4 \
      -o $OUT/${TAG}_fe -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_fe.log 2>&1
  for r in tc fe; do
    ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_raw.csv 2>/dev/null
    python tools/ncu_summary.py $OUT/${TAG}_${r}_raw.csv > $OUT/${TAG}_${r}_summary.txt 2>&1
  done
fi
ls -la $OUT | tail -20
