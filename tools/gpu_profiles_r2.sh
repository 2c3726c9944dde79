#!/bin/bash
# Round-2 evidence: launch list of the bench command + ncu --set full of one warm step of every hot kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2q}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --e2e-repeats 1"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 120 python tools/c1prof.py > /dev/null 2>&1
$CMD > /dev/null 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"logmel_frontend_tc|frontend_finish|conv1_|conv3x3_persistent|conv3x3_stream|gemm_persistent|gru_layer_pp" --launch-skip 30 -c 10 \
    -o $OUT/${TAG}_full -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_full_raw.csv > $OUT/${TAG}_full_summary.txt 2>&1
ls -la $OUT | grep $TAG
