#!/bin/bash
# ncu --set full of one warm conv1 launch (source page + raw metrics)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-c1}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --e2e-repeats 1"
$CMD > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_" --launch-skip 6 -c 1 -o $OUT/${TAG}_c1 -f $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/${TAG}_c1.ncu-rep --page raw --csv > $OUT/${TAG}_c1_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_c1.ncu-rep --page source --csv > $OUT/${TAG}_c1_source.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_c1_raw.csv > $OUT/${TAG}_c1_summary.txt 2>&1
cat $OUT/${TAG}_c1_summary.txt
