#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --workload config4 --steps 6 --warmup 3 --no-graph --no-cpu-baseline > $OUT/tn_plain.json 2> $OUT/tn_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/tn_launches.csv python bench.py --workload config4 --steps 6 --warmup 3 --no-graph --no-cpu-baseline > $OUT/tn_ncu.log 2>&1
echo rc=$?; wc -l $OUT/tn_launches.csv
