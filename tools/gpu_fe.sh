#!/bin/bash
# frontend only: parity tests of the frontend, timings at three shapes, optional ncu capture with source (TAG, NCU=1)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-fe}; NCU=${2:-0}
timeout 240 python -m pytest -m gpu -x -q tests/test_gpu_parity.py -k "frontend or golden" > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/fe_time.py > $OUT/${TAG}_fe_time.log 2>&1; cat $OUT/${TAG}_fe_time.log
if [ "$NCU" = "1" ]; then
timeout 120 python tools/fe_prof.py 2048 48000 5 > $OUT/${TAG}_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel_frontend_tc -s 3 -c 1 -o $OUT/${TAG}_fe -f python tools/fe_prof.py 2048 48000 5 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
fi
