#!/bin/bash
# A/B of build-time switches of the frontend kernel: rebuilds on the box per variant, prints the three timings
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  SIR_NVCC_EXTRA="$v" python speech-intent-recognizer_b200/build.py --force > /dev/null 2>&1
  echo "== $v"; timeout 300 python tools/fe_time.py 2>&1 | tail -3
done
