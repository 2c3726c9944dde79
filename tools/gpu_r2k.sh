#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in "-DSIR_FE_PREFETCH=0" "-DSIR_FE_PREFETCH=1" "-DSIR_FE_PREFETCH=2"; do
  SIR_NVCC_EXTRA="$v" python speech-intent-recognizer_b200/build.py --force > /dev/null 2>&1
  echo "== $v"; timeout 300 python tools/fe_time.py 2>&1 | tail -3
done
