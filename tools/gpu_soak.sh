#!/bin/bash
# robustness: the GPU suite several times in a row (the frontend's pipeline is all barriers: a rare race would show as a
# hang - every wait is bounded and traps - or as a mismatch).  (compute-sanitizer is closed on this pool.)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-soak}; N=${2:-4}
for i in $(seq 1 $N); do
  timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_pytest_$i.log 2>&1; echo "run $i rc=$? $(tail -1 $OUT/${TAG}_pytest_$i.log)"
done
