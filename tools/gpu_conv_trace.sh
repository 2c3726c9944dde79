#!/bin/bash
set -u
SIR_NVCC_EXTRA="-DSIR_CONV_TRACE" python speech-intent-recognizer_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/conv_trace.py 2>&1 | tail -12
