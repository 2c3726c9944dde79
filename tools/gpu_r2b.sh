#!/bin/bash
# Round-2 visit B: first run of the tensor-core DFT frontend: a direct check against the CUDA-core kernel, parity tests, timing.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2b}
timeout 300 python - > $OUT/${TAG}_first.log 2>&1 <<'PY'
import importlib, os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, os.getcwd())
native = importlib.import_module("speech-intent-recognizer_b200._native")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
from oracle import logmel_np
fe = native.Frontend()
w = synth.speech_like(5, 4, 48000)
for mode, name in ((native.OUT_MEL_POWER, "power"), (native.OUT_MEL_DB, "db"), (native.OUT_LOGMEL_NORM, "norm")):
    out = fe.forward(torch.from_numpy(w).cuda(), mode=mode).cpu().numpy()
    torch.cuda.synchronize()
    for i in range(2):
        p = logmel_np.mel_power(w[i])
        want = p if name == "power" else (logmel_np.amplitude_to_db(p) if name == "db" else logmel_np.extract_features(w[i]))
        err = np.abs(out[i] - want)
        print(name, i, "max abs err", err.max(), "scale", np.abs(want).max(), "rel", err.max() / np.abs(want).max(), "argmax", np.unravel_index(err.argmax(), err.shape), flush=True)
print("first check done")
PY
echo "first rc=$?"; tail -12 $OUT/${TAG}_first.log
timeout 900 python -m pytest -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/fe_time.py > $OUT/${TAG}_fe_time.log 2>&1; cat $OUT/${TAG}_fe_time.log
SIR_FRONTEND_KERNEL=cuda timeout 300 python tools/fe_time.py > $OUT/${TAG}_fe_time_cuda.log 2>&1; cat $OUT/${TAG}_fe_time_cuda.log
timeout 600 python bench.py --no-train > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['frontend_roofline']['frac'], {k:v['ms_per_step'] for k,v in d['stages'].items()})
except Exception as e: print('ERR', e)
PY
