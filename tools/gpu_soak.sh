#!/bin/bash
# robustness: the GPU suite several times in a row (the frontend's pipeline is all barriers: a rare race would show as a
# hang - every wait is bounded and traps - or as a mismatch), then compute-sanitizer memcheck over a small ragged batch
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-soak}; N=${2:-4}
for i in $(seq 1 $N); do
  timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_pytest_$i.log 2>&1; echo "run $i rc=$? $(tail -1 $OUT/${TAG}_pytest_$i.log)"
done
cat > /tmp/soak_small.py <<'PY'
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
native = importlib.import_module("speech-intent-recognizer_b200._native")
fe = native.Frontend()
rng = np.random.default_rng(3)
for B, L in ((5, 30000), (1, 700), (33, 48000)):
    w = torch.from_numpy((rng.standard_normal((B, L)) * 0.1).astype(np.float32)).cuda()
    lens = torch.from_numpy(rng.integers(600, L + 1, size=B).astype(np.int32)).cuda()
    out = fe.forward(w, lengths=lens, out_frames=200)
    pcm = (w * 32767).round().clamp(-32768, 32767).to(torch.int16)
    out2 = fe.forward_pcm16(pcm, lengths=lens, out_frames=200) if hasattr(fe, "forward_pcm16") else out
    torch.cuda.synchronize()
    print(B, L, float(out.abs().max()), float(out2.abs().max()))
print("small ok")
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/soak_small.py > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/${TAG}_memcheck.log
