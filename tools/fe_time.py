"""Scratch: frontend timing only (B x L from argv)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
fe = native.Frontend()
g = torch.Generator(device="cuda").manual_seed(0)
cases = [(256, 48000, 200), (4096, 48000, 94), (2048, 160000, 313)]
for B, L, F in cases:
    w = (torch.rand(B, L, device="cuda", generator=g) - 0.5) * 0.2
    out = torch.empty(B, 64, F, device="cuda")
    for _ in range(3):
        fe.forward(w, out=out, out_frames=F)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fe.forward(w, out=out, out_frames=F); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    T = 1 + L // 512
    print(f"{os.environ.get('SIR_FE_DEBUG','0')}: B={B} L={L} F={F}: {np.median(ts):.3f} ms  {B*(4*L+4*64*T)/np.median(ts)/1e6:.0f} GB/s", flush=True)
    del w, out
