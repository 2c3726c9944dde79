"""Scratch for ncu: a few launches of the frontend alone.  usage: fe_prof.py B L [iters]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
B, L = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fe = native.Frontend()
g = torch.Generator(device="cuda").manual_seed(0)
w = (torch.rand(B, L, device="cuda", generator=g) - 0.5) * 0.2
out = torch.empty(B, 64, 200, device="cuda")
for _ in range(iters):
    fe.forward(w, out=out, out_frames=200)
torch.cuda.synchronize()
print("done")
