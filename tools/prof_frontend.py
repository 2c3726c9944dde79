"""Small driver for ncu captures: a few launches of the frontend (and optionally the model) at bench shapes."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "frontend"
fe = native.Frontend()
g = torch.Generator(device="cuda").manual_seed(0)
w = (torch.rand(B, 48000, device="cuda", generator=g) - 0.5) * 0.2
out = torch.empty(B, 64, 200, device="cuda")
if what == "frontend":
    for _ in range(4):
        fe.forward(w, out_frames=200, out=out)
else:
    m = native.Model(31, 64)
    m.load_weights(torch.from_numpy(synth.flatten_weights(synth.make_weights(1234))))
    for _ in range(3):
        m.pipeline(fe, w, out_frames=200, features=out)
torch.cuda.synchronize()
print("done")
