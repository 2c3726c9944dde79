import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
m = native.Model(31, 64); m.load_weights(torch.from_numpy(synth.flatten_weights(synth.make_weights(1234))))
x = torch.randn(256, 64, 200, device="cuda")
for _ in range(4): y = m.forward(x)
torch.cuda.synchronize(); print("ok")
