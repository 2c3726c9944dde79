"""Timeline of conv2's MMA-issuing warp (library built with SIR_NVCC_EXTRA=-DSIR_CONV_TRACE): per tile of CTA 0, cycles spent
waiting for the accumulator, waiting for each activation stage, and issuing each tap's MMAs + commit."""
import ctypes, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
m = native.Model(31, 64)
m.load_weights(torch.from_numpy(synth.flatten_weights(synth.make_weights(1234))))
x = torch.randn(256, 64, 200, device="cuda")
for _ in range(3):
    m.forward(x)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
assert native.load_library().sir_debug_conv_trace(buf, 512) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(64, 8).copy()
ok = t[2:44]
per_tile = np.diff(t[2:45, 0])
print("tile period: median %d cycles (min %d, max %d)" % (np.median(per_tile), per_tile.min(), per_tile.max()))
print("wait go:  median %d   issue 36 MMAs + commits: median %d   to next tile: median %d" % (
    np.median(ok[:, 1] - ok[:, 0]), np.median(ok[:, 7] - ok[:, 1]), np.median(t[3:45, 0] - ok[:, 7])))
