"""Timeline of the tensor-core frontend's pipeline (library built with SIR_NVCC_EXTRA=-DSIR_FE_TRACE): per item of CTA 0,
the SM-clock times of every role's hand-offs relative to the item's start in the A warps."""
import ctypes, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 48000
fe = native.Frontend()
g = torch.Generator(device="cuda").manual_seed(0)
w = (torch.rand(B, L, device="cuda", generator=g) - 0.5) * 0.2
out = torch.empty(B, 64, 94, device="cuda")
for _ in range(3):
    fe.forward(w, out=out, out_frames=94)
torch.cuda.synchronize()
lib = native.load_library()
C, I, E = 4, 96, 32
buf = (ctypes.c_longlong * (C * I * E))()
rc = lib.sir_debug_fe_trace(buf, C * I * E)
assert rc == 0, rc
t = np.frombuffer(buf, dtype=np.int64).reshape(C, I, E).copy()
names = {0: "A item", 1: "A peeked", 2: "A f0 go", 3: "A f0 done", 4: "A f1 go", 5: "A f1 done", 6: "A f2 go", 7: "A f2 done", 8: "A f3 go",
         9: "A f3 done", 10: "MMA1 s0", 11: "MMA1 s1", 12: "MMA1 s2", 13: "MMA1 s3", 14: "MMA2 m0", 15: "MMA2 m1", 16: "C d1[0]", 17: "C d1[1]",
         18: "C d1[2]", 19: "C d1[3]", 20: "C a2empty", 21: "C done", 22: "DE d2[0]", 23: "DE d2[1]", 24: "DE pw done", 25: "DE mel go",
         26: "DE mel done", 27: "F tile", 28: "F done", 29: "A0 published", 30: "A0 top", 31: "A0 slot free"}
c = 0
period = np.diff(t[c, 20:80, 0])
print("item period (A item start to next): median %.0f cycles, min %.0f max %.0f" % (np.median(period), period.min(), period.max()))
for it in (40, 41, 42):
    base = t[c, it, 0]
    ev = sorted((t[c, it, e] - base, names[e]) for e in names if t[c, it, e] > 0)
    print("item %d (next item starts at +%d):" % (it, t[c, it + 1, 0] - base))
    print("   " + "  ".join("%s %+d" % (n, d) for d, n in ev))
