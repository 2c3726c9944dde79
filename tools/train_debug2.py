"""Intermediate gradients of the CUDA backward (SIR_TRAIN_DUMP) against torch autograd on the CPU (debug aid)."""
import importlib, os, sys
os.environ["SIR_TRAIN_DUMP"] = "/tmp/sirdump_"
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import golden, golden_keep, synth, train_inputs
from oracle import train_port
from oracle.torch_port import ClassifierPort, load_numpy_state
models = importlib.import_module("speech-intent-recognizer_b200.models.models")
g = golden("train")
x, labels = train_inputs()
sd = synth.make_weights(int(g["weight_seed"]))
m = models.CNNAudioGRU(31)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
m = m.cuda().train()
keep = golden_keep(g)
m._next_dropout_keep = torch.from_numpy(keep).cuda()
out = m(torch.from_numpy(x).cuda())
F.cross_entropy(out, torch.from_numpy(labels).cuda()).backward()
torch.cuda.synchronize()
# CPU autograd with retained intermediates
port = load_numpy_state(ClassifierPort(31), sd).train()
xt = torch.from_numpy(x).unsqueeze(1)
acts, zs = [], []
h = xt
for i in (1, 2, 3):
    z = getattr(port, f"conv{i}")(h); z.retain_grad(); zs.append(z)
    h = F.max_pool2d(F.relu(getattr(port, f"bn{i}")(z)), 2); h.retain_grad(); acts.append(h)
b, c, hh, w = h.shape
gin = h.permute(0, 3, 1, 2).contiguous().view(b, w, c * hh); gin.retain_grad()
y0 = train_port._gru_layer(gin, port.gru, 0) * torch.from_numpy(keep).float() * 2.0
y1 = train_port._gru_layer(y0, port.gru, 1)
wts = F.softmax(port.attention(y1), dim=1)
logits = port.fc((y1 * wts).sum(1))
F.cross_entropy(logits, torch.from_numpy(labels)).backward()
def load(name, shape):
    return np.fromfile(f"/tmp/sirdump_{name}.bin", dtype=np.float32).reshape(shape)
def report(name, got, want):
    want = want.detach().numpy()
    d = np.abs(got - want)
    idx = np.unravel_index(np.argmax(d), d.shape)
    print(f"{name:6s} max err {d.max():.3e} rel {d.max() / np.abs(want).max():.3e} at {idx} got {got[idx]:.5e} want {want[idx]:.5e}; frac>1e-3*max: {(d > 1e-3 * np.abs(want).max()).mean():.4f}")
    return d
B = x.shape[0]
report("dgin", load("dgin", (B, 25, 1024)), gin.grad)
report("dz3", load("dz3", (B, 16, 50, 128)), zs[2].grad.permute(0, 2, 3, 1))
d = report("da2", load("da2", (B, 16, 50, 64)), acts[1].grad.permute(0, 2, 3, 1))
print("   da2 err by row y:", d.max(axis=(0, 2, 3)).round(5))
print("   da2 err by col x:", d.max(axis=(0, 1, 3)).round(5))
print("   da2 err by batch:", d.max(axis=(1, 2, 3)).round(5))
d = report("dz2", load("dz2", (B, 32, 100, 64)), zs[1].grad.permute(0, 2, 3, 1))
d = report("da1", load("da1", (B, 32, 100, 32)), acts[0].grad.permute(0, 2, 3, 1))
print("   da1 err by row y:", d.max(axis=(0, 2, 3)).round(5))
