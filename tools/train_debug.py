"""Per-parameter gradient error of the CUDA training step against the golden step (debug aid)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import golden, golden_keep, sample_positions, synth, train_inputs
models = importlib.import_module("speech-intent-recognizer_b200.models.models")
g = golden("train")
x, labels = train_inputs()
sd = synth.make_weights(int(g["weight_seed"]))
m = models.CNNAudioGRU(31)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
m = m.cuda().train()
m._next_dropout_keep = torch.from_numpy(golden_keep(g)).cuda()
out = m(torch.from_numpy(x).cuda())
loss = torch.nn.functional.cross_entropy(out, torch.from_numpy(labels).cuda())
loss.backward()
print("logit err", np.max(np.abs(out.detach().cpu().numpy() - g["logits"])), "loss", float(loss), float(g["loss"]))
pos = sample_positions(sd)
for k, p in m.named_parameters():
    flat = p.grad.cpu().numpy().reshape(-1).astype(np.float64)
    want = g[f"gsamp/{k}"]
    err = np.max(np.abs(flat[pos[k]] - want)) / max(np.max(np.abs(want)), 1e-12)
    nrm = np.sqrt((flat * flat).sum())
    print(f"{k:32s} samp_err {err:9.2e}  norm {nrm:12.5e} want {float(g['gnorm/' + k]):12.5e}  sum {flat.sum():12.5e} want {float(g['gsum/' + k]):12.5e}")
for k, b in m.named_buffers():
    if "num_batches" not in k:
        print(f"{k:32s} buf_err {np.max(np.abs(b.cpu().numpy() - g['buf/' + k])):9.2e}")
