"""Generate ``train.npz`` FROM THE REFERENCE ITSELF: one training step of the unmodified ``CNNAudioGRU``.

Run once in the build container (needs /root/reference, which the GPU box does not have):

    python tests/golden/make_golden_train.py

It imports the reference's own ``models.models.CNNAudioGRU`` (/root/reference/models/models.py:5-68), puts it in
train mode, seeds torch's global generator, and replays the fp32 branch of the reference's step
(/root/reference/scripts/train.py:90-108: zero_grad, forward, nn.CrossEntropyLoss, backward, optim.Adam step with
weight_decay - the CPU / scaler=None branch at :103-108).  The GRU dropout mask the reference drew is recovered by
re-seeding (oracle/train_port.py:recover_gru_dropout_keep) and the script ASSERTS that the oracle restatement
with that mask reproduces the reference's logits, loss and gradients before anything is written.

Stored (small): input seed/checksums, the keep mask (bit-packed), logits, loss, BatchNorm running statistics after
the step, and for every parameter: gradient L2 norm, sum, and 48 sampled entries (same for the Adam-updated
parameters).  Sample positions come from a seeded generator stored alongside.
"""
import hashlib
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
from models.models import CNNAudioGRU  # noqa: E402  (the reference)
from oracle import train_port  # noqa: E402
from oracle.torch_port import ClassifierPort, load_numpy_state  # noqa: E402

from tests.util import (TRAIN_B as B, TRAIN_CLASSES as CLASSES, TRAIN_FRAMES as FRAMES, TRAIN_SEED as SEED,  # noqa: E402
                        pool_tie_margin, sample_positions, train_inputs as make_inputs)

torch.set_num_threads(4)
WEIGHT_SEED, LR, WD = 1234, 1e-3, 1e-4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    x, labels = make_inputs()
    sd = synth.make_weights(WEIGHT_SEED)
    xt, lt = torch.from_numpy(x), torch.from_numpy(labels)

    ref = CNNAudioGRU(num_classes=CLASSES)
    ref.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    ref.train()
    opt = torch.optim.Adam(ref.parameters(), lr=LR, weight_decay=WD)
    crit = torch.nn.CrossEntropyLoss()
    opt.zero_grad(set_to_none=True)
    torch.manual_seed(SEED)
    out = ref(xt)                                                  # the reference's own forward, its own dropout draw
    loss = crit(out, lt)
    loss.backward()
    ref_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    opt.step()

    keep = train_port.recover_gru_dropout_keep(SEED, B, FRAMES // 8)
    port = load_numpy_state(ClassifierPort(CLASSES), sd)
    margin = pool_tie_margin(port, xt)
    assert margin > 1e-5, f"seed {SEED} has a near-tied max-pool window (gap {margin:.2e}); scan for another seed"
    p_loss, p_logits, p_grads = train_port.loss_and_grads(port, xt, lt, keep)
    assert torch.allclose(p_logits, out.detach(), atol=2e-5, rtol=1e-5), (p_logits - out).abs().max()
    assert abs(p_loss - float(loss)) < 1e-5
    worst = 0.0
    for k, g in ref_grads.items():
        if k == "attention.bias":          # softmax is shift-invariant: the exact gradient is 0, both sides hold rounding noise
            assert float(g.abs().max()) < 1e-5 and float(p_grads[k].abs().max()) < 1e-5
            continue
        err = float((p_grads[k] - g).abs().max() / float(g.abs().max()))
        worst = max(worst, err)
        assert err < 2e-4, (k, err)
    for k, b in ref.named_buffers():
        if "num_batches" not in k:
            assert torch.allclose(dict(port.named_buffers())[k], b, atol=1e-6), k
    print(f"oracle port reproduces the reference step: loss {float(loss):.6f}, worst gradient error {worst:.2e} (rel. to scale)")

    pos = sample_positions(sd)
    store = {"seed": SEED, "weight_seed": WEIGHT_SEED, "lr": LR, "weight_decay": WD, "x_sha": sha(x), "labels": labels,
             "keep_bits": np.packbits(keep.numpy().reshape(-1)), "logits": out.detach().numpy(), "loss": np.float32(float(loss)),
             "pool_tie_margin": np.float64(margin)}
    new_sd = ref.state_dict()
    for k, g in ref_grads.items():
        g = g.numpy().reshape(-1).astype(np.float64)
        store[f"gnorm/{k}"] = np.float64(np.sqrt((g * g).sum()))
        store[f"gsum/{k}"] = np.float64(g.sum())
        store[f"gsamp/{k}"] = g[pos[k]].astype(np.float32)
        store[f"psamp/{k}"] = new_sd[k].numpy().reshape(-1)[pos[k]]
    for k, b in ref.named_buffers():
        if "num_batches" not in k:
            store[f"buf/{k}"] = b.numpy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **store)
    print("wrote train.npz", os.path.getsize(os.path.join(HERE, "train.npz")), "bytes")


if __name__ == "__main__":
    main()
