"""Shared helpers for the tests: package import, golden loading, the parity metrics."""
import hashlib
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

pkg = importlib.import_module("speech-intent-recognizer_b200")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")

# Parity bars from BASELINE.json north_star / BASELINE.md section 5.
FEATURE_REL_TOL = 1e-4     # max|a-b| / max|b|  (relative to the tensor's scale), fp32
LOGIT_ABS_TOL = 1e-3       # max|a-b| on logits


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel_to_scale(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def golden_waves():
    g = golden("frontend")
    lengths = [int(v) for v in g["lengths"]]
    waves = synth.speech_like(int(g["speech_seed"]), len(lengths), max(lengths), lengths)
    noise = synth.white_noise(int(g["noise_seed"]), 1, 48000)
    assert sha(waves) == str(g["speech_sha"]), "synthetic speech generator drifted from the golden inputs"
    assert sha(noise) == str(g["noise_sha"]), "synthetic noise generator drifted from the golden inputs"
    return g, waves, lengths, noise


# ---- training-step fixtures (shared with tests/golden/make_golden_train.py) -----------------------------------
TRAIN_B, TRAIN_FRAMES, TRAIN_VALID, TRAIN_CLASSES = 6, 200, 94, 31
# seeds picked by scanning for fixtures without near-tied max-pool windows (pool_tie_margin below)
TRAIN_SEED, TRAIN_N_SAMPLES = 20261077, 48
TRAIN_SEED_B16 = 179


def train_inputs(seed=TRAIN_SEED, batch=TRAIN_B):
    """Normalised log-mel-like features: unit-variance noise with slow structure on the valid frames, zero tail."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, 64, TRAIN_FRAMES)).astype(np.float32)
    x += (0.8 * np.sin(np.linspace(0, 6, 64, dtype=np.float32))[None, :, None]
          * rng.uniform(0.5, 1.5, (batch, 1, 1)).astype(np.float32))
    x[:, :, TRAIN_VALID:] = 0.0
    labels = rng.integers(0, TRAIN_CLASSES, size=batch).astype(np.int64)
    return x, labels


def sample_positions(sd, seed=TRAIN_SEED):
    rng = np.random.default_rng(seed + 1)
    return {k: np.sort(rng.choice(v.size, size=min(TRAIN_N_SAMPLES, v.size), replace=False)) for k, v in sd.items()}


def golden_keep(g, batch=TRAIN_B, frames=TRAIN_FRAMES):
    n = batch * (frames // 8) * 512
    return np.unpackbits(g["keep_bits"])[:n].reshape(batch, frames // 8, 512).astype(np.uint8)


def pool_tie_margin(port, x):
    """Smallest gap between the two largest candidates of any 2x2 max-pool window that passes its ReLU, over the
    three conv stages of ``port`` (a ClassifierPort) in train mode on ``x [B,64,T]`` (torch tensor).

    Max-pool routes the whole gradient of a window to its arg-max.  Two implementations whose conv outputs differ
    by fp32 rounding (~1e-6) pick different arg-maxes where the top two candidates tie to that precision - both
    are valid sub-gradients, but the gradients differ visibly.  Gradient-parity fixtures are therefore chosen (by
    seed) so that no window ties, and the tests assert this margin to say so.
    """
    import copy
    import torch
    import torch.nn.functional as F
    m = copy.deepcopy(port).train()
    h, margin = x.unsqueeze(1), float("inf")
    with torch.no_grad():
        for i in (1, 2, 3):
            u = getattr(m, f"bn{i}")(getattr(m, f"conv{i}")(h))
            B, C, H, W = u.shape
            win = u[:, :, :H // 2 * 2, :W // 2 * 2].reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5)
            top2 = win.reshape(B, C, H // 2, W // 2, 4).topk(2, dim=-1).values
            on = top2[..., 0] > 0
            gap = (top2[..., 0] - torch.clamp(top2[..., 1], min=0.0))[on]       # a runner-up below 0 ties with ReLU's 0
            gap = gap[gap > 0]          # exact ties (the all-zero padding frames) resolve identically everywhere: first wins
            margin = min(margin, float(gap.min()))
            h = F.max_pool2d(F.relu(u), 2)
    return margin
