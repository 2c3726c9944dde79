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
TRAIN_SEED, TRAIN_N_SAMPLES = 20261018, 48


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
