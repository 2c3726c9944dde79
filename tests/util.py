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
