"""Numerics study for a tensor-core DFT frontend ("Tensor-core DFT, costed"): does a 32 x 32
two-stage DFT with fp16 (hi, lo) operand splits and fp32 accumulation hold the 1e-4 feature bar?  Pure numpy, CPU only.

  x[n], n = 32 n1 + n2:  Y[n2, k1] = sum_n1 (w x)[32 n1 + n2] W32^(n1 k1)              (stage 1, window folded in)
                         X[k1 + 32 k2] = sum_n2 Y[n2, k1] W1024^(n2 k1) W32^(n2 k2)    (stage 2, twiddle folded in)
Operands are rounded to fp16; a split operand a = a_hi + a_lo contributes a_hi b_hi + a_hi b_lo + a_lo b_hi.
Frames are scaled by a power of two (exact) so that max |x| is in [0.5, 1): keeps a_lo out of the fp16 subnormals.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import logmel_np  # noqa: E402  (test infrastructure: this is a study script, not the product)
import importlib  # noqa: E402

synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")


def f16(a):
    return a.astype(np.float16).astype(np.float32)


def split(a):
    hi = f16(a)
    return hi, f16(a - hi)


def mm(a, b, passes):
    """a @ b with fp16 operands, fp32 accumulation; passes: 1 (hi.hi), 2 (+ lo.hi), 3 (+ hi.lo)."""
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    out = a_hi @ b_hi
    if passes >= 2:
        out = out + a_lo @ b_hi
    if passes >= 3:
        out = out + a_hi @ b_lo
    return out.astype(np.float32)


def build_matrices():
    n1 = np.arange(32)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024)
    # stage 1: per n2 a [32 (n1) x 32] matrix: columns Re k1 = 0..16, Im k1 = 1..15
    s1 = np.zeros((32, 32, 32))
    for n2 in range(32):
        w = win[32 * n1 + n2]
        for k1 in range(17):
            s1[n2, :, k1] = w * np.cos(2 * np.pi * n1 * k1 / 32)
        for k1 in range(1, 16):
            s1[n2, :, 16 + k1] = -w * np.sin(2 * np.pi * n1 * k1 / 32)
    # stage 2: per k1 a [64 (Re n2, Im n2) x 64 (Re k2, Im k2)] matrix
    n2 = np.arange(32)[:, None]
    k2 = np.arange(32)[None, :]
    s2 = np.zeros((17, 64, 64))
    for k1 in range(17):
        ang = -2 * np.pi * (n2 * k1 / 1024 + n2 * k2 / 32)
        c, s = np.cos(ang), np.sin(ang)
        s2[k1, :32, :32] = c
        s2[k1, :32, 32:] = s
        s2[k1, 32:, :32] = -s
        s2[k1, 32:, 32:] = c
    return s1.astype(np.float32), s2.astype(np.float32)


def power_spectrum_tc(frames, s1, s2, passes, scale_frames=True):
    """frames [F, 1024] fp32 -> power [F, 513] through the two-stage fp16-split DFT."""
    F = frames.shape[0]
    x = frames.astype(np.float32)
    sc = np.ones((F, 1), np.float32)
    if scale_frames:
        mx = np.maximum(np.abs(x).max(1, keepdims=True), 1e-30)
        sc = np.exp2(-np.ceil(np.log2(mx))).astype(np.float32)          # power of two: exact
        x = x * sc
    xr = x.reshape(F, 32, 32)                                             # [F, n1, n2]
    Y = np.zeros((F, 32, 32), np.float32)                                 # [F, n2, (Re 0..16, Im 1..15)]
    for n2 in range(32):
        Y[:, n2, :] = mm(xr[:, :, n2], s1[n2], passes)
    P = np.zeros((F, 513), np.float64)
    for k1 in range(17):
        yr = Y[:, :, k1]
        yi = Y[:, :, 16 + k1] if 1 <= k1 <= 15 else np.zeros_like(yr)
        Z = mm(np.concatenate([yr, yi], 1), s2[k1], passes)              # [F, (Re k2, Im k2)]
        pw = Z[:, :32].astype(np.float64) ** 2 + Z[:, 32:].astype(np.float64) ** 2
        for k2_ in range(32):
            k = k1 + 32 * k2_
            kk = k if k <= 512 else 1024 - k
            P[:, kk] = pw[:, k2_]
    return (P / (sc.astype(np.float64) ** 2)).astype(np.float32)


def features_from_power(P, n_mels=64):
    fb = logmel_np.melscale_fbanks(n_mels=n_mels, dtype=np.float64)       # [513, n_mels]
    mel = P.astype(np.float64) @ fb
    db = 10.0 * np.log10(np.maximum(mel, 1e-10))
    db = db.T                                                            # [n_mels, T]
    return (db - db.mean()) / (db.std(ddof=1) + 1e-5)


def frames_of(wave):
    x = np.pad(wave, (512, 512), mode="reflect")
    T = 1 + len(wave) // 512
    return np.stack([x[512 * t: 512 * t + 1024] for t in range(T)])


def main():
    s1, s2 = build_matrices()
    cases = {"speech_like": synth.speech_like(5, 4, 48000), "white": synth.white_noise(6, 4, 48000),
             "quiet (x 1e-3)": synth.speech_like(7, 2, 48000) * 1e-3}
    for name, waves in cases.items():
        for passes in (1, 2, 3):
            worst = 0.0
            for w in waves:
                want = logmel_np.extract_features(w, dtype=np.float64)
                got = features_from_power(power_spectrum_tc(frames_of(w), s1, s2, passes))
                worst = max(worst, float(np.max(np.abs(got - want)) / np.max(np.abs(want))))
            print(f"{name:16s} passes {passes}: max |err| / max |feature| = {worst:.2e}   (bar 1e-4)")


if __name__ == "__main__":
    main()
