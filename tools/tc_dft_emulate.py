"""Numpy emulation of the tensor-core DFT frontend's arithmetic (frontend_tc.cu) on whole utterances, against the fp32
oracle (oracle/logmel_np.py): which utterances / frames / bands carry the largest error, and what a variant would change.
Usage: python tools/tc_dft_emulate.py [n_utts] [variant]   variants: tf32 (the kernel as built), base (fp16 pieces + per-frame scale: the first version), lolo (4th pass), trunc, rawmax (scale from the un-windowed maximum), fwin (Hann window as a 3-tap in the frequency domain: rejected, 4.6e-5)"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import logmel_np  # noqa: E402

synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")
f32, f16 = np.float32, np.float16


def split(v):
    hi = v.astype(f16).astype(f32)
    lo = (v - hi).astype(f16).astype(f32)
    return hi, lo


def tables():
    n1 = np.arange(32)[:, None]
    o = np.arange(32)[None, :]
    j = o >> 1
    a = 2 * np.pi * ((n1 * j) % 32) / 32
    b1 = np.where(o & 1, -np.sin(a), np.cos(a))
    b1[:, 0] = 1.0
    b1[:, 1] = np.where(np.arange(32) & 1, -1.0, 1.0)
    kap = np.arange(64)[:, None]
    nu = np.arange(64)[None, :]
    n2, c, k2, cp = kap >> 1, kap & 1, nu >> 1, nu & 1
    a = 2 * np.pi * ((n2 * k2) % 32) / 32
    b2 = np.where((c == 0) & (cp == 0), np.cos(a), np.where((c == 1) & (cp == 0), np.sin(a), np.where((c == 0) & (cp == 1), -np.sin(a), np.cos(a))))
    n2 = np.arange(32)[:, None]
    k1 = np.arange(17)[None, :]
    tw = np.exp(-2j * np.pi * n2 * k1 / 1024)
    def sp(m):
        hi = m.astype(f32).astype(f16).astype(f32)
        lo = (m - hi.astype(np.float64)).astype(f32).astype(f16).astype(f32)
        return hi, lo
    return sp(b1), sp(b2), tw.astype(np.complex64)


(B1H, B1L), (B2H, B2L), TW = tables()
WIN = logmel_np.hann_window()
FB = logmel_np.melscale_fbanks()


def mm3(ah, al, bh, bl, lolo):
    d = ah @ bh + ah @ bl + al @ bh
    if lolo:
        d = d + al @ bl
    return d.astype(f32)


def features_fwin(w, item_frames=15):
    """Variant `fwin`: NO window on the samples (the stage-1 operand of a 512-sample block is built once and serves both
    frames that overlap it), one power-of-two scale per work item from the raw maximum, and the periodic Hann window as the
    3-tap  Z[k1] = Y'[k1] / 2 - (Y'[k1 - 1] + Y'[k1 + 1]) / 4  on the twiddled stage-1 output (Y'[-1] = conj Y'[1],
    Y'[17] = conj Y'[15] W32^n2), in fp32, before the stage-2 split."""
    L = len(w)
    x = np.pad(w, (512, 512), mode="reflect")
    T = 1 + L // 512
    fr = np.stack([x[512 * t:512 * t + 1024] for t in range(T)]).astype(f32)
    s = np.zeros(T, f32)
    for t0 in range(0, T, item_frames):
        t1 = min(T, t0 + item_frames)
        m = np.abs(x[512 * t0:512 * (t1 + 1)]).max().astype(f32)
        eb = int(np.clip(int(np.array([m], f32).view(np.uint32)[0] >> 23), 65, 187))
        s[t0:t1] = 2.0 ** (127 + 14 - eb)                                  # item maximum in [2^14, 2^15)
    t = (fr * s[:, None]).astype(f32)
    hi = t.astype(f16).astype(f32)
    lo = (t - hi).astype(f16).astype(f32)
    ah = hi.reshape(T, 32, 32).transpose(0, 2, 1).reshape(T * 32, 32)
    al = lo.reshape(T, 32, 32).transpose(0, 2, 1).reshape(T * 32, 32)
    y = mm3(ah, al, B1H, B1L, False).reshape(T, 32, 32)
    Y = np.zeros((T, 32, 17), np.complex64)
    Y[:, :, 0] = y[:, :, 0]
    Y[:, :, 16] = y[:, :, 1]
    Y[:, :, 1:16] = y[:, :, 2::2][:, :, :15] + 1j * y[:, :, 3::2][:, :, :15]
    c0 = f32(2.0 ** -7)
    Yp = (Y * (TW * c0)[None]).astype(np.complex64)                        # [f, n2, k1], k1 = 0..16
    ext = np.zeros((T, 32, 19), np.complex64)                              # k1 = -1..17
    ext[:, :, 1:18] = Yp
    ext[:, :, 0] = np.conj(Yp[:, :, 1])
    w32 = np.exp(-2j * np.pi * np.arange(32) / 32).astype(np.complex64)
    ext[:, :, 18] = (np.conj(Yp[:, :, 15]) * w32[None, :]).astype(np.complex64)
    Z = (ext[:, :, 1:18] - f32(0.5) * (ext[:, :, 0:17] + ext[:, :, 2:19]).astype(np.complex64)).astype(np.complex64)
    a2 = np.zeros((T, 17, 64), f32)
    a2[:, :, 0::2] = Z.real.transpose(0, 2, 1)
    a2[:, :, 1::2] = Z.imag.transpose(0, 2, 1)
    a2h, a2l = split(a2.reshape(T * 17, 64))
    X = mm3(a2h, a2l, B2H, B2L, False).reshape(T, 17, 32, 2)
    pw = (X[..., 0] ** 2 + X[..., 1] ** 2).astype(f32)
    P = np.zeros((T, 513), f32)
    for k1 in range(17):
        for k2 in range(32):
            if k1 == 0:
                k = 32 * k2 if k2 <= 16 else -1
            elif k1 == 16:
                k = 16 + 32 * k2 if k2 < 16 else -1
            else:
                k = k1 + 32 * k2 if k2 < 16 else 1024 - (k1 + 32 * k2)
            if k >= 0:
                P[:, k] = pw[:, k1, k2]
    inv2 = (1.0 / (s.astype(np.float64) * 2.0 ** -6) ** 2).astype(f32)     # Z = X_w s 2^-6
    mel = ((P @ FB) * inv2[:, None]).astype(f32).T
    db = logmel_np.amplitude_to_db(mel)
    return logmel_np.normalize(db), db


def trunc19(v):
    return (np.ascontiguousarray(v, f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def split_tf32(v):
    """(hi, lo) as the tensor core reads them with kind::tf32: the 13 low mantissa bits of an operand are ignored."""
    hi = trunc19(v)
    return hi, trunc19((v - hi).astype(f32))


def tables_tf32():
    def sp(m):
        def rn(x):
            u = np.ascontiguousarray(x, f32).view(np.uint32).astype(np.uint64)
            u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
            return u.astype(np.uint32).view(f32)
        hi = rn(m.astype(f32))
        lo = rn((m - hi.astype(np.float64)).astype(f32))
        return hi, lo
    n1 = np.arange(32)[:, None]
    o = np.arange(32)[None, :]
    a = 2 * np.pi * ((n1 * (o >> 1)) % 32) / 32
    b1 = np.where(o & 1, -np.sin(a), np.cos(a))
    b1[:, 0] = 1.0
    b1[:, 1] = np.where(np.arange(32) & 1, -1.0, 1.0)
    kap = np.arange(64)[:, None]
    nu = np.arange(64)[None, :]
    n2, c, k2, cp = kap >> 1, kap & 1, nu >> 1, nu & 1
    a = 2 * np.pi * ((n2 * k2) % 32) / 32
    b2 = np.where((c == 0) & (cp == 0), np.cos(a), np.where((c == 1) & (cp == 0), np.sin(a), np.where((c == 0) & (cp == 1), -np.sin(a), np.cos(a))))
    return sp(b1), sp(b2)


def bf16_rn(v):
    u = np.ascontiguousarray(v, f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(f32)


def features_tf32(w, lo_bf16=False):
    """The kernel as built (round 2, second version): TF32 pieces, no scaling, the window on the samples.
    lo_bf16: the stage-2 lo pass as a kind::f16 MMA on bf16 operands (lo piece and B2 rounded to bf16)."""
    (b1h, b1l), (b2h, b2l) = tables_tf32()
    L = len(w)
    x = np.pad(w, (512, 512), mode="reflect")
    T = 1 + L // 512
    fr = np.stack([x[512 * t:512 * t + 1024] for t in range(T)]).astype(f32)
    t = (fr * WIN[None, :]).astype(f32)
    hi, lo = split_tf32(t)
    ah = hi.reshape(T, 32, 32).transpose(0, 2, 1).reshape(T * 32, 32)
    al = lo.reshape(T, 32, 32).transpose(0, 2, 1).reshape(T * 32, 32)
    y = mm3(ah, al, b1h, b1l, False).reshape(T, 32, 32)
    Y = np.zeros((T, 32, 17), np.complex64)
    Y[:, :, 0] = y[:, :, 0]
    Y[:, :, 16] = y[:, :, 1]
    Y[:, :, 1:16] = y[:, :, 2::2][:, :, :15] + 1j * y[:, :, 3::2][:, :, :15]
    Yp = (Y * TW[None]).astype(np.complex64)
    a2 = np.zeros((T, 17, 64), f32)
    a2[:, :, 0::2] = Yp.real.transpose(0, 2, 1)
    a2[:, :, 1::2] = Yp.imag.transpose(0, 2, 1)
    a2 = a2.reshape(T * 17, 64)
    a2h, a2l = split_tf32(a2)
    if lo_bf16:
        X = (a2h @ b2h + a2h @ b2l + bf16_rn((a2 - a2h).astype(f32)) @ bf16_rn((b2h.astype(np.float64) + b2l).astype(f32))).astype(f32)
        X = X.reshape(T, 17, 32, 2)
    else:
        X = mm3(a2h, a2l, b2h, b2l, False).reshape(T, 17, 32, 2)
    pw = (X[..., 0] ** 2 + X[..., 1] ** 2).astype(f32)
    P = np.zeros((T, 513), f32)
    for k1 in range(17):
        for k2 in range(32):
            if k1 == 0:
                k = 32 * k2 if k2 <= 16 else -1
            elif k1 == 16:
                k = 16 + 32 * k2 if k2 < 16 else -1
            else:
                k = k1 + 32 * k2 if k2 < 16 else 1024 - (k1 + 32 * k2)
            if k >= 0:
                P[:, k] = pw[:, k1, k2]
    mel = (P @ FB).astype(f32).T
    db = logmel_np.amplitude_to_db(mel)
    return logmel_np.normalize(db), db


def features_tc(w, variant="base"):
    if variant == "tf32b":
        return features_tf32(w, lo_bf16=True)
    if variant == "fwin":
        return features_fwin(w)
    if variant == "tf32":
        return features_tf32(w)
    L = len(w)
    x = np.pad(w, (512, 512), mode="reflect")
    T = 1 + L // 512
    fr = np.stack([x[512 * t:512 * t + 1024] for t in range(T)]).astype(f32)
    tw = (fr * WIN[None, :]).astype(f32)                       # windowed frame, fp32 like the reference
    m = np.abs(fr if variant == "rawmax" else tw).max(axis=1)  # the scale comes from the WINDOWED maximum
    eb = np.clip((m.view(np.uint32) >> 23).astype(np.int64), 65, 187)
    s = (2.0 ** (127 - eb)).astype(f32)
    inv2 = (2.0 ** (2 * (eb - 127))).astype(f32)
    t = (tw * s[:, None]).astype(f32)                          # exact (power of two)
    hi = t.astype(f16).astype(f32)
    if variant == "trunc":
        hi = (t.view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)
    lo = (t - hi).astype(f16).astype(f32)
    lolo = variant == "lolo"
    # stage 1: rows (frame, n2), K = n1
    ah = hi.reshape(T, 32, 32).transpose(0, 2, 1).reshape(T * 32, 32)      # [f, n1, n2] -> [(f, n2), n1]
    al = lo.reshape(T, 32, 32).transpose(0, 2, 1).reshape(T * 32, 32)
    y = mm3(ah, al, B1H, B1L, lolo).reshape(T, 32, 32)                    # [f, n2, o]
    Y = np.zeros((T, 32, 17), np.complex64)
    Y[:, :, 0] = y[:, :, 0]
    Y[:, :, 16] = y[:, :, 1]
    Y[:, :, 1:16] = y[:, :, 2::2][:, :, :15] + 1j * y[:, :, 3::2][:, :, :15]
    Yp = (Y * TW[None]).astype(np.complex64)                              # [f, n2, k1]
    a2 = np.zeros((T, 17, 64), f32)
    a2[:, :, 0::2] = Yp.real.transpose(0, 2, 1)
    a2[:, :, 1::2] = Yp.imag.transpose(0, 2, 1)
    a2h, a2l = split(a2.reshape(T * 17, 64))
    X = mm3(a2h, a2l, B2H, B2L, lolo).reshape(T, 17, 32, 2)               # [f, k1, k2, c]
    pw = (X[..., 0] ** 2 + X[..., 1] ** 2).astype(f32)
    P = np.zeros((T, 513), f32)
    for k1 in range(17):
        for k2 in range(32):
            if k1 == 0:
                k = 32 * k2 if k2 <= 16 else -1
            elif k1 == 16:
                k = 16 + 32 * k2 if k2 < 16 else -1
            else:
                k = k1 + 32 * k2 if k2 < 16 else 1024 - (k1 + 32 * k2)
            if k >= 0:
                P[:, k] = pw[:, k1, k2]
    mel = ((P @ FB) * inv2[:, None]).astype(f32).T
    db = logmel_np.amplitude_to_db(mel)
    return logmel_np.normalize(db), db


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    variant = sys.argv[2] if len(sys.argv) > 2 else "base"
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 2026
    w = synth.speech_like(seed, n, 48000)
    worst = (0, -1)
    for i in range(n):
        want = logmel_np.extract_features(w[i])
        want_db = logmel_np.amplitude_to_db(logmel_np.mel_power(w[i]))
        got, got_db = features_tc(w[i], variant)
        err = np.abs(got - want)
        rel = err.max() / np.abs(want).max()
        if rel > worst[0]:
            worst = (rel, i)
        if rel > 3e-5:
            mb, t = np.unravel_index(err.argmax(), err.shape)
            print(f"utt {i}: rel {rel:.2e} at band {mb} frame {t}; dB err there {abs(got_db[mb, t] - want_db[mb, t]):.2e}, max dB err {np.abs(got_db - want_db).max():.2e}, "
                  f"mean shift {abs(got_db.mean() - want_db.mean()):.2e} std {want_db.std():.2f}")
    print(f"variant {variant}: worst rel_to_scale {worst[0]:.3e} (utterance {worst[1]}) over {n} utterances")


if __name__ == "__main__":
    main()
