"""Scratch timing of the individual stages on one GPU (CUDA events, L2-exceeding inputs or rotation)."""
import importlib
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(ts)), float(np.min(ts))


def main():
    fe = native.Frontend()
    model = native.Model(31, 64)
    model.load_weights(torch.from_numpy(synth.flatten_weights(synth.make_weights(1234))))
    g = torch.Generator(device="cuda").manual_seed(0)
    for B, L in ((256, 48000), (4096, 48000), (30043, 48000), (2048, 160000)):
        w = (torch.rand(B, L, device="cuda", generator=g) - 0.5) * 0.2
        T = 1 + L // 512
        out = torch.empty(B, 64, T, device="cuda")
        med, best = timeit(lambda: fe.forward(w, out=out))
        byt = B * (4 * L + 4 * 64 * T)
        print(f"frontend B={B} L={L}: median {med:.3f} ms best {best:.3f} ms -> {B / med * 1e3:.0f} utt/s, "
              f"{byt / med / 1e6:.1f} GB/s algorithmic")
        del w, out
    for B in (16, 256, 1024):
        x = torch.randn(B, 64, 200, device="cuda")
        med, best = timeit(lambda: model.forward(x), iters=5)
        print(f"classifier B={B}: median {med:.3f} ms -> {B / med * 1e3:.0f} utt/s, {B * 400.6e6 / med / 1e9:.1f} TFLOP/s")
    w = (torch.rand(256, 48000, device="cuda", generator=g) - 0.5) * 0.2
    med, best = timeit(lambda: model.pipeline(fe, w), iters=5)
    print(f"pipeline B=256: median {med:.3f} ms -> {256 / med * 1e3:.0f} utt/s")


if __name__ == "__main__":
    main()
