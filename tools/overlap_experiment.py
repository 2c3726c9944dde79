"""Scratch: does running consecutive batches on alternating streams (two handles) overlap the latency-bound GRU
recurrence with the next batch's frontend / conv stack?  Also: pcm16 frontend timing and pinned H2D bandwidth."""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
native = importlib.import_module("speech-intent-recognizer_b200._native")
synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")


def main():
    B, L, K = 256, 48000, 60
    flat = torch.from_numpy(synth.flatten_weights(synth.make_weights(1234)))
    n_streams = int(os.environ.get("NSTREAMS", "2"))
    fes = [native.Frontend() for _ in range(n_streams)]
    models = [native.Model(31, 64) for _ in range(n_streams)]
    for m in models:
        m.load_weights(flat)
    g = torch.Generator(device="cuda").manual_seed(0)
    waves = [(torch.rand(B, L, device="cuda", generator=g) - 0.5) * 0.2 for _ in range(4)]
    feats = [torch.empty(B, 64, 200, device="cuda") for _ in range(n_streams)]
    logits = [torch.empty(B, 31, device="cuda") for _ in range(n_streams)]
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    main_s = torch.cuda.current_stream()

    def step(i, s):
        with torch.cuda.stream(streams[s]):
            fes[s].forward(waves[i % 4], out=feats[s], out_frames=200)
            logits[s] = models[s].forward(feats[s])

    def run(n_used):
        for i in range(6):
            step(i, i % n_used)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for s in range(n_used):
            streams[s].wait_stream(main_s)
        for i in range(K):
            step(i, i % n_used)
        for s in range(n_used):
            main_s.wait_stream(streams[s])
        e1.record(main_s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{n_used} stream(s): {ms:.3f} ms/step -> {B / ms * 1e3:.0f} utt/s", flush=True)

    for n in range(1, n_streams + 1):
        run(n)
        run(n)

    # pcm16 vs fp32 frontend
    fe = fes[0]
    w = waves[0]
    pcm = (w * 32767.0).round().to(torch.int16)
    out = feats[0]
    for name, src in (("fp32", w), ("pcm16", pcm)):
        for _ in range(3):
            fe.forward(src, out=out, out_frames=200)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fe.forward(src, out=out, out_frames=200)
        e1.record()
        torch.cuda.synchronize()
        print(f"frontend {name}: {e0.elapsed_time(e1) / 20:.4f} ms")

    # pinned H2D bandwidth
    for dt, nbytes in ((torch.float32, 4), (torch.int16, 2)):
        h = torch.empty(B, L, dtype=dt).pin_memory()
        d = torch.empty(B, L, dtype=dt, device="cuda")
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"H2D {dt}: {ms:.3f} ms per {B * L * nbytes / 1e6:.1f} MB -> {B * L * nbytes / ms / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
