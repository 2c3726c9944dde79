"""cProfile of the host side of IntentPipeline.submit/collect (the e2e loop of bench.py): where the enqueue time goes."""
import cProfile
import importlib
import io
import pstats
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pre = importlib.import_module("speech-intent-recognizer_b200.scripts.precompute_features")
models = importlib.import_module("speech-intent-recognizer_b200.models.models")
pipe_mod = importlib.import_module("speech-intent-recognizer_b200.pipeline")

B, L = 256, 48000
extractor = pre.AudioFeatureExtractor(sample_rate=16000, n_mels=64)
model = models.CNNAudioGRU(31).cuda().eval()
pipe = pipe_mod.IntentPipeline(extractor, model, sub_batches=1, out_frames=200, max_duration=5.0, depth=4)
host = (torch.randn(B, L) * 3000).to(torch.int16).pin_memory()
pipe.reserve(B, L)


def loop(n):
    pending = []
    for _ in range(n):
        pending.append(pipe.submit(host))
        if len(pending) == 4:
            pipe.collect(pending.pop(0))
    while pending:
        pipe.collect(pending.pop(0))


loop(100)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
loop(300)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
