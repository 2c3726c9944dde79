"""Waveform -> intent for batches that start in HOST memory: the end-to-end entry a serving / evaluation loop calls.

The reference moves every batch host -> device and only then starts computing (scripts/train.py:86-87,
scripts/evaluate.py:79-82, scripts/test_model.py:121-127).  On a B200 a 256 x 3 s fp32 batch is 49 MB, i.e.
~0.9 ms of PCIe time next to ~1 ms of compute, so this entry splits the batch into sub-batches and overlaps the
H2D copy of sub-batch i+1 (copy stream) with the feature frontend and the conv stack of sub-batch i (compute
stream); the GRU layers, attention pooling and fc then run once over the whole batch (the recurrence is latency
bound - splitting it would multiply that latency) and the logits are copied back to pinned host memory.

Same results as ``model(extractor.extract_batch(waves.cuda()))``: identical kernels, identical order per utterance.
"""
from __future__ import annotations

import torch

from . import _native


class IntentPipeline:
    def __init__(self, extractor, model, sub_batches: int = 4, out_frames: int = 200, max_duration=5.0):
        self.extractor, self.model = extractor, model
        self.sub_batches, self.out_frames, self.max_duration = int(sub_batches), int(out_frames), max_duration
        self._copy_stream = torch.cuda.Stream()
        self._bufs = None

    def _buffers(self, B, L, num_classes, device):
        key = (B, L)
        if self._bufs is None or self._bufs[0] != key:
            self._bufs = (key,
                          torch.empty((B, L), device=device, dtype=torch.float32),
                          torch.empty((B, self.extractor.n_mels, self.out_frames), device=device, dtype=torch.float32),
                          torch.empty((B, num_classes), device=device, dtype=torch.float32),
                          torch.empty((B, num_classes), dtype=torch.float32).pin_memory())
        return self._bufs[1:]

    @torch.no_grad()
    def infer_host(self, waves: torch.Tensor, lengths: torch.Tensor = None, out: torch.Tensor = None) -> torch.Tensor:
        """``waves [B, L]`` fp32 in (ideally pinned) host memory -> logits ``[B, num_classes]`` in pinned host memory.

        Synchronises the compute stream before returning (the caller reads the result).
        """
        if waves.is_cuda:
            raise _native.NativeError("infer_host takes host tensors; use extract_batch + model() for device tensors")
        model = self.model
        if model.training:
            raise _native.NativeError("infer_host is an inference entry: call model.eval() first")
        if model._native_model is None or model._native_dirty or model._uploaded_versions != model._versions():
            model.refresh_weights()
        B, L = waves.shape
        dev = torch.device("cuda", torch.cuda.current_device())
        d_wave, feats, logits, host_logits = self._buffers(B, L, model.num_classes, dev)
        if out is not None:
            host_logits = out
        if B > _native.Model.MAX_STAGED_BATCH:                       # larger than one workspace pass: plain path
            d_wave.copy_(waves, non_blocking=True)
            self.extractor.extract_batch(d_wave, lengths=lengths, max_duration=self.max_duration,
                                         out_frames=self.out_frames, out=feats)
            host_logits.copy_(model(feats), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host_logits
        compute = torch.cuda.current_stream()
        copy = self._copy_stream
        copy.wait_stream(compute)                                    # the previous call's readers of d_wave are done
        n_sub = max(1, min(self.sub_batches, B))
        bounds = [(i * B) // n_sub for i in range(n_sub + 1)]
        events = []
        with torch.cuda.stream(copy):
            for i in range(n_sub):
                a, b = bounds[i], bounds[i + 1]
                d_wave[a:b].copy_(waves[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                events.append(ev)
        d_len = lengths
        for i in range(n_sub):
            a, b = bounds[i], bounds[i + 1]
            if b == a:
                continue
            compute.wait_event(events[i])
            self.extractor.extract_batch(d_wave[a:b], lengths=None if d_len is None else d_len[a:b],
                                         max_duration=self.max_duration, out_frames=self.out_frames, out=feats[a:b])
            model._native_model.forward_convs(feats[a:b], B, a)
        model._native_model.forward_head(B, self.out_frames, logits)
        host_logits.copy_(logits, non_blocking=True)
        compute.synchronize()
        return host_logits
