"""Waveform -> intent for batches that start in HOST memory: the end-to-end entry a serving / evaluation loop calls.

The reference moves every batch host -> device and only then starts computing (scripts/train.py:86-87,
scripts/evaluate.py:79-82, scripts/test_model.py:121-127).  On a B200 a 256 x 3 s fp32 batch is 49 MB, i.e.
~0.9 ms of PCIe time next to ~1 ms of compute, so this entry hides the copy twice over:

* inside one batch, the batch is split into sub-batches and the H2D copy of sub-batch i+1 (copy stream) overlaps
  the feature frontend and the conv stack of sub-batch i (compute stream); the GRU layers, attention pooling and
  fc then run once over the whole batch (the recurrence is latency bound - splitting it would multiply that
  latency) and the logits are copied back to pinned host memory;
* across batches, ``submit`` / ``collect`` keep ``depth`` batches in flight on rotating device buffers, each slot
  on its OWN compute stream (the model handle keeps one workspace per stream), so the copy of batch k+1 overlaps
  the GRU / head of batch k and the latency-bound GRU recurrence of batch k (96 of 148 SMs, mostly waiting)
  overlaps the frontend and conv stack of batch k+1 (``infer_stream`` wraps that for an iterable of batches).

Measured on a B200 (256 x 3 s per batch): synchronous single batches 123 k utt/s with one copy, 150 k with 4
sub-batches; streaming with depth 2 and whole-batch copies 230 k utt/s (the device-resident path does 241 k) -
sub-batching costs small-batch efficiency, so it only pays when nothing else is in flight (default: 1).

Same results as ``model(extractor.extract_batch(waves.cuda()))``: identical kernels, identical order per utterance.
"""
from __future__ import annotations

from collections import deque

import torch

from . import _native


def bind_host_to_gpu(device_index: int = None):
    """Pin the calling process to the CPU cores next to GPU ``device_index`` (NVML's ideal CPU affinity, i.e. the
    socket / NUMA node the GPU's PCIe root hangs off).  Pinned host buffers allocated AFTERWARDS are first-touched
    on that node, so the H2D copies of 8 ranks do not all cross the socket interconnect.  Returns the CPU list, or
    None when NVML or the affinity call is unavailable (the caller carries on unbound)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if idx < len(visible) and visible[idx].strip().isdigit():
            idx = int(visible[idx])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - binding is an optimisation, never a requirement
        return None


class _Slot:
    def __init__(self):
        self.key = None
        self.waves = {}                          # (B, L, dtype) -> device staging buffer, kept across dtype changes
        self.d_wave = self.feats = self.logits = self.host_logits = None
        self.done = torch.cuda.Event()
        self.copied = []                         # one event per sub-batch copy, reused from step to step
        self.busy = False
        self.stream = torch.cuda.Stream()        # every slot computes on its own stream (own workspace in the handle)


class IntentPipeline:
    def __init__(self, extractor, model, sub_batches: int = 1, out_frames: int = 200, max_duration=5.0, depth: int = 2):
        self.extractor, self.model = extractor, model
        self.sub_batches, self.out_frames, self.max_duration = int(sub_batches), int(out_frames), max_duration
        self._copy_stream = torch.cuda.Stream()
        self._slots = [_Slot() for _ in range(max(1, int(depth)))]
        self._next = 0

    def _prepare(self, slot, B, L, num_classes, device, dtype):
        """Slot buffers for a ``[B, L]`` batch of ``dtype``.  The waveform staging buffer is kept per (B, L, dtype), so a
        caller that alternates fp32 and PCM16 batches (or ``reserve``s both up front) never re-allocates in the loop."""
        if (B, L, dtype) not in slot.waves:
            if len(slot.waves) >= 4:
                slot.waves.clear()
            slot.waves[(B, L, dtype)] = torch.empty((B, L), device=device, dtype=dtype)
        slot.d_wave = slot.waves[(B, L, dtype)]
        if slot.key != (B, num_classes):
            slot.key = (B, num_classes)
            slot.feats = torch.empty((B, self.extractor.n_mels, self.out_frames), device=device, dtype=torch.float32)
            slot.logits = torch.empty((B, num_classes), device=device, dtype=torch.float32)
            slot.host_logits = torch.empty((B, num_classes), dtype=torch.float32).pin_memory()

    def reserve(self, B, L, dtypes=(torch.float32, torch.int16)):
        """Allocate every slot's buffers for ``[B, L]`` batches of each of ``dtypes`` before the first ``submit``."""
        dev = torch.device("cuda", torch.cuda.current_device())
        for slot in self._slots:
            for dt in dtypes:
                self._prepare(slot, B, L, self.model.num_classes, dev, dt)

    @torch.no_grad()
    def submit(self, waves: torch.Tensor, lengths: torch.Tensor = None):
        """Enqueue one batch (``waves [B, L]``, fp32 or int16 PCM, in - ideally pinned - host memory); returns a ticket
        for ``collect``.  PCM16 input is scaled by 1/32768 on the device (what torchaudio.load does on the host) and
        halves the PCIe bytes per utterance.

        ``lengths`` (optional) is a CUDA int32 tensor ``[B]`` of valid samples per row (the kernels read it on the
        device; ``waves`` itself stays on the host).  ``waves`` is copied asynchronously on a side stream: the caller
        must not overwrite or free it before ``collect`` has returned for this ticket.

        Nothing here waits for the GPU unless all ``depth`` slots are still in flight.
        """
        if waves.is_cuda:
            raise _native.NativeError("IntentPipeline takes host tensors; use extract_batch + model() for device tensors")
        model = self.model
        if model.training:
            raise _native.NativeError("IntentPipeline is an inference entry: call model.eval() first")
        if model._native_model is None or model._native_dirty or model._uploaded_versions != model._versions():
            model.refresh_weights()
        slot = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        if slot.busy:
            self.collect(slot)                                       # all slots in flight: drain the oldest
        B, L = waves.shape
        dev = torch.device("cuda", torch.cuda.current_device())
        if waves.dtype not in (torch.float32, torch.int16):
            raise _native.NativeError(f"waves must be float32 or int16 PCM, got {waves.dtype}")
        self._prepare(slot, B, L, model.num_classes, dev, waves.dtype)
        compute, copy = slot.stream, self._copy_stream
        compute.wait_stream(torch.cuda.current_stream())             # weight uploads / caller work enqueued so far
        copy.wait_event(slot.done)                                   # the slot's previous readers of d_wave are done
        staged = B <= _native.Model.MAX_STAGED_BATCH
        n_sub = max(1, min(self.sub_batches, B)) if staged else 1
        bounds = [(i * B) // n_sub for i in range(n_sub + 1)]
        while len(slot.copied) < n_sub:
            slot.copied.append(torch.cuda.Event())
        events = slot.copied
        with torch.cuda.stream(copy):
            for i in range(n_sub):
                a, b = bounds[i], bounds[i + 1]
                if n_sub == 1:
                    slot.d_wave.copy_(waves, non_blocking=True)
                else:
                    slot.d_wave[a:b].copy_(waves[a:b], non_blocking=True)
                events[i].record(copy)
        with torch.cuda.stream(compute):
            for i in range(n_sub):
                a, b = bounds[i], bounds[i + 1]
                if b == a:
                    continue
                compute.wait_event(events[i])
                self.extractor.extract_batch(slot.d_wave[a:b], lengths=None if lengths is None else lengths[a:b],
                                             max_duration=self.max_duration, out_frames=self.out_frames,
                                             out=slot.feats[a:b])
                if staged:
                    model._native_model.forward_convs(slot.feats[a:b], B, a)
            if staged:
                model._native_model.forward_head(B, self.out_frames, slot.logits)
            else:                                                    # larger than one workspace pass: plain forward
                slot.logits.copy_(model(slot.feats))
            slot.host_logits.copy_(slot.logits, non_blocking=True)
            slot.done.record(compute)
        slot.busy = True
        return slot

    def collect(self, ticket) -> torch.Tensor:
        """Wait for a submitted batch; returns its logits ``[B, num_classes]`` in pinned host memory (valid until the
        slot is reused, ``depth`` submits later)."""
        ticket.done.synchronize()
        ticket.busy = False
        return ticket.host_logits

    def infer_host(self, waves: torch.Tensor, lengths: torch.Tensor = None) -> torch.Tensor:
        """One batch, synchronously: ``collect(submit(waves))``."""
        return self.collect(self.submit(waves, lengths))

    def infer_stream(self, batches):
        """Yield the logits of every batch of ``batches`` (an iterable of host tensors or ``(waves, lengths)``
        pairs) in order, keeping up to ``depth`` batches in flight so copies and compute of consecutive batches
        overlap.  A yielded tensor is valid until the generator is advanced again (its slot is then reused)."""
        pending = deque()
        for item in batches:
            waves, lengths = item if isinstance(item, tuple) else (item, None)
            if len(pending) == len(self._slots):
                yield self.collect(pending.popleft())
            pending.append(self.submit(waves, lengths))
        while pending:
            yield self.collect(pending.popleft())
