"""Mirror of the hot-path pieces of the reference's ``scripts/train.py``.

* ``collate_fn`` (:49-70), ``train_epoch`` (:72-118) and ``validate`` (:120-155) keep the reference signatures:
  they run unmodified against the CUDA ``CNNAudioGRU`` (whose backward is bridged into autograd), a torch
  optimizer, ``nn.CrossEntropyLoss`` and an optional ``GradScaler``.
* ``DataParallelTrainer`` is the same step as fused kernels over flat buffers - train-mode forward, cross-entropy,
  backward, ONE gradient all-reduce, fused unscale + Adam - one process per GPU, utterances sharded by batch.
  The reference is single-GPU (``CUDA_VISIBLE_DEVICES=0`` at :17); data parallelism is new here and follows
  SURVEY.md section 8(e): per-rank batch, per-GPU BatchNorm statistics (no SyncBN in the reference), one flat
  fp32 all-reduce per step, found-inf decided collectively so all ranks skip together.

Importing the reference's own module sets ``CUDA_VISIBLE_DEVICES=0`` as a side effect, which would hide GPUs 1-7
in a data-parallel job - this mirror has no import-time side effects.
"""
from __future__ import annotations

import torch

from .. import _native


def collate_fn(batch):
    """Drop ``None``/empty items, re-pad/trim the time axis to 200, stack; ``(None, None)`` if nothing is left.

    Works on CPU or CUDA tensors (pure data movement, no arithmetic).
    """
    max_length = 200
    mel_specs, labels = [], []
    for mel, label in batch:
        if mel is None or mel.shape[0] == 0 or mel.shape[1] == 0:
            continue
        if mel.size(1) > max_length:
            mel = mel[:, :max_length]
        elif mel.size(1) < max_length:
            mel = torch.nn.functional.pad(mel, (0, max_length - mel.size(1)))
        mel_specs.append(mel)
        labels.append(label)
    if not mel_specs:
        return None, None
    return torch.stack(mel_specs), torch.tensor(labels, dtype=torch.long)


def train_epoch(model, train_loader, optimizer, criterion, device, scaler=None):
    """One epoch with the reference's loop structure (scripts/train.py:72-118); returns the mean loss.

    ``model`` is the CUDA ``CNNAudioGRU``: its forward/backward are the hand-written kernels, ``criterion`` and
    ``optimizer`` are whatever the caller built (``nn.CrossEntropyLoss``, ``optim.Adam``).  The contractions are
    fp32-equivalent, so ``scaler`` (a ``GradScaler``) is honoured for its skip semantics but changes no numerics.
    """
    model.train()
    train_losses = []
    for mel, label in train_loader:
        if mel is None or label is None or mel.size(0) == 0:
            continue
        mel = mel.to(device, non_blocking=True)
        label = label.to(device, non_blocking=True)
        optimizer.zero_grad(set_to_none=True)
        output = model(mel)
        loss = criterion(output, label)
        if scaler is not None:
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()
        else:
            loss.backward()
            optimizer.step()
        train_losses.append(loss.item())
    return sum(train_losses) / max(len(train_losses), 1)


def validate(model, val_loader, criterion, device, scaler=None):
    """scripts/train.py:120-155: eval-mode loss and accuracy."""
    model.eval()
    val_losses, correct, total = [], 0, 0
    with torch.no_grad():
        for mel, label in val_loader:
            if mel is None or label is None or mel.size(0) == 0:
                continue
            mel = mel.to(device, non_blocking=True)
            label = label.to(device, non_blocking=True)
            output = model(mel)
            loss = criterion(output, label)
            predicted = output.argmax(1)
            total += label.size(0)
            correct += (predicted == label).sum().item()
            val_losses.append(loss.item())
    return sum(val_losses) / max(len(val_losses), 1), correct / max(total, 1)


# ---- data parallelism ---------------------------------------------------------------------------------------
def shard_range(n_items: int, rank: int, world: int):
    """Contiguous split of ``n_items`` over ``world`` ranks: rank r takes [r*n/W, (r+1)*n/W)."""
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def sync_flat_gradients(flat_grad: torch.Tensor, group=None):
    """Sum ``flat_grad`` (gradients + trailing found-inf flag) over the ranks of ``group``: ONE all-reduce.

    Works on CUDA tensors over NCCL (the training job) and on CPU tensors over gloo (the host-logic tests).
    Returns the world size the caller divides by (folded into Adam's ``inv_scale`` on the GPU path).
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return world


def broadcast_initial_state(flat: torch.Tensor, group=None, src_rank_in_group: int = 0):
    """Make every rank of ``group`` start from rank 0's parameters and buffers (what DistributedDataParallel's
    constructor does): ONE broadcast of the flat buffer.  CUDA tensors over NCCL, CPU tensors over gloo."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return False
    src = dist.get_global_rank(group, src_rank_in_group) if group is not None else src_rank_in_group
    dist.broadcast(flat, src=src, group=group)
    return True


class LossScaler:
    """Host half of ``torch.cuda.amp.GradScaler`` (scripts/train.py:258): scale growth / back-off bookkeeping."""

    def __init__(self, enabled=True, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.enabled = enabled
        self.scale = float(init_scale) if enabled else 1.0
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self._good_steps = 0

    def update(self, found_inf: bool):
        if not self.enabled:
            return
        if found_inf:
            self.scale *= self.backoff_factor
            self._good_steps = 0
        else:
            self._good_steps += 1
            if self._good_steps == self.growth_interval:
                self.scale *= self.growth_factor
                self._good_steps = 0


class DataParallelTrainer:
    """Fused training step of ``scripts/train.py:80-116`` for one rank of a data-parallel job.

    ``step(mel, label)``: train-mode forward -> cross-entropy (scaled) -> backward into the flat gradient buffer
    -> inf/nan flag -> ONE all-reduce (gradients + flag) -> fused unscale + Adam (skipped on every rank if any
    rank saw inf/nan) -> loss read-back (the reference's ``loss.item()`` sync at :112-116).
    """

    def __init__(self, model, lr=5e-5, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, use_amp=True, process_group=None,
                 seed=0, use_graph=False):
        import torch.distributed as dist
        self.model = model
        self.group = process_group
        self.rank = dist.get_rank(process_group) if dist.is_available() and dist.is_initialized() else 0
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.scaler = LossScaler(enabled=use_amp)
        model._ensure_native()
        flat = model.flatten_parameters_()
        # Replicas must START identical: like DistributedDataParallel's constructor, broadcast rank 0's parameters and
        # buffers (the flat buffer holds both, BatchNorm running statistics included) - ranks that built the model with an
        # unseeded init would otherwise apply the same averaged gradient to different weights and drift apart silently.
        # The Adam moments start at zero everywhere.
        if broadcast_initial_state(flat, process_group):
            model._native_dirty = True
        model.dropout_seed = (int(seed) << 8) + self.rank         # distinct dropout streams per rank
        self.n = model.weight_count()
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.bucket = model._native_model.gru_grad_offset()        # conv / BatchNorm gradients lie in front of it
        self._side = torch.cuda.Stream() if self.world > 1 else None
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.segments = model.param_segments()
        self.adam_steps = 0
        self.skipped_steps = 0
        self._scalars = torch.zeros(2, device="cuda", dtype=torch.float32)        # [loss, found_inf]
        self._host_scalars = torch.zeros(2, dtype=torch.float32).pin_memory()
        # use_graph: the step's per-step scalars live on the device (``_native.TrainState``) and the launches before / after
        # the all-reduce are captured ONCE as two CUDA graphs and replayed - at batch 16 the ~60 launches of a step cost more
        # host time than device time.  Built lazily for the first (batch, frames) shape; other shapes run eagerly.
        self.use_graph = bool(use_graph)
        self._graph = None
        # optional device timing of the collective (bench.py): CUDA events around the all-reduce of every step
        self.time_collective = False
        self.collective_ms = []
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    # -- CUDA-graph path ------------------------------------------------------------------------------------------------
    def _graph_body_pre(self, g):
        """begin -> forward -> loss -> backward.  With more than one rank the backward stops after its first part (head + GRU
        layers: every gradient behind ``self.bucket`` = 96 % of the bytes) and checks that bucket for inf/nan, so that its
        all-reduce can run while ``_graph_body_mid`` computes the conv gradients in front of it."""
        m = self.model
        g["state"].begin(self.betas, self.world)
        m._flat_grad[self.n:].zero_()
        m._native_model.train_forward(m._flat, g["feats"], seed=m.dropout_seed, bn_momentum=m.bn1.momentum, bn_eps=m.bn1.eps,
                                      logits=g["logits"])
        _native.cross_entropy_state(g["logits"], g["labels"], g["state"], self._scalars[0:1], g["dlogits"])
        if self.world > 1:
            m._native_model.backward_part(m._flat, g["dlogits"], m._flat_grad, 1)
            _native.grad_nonfinite(m._flat_grad[self.bucket:], self.n - self.bucket, m._flat_grad[self.n:])
        else:
            m._native_model.backward(m._flat, g["dlogits"], m._flat_grad)
            _native.grad_nonfinite(m._flat_grad, self.n, m._flat_grad[self.n:])

    def _graph_body_mid(self, g):
        m = self.model
        m._native_model.backward_part(m._flat, None, m._flat_grad, 2)

    def _graph_body_post(self, g):
        m = self.model
        if self.world > 1:
            # the conv bucket is checked AFTER its all-reduce: the reduced values are the same bits on every rank, an inf / nan
            # of any rank survives the sum, and the flag needs no collective of its own
            _native.grad_nonfinite(m._flat_grad, self.bucket, m._flat_grad[self.n:])
        _native.adam_step_state(m._flat, m._flat_grad, self.exp_avg, self.exp_avg_sq, self.segments, self.lr, self.betas, self.eps,
                                self.weight_decay, g["state"], m._flat_grad[self.n:])
        g["state"].end(m._flat_grad[self.n:], g["offset_inc"])
        self._scalars[1:2].copy_(m._flat_grad[self.n:])
        self._host_scalars.copy_(self._scalars, non_blocking=True)

    def _build_graph(self, B, n_mels, T):
        import torch.distributed as dist
        m = self.model
        self.world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        g = {"key": (B, T), "offset_inc": (B * (T // 8) * 512 + 3) // 4,
             "state": _native.TrainState(self.scaler.scale, self.adam_steps, m._dropout_offset),
             "feats": torch.zeros((B, n_mels, T), device="cuda"), "labels": torch.zeros(B, device="cuda", dtype=torch.int64),
             "logits": torch.zeros((B, m.num_classes), device="cuda"), "dlogits": torch.zeros((B, m.num_classes), device="cuda")}
        m._native_model.set_train_state(g["state"])
        # a warm-up pass outside capture sizes every workspace and opts the kernels in; it must not change the training state:
        # parameters, moments, running statistics and the step state are restored afterwards
        saved = (m._flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), g["state"].buf.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._graph_body_pre(g)
            if self.world > 1:
                self._graph_body_mid(g)
            self._graph_body_post(g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        m._flat.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2]); g["state"].buf.copy_(saved[3])
        before = _native.launch_count()
        g["pre"], g["post"] = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g["pre"]):
            self._graph_body_pre(g)
        if self.world > 1:
            g["mid"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g["mid"], pool=g["pre"].pool()):
                self._graph_body_mid(g)
        with torch.cuda.graph(g["post"], pool=g["pre"].pool()):
            self._graph_body_post(g)
        g["launches"] = _native.launch_count() - before          # kernels of libsir_b200 inside one replay of both graphs
        self._graph = g

    def _step_graph(self, feats, label) -> float:
        m, g = self.model, self._graph
        g["feats"].copy_(feats, non_blocking=True)
        g["labels"].copy_(label, non_blocking=True)
        g["pre"].replay()
        if self.world > 1:
            # two gradient buckets (what DistributedDataParallel does behind scripts/train.py:107): the GRU / head bucket +
            # found-inf flag is reduced on a side stream while the conv backward runs, the small conv bucket follows it
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                sync_flat_gradients(m._flat_grad[self.bucket:self.n + 1], self.group)
            g["mid"].replay()
            if self.time_collective:
                self._ev[0].record()                                # from here on the step only waits for the collectives
            sync_flat_gradients(m._flat_grad[:self.bucket], self.group)
            main.wait_stream(self._side)
            if self.time_collective:
                self._ev[1].record()
        elif self.time_collective:
            self._ev[0].record()
            self._ev[1].record()
        g["post"].replay()
        torch.cuda.current_stream().synchronize()                 # the reference reads loss.item() every step
        m._dropout_offset += g["offset_inc"]
        m._native_dirty = True
        m._train_generation += 1
        self.graph_replays = getattr(self, "graph_replays", 0) + 1
        return self._finish_step()

    def _finish_step(self) -> float:
        loss, found_inf = float(self._host_scalars[0]), bool(self._host_scalars[1] != 0)
        if self.time_collective:
            self.collective_ms.append(self._ev[0].elapsed_time(self._ev[1]))
        if found_inf:
            self.skipped_steps += 1
        else:
            self.adam_steps += 1
        old_scale = self.scaler.scale
        self.scaler.update(found_inf)
        if self._graph is not None and self.scaler.scale != old_scale:
            self._graph["state"].set_scale(self.scaler.scale)     # GradScaler growth / back-off between replays
        with torch.no_grad():
            for bn in (self.model.bn1, self.model.bn2, self.model.bn3):
                bn.num_batches_tracked += 1
        return loss

    def step(self, mel: torch.Tensor, label: torch.Tensor, dropout_keep: torch.Tensor = None) -> float:
        m = self.model
        m.train()
        feats = m._check_input(mel).to(device="cuda", dtype=torch.float32).contiguous()
        label = label.to(device="cuda", dtype=torch.int64)
        if self.use_graph and dropout_keep is None:
            if self._graph is None:
                self._build_graph(feats.shape[0], feats.shape[1], feats.shape[2])
            if self._graph["key"] == (feats.shape[0], feats.shape[2]):
                return self._step_graph(feats, label)
        if self._graph is not None:
            # an eager step between replays: move the device state's counters over to the host arguments and back afterwards
            step_d, _, off_d = self._graph["state"].read()
            m._dropout_offset = off_d
            m._native_model.set_train_state(None)
        flat, grads = m.flatten_parameters_(), m._flat_grad
        B, _, T = feats.shape
        logits = m._native_model.train_forward(flat, feats, dropout_keep=dropout_keep, seed=m.dropout_seed,
                                               offset=m._dropout_offset, bn_momentum=m.bn1.momentum, bn_eps=m.bn1.eps)
        m._dropout_offset += (B * (T // 8) * 512 + 3) // 4
        m._native_dirty = True
        m._train_generation += 1                                   # an older autograd graph's activations are gone
        _, dlogits = _native.cross_entropy(logits, label, scale=self.scaler.scale, loss_out=self._scalars[0:1])
        m._native_model.backward(flat, dlogits, grads)
        grads[self.n:].zero_()
        _native.grad_nonfinite(grads, self.n, grads[self.n:])
        if self.time_collective:
            self._ev[0].record()
        world = sync_flat_gradients(grads, self.group)
        if self.time_collective:
            self._ev[1].record()
        _native.adam_step(flat, grads, self.exp_avg, self.exp_avg_sq, self.segments, self.lr, self.betas, self.eps,
                          self.weight_decay, step=self.adam_steps + 1, inv_scale=1.0 / (self.scaler.scale * world),
                          found_inf=grads[self.n:])
        self._scalars[1:2].copy_(grads[self.n:])
        self._host_scalars.copy_(self._scalars, non_blocking=True)
        torch.cuda.current_stream().synchronize()                 # the reference reads loss.item() every step
        loss = self._finish_step()
        if self._graph is not None:                               # hand the counters back to the device state
            g = self._graph
            g["state"] = _native.TrainState(self.scaler.scale, self.adam_steps, m._dropout_offset)
            self._graph = None                                    # (the graphs captured the old state buffer: rebuild lazily)
        return loss
