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
                 seed=0):
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
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.segments = model.param_segments()
        self.adam_steps = 0
        self.skipped_steps = 0
        self._scalars = torch.zeros(2, device="cuda", dtype=torch.float32)        # [loss, found_inf]
        self._host_scalars = torch.zeros(2, dtype=torch.float32).pin_memory()
        # optional device timing of the collective (bench.py): CUDA events around the all-reduce of every step
        self.time_collective = False
        self.collective_ms = []
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def step(self, mel: torch.Tensor, label: torch.Tensor, dropout_keep: torch.Tensor = None) -> float:
        m = self.model
        m.train()
        feats = m._check_input(mel).to(device="cuda", dtype=torch.float32).contiguous()
        label = label.to(device="cuda", dtype=torch.int64)
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
        loss, found_inf = float(self._host_scalars[0]), bool(self._host_scalars[1] != 0)
        if self.time_collective:
            self.collective_ms.append(self._ev[0].elapsed_time(self._ev[1]))
        if found_inf:
            self.skipped_steps += 1
        else:
            self.adam_steps += 1
        self.scaler.update(found_inf)
        with torch.no_grad():
            for bn in (m.bn1, m.bn2, m.bn3):
                bn.num_batches_tracked += 1
        return loss
