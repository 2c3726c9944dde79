"""B200-native mirror of the reference's ``models/models.py`` (``CNNAudioGRU``).

Same constructor, same submodule names - hence the same ``state_dict`` keys and shapes, so reference
checkpoints load with ``load_state_dict`` - and the same ``forward(x)`` contract
(/root/reference/models/models.py:6-68): ``x`` is ``[B, n_mels, T]`` or ``[B, 1, n_mels, T]``, the result is
``[B, num_classes]`` raw logits.  The torch submodules are parameter containers only: ``forward`` hands the
parameters and the input to the hand-written sm_100a kernels behind ``sir_model_forward`` (eval) or
``sir_model_train_forward`` / ``sir_model_backward`` (train; bridged into autograd so that the reference's
``loss.backward(); optimizer.step()`` at scripts/train.py:99-108 keeps working).

On a CUDA device all parameters and BatchNorm running statistics are views into ONE flat fp32 buffer in
``state_dict_spec`` order: the kernels read it directly, the training step writes one flat gradient buffer, and
the data-parallel trainer (scripts/train.py) all-reduces that buffer with a single NCCL call.
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn

from .. import _native
from ..utils.synth import state_dict_spec


class _TrainStep(torch.autograd.Function):
    """Train-mode forward / backward of the whole classifier as one autograd node."""

    @staticmethod
    def forward(ctx, model, x, *params):
        feats = x.detach().contiguous()
        keep, model._next_dropout_keep = model._next_dropout_keep, None
        logits = model._native_model.train_forward(model._flat, feats, dropout_keep=keep, seed=model.dropout_seed,
                                                   offset=model._dropout_offset, bn_momentum=model.bn1.momentum,
                                                   bn_eps=model.bn1.eps)
        model._dropout_offset += (feats.shape[0] * (feats.shape[2] // 8) * 512 + 3) // 4
        model._native_dirty = True                      # running statistics changed, eval weights are stale
        ctx.model, ctx.feats = model, feats             # the kernels read `feats` again in the backward
        # the handle keeps ONE set of saved activations: remember which forward they belong to
        model._train_generation += 1
        ctx.generation = model._train_generation
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        if ctx.generation != model._train_generation:
            raise _native.NativeError("CNNAudioGRU.backward: the activations saved by this forward were overwritten by a "
                                      "later train-mode forward (the native handle keeps one set); run backward before "
                                      "the next training forward")
        model._native_model.backward(model._flat, dlogits.contiguous(), model._flat_grad)
        # One copy of the flat buffer, then views into the COPY: sir_model_backward overwrites `_flat_grad` on every call,
        # and autograd may keep what it is handed here as p.grad - views into the live buffer would make a second
        # backward (gradient accumulation, zero_grad(set_to_none=False)) add the buffer to itself.
        snap = model._flat_grad.clone()
        grads = [snap[o:o + n].view(shape) for (o, n, shape) in model._param_slices]
        return (None, None, *grads)


class CNNAudioGRU(nn.Module):
    def __init__(self, num_classes, input_channels=1, n_mels=64):
        super().__init__()
        if input_channels != 1:
            raise _native.NativeError("the CUDA path implements input_channels=1 (every reference call site)")
        self.conv1 = nn.Conv2d(input_channels, 32, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        self.conv3 = nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(128)
        self.relu = nn.ReLU(inplace=True)
        self.pool = nn.MaxPool2d(2)
        self.dropout = nn.Dropout(0.5)              # declared but never applied in the reference forward (:20)
        self.gru_input_size = 128 * (n_mels // 8)   # 1024 for the reference's 64 mels (:23)
        self.gru = nn.GRU(input_size=self.gru_input_size, hidden_size=256, num_layers=2, batch_first=True,
                          bidirectional=True, dropout=0.5)
        self.attention = nn.Linear(512, 1)
        self.fc = nn.Linear(512, num_classes)
        self.num_classes, self.n_mels = num_classes, n_mels
        self.dropout_seed = 0                       # Philox key of the GRU dropout; set per rank for data parallelism
        self._dropout_offset = 0
        self._next_dropout_keep = None              # tests: explicit keep mask for the next training forward
        self._train_generation = 0                  # which train-mode forward the handle's saved activations belong to
        self._native_model = None
        self._uploaded_versions = None
        self._native_dirty = True
        self._flat = self._flat_grad = None
        self._param_slices = self._param_list = None

    # -- flat parameter buffer --------------------------------------------------------------------------------
    def _spec(self):
        return state_dict_spec(self.num_classes, self.n_mels)

    def _named_tensors(self):
        """(key, tensor) in flat order; tensors are the live Parameters / buffers."""
        params, buffers = dict(self.named_parameters()), dict(self.named_buffers())
        return [(key, params[key] if key in params else buffers[key]) for key, _ in self._spec()]

    def weight_count(self) -> int:
        return sum(int(torch.Size(shape).numel()) for _, shape in self._spec())

    def _is_flat(self) -> bool:
        if self._flat is None:
            return False
        base, off = self._flat.data_ptr(), 0
        for key, t in self._named_tensors():
            if t.data_ptr() != base + 4 * off or t.device != self._flat.device or t.dtype != torch.float32:
                return False
            off += t.numel()
        return True

    def flatten_parameters_(self):
        """Re-home every parameter / running statistic as a view into one flat CUDA buffer (idempotent)."""
        if self._is_flat():
            return self._flat
        named = self._named_tensors()
        for key, shape in self._spec():
            t = dict(named)[key]
            if tuple(t.shape) != tuple(shape):
                raise _native.NativeError(f"{key}: expected {shape}, got {tuple(t.shape)}")
        n = self.weight_count()
        flat = torch.empty(n, device="cuda", dtype=torch.float32)
        # one extra element behind the gradients carries the found-inf flag through the all-reduce
        flat_grad = torch.zeros(n + 1, device="cuda", dtype=torch.float32)
        off, slices, plist = 0, [], []
        with torch.no_grad():
            for key, t in named:
                k = t.numel()
                view = flat[off:off + k].view(t.shape)
                view.copy_(t.detach().to(device="cuda", dtype=torch.float32))
                t.data = view
                if isinstance(t, nn.Parameter):
                    slices.append((off, k, tuple(t.shape)))
                    plist.append(t)
                    t.grad = None
                off += k
        self._flat, self._flat_grad, self._param_slices, self._param_list = flat, flat_grad, slices, plist
        self._native_dirty = True
        return flat

    def param_segments(self):
        """Contiguous (offset, count) ranges of trainable parameters inside the flat buffer (<= 4 ranges)."""
        segs = []
        for o, k, _ in self._param_slices:
            if segs and segs[-1][0] + segs[-1][1] == o:
                segs[-1] = (segs[-1][0], segs[-1][1] + k)
            else:
                segs.append((o, k))
        return segs

    # -- weight hand-off ------------------------------------------------------------------------------------
    def _flat_weights(self) -> torch.Tensor:
        sd = self.state_dict()
        parts = []
        for key, shape in self._spec():
            t = sd[key]
            if tuple(t.shape) != tuple(shape):
                raise _native.NativeError(f"{key}: expected {shape}, got {tuple(t.shape)}")
            parts.append(t.detach().reshape(-1).to(torch.float32))
        return torch.cat(parts).contiguous()

    def _versions(self):
        # (address, in-place version) of every parameter and buffer: changes when a tensor is written in place or replaced.
        # Walks the module tree itself - state_dict() builds a detached copy of every entry and cost a third of the host
        # time of IntentPipeline.submit.
        return tuple((t.data_ptr(), t._version) for t in itertools.chain(self.parameters(), self.buffers()))

    def _ensure_native(self):
        if self._native_model is None:
            self._native_model = _native.Model(self.num_classes, self.n_mels)

    def refresh_weights(self):
        """(Re)upload the parameters to the native handle; called automatically when they change."""
        self._ensure_native()
        src = self._flat if self._is_flat() else self._flat_weights()
        self._native_model.load_weights(src, bn_eps=self.bn1.eps)
        self._uploaded_versions = self._versions()
        self._native_dirty = False

    def _check_input(self, x):
        if x.dim() == 4:
            if x.size(1) != 1:
                raise _native.NativeError("expected [B, 1, n_mels, T]")
            x = x[:, 0]
        if x.dim() != 3 or x.size(1) != self.n_mels:
            raise _native.NativeError(f"expected [B, {self.n_mels}, T] or [B, 1, {self.n_mels}, T], got {tuple(x.shape)}")
        return x

    def forward(self, x):
        x = self._check_input(x)
        dev_in = x.device
        if self.training:
            if not torch.cuda.is_available():
                raise _native.NativeError("CNNAudioGRU.forward (train mode): no CUDA device and no CPU path")
            self._ensure_native()
            self.flatten_parameters_()
            y = _TrainStep.apply(self, x.to(device="cuda", dtype=torch.float32), *self._param_list)
            with torch.no_grad():
                for bn in (self.bn1, self.bn2, self.bn3):
                    bn.num_batches_tracked += 1
            return y.to(dev_in)
        if self._native_model is None or self._native_dirty or self._uploaded_versions != self._versions():
            self.refresh_weights()
        y = self._native_model.forward(x.to(device="cuda", dtype=torch.float32))
        return y.to(dev_in)


if __name__ == "__main__":
    model = CNNAudioGRU(num_classes=31).cuda().eval()
    print("Output shape:", model(torch.randn(4, 64, 200, device="cuda")).shape)
