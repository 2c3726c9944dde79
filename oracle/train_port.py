"""ORACLE (test infrastructure only) - the reference's training step restated in plain fp32 torch on the CPU.

Only ``tests/`` (and the golden generator under ``tests/golden/``) may import this module; the product path
never does.

What it follows
  * ``CNNAudioGRU.forward`` under ``model.train()``        /root/reference/models/models.py:41-68
    (BatchNorm2d with batch statistics + running-stat update, nn.GRU inter-layer dropout p = 0.5 at :27)
  * the step of ``train_epoch``                            /root/reference/scripts/train.py:90-108
    (zero_grad, forward, ``nn.CrossEntropyLoss``, backward, ``optim.Adam(lr, weight_decay)`` at :246-250)

The only liberty: nn.GRU draws its dropout mask from torch's global RNG inside the fused op, so the 2-layer GRU
is restated as an explicit loop (same gate equations as torch.nn.GRU) with the keep mask passed IN.  Pinning:
``tests/golden/make_golden_train.py`` runs the unmodified reference class with a known seed, recovers the mask it
drew, and checks that this port reproduces its logits and gradients (``tests/golden/train.npz``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle.torch_port import ClassifierPort

DROPOUT_P = 0.5   # models/models.py:27


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """torch.nn.GRU equations, one direction: r,z,n gate order; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gh = h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi[:, t].chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1.0 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)


def _gru_layer(x, gru, layer):
    outs = []
    for suffix, rev in (("", False), ("_reverse", True)):
        p = [getattr(gru, f"{n}_l{layer}{suffix}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        outs.append(_gru_direction(x, *p, reverse=rev))
    return torch.cat(outs, dim=2)


def train_forward(model: ClassifierPort, x: torch.Tensor, keep: torch.Tensor) -> torch.Tensor:
    """``model`` in train mode, ``x [B,64,T]``, ``keep [B, T/8, 512]`` (1 = keep) -> logits (autograd-tracked)."""
    assert model.training
    if x.dim() == 3:
        x = x.unsqueeze(1)
    for i in (1, 2, 3):
        x = F.max_pool2d(F.relu(getattr(model, f"bn{i}")(getattr(model, f"conv{i}")(x))), 2)
    b, c, h, w = x.shape
    x = x.permute(0, 3, 1, 2).contiguous().view(b, w, c * h)
    y0 = _gru_layer(x, model.gru, 0)
    y0 = y0 * keep.to(y0.dtype) / (1.0 - DROPOUT_P)
    y1 = _gru_layer(y0, model.gru, 1)
    weights = F.softmax(model.attention(y1), dim=1)
    context = (y1 * weights).sum(dim=1)
    return model.fc(context)


def loss_and_grads(model: ClassifierPort, x, labels, keep):
    """-> (loss, logits, {param name: grad}); leaves the running statistics updated like one reference step."""
    model.train()
    model.zero_grad(set_to_none=True)
    logits = train_forward(model, x, keep)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    return float(loss.detach()), logits.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def recover_gru_dropout_keep(seed: int, batch: int, steps: int) -> torch.Tensor:
    """The keep mask nn.GRU(dropout=0.5) draws on the CPU right after ``torch.manual_seed(seed)``.

    ATen transposes a batch_first input to time-major and applies ``dropout`` to the layer-0 output
    ``[T, B, 512]`` with the global CPU generator; nothing before it in CNNAudioGRU.forward consumes random
    numbers.  Returned batch-major ``[B, T, 512]``.
    """
    torch.manual_seed(seed)
    keep = F.dropout(torch.ones(steps, batch, 512), DROPOUT_P, True) > 0
    return keep.transpose(0, 1).contiguous().to(torch.uint8)
