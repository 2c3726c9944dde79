"""ORACLE (test infrastructure only) - numpy restatement of ``CNNAudioGRU.forward`` in eval mode.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  Follows /root/reference/models/models.py:41-68 line by line; the layer arithmetic is
torch.nn's (Conv2d 3x3 s1 p1 no bias, BatchNorm2d eval with eps 1e-5, ReLU, MaxPool2d(2), 2-layer
bidirectional GRU(hidden 256, gate order r,z,n), Linear).  Pinned against the reference class imported in
the build container through tests/golden/classifier_*.npz (tests/test_oracle_cpu.py).
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, ref: models/models.py:11,13,15


def _conv3x3(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """x [B,C,H,W], w [O,C,3,3] -> [B,O,H,W]; stride 1, zero padding 1, no bias (ref: models/models.py:10-14)."""
    B, C, H, W = x.shape
    O = w.shape[0]
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    s = xp.strides
    win = np.lib.stride_tricks.as_strided(xp, (B, H, W, C, 3, 3), (s[0], s[2], s[3], s[1], s[2], s[3]))
    cols = win.reshape(B * H * W, C * 9)
    y = cols @ w.reshape(O, C * 9).T
    return np.ascontiguousarray(y.reshape(B, H, W, O).transpose(0, 3, 1, 2))


def _bn_eval(x, weight, bias, mean, var):
    """BatchNorm2d in eval mode: (x - running_mean) / sqrt(running_var + eps) * weight + bias."""
    dt = x.dtype
    scale = (weight.astype(np.float64) / np.sqrt(var.astype(np.float64) + BN_EPS))
    shift = bias.astype(np.float64) - mean.astype(np.float64) * scale
    return x * scale.astype(dt)[None, :, None, None] + shift.astype(dt)[None, :, None, None]


def _maxpool2(x):
    """MaxPool2d(2): floor mode, odd trailing row/column dropped (ref: models/models.py:19)."""
    B, C, H, W = x.shape
    x = x[:, :, : H // 2 * 2, : W // 2 * 2]
    return x.reshape(B, C, H // 2, 2, W // 2, 2).max(axis=(3, 5))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one GRU layer, x [B,T,I] -> [B,T,256].

    torch.nn.GRU: r = s(W_ir x + b_ir + W_hr h + b_hr); z likewise;
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h.  h0 = 0.
    """
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x.reshape(B * T, -1) @ w_ih.T + b_ih
    gi = gi.reshape(B, T, 3 * H)
    h = np.zeros((B, H), dtype=x.dtype)
    out = np.zeros((B, T, H), dtype=x.dtype)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, t, :H] + gh[:, :H])
        z = _sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        out[:, t] = h
    return out


def forward(x: np.ndarray, sd, dtype=np.float32, return_intermediates: bool = False):
    """Logits ``[B, num_classes]`` from features ``[B, n_mels, T]`` or ``[B, 1, n_mels, T]``.

    ``sd`` maps the reference ``state_dict`` keys to arrays.
    ref: models/models.py:41-68.  Dropout (:20) is never called in forward; GRU inter-layer dropout is
    inactive in eval mode.
    """
    P = {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v).astype(dtype)
         for k, v in sd.items() if "num_batches_tracked" not in k}
    x = np.asarray(x, dtype=dtype)
    if x.ndim == 3:                                   # :46-47
        x = x[:, None]
    inter = {}
    for i in (1, 2, 3):                               # :50-52
        x = _conv3x3(x, P[f"conv{i}.weight"])
        x = _bn_eval(x, P[f"bn{i}.weight"], P[f"bn{i}.bias"], P[f"bn{i}.running_mean"], P[f"bn{i}.running_var"])
        x = _maxpool2(np.maximum(x, 0))
        inter[f"pool{i}"] = x
    b, c, h, w = x.shape                              # :55-57  feature index = c*h_dim + h
    x = np.ascontiguousarray(x.transpose(0, 3, 1, 2)).reshape(b, w, c * h)
    inter["gru_in"] = x
    for layer in (0, 1):                              # :60
        outs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            outs.append(_gru_direction(x, P[f"gru.weight_ih_l{layer}{sfx}"], P[f"gru.weight_hh_l{layer}{sfx}"],
                                       P[f"gru.bias_ih_l{layer}{sfx}"], P[f"gru.bias_hh_l{layer}{sfx}"], rev))
        x = np.concatenate(outs, axis=2)
        inter[f"gru_l{layer}"] = x
    a = x @ P["attention.weight"].T + P["attention.bias"]          # :63  [B,T,1]
    a = a - a.max(axis=1, keepdims=True)
    wts = np.exp(a)
    wts = wts / wts.sum(axis=1, keepdims=True)
    ctx = (x * wts).sum(axis=1)                                     # :64
    logits = ctx @ P["fc.weight"].T + P["fc.bias"]                  # :67
    if return_intermediates:
        inter["context"] = ctx
        return logits.astype(dtype), inter
    return logits.astype(dtype)
