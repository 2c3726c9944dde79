"""B200-native mirror of the reference's ``models/models.py`` (``CNNAudioGRU``).

Same constructor, same submodule names - hence the same ``state_dict`` keys and shapes, so reference
checkpoints load with ``load_state_dict`` - and the same ``forward(x)`` contract
(/root/reference/models/models.py:6-68): ``x`` is ``[B, n_mels, T]`` or ``[B, 1, n_mels, T]``, the result is
``[B, num_classes]`` raw logits.  The torch submodules are parameter containers only: ``forward`` hands the
flattened parameters and the input to the hand-written sm_100a kernels behind ``sir_model_forward``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native
from ..utils.synth import state_dict_spec


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
        self._native_model = None
        self._uploaded_versions = None

    # -- weight hand-off ------------------------------------------------------------------------------------
    def _flat_weights(self) -> torch.Tensor:
        sd = self.state_dict()
        parts = []
        for key, shape in state_dict_spec(self.num_classes, self.n_mels):
            t = sd[key]
            if tuple(t.shape) != tuple(shape):
                raise _native.NativeError(f"{key}: expected {shape}, got {tuple(t.shape)}")
            parts.append(t.detach().reshape(-1).to(torch.float32))
        return torch.cat(parts).contiguous()

    def _versions(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict().values())

    def refresh_weights(self):
        """(Re)upload the parameters to the native handle; called automatically when they change."""
        if self._native_model is None:
            self._native_model = _native.Model(self.num_classes, self.n_mels)
        self._native_model.load_weights(self._flat_weights(), bn_eps=self.bn1.eps)
        self._uploaded_versions = self._versions()

    def forward(self, x):
        if self.training:
            raise _native.NativeError("CNNAudioGRU.forward: the CUDA path implements eval mode; call .eval() "
                                      "(training-mode kernels are not built yet, see DESIGN.md)")
        if x.dim() == 4:
            if x.size(1) != 1:
                raise _native.NativeError("expected [B, 1, n_mels, T]")
            x = x[:, 0]
        if x.dim() != 3 or x.size(1) != self.n_mels:
            raise _native.NativeError(f"expected [B, {self.n_mels}, T] or [B, 1, {self.n_mels}, T], got {tuple(x.shape)}")
        dev_in = x.device
        if self._native_model is None or self._uploaded_versions != self._versions():
            self.refresh_weights()
        y = self._native_model.forward(x.to(device="cuda", dtype=torch.float32))
        return y.to(dev_in)


if __name__ == "__main__":
    model = CNNAudioGRU(num_classes=31).cuda().eval()
    print("Output shape:", model(torch.randn(4, 64, 200, device="cuda")).shape)
