"""ORACLE (test infrastructure only) - the reference's CPU path restated over the SAME third-party calls.

The reference is pure Python and cannot travel to the GPU box (/root/reference does not exist there), but the
libraries that hold its arithmetic (torchaudio 2.11.0 transforms, torch.nn 2.11.0) are part of the image.
This port issues exactly the calls the reference issues, in the reference's order, so that timing it on the
box's host cores is timing the reference's own CPU implementation (``bench.py`` cpu_baseline kind "port" and
``--impl reference``), and so that tests have a second, independent checker next to the numpy restatement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchaudio


class FeaturePort:
    """ref: scripts/precompute_features.py:21-36 (constructor) and :49-75 (per-utterance body)."""

    def __init__(self, sample_rate=16000, n_mels=64, n_fft=1024, hop_length=512):
        self.sample_rate = sample_rate
        self.mel = torchaudio.transforms.MelSpectrogram(sample_rate=sample_rate, n_fft=n_fft,
                                                        hop_length=hop_length, n_mels=n_mels)
        self.to_db = torchaudio.transforms.AmplitudeToDB()

    def one(self, waveform: torch.Tensor, max_duration=5.0) -> torch.Tensor:
        """waveform [C, L] -> normalised log-mel [n_mels, T]; one call per utterance like the loop at :124-130."""
        if waveform.shape[0] > 1:
            waveform = waveform.mean(dim=0, keepdim=True)
        if max_duration is not None:
            limit = int(max_duration * self.sample_rate)
            if waveform.shape[1] > limit:
                waveform = waveform[:, :limit]
        m = self.to_db(self.mel(waveform)).squeeze(0)
        return (m - m.mean()) / (m.std() + 1e-5)

    def batch_padded(self, waves: torch.Tensor, lengths=None, target=200, max_duration=5.0) -> torch.Tensor:
        """Per-utterance loop + pad/trim to ``target`` frames (ref: scripts/dataset.py:109-113) -> [B,n_mels,target]."""
        out = []
        for i in range(waves.shape[0]):
            n = int(lengths[i]) if lengths is not None else waves.shape[1]
            m = self.one(waves[i:i + 1, :n], max_duration)
            if m.shape[1] > target:
                m = m[:, :target]
            elif m.shape[1] < target:
                m = F.pad(m, (0, target - m.shape[1]))
            out.append(m)
        return torch.stack(out)


class ClassifierPort(nn.Module):
    """ref: models/models.py:6-68.  Same submodule names, hence the same state_dict keys."""

    def __init__(self, num_classes, input_channels=1, n_mels=64):
        super().__init__()
        chans = (input_channels, 32, 64, 128)
        for i in range(3):
            setattr(self, f"conv{i + 1}", nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1, bias=False))
            setattr(self, f"bn{i + 1}", nn.BatchNorm2d(chans[i + 1]))
        self.gru = nn.GRU(128 * (n_mels // 8), 256, num_layers=2, batch_first=True, bidirectional=True, dropout=0.5)
        self.attention = nn.Linear(512, 1)
        self.fc = nn.Linear(512, num_classes)

    def forward(self, x, return_context=False):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        for i in (1, 2, 3):
            x = F.max_pool2d(F.relu(getattr(self, f"bn{i}")(getattr(self, f"conv{i}")(x))), 2)
        b, c, h, w = x.shape
        x = x.permute(0, 3, 1, 2).contiguous().view(b, w, c * h)
        x, _ = self.gru(x)
        weights = F.softmax(self.attention(x), dim=1)
        context = (x * weights).sum(dim=1)
        return context if return_context else self.fc(context)


def load_numpy_state(model: nn.Module, sd) -> nn.Module:
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=False)
    return model
