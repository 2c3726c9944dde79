"""Mirror of the one live function of the reference's ``scripts/augment.py``: ``apply_spec_augmentation``.

Reference: /root/reference/scripts/augment.py:137-165 (identical semantics to
``FSCIntentDataset.augment_features``, scripts/dataset.py:160-176).  The waveform augmentations of that file
(:6-135) are dead code in the reference (never imported, sox-dependent) and are out of scope.

The random draws are made on the host exactly in the reference's order - ``random.random()`` gates, then
``torch.rand(1)`` twice per applied mask inside torchaudio's ``mask_along_axis`` - so a seeded run reproduces
the reference's masks; the masking itself runs in the ``sir_features_finalize`` kernel.
"""
from __future__ import annotations

import random

import torch

from .. import _native


def draw_mask_params(n_mels: int, n_frames: int, time_mask_param: int = 20, freq_mask_param: int = 10,
                     gate=random.random):
    """[t_start, t_end, f_start, f_end] with the reference's RNG call order and fp32 arithmetic."""
    params = [0, 0, 0, 0]
    for k, (param, size) in enumerate(((time_mask_param, n_frames), (freq_mask_param, n_mels))):
        if gate() < 0.5:
            if param < 1:
                continue
            # TA:functional/functional.py:930-934
            value = torch.rand(1) * param
            min_value = torch.rand(1) * (size - value)
            start = int(min_value.long())
            params[2 * k] = start
            params[2 * k + 1] = start + int(value.long())
    return params


def apply_mask_params(mel_spec: torch.Tensor, params) -> torch.Tensor:
    """Zero the bands on a ``[n_mels, T]`` map with the CUDA kernel; result on the input's device."""
    dev_in = mel_spec.device
    x = mel_spec.to(device="cuda", dtype=torch.float32)[None]
    masks = torch.tensor([params], dtype=torch.int32, device="cuda")
    return _native.features_finalize(x, x.shape[-1], masks=masks)[0].to(dev_in)


def apply_spec_augmentation(mel_spec, time_mask_param=20, freq_mask_param=10):
    """Apply spectrogram augmentation to ``mel_spec [freq, time]`` (reference :137-165)."""
    params = draw_mask_params(mel_spec.shape[0], mel_spec.shape[1], time_mask_param, freq_mask_param)
    return apply_mask_params(mel_spec, params)
