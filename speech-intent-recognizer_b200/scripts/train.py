"""Mirror of the hot-path pieces of the reference's ``scripts/train.py``: ``collate_fn`` (:49-70).

Importing the reference's own module sets ``CUDA_VISIBLE_DEVICES=0`` as a side effect (scripts/train.py:17),
which would hide GPUs 1-7 in a data-parallel job - this mirror has no import-time side effects.
The data-parallel training step (:72-118 plus the gradient all-reduce the reference lacks) is not built yet;
see DESIGN.md "what comes next".
"""
from __future__ import annotations

import torch


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
