"""Mirror of the evaluation loop of the reference's ``scripts/evaluate.py`` (:76-98): arg-max predictions, accuracy and
the confusion matrix, with the counting done on the device (``sir_predict``) instead of collecting every prediction on
the host for sklearn.  Report formatting / plotting (:90-117) is out of scope.
"""
from __future__ import annotations

import torch

from .. import _native


def evaluate_loader(model, loader, num_classes: int, device="cuda"):
    """-> ``(accuracy, confusion [C, C] int64 CPU tensor; rows = true label, columns = prediction)``.

    ``loader`` yields ``(mel [B, n_mels, T], label [B])`` like the reference's DataLoader with ``collate_fn``;
    ``(None, None)`` batches are skipped as in scripts/train.py:82-83.
    """
    model.eval()
    confusion = torch.zeros((num_classes, num_classes), device="cuda", dtype=torch.int64)
    correct = torch.zeros(1, device="cuda", dtype=torch.int64)
    total = 0
    with torch.no_grad():
        for mel, label in loader:
            if mel is None or label is None or mel.size(0) == 0:
                continue
            logits = model(mel.to(device="cuda", dtype=torch.float32, non_blocking=True))
            _native.predict(logits, k=0, labels=label.to(device="cuda", dtype=torch.int64, non_blocking=True),
                            confusion=confusion, correct=correct)
            total += int(label.size(0))
    return int(correct.item()) / max(total, 1), confusion.cpu()
