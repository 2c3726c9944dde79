"""ORACLE (test infrastructure only) - numpy restatement of the reference's evaluation head.

Only ``tests/`` may import this module.  Follows /root/reference/scripts/test_model.py:121-156 (softmax, argmax,
confidence, ``get_top_predictions``: ``np.argsort(probs)[::-1][:k]``) and /root/reference/scripts/evaluate.py:79-98
(argmax, then sklearn ``accuracy_score`` / ``confusion_matrix`` over the collected predictions - restated as the counts
those functions are defined by, since sklearn's plotting stack is not part of the path).
"""
from __future__ import annotations

import numpy as np


def softmax(logits: np.ndarray) -> np.ndarray:
    """torch.nn.functional.softmax(output, dim=1) in fp32 (test_model.py:124)."""
    x = logits.astype(np.float32)
    e = np.exp(x - x.max(axis=1, keepdims=True), dtype=np.float32)
    return (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)


def predict(logits: np.ndarray, k: int = 3):
    """-> (pred [B], confidence [B], topk_idx [B,k], topk_prob [B,k]) per test_model.py:123-156."""
    probs = softmax(logits)
    pred = logits.argmax(axis=1)                                 # torch.argmax: first maximum
    conf = probs[np.arange(len(pred)), pred]
    idx = np.stack([np.argsort(p)[::-1][:k] for p in probs])     # get_top_predictions
    return pred, conf, idx, np.take_along_axis(probs, idx, axis=1)


def accuracy_and_confusion(pred: np.ndarray, labels: np.ndarray, num_classes: int):
    """accuracy_score and confusion_matrix (rows = true label, columns = prediction) of evaluate.py:88,96."""
    cm = np.zeros((num_classes, num_classes), dtype=np.int64)
    np.add.at(cm, (labels, pred), 1)
    return float((pred == labels).mean()) if len(pred) else 0.0, cm
