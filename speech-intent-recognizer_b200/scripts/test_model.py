"""Mirror of the inference callers in the reference's ``scripts/test_model.py`` (:106-156): ``predict`` keeps the
reference signature and result dictionary; ``predict_batch`` is the batched form (features or waveforms already on the
GPU).  Softmax, arg-max, confidence and top-k run in one kernel (``sir_predict``) right behind the logits.
"""
from __future__ import annotations

import logging

import torch

from .. import _native

logger = logging.getLogger(__name__)
_extractor = None


def _features_of(audio_path):
    """The reference's module-level ``extract_features`` (:50-104): no 5 s truncation, ``[1, n_mels, T]`` or None."""
    global _extractor
    from .precompute_features import AudioFeatureExtractor
    if _extractor is None:
        _extractor = AudioFeatureExtractor()
    feat = _extractor.extract_features(audio_path, max_duration=None)
    return None if feat is None else feat.unsqueeze(0)


def get_top_predictions(topk_idx, topk_prob, inv_label_map):
    """One utterance's top-k as the reference formats it (:145-156)."""
    return [{"label": inv_label_map.get(int(i), "Unknown"), "probability": float(p)} for i, p in zip(topk_idx, topk_prob)]


def predict_batch(model, mel_specs: torch.Tensor, label_map, k: int = 3):
    """``mel_specs [B, n_mels, T]`` (CUDA) -> list of the reference's result dictionaries, one per utterance."""
    max_length = 200                                              # pad / truncate like :113-119
    if mel_specs.size(-1) > max_length:
        mel_specs = mel_specs[..., :max_length]
    elif mel_specs.size(-1) < max_length:
        mel_specs = torch.nn.functional.pad(mel_specs, (0, max_length - mel_specs.size(-1)))
    with torch.no_grad():
        logits = model(mel_specs.to(device="cuda", dtype=torch.float32))
    pred, conf, tk_i, tk_p = _native.predict(logits, k=min(k, logits.shape[1]))
    pred, conf, tk_i, tk_p = pred.cpu(), conf.cpu(), tk_i.cpu(), tk_p.cpu()
    inv = {v: key for key, v in label_map.items()}
    return [{"predicted_label": inv.get(int(pred[b]), "Unknown"), "confidence": float(conf[b]),
             "top_predictions": get_top_predictions(tk_i[b], tk_p[b], inv)} for b in range(logits.shape[0])]


def predict(model, audio_path, label_map, device=None):
    """Make a prediction on a single audio file (reference signature; ``None`` on any failure like :141-143)."""
    try:
        mel_spec = _features_of(audio_path)
        if mel_spec is None:
            return None
        return predict_batch(model, mel_spec, label_map)[0]
    except Exception as e:  # noqa: BLE001 - the reference swallows every error here
        logger.error(f"Error during prediction: {str(e)}")
        return None
