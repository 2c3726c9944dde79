"""B200-native hot path of avi2924/Speech-Intent-Recognizer: batched log-mel frontend + CNNAudioGRU forward/backward.

The directory name carries a hyphen (the project's name), so import it with
``importlib.import_module("speech-intent-recognizer_b200")`` or through the ``sir_b200`` alias module at the
repo root.  Layout:
    csrc/                       hand-written sm_100a kernels + the C ABI (include/sir_b200.h)
    _native.py                  ctypes binding (no fallback)
    scripts/, models/           host-side mirrors of the reference's Python interface for this path
    pipeline.py                 host-buffer entry: H2D copy overlapped with the frontend + conv stack
    utils/                      synthetic inputs/weights, audio file I/O

The reference-facing names are re-exported lazily (``from sir_b200 import CNNAudioGRU``).
"""
import importlib

__version__ = "0.2.0"

_EXPORTS = {
    "AudioFeatureExtractor": "scripts.precompute_features", "precompute_dataset_features": "scripts.precompute_features",
    "FSCIntentDataset": "scripts.dataset", "apply_spec_augmentation": "scripts.augment",
    "collate_fn": "scripts.train", "train_epoch": "scripts.train", "validate": "scripts.train",
    "DataParallelTrainer": "scripts.train", "CNNAudioGRU": "models.models", "IntentPipeline": "pipeline",
    "predict": "scripts.test_model", "predict_batch": "scripts.test_model", "evaluate_loader": "scripts.evaluate",
}


def __getattr__(name):
    if name in _EXPORTS:
        return getattr(importlib.import_module(f"{__name__}.{_EXPORTS[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
