"""B200-native hot path of avi2924/Speech-Intent-Recognizer: batched log-mel frontend + CNNAudioGRU forward.

The directory name carries a hyphen (the project's name), so import it with
``importlib.import_module("speech-intent-recognizer_b200")`` or through the ``sir_b200`` alias module at the
repo root.  Layout:
    csrc/                       hand-written sm_100a kernels + the C ABI (include/sir_b200.h)
    _native.py                  ctypes binding (no fallback)
    scripts/, models/           host-side mirrors of the reference's Python interface for this path
    utils/                      synthetic inputs/weights, audio file I/O
"""
__version__ = "0.1.0"
