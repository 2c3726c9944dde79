"""Import alias: ``import sir_b200`` == ``importlib.import_module("speech-intent-recognizer_b200")``."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("speech-intent-recognizer_b200")
