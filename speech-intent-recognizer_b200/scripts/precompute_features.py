"""B200-native mirror of the reference's ``scripts/precompute_features.py`` (same names, same contracts).

``AudioFeatureExtractor`` keeps the reference constructor and attributes
(/root/reference/scripts/precompute_features.py:21-36): ``.mel_transform`` and ``.amplitude_to_db`` are
callables with torchaudio's semantics, ``.extract_features(audio_path, max_duration=5.0)`` returns a CPU
``Tensor[n_mels, T]`` or ``None`` on any error (:38-79).  Underneath, every number is produced by the fused
sm_100a kernel behind ``sir_frontend_forward``; there is no CPU path.  ``extract_batch`` is the batched entry
the reference lacks: one launch for a whole batch of (ragged) utterances already resident in HBM.
``precompute_dataset_features`` writes the reference's on-disk cache format (:98-101,134-142) but runs the
CSV through the GPU in batches instead of one utterance at a time (:124-139).
"""
from __future__ import annotations

import argparse
import json
import logging
import os

import torch

from .. import _native
from ..utils.audio_io import load_audio

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
logger = logging.getLogger(__name__)


class _MelTransform:
    """Callable standing in for ``torchaudio.transforms.MelSpectrogram(sample_rate, n_fft, hop_length, n_mels)``.

    ``waveform [..., L]`` -> power mel spectrogram ``[..., n_mels, 1 + L // hop]`` on the input's device.
    """

    def __init__(self, frontend: "_native.Frontend"):
        self._fe = frontend

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        return _run(self._fe, waveform, _native.OUT_MEL_POWER)

    forward = __call__


class _AmplitudeToDB:
    """Callable standing in for ``torchaudio.transforms.AmplitudeToDB()`` (power, amin 1e-10, no top_db)."""

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        y = _native.amplitude_to_db(x.to(device="cuda", dtype=torch.float32))
        return y.to(x.device)

    forward = __call__


def _run(fe, waveform, mode, max_samples=0):
    lead = waveform.shape[:-1]
    w = waveform.reshape(-1, waveform.shape[-1]).to(device="cuda", dtype=torch.float32)
    if w.shape[-1] <= fe.hop:
        # torch.stft raises for reflect padding >= signal length; keep that error behaviour
        raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, "
                           f"but got: padding ({fe.hop}, {fe.hop}) at dimension 2 of input {list(waveform.shape)}")
    out = fe.forward(w.contiguous(), mode=mode, max_samples=max_samples)
    return out.reshape(*lead, out.shape[-2], out.shape[-1]).to(waveform.device)


class AudioFeatureExtractor:
    """Extract log-mel features on the GPU (reference: scripts/precompute_features.py:18-79)."""

    def __init__(self, sample_rate=16000, n_mels=64, n_fft=1024, hop_length=512):
        self.sample_rate = sample_rate
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.hop_length = hop_length
        self._fe = _native.Frontend(sample_rate, n_mels, n_fft, hop_length)
        self.mel_transform = _MelTransform(self._fe)
        self.amplitude_to_db = _AmplitudeToDB()
        self._resamplers = {}

    # -- batched entry (new) ------------------------------------------------------------------------------
    def extract_batch(self, waveforms: torch.Tensor, lengths: torch.Tensor = None, max_duration=5.0,
                      out_frames: int = None, masks: torch.Tensor = None, status: torch.Tensor = None,
                      out: torch.Tensor = None) -> torch.Tensor:
        """``waveforms [B, Lmax]`` fp32 CUDA, ``lengths [B]`` int32 CUDA -> ``[B, n_mels, out_frames]`` CUDA.

        Per utterance: truncate to ``int(max_duration * sample_rate)`` samples, log-mel in dB, normalise with
        the utterance's own mean / unbiased std over its ``T_i = 1 + L_i // hop`` valid frames, optional
        SpecAugment bands ``masks [B, 4]`` on the valid frames, zero tail / trim to ``out_frames``.
        """
        max_samples = int(max_duration * self.sample_rate) if max_duration is not None else 0
        return self._fe.forward(waveforms, lengths=lengths, max_samples=max_samples, mode=_native.OUT_LOGMEL_NORM,
                                out_frames=out_frames, masks=masks, status=status, out=out)

    # -- reference per-file entry -------------------------------------------------------------------------
    def load_waveform(self, audio_path):
        """File -> mono fp32 ``[1, L]`` at ``self.sample_rate`` (reference lines :47-56)."""
        waveform, sr = load_audio(audio_path)
        if waveform.shape[0] > 1:
            waveform = torch.mean(waveform, dim=0, keepdim=True)
        if sr != self.sample_rate:                               # reference: torchaudio.transforms.Resample (:54-56)
            key = (int(sr), int(self.sample_rate))
            if key not in self._resamplers:
                self._resamplers[key] = _native.Resampler(*key)
            waveform = self._resamplers[key](waveform.to(device="cuda", dtype=torch.float32))
        return waveform

    def extract_features(self, audio_path, max_duration=5.0):
        """Extract mel spectrogram features from an audio file -> CPU ``Tensor[n_mels, T]`` or ``None``."""
        try:
            if not os.path.exists(audio_path):
                logger.error(f"File not found: {audio_path}")
                return None
            waveform = self.load_waveform(audio_path).to(device="cuda", dtype=torch.float32)
            if waveform.shape[1] <= self.hop_length:
                raise RuntimeError("audio shorter than the reflect padding")
            return self.extract_batch(waveform.contiguous(), max_duration=max_duration)[0].cpu()
        except Exception as e:  # noqa: BLE001 - the reference swallows every error here (:77-79)
            logger.error(f"Error processing {audio_path}: {str(e)}")
            return None


def _label_column(df):
    """Label-column selection of the reference (:108-120)."""
    if "label" in df.columns:
        return "label"
    if "intent" in df.columns:
        return "intent"
    if "action" in df.columns and "object" in df.columns:
        df["label"] = df["action"] + "_" + df["object"]
        return "label"
    df["label"] = "unknown"
    logger.warning("Could not find label column, using 'unknown' as label")
    return "label"


def precompute_dataset_features(csv_path, output_dir, label_map_path=None, max_duration=5.0, batch_size=256):
    """Precompute and cache all features of a CSV; returns the cache path (reference :81-147).

    The cache is the reference's format - ``torch.save({path: {'features': Tensor[n_mels, T] (CPU),
    'label': str}})`` to ``<output_dir>/<csv basename>_features.pt`` - consumed unchanged by
    ``FSCIntentDataset`` (scripts/dataset.py:44-56,87-94).
    """
    import pandas as pd

    df = pd.read_csv(csv_path)
    logger.info(f"Loaded {len(df)} samples from {csv_path}")
    extractor = AudioFeatureExtractor()
    os.makedirs(output_dir, exist_ok=True)
    dataset_name = os.path.basename(csv_path).replace(".csv", "")
    cache_file = os.path.join(output_dir, f"{dataset_name}_features.pt")
    label_column = _label_column(df)
    logger.info(f"Using '{label_column}' column for labels")

    features_dict, error_count = {}, 0
    max_samples = int(max_duration * extractor.sample_rate)
    for start in range(0, len(df), batch_size):
        rows = df.iloc[start:start + batch_size]
        waves, metas = [], []
        for _, row in rows.iterrows():
            path = row["path"]
            try:
                if not os.path.exists(path):
                    raise FileNotFoundError(f"File not found: {path}")
                w = extractor.load_waveform(path)[0, :max_samples]
                if w.shape[0] <= extractor.hop_length:
                    raise RuntimeError("audio shorter than the reflect padding")
                waves.append(w)
                metas.append((path, row[label_column]))
            except Exception as e:  # noqa: BLE001
                logger.error(f"Error processing {path}: {str(e)}")
                error_count += 1
        if not waves:
            continue
        lens = torch.tensor([w.shape[0] for w in waves], dtype=torch.int32)
        lmax = (int(lens.max()) + 3) // 4 * 4
        host = torch.zeros((len(waves), lmax), dtype=torch.float32).pin_memory()
        for i, w in enumerate(waves):
            host[i, : w.shape[0]] = w.to("cpu", torch.float32)
        feats = extractor.extract_batch(host.cuda(non_blocking=True), lens.cuda(non_blocking=True),
                                        max_duration=max_duration).cpu()
        for i, (path, label) in enumerate(metas):
            t = 1 + int(lens[i]) // extractor.hop_length
            features_dict[path] = {"features": feats[i, :, :t].clone(), "label": label}

    torch.save(features_dict, cache_file)
    logger.info(f"Saved {len(features_dict)} features to {cache_file}")
    logger.info(f"Failed to process {error_count} files")
    return cache_file


def main():
    parser = argparse.ArgumentParser(description="Precompute audio features on the GPU")
    parser.add_argument("--train_csv", type=str, required=True, help="Path to training CSV file")
    parser.add_argument("--valid_csv", type=str, required=True, help="Path to validation CSV file")
    parser.add_argument("--test_csv", type=str, required=True, help="Path to test CSV file")
    parser.add_argument("--output_dir", type=str, default="data/cached_features")
    parser.add_argument("--label_map", type=str, default=None, help="Path to label map JSON file")
    args = parser.parse_args()
    os.makedirs(args.output_dir, exist_ok=True)
    logger.info("Starting feature precomputation...")
    cache_info = {
        "train_features": precompute_dataset_features(args.train_csv, args.output_dir, args.label_map),
        "valid_features": precompute_dataset_features(args.valid_csv, args.output_dir, args.label_map),
        "test_features": precompute_dataset_features(args.test_csv, args.output_dir, args.label_map),
    }
    with open(os.path.join(args.output_dir, "cache_info.json"), "w") as f:
        json.dump(cache_info, f, indent=2)
    logger.info("Feature precomputation complete!")


if __name__ == "__main__":
    main()
