"""B200-native mirror of the reference's ``scripts/dataset.py`` (``FSCIntentDataset``).

Same constructor, ``__len__``, ``__getitem__ -> (Tensor[64, 200] fp32 on CPU, int)``, ``extract_features`` and
``augment_features`` as /root/reference/scripts/dataset.py:12-176; the on-disk feature cache written by
``precompute_dataset_features`` is read unchanged (:44-56).  All arithmetic (features, masking, pad/trim) runs
in the CUDA kernels of libsir_b200; CUDA cannot be used from DataLoader worker processes, so use
``num_workers=0`` with the per-item API, or - the B200-native way - ``get_batch(indices)``, which serves a
whole batch from an HBM-resident copy of the cache with device-side SpecAugment sampling in two launches.
"""
from __future__ import annotations

import json
import logging
import os

import numpy as np
import torch

from .. import _native
from .augment import apply_mask_params, draw_mask_params
from .precompute_features import AudioFeatureExtractor

logger = logging.getLogger(__name__)


class FSCIntentDataset(torch.utils.data.Dataset):
    """Fluent Speech Commands dataset with feature caching; features come from the GPU frontend."""

    def __init__(self, csv_path, label_map_path, is_training=True, augment_prob=0.5,
                 use_cache=True, cache_dir="data/cached_features", mel_spec_length=200):
        import pandas as pd

        self.data = pd.read_csv(csv_path)
        self.sample_rate = 16000
        self.is_training = is_training
        self.augment_prob = augment_prob if is_training else 0.0
        self.n_mels = 64
        self.mel_spec_length = mel_spec_length
        self.use_cache = use_cache
        with open(label_map_path, "r") as f:
            self.label_map = json.load(f)
        self.in_memory_cache = {}
        if use_cache:
            dataset_name = os.path.basename(csv_path).replace(".csv", "")
            self.cache_file = os.path.join(cache_dir, f"{dataset_name}_features.pt")
            if os.path.exists(self.cache_file):
                logger.info(f"Loading cached features from {self.cache_file}")
                self.features_dict = torch.load(self.cache_file)
                logger.info(f"Loaded {len(self.features_dict)} cached features")
            else:
                logger.info(f"No cached features found at {self.cache_file}")
                self.features_dict = {}
        else:
            self.features_dict = {}
        self._extractor = None          # created lazily: needs a CUDA device
        self._resident = None           # HBM-resident [N, n_mels, mel_spec_length] cache for get_batch
        self.time_mask_param, self.freq_mask_param = 20, 10      # reference :70-71
        logger.info(f"Initialized dataset with {len(self.data)} samples, {len(self.label_map)} classes")

    def __len__(self):
        return len(self.data)

    @property
    def extractor(self):
        if self._extractor is None:
            self._extractor = AudioFeatureExtractor(self.sample_rate, self.n_mels, 1024, 512)
        return self._extractor

    def _lookup(self, audio_path):
        if audio_path in self.in_memory_cache:
            return self.in_memory_cache[audio_path]
        if audio_path in self.features_dict:
            mel_spec = self.features_dict[audio_path]["features"]
            self.in_memory_cache[audio_path] = mel_spec
            return mel_spec
        mel_spec = self.extract_features(audio_path)
        if mel_spec is not None:
            self.in_memory_cache[audio_path] = mel_spec
        return mel_spec

    def __getitem__(self, idx):
        """(features [n_mels, mel_spec_length] on CPU, label id) - reference :78-115."""
        audio_path = self.data.iloc[idx]["path"]
        label = self.data.iloc[idx]["label"]
        label_id = self.label_map.get(label, 0)
        mel_spec = self._lookup(audio_path)
        params = None
        if self.is_training and np.random.random() < self.augment_prob:
            params = draw_mask_params(mel_spec.shape[0], mel_spec.shape[1], self.time_mask_param,
                                      self.freq_mask_param, gate=np.random.random)
        x = mel_spec.to(device="cuda", dtype=torch.float32)[None]
        masks = torch.tensor([params], dtype=torch.int32, device="cuda") if params is not None else None
        out = _native.features_finalize(x, self.mel_spec_length, masks=masks)[0]
        return out.cpu(), label_id

    def extract_features(self, audio_path):
        """Features of one file; zeros ``[n_mels, mel_spec_length]`` on any error (reference :117-158)."""
        feat = self.extractor.extract_features(audio_path, max_duration=5.0)
        if feat is None:
            return torch.zeros((self.n_mels, self.mel_spec_length))
        return feat

    def augment_features(self, mel_spec):
        """SpecAugment on an unpadded ``[n_mels, T]`` map (reference :160-176), host RNG in reference order."""
        params = draw_mask_params(mel_spec.shape[0], mel_spec.shape[1], self.time_mask_param,
                                  self.freq_mask_param, gate=np.random.random)
        return apply_mask_params(mel_spec, params)

    # -- batched, HBM-resident path (new) -------------------------------------------------------------------
    def make_resident(self, device="cuda"):
        """Upload every cached feature map once: ``[N, n_mels, mel_spec_length]`` + valid frame counts + labels."""
        n = len(self.data)
        host = torch.zeros((n, self.n_mels, self.mel_spec_length), dtype=torch.float32)
        frames = torch.zeros(n, dtype=torch.int32)
        labels = torch.zeros(n, dtype=torch.int64)
        for i in range(n):
            row = self.data.iloc[i]
            m = self._lookup(row["path"])
            t = min(m.shape[1], self.mel_spec_length)
            host[i, :, :t] = m[:, :t]
            # masks are drawn on the unpadded map, trimming happens afterwards (reference :105-113)
            frames[i] = m.shape[1]
            labels[i] = self.label_map.get(row["label"], 0)
        self._resident = (host.to(device), frames.to(device), labels.to(device))
        return self._resident

    def get_batch(self, indices: torch.Tensor, seed: int = 0, epoch: int = 0):
        """``indices [B]`` (CUDA int64) -> ``(features [B, n_mels, mel_spec_length], labels [B])`` on the GPU.

        SpecAugment parameters come from the counter-based device sampler keyed on ``(seed, epoch, index)``,
        so a sample's mask does not depend on how the epoch is batched or sharded across ranks.
        """
        if self._resident is None:
            self.make_resident()
        feats, frames, labels = self._resident
        idx = indices.to(feats.device)
        x = feats.index_select(0, idx)
        fr = frames.index_select(0, idx)
        masks = None
        if self.is_training and self.augment_prob > 0:
            # one Philox counter per (epoch, sample): drawn once per epoch for the whole dataset (one launch),
            # then gathered, so a sample's mask is independent of batching and of the rank that serves it
            key = (int(seed), int(epoch))
            if getattr(self, "_epoch_masks_key", None) != key:
                self._epoch_masks = _native.specaugment_sample(
                    seed, epoch * len(self.data), len(self.data), self.n_mels, 0, frames=frames,
                    augment_prob=self.augment_prob, time_mask_param=self.time_mask_param,
                    freq_mask_param=self.freq_mask_param, device=feats.device)
                self._epoch_masks_key = key
            masks = self._epoch_masks.index_select(0, idx).contiguous()
        out = _native.features_finalize(x, self.mel_spec_length, frames=torch.clamp(fr, max=self.mel_spec_length),
                                        masks=masks)
        return out, labels.index_select(0, idx)
