"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run once in the build container (the reference tree is mounted read-only at /root/reference and is not
available on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules unmodified -
``scripts.precompute_features.AudioFeatureExtractor`` (/root/reference/scripts/precompute_features.py:18-79),
``scripts.dataset.FSCIntentDataset.augment_features`` (/root/reference/scripts/dataset.py:160-176) and
``models.models.CNNAudioGRU`` (/root/reference/models/models.py:5-68) - applies them to the seeded synthetic
inputs of ``utils/synth.py`` and stores inputs' checksums + the reference outputs as small ``.npz`` files.
``torchaudio.load`` cannot run in this image (no torchcodec), so the file-reading lines of
``extract_features`` (:41-56) are skipped and lines :49-75 are replayed on in-memory waveforms.
"""
import hashlib
import importlib
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

synth = importlib.import_module("speech-intent-recognizer_b200.utils.synth")

from scripts.precompute_features import AudioFeatureExtractor  # noqa: E402  (the reference)
from scripts.dataset import FSCIntentDataset  # noqa: E402
from models.models import CNNAudioGRU  # noqa: E402
import torchaudio  # noqa: E402

torch.set_num_threads(1)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_features(ex, wave_1xL: torch.Tensor, max_duration=5.0):
    """Lines 49-75 of the reference's extract_features, minus file I/O."""
    waveform = wave_1xL
    if waveform.shape[0] > 1:
        waveform = torch.mean(waveform, dim=0, keepdim=True)
    if max_duration is not None:
        max_samples = int(max_duration * ex.sample_rate)
        if waveform.shape[1] > max_samples:
            waveform = waveform[:, :max_samples]
    mel_spec = ex.mel_transform(waveform)
    mel_db = ex.amplitude_to_db(mel_spec)
    m = mel_db.squeeze(0)
    return mel_spec.squeeze(0), mel_db.squeeze(0), (m - m.mean()) / (m.std() + 1e-5)


def frontend_golden():
    ex = AudioFeatureExtractor()
    lengths = [48000, 20352, 90000, 5000, 53760, 1000]
    L = max(lengths)
    waves = synth.speech_like(11, len(lengths), L, lengths)
    noise = synth.white_noise(12, 1, 48000)
    out = {"lengths": np.asarray(lengths, np.int64), "speech_seed": 11, "noise_seed": 12,
           "speech_sha": sha(waves), "noise_sha": sha(noise)}
    for i, n in enumerate(lengths):
        p, d, f = ref_features(ex, torch.from_numpy(waves[i:i + 1, :n]))
        out[f"feat_{i}"] = f.numpy()
        if i == 0:
            out["mel_power_0"] = p.numpy()
            out["mel_db_0"] = d.numpy()
    _, _, f = ref_features(ex, torch.from_numpy(noise))
    out["feat_noise"] = f.numpy()
    # un-truncated twin (scripts/test_model.py:50-104) on the 90000-sample clip
    _, _, f = ref_features(ex, torch.from_numpy(waves[2:3, :90000]), max_duration=None)
    out["feat_2_untruncated"] = f.numpy()
    # all-zero (digital silence) utterance: dB is -100 everywhere, std 0 -> features exactly 0
    _, _, f = ref_features(ex, torch.zeros(1, 16000))
    out["feat_silence"] = f.numpy()
    # 80-mel frontend for config 5 (first 2 s of the clip to keep the file small)
    ex80 = AudioFeatureExtractor(n_mels=80)
    _, _, f = ref_features(ex80, torch.from_numpy(waves[0:1, :32000]))
    out["feat80_0"] = f.numpy()
    out["fb"] = ex.mel_transform.mel_scale.fb.numpy()
    out["window"] = ex.mel_transform.spectrogram.window.numpy()
    np.savez_compressed(os.path.join(HERE, "frontend.npz"), **out)
    return waves, lengths, ex


def augment_golden(waves, lengths, ex):
    """Replay FSCIntentDataset.augment_features with seeded global RNGs and record the uniforms it drew."""
    with tempfile.TemporaryDirectory() as td:
        csv = os.path.join(td, "d.csv")
        with open(csv, "w") as f:
            f.write("path,label\nx.wav,a\n")
        lm = os.path.join(td, "lm.json")
        with open(lm, "w") as f:
            json.dump({"a": 0}, f)
        ds = FSCIntentDataset(csv, lm, is_training=True, augment_prob=1.0, use_cache=False)
    _, _, feat = ref_features(ex, torch.from_numpy(waves[0:1, :lengths[0]]))
    cases = []
    seed = 0
    while len(cases) < 12:
        seed += 1
        np.random.seed(seed)
        torch.manual_seed(seed)
        got = ds.augment_features(feat.clone())
        # replay the draws: gate, [U1, U2], gate, [U1, U2]
        np.random.seed(seed)
        torch.manual_seed(seed)
        u = np.ones(6, np.float32)
        u[0] = np.random.random()
        if u[0] < 0.5:
            u[1] = torch.rand(1).item()
            u[2] = torch.rand(1).item()
        u[3] = np.random.random()
        if u[3] < 0.5:
            u[4] = torch.rand(1).item()
            u[5] = torch.rand(1).item()
        cases.append((seed, u, got.numpy()))
    np.savez_compressed(os.path.join(HERE, "augment.npz"),
                        seeds=np.asarray([c[0] for c in cases]),
                        uniforms=np.stack([c[1] for c in cases]),
                        outputs=np.stack([c[2] for c in cases]),
                        base=feat.numpy())


def classifier_golden(waves, lengths, ex):
    sd = synth.make_weights(1234)
    model = CNNAudioGRU(num_classes=31).eval()
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert all("num_batches_tracked" in k for k in missing.missing_keys), missing
    feats = []
    for i in (0, 1, 2, 4):
        _, _, f = ref_features(ex, torch.from_numpy(waves[i:i + 1, :lengths[i]]))
        f = f[:, :200] if f.shape[1] > 200 else torch.nn.functional.pad(f, (0, 200 - f.shape[1]))
        feats.append(f)
    x = torch.stack(feats)
    with torch.no_grad():
        logits = model(x)
        # config-1 call pattern: 4-D [1,1,64,T] with variable T, no padding (scripts/test_tts_samples.py:77-87)
        _, _, fvar = ref_features(ex, torch.from_numpy(waves[1:2, :lengths[1]]))
        logits_var = model(fvar[None, None])
    np.savez_compressed(os.path.join(HERE, "classifier.npz"), weight_seed=1234, weights_sha=sha(synth.flatten_weights(sd)),
                        x=x.numpy(), logits=logits.numpy(), x_var=fvar.numpy(), logits_var=logits_var.numpy())


if __name__ == "__main__":
    waves, lengths, ex = frontend_golden()
    augment_golden(waves, lengths, ex)
    classifier_golden(waves, lengths, ex)
    meta = {"torch": torch.__version__, "torchaudio": torchaudio.__version__, "numpy": np.__version__,
            "reference": "avi2924/Speech-Intent-Recognizer mounted at /root/reference"}
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
