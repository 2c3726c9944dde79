"""GPU tests of the reference-facing mirrors END TO END: files on disk -> cache -> dataset -> batches -> predictions.

The kernels underneath are covered by tests/test_gpu_parity.py; these tests run the Python classes a user of the
reference would switch to (``AudioFeatureExtractor.extract_features(path)``, ``precompute_dataset_features``,
``FSCIntentDataset.__getitem__ / make_resident / get_batch``, ``apply_spec_augmentation``, ``test_model.predict``)
against the oracle, including the full 256-utterance config-2 batch (the batch size that takes the chained GRU kernel).

References: /root/reference/scripts/precompute_features.py:38-147, scripts/dataset.py:15-176, scripts/augment.py:137-165,
scripts/test_model.py:106-143, models/models.py:41-68.
"""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import classifier_np, logmel_np
from oracle.torch_port import ClassifierPort, FeaturePort, load_numpy_state
from tests.util import FEATURE_REL_TOL, LOGIT_ABS_TOL, rel_to_scale, synth

pytestmark = pytest.mark.gpu
native = importlib.import_module("speech-intent-recognizer_b200._native")
pre = importlib.import_module("speech-intent-recognizer_b200.scripts.precompute_features")
dataset = importlib.import_module("speech-intent-recognizer_b200.scripts.dataset")
augment = importlib.import_module("speech-intent-recognizer_b200.scripts.augment")
test_model = importlib.import_module("speech-intent-recognizer_b200.scripts.test_model")
models = importlib.import_module("speech-intent-recognizer_b200.models.models")
audio_io = importlib.import_module("speech-intent-recognizer_b200.utils.audio_io")


def cuda_model(seed=1234):
    m = models.CNNAudioGRU(31)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(seed).items()}, strict=False)
    return m.cuda().eval()


def pcm_roundtrip(wave):
    """What a 16-bit WAV holds and torchaudio.load hands on: round(x * 32768) clipped, / 32768."""
    return (np.clip(np.round(np.asarray(wave, np.float32) * 32768.0), -32768, 32767) / 32768.0).astype(np.float32)


def test_config2_full_batch_matches_oracle():
    """BASELINE configs[1] at FULL size: 256 utterances x 3 s through ``extract_batch`` + ``CNNAudioGRU`` against the
    oracle on all 256 x 31 logits.  At this batch the recurrence runs ``gru_layer_pp_kernel<2>`` (two 32-utterance chains per
    cluster), which smaller parity batches never reach."""
    B = 256
    w = synth.speech_like(2026, B, 48000)
    ex = pre.AudioFeatureExtractor()
    model = cuda_model()
    feats = ex.extract_batch(torch.from_numpy(w).cuda(), out_frames=200)
    logits = model(feats).cpu().numpy()
    feats = feats.cpu().numpy()
    torch.set_num_threads(os.cpu_count() or 1)
    want_f = FeaturePort().batch_padded(torch.from_numpy(w), target=200)          # the reference's own torchaudio calls
    with torch.no_grad():
        want = load_numpy_state(ClassifierPort(31).eval(), synth.make_weights(1234))(want_f).numpy()
    want_f = want_f.numpy()
    for i in range(B):
        assert rel_to_scale(feats[i], want_f[i]) < FEATURE_REL_TOL, i
    assert np.max(np.abs(logits - want)) < LOGIT_ABS_TOL, float(np.max(np.abs(logits - want)))
    top = np.sort(want, axis=1)
    safe = (top[:, -1] - top[:, -2]) > 4 * LOGIT_ABS_TOL
    assert safe.sum() >= 250, int(safe.sum())                 # how many rows have an arg-max that 1e-3 cannot flip
    assert np.array_equal(logits.argmax(1)[safe], want.argmax(1)[safe])
    assert len(set(want.argmax(1).tolist())) >= 10            # not vacuous: many classes predicted
    # the numpy restatement agrees with the torchaudio port on a sample of rows (two independent checkers)
    for i in (0, 100, 255):
        assert rel_to_scale(logmel_np.dataset_item(w[i]), want_f[i]) < 1e-5


@pytest.fixture(scope="module")
def wav_corpus(tmp_path_factory):
    """Six PCM16 WAV files (16 kHz mono, 16 kHz stereo, 24 kHz -> resampler branch, a 6 s clip -> 5 s truncation, one too
    short, one missing) + the reference's CSV (columns path, label) + label map."""
    root = tmp_path_factory.mktemp("corpus")
    base = synth.speech_like(77, 5, 96000)
    entries = []

    def add(name, wave, rate, label):
        path = str(root / name)
        audio_io.write_wav_pcm16(path, wave, rate)
        entries.append({"path": path, "label": label, "wave": pcm_roundtrip(wave), "rate": rate})

    add("a_16k.wav", base[0, :48000], 16000, "activate_lights")
    add("b_16k_stereo.wav", np.stack([base[1, :40000], 0.5 * base[2, :40000]]), 16000, "increase_volume")
    add("c_24k.wav", base[2, :60000], 24000, "decrease_heat")
    add("d_long.wav", base[3, :96000], 16000, "activate_lights")
    add("e_short.wav", base[4, :400], 16000, "increase_volume")        # <= 512 samples: the reference's stft raises -> skipped
    entries.append({"path": str(root / "missing.wav"), "label": "decrease_heat", "wave": None, "rate": 16000})
    import pandas as pd
    csv = str(root / "train_data.csv")
    pd.DataFrame([{"path": e["path"], "label": e["label"]} for e in entries]).to_csv(csv, index=False)
    label_map = {"activate_lights": 0, "decrease_heat": 1, "increase_volume": 2}
    lm = str(root / "label_map.json")
    with open(lm, "w") as f:
        json.dump(label_map, f)
    return {"root": str(root), "csv": csv, "label_map_path": lm, "label_map": label_map, "entries": entries}


def oracle_features(entry, max_duration=5.0):
    """The reference's per-file path on the decoded samples: mono mean -> resample -> truncate -> log-mel -> normalise."""
    w = entry["wave"]
    if w.ndim == 2:
        w = w.mean(axis=0).astype(np.float32)                 # scripts/precompute_features.py:50-51
    if entry["rate"] != 16000:
        w = logmel_np.resample(w, entry["rate"], 16000)       # :54-56
    return logmel_np.extract_features(w, max_duration=max_duration)


def test_extract_features_from_wav_files(wav_corpus):
    """``AudioFeatureExtractor.extract_features(path)`` (scripts/precompute_features.py:38-79) on real PCM16 files."""
    ex = pre.AudioFeatureExtractor()
    for e in wav_corpus["entries"][:4]:
        got = ex.extract_features(e["path"])
        want = oracle_features(e)
        assert got.device.type == "cpu" and tuple(got.shape) == want.shape, e["path"]
        assert rel_to_scale(got.numpy(), want) < FEATURE_REL_TOL, e["path"]
    assert oracle_features(wav_corpus["entries"][3]).shape[1] == 157           # 5 s truncation: 1 + 80000 // 512
    assert ex.extract_features(wav_corpus["entries"][4]["path"]) is None       # too short -> None (:77-79)
    assert ex.extract_features(wav_corpus["entries"][5]["path"]) is None       # missing file -> None (:42-44)
    # the un-truncated twin (scripts/test_model.py:50-104)
    got = ex.extract_features(wav_corpus["entries"][3]["path"], max_duration=None)
    assert rel_to_scale(got.numpy(), oracle_features(wav_corpus["entries"][3], None)) < FEATURE_REL_TOL


def test_precompute_cache_roundtrip_and_dataset(wav_corpus):
    """``precompute_dataset_features`` writes the reference's ``.pt`` format (scripts/precompute_features.py:98-101,134-142);
    ``FSCIntentDataset`` reads it (scripts/dataset.py:44-56) and serves items (:78-115) and device batches."""
    out_dir = os.path.join(wav_corpus["root"], "cached_features")
    cache = pre.precompute_dataset_features(wav_corpus["csv"], out_dir, wav_corpus["label_map_path"], batch_size=3)
    assert cache == os.path.join(out_dir, "train_data_features.pt") and os.path.exists(cache)
    blob = torch.load(cache)
    good = wav_corpus["entries"][:4]
    assert sorted(blob) == sorted(e["path"] for e in good)                      # the short and the missing file are dropped
    for e in good:
        item = blob[e["path"]]
        assert set(item) == {"features", "label"} and item["label"] == e["label"]
        f = item["features"]
        assert f.device.type == "cpu" and f.dtype == torch.float32
        want = oracle_features(e)
        assert tuple(f.shape) == want.shape and rel_to_scale(f.numpy(), want) < FEATURE_REL_TOL, e["path"]

    ds = dataset.FSCIntentDataset(wav_corpus["csv"], wav_corpus["label_map_path"], is_training=False, cache_dir=out_dir)
    assert len(ds) == 6 and len(ds.features_dict) == 4
    for i, e in enumerate(good):
        x, y = ds[i]
        assert x.device.type == "cpu" and tuple(x.shape) == (64, 200) and y == wav_corpus["label_map"][e["label"]]
        assert rel_to_scale(x.numpy(), logmel_np.pad_or_trim(oracle_features(e), 200)) < FEATURE_REL_TOL
    x, y = ds[4]                                              # cache miss + extraction failure -> zeros (dataset.py:123,156-158)
    assert tuple(x.shape) == (64, 200) and not x.any() and y == wav_corpus["label_map"]["increase_volume"]

    # training items: the reference's host RNG order (np.random gate, then torch.rand inside mask_along_axis)
    tr = dataset.FSCIntentDataset(wav_corpus["csv"], wav_corpus["label_map_path"], is_training=True, augment_prob=1.0,
                                  cache_dir=out_dir)
    masked = 0
    for i, e in enumerate(good):
        np.random.seed(5 + i)
        torch.manual_seed(5 + i)
        x, _ = tr[i]
        np.random.seed(5 + i)
        torch.manual_seed(5 + i)
        base = oracle_features(e)
        assert np.random.random() < 1.0                       # the augment gate of dataset.py:105 consumes one draw
        params = augment.draw_mask_params(64, base.shape[1], 20, 10, gate=np.random.random)
        want = logmel_np.pad_or_trim(logmel_np.apply_masks(base, params), 200)
        assert rel_to_scale(x.numpy(), want) < FEATURE_REL_TOL and np.array_equal(x.numpy() == 0, want == 0), i
        masked += int(params[1] > params[0]) + int(params[3] > params[2])
    assert masked >= 2                                        # some bands were actually drawn

    # HBM-resident batches: same values as the per-item path, labels from the map, masks from the device sampler
    sub = dataset.FSCIntentDataset(wav_corpus["csv"], wav_corpus["label_map_path"], is_training=False, cache_dir=out_dir)
    sub.data = sub.data.iloc[:4].reset_index(drop=True)
    feats, frames, labels = sub.make_resident()
    assert tuple(feats.shape) == (4, 64, 200) and frames.cpu().tolist() == [oracle_features(e).shape[1] for e in good]
    idx = torch.tensor([3, 0, 2, 2], device="cuda")
    xb, yb = sub.get_batch(idx)
    assert xb.is_cuda and tuple(xb.shape) == (4, 64, 200)
    for k, i in enumerate(idx.cpu().tolist()):
        assert torch.equal(xb[k].cpu(), sub[i][0]) and int(yb[k]) == sub[i][1]
    trb = dataset.FSCIntentDataset(wav_corpus["csv"], wav_corpus["label_map_path"], is_training=True, augment_prob=0.7,
                                   cache_dir=out_dir)
    trb.data = trb.data.iloc[:4].reset_index(drop=True)
    xa, _ = trb.get_batch(idx, seed=11, epoch=3)
    fr = np.asarray([oracle_features(e).shape[1] for e in good], np.int32)
    all_masks = logmel_np.sample_masks(11, 3 * 4, 4, 64, fr, augment_prob=0.7)   # Philox counter = epoch * N + index
    for k, i in enumerate(idx.cpu().tolist()):
        want = logmel_np.pad_or_trim(logmel_np.apply_masks(oracle_features(good[i]), all_masks[i]), 200)
        assert rel_to_scale(xa[k].cpu().numpy(), want) < FEATURE_REL_TOL, (k, i)
        assert np.array_equal(xa[k].cpu().numpy() == 0, want == 0), (k, i)
    assert torch.equal(xa[2], xa[3])                          # same sample, same epoch -> same mask however it is batched
    xa2, _ = trb.get_batch(idx, seed=11, epoch=4)
    assert not torch.equal(xa, xa2)                           # a new epoch draws new masks


def test_apply_spec_augmentation_matches_reference_draw_order():
    """scripts/augment.py:137-165: seeded host RNG -> the same bands the reference would zero."""
    import random
    base = logmel_np.extract_features(synth.speech_like(3, 1, 48000)[0])
    hits = 0
    for seed in range(12):
        random.seed(seed)
        torch.manual_seed(seed)
        got = augment.apply_spec_augmentation(torch.from_numpy(base))
        random.seed(seed)
        torch.manual_seed(seed)
        params = augment.draw_mask_params(64, base.shape[1], 20, 10)
        want = logmel_np.apply_masks(base, params)
        assert got.device.type == "cpu" and np.array_equal(got.numpy(), want), seed
        hits += int(params[1] > params[0]) + int(params[3] > params[2])
    assert hits >= 6


def test_predict_from_file_matches_oracle(wav_corpus):
    """``test_model.predict(model, path, label_map)`` (scripts/test_model.py:106-143): un-truncated features -> pad/trim to
    200 -> forward -> softmax / arg-max / top-3, on 16 kHz and 24 kHz files; None on a bad file."""
    from oracle import eval_np
    model = cuda_model()
    label_map = {f"intent_{i}": i for i in range(31)}
    sd = synth.make_weights(1234)
    for e in (wav_corpus["entries"][0], wav_corpus["entries"][2], wav_corpus["entries"][3]):
        res = test_model.predict(model, e["path"], label_map)
        feat = logmel_np.pad_or_trim(oracle_features(e, max_duration=None), 200)
        want = classifier_np.forward(feat[None], sd)
        pred, conf, idx, prob = eval_np.predict(want, 3)
        top = np.sort(want[0])
        if top[-1] - top[-2] > 4 * LOGIT_ABS_TOL:
            assert res["predicted_label"] == f"intent_{int(pred[0])}", e["path"]
            assert abs(res["confidence"] - float(conf[0])) < 2e-3
        assert len(res["top_predictions"]) == 3
        assert abs(sum(p["probability"] for p in res["top_predictions"]) - float(prob[0].sum())) < 3e-3
    assert test_model.predict(model, wav_corpus["entries"][5]["path"], label_map) is None
    assert test_model.predict(model, wav_corpus["entries"][4]["path"], label_map) is None
