"""CPU tests: the oracle restatements against golden vectors produced by the reference itself
(tests/golden/make_golden.py imported /root/reference unmodified)."""
import collections

import numpy as np
import pytest

from oracle import classifier_np, logmel_np
from tests.util import FEATURE_REL_TOL, LOGIT_ABS_TOL, golden, golden_waves, rel_to_scale, synth


def test_constants_match_reference():
    g = golden("frontend")
    assert np.max(np.abs(logmel_np.hann_window() - g["window"])) < 5e-7   # torch builds it in fp32
    fb = logmel_np.melscale_fbanks()
    assert fb.shape == (513, 64)
    assert np.max(np.abs(fb - g["fb"])) < 1e-5     # torch evaluates the triangles in fp32; ours is exact
    # 97 % sparse: every frequency bin feeds at most two adjacent mels
    assert int((fb > 0).sum(axis=1).max()) <= 2


def test_frontend_stages_match_reference():
    g, waves, lengths, _ = golden_waves()
    p = logmel_np.mel_power(waves[0, :lengths[0]])
    assert rel_to_scale(p, g["mel_power_0"]) < 1e-5
    d = logmel_np.amplitude_to_db(g["mel_power_0"])
    assert np.max(np.abs(d - g["mel_db_0"])) < 1e-4          # dB units, |dB| up to 100


@pytest.mark.parametrize("i", range(6))
def test_features_match_reference(i):
    g, waves, lengths, _ = golden_waves()
    f = logmel_np.extract_features(waves[i, :lengths[i]])
    want = g[f"feat_{i}"]
    assert f.shape == want.shape == (64, 1 + min(lengths[i], 80000) // 512)
    assert rel_to_scale(f, want) < FEATURE_REL_TOL


def test_features_edge_cases():
    g, waves, lengths, noise = golden_waves()
    assert rel_to_scale(logmel_np.extract_features(noise[0]), g["feat_noise"]) < FEATURE_REL_TOL
    f = logmel_np.extract_features(waves[2, :90000], max_duration=None)
    assert f.shape == (64, 176) and rel_to_scale(f, g["feat_2_untruncated"]) < FEATURE_REL_TOL
    sil = logmel_np.extract_features(np.zeros(16000, np.float32))
    assert np.array_equal(sil, g["feat_silence"]) and not sil.any()
    f80 = logmel_np.extract_features(waves[0, :32000], n_mels=80)
    assert f80.shape == (80, 63) and rel_to_scale(f80, g["feat80_0"]) < FEATURE_REL_TOL
    with pytest.raises(ValueError):
        logmel_np.mel_power(np.zeros(512, np.float32))
    # stereo input is averaged to mono first
    st = np.stack([waves[0, :48000], waves[1, :48000]])
    assert rel_to_scale(logmel_np.extract_features(st), logmel_np.extract_features(st.mean(0))) < 1e-6


def test_fp64_oracle_bounds_fp32():
    """The fp32 path (reference and restatement) sits within 1e-4 of the exact (float64) answer."""
    g, waves, lengths, _ = golden_waves()
    f64 = logmel_np.extract_features(waves[0, :lengths[0]].astype(np.float64), dtype=np.float64)
    assert rel_to_scale(g["feat_0"], f64) < FEATURE_REL_TOL


def test_specaugment_matches_reference():
    g = golden("augment")
    base = g["base"]
    seen_t = seen_f = 0
    for u, want in zip(g["uniforms"], g["outputs"]):
        params = logmel_np.sample_mask_params(u, base.shape[0], base.shape[1])
        assert np.array_equal(logmel_np.apply_masks(base, params), want)
        seen_t += params[1] > params[0]
        seen_f += params[3] > params[2]
    assert seen_t >= 2 and seen_f >= 2          # the fixture exercises both mask kinds


def test_mask_sampler_distribution():
    rng = np.random.default_rng(5)
    u = rng.random((20000, 6), dtype=np.float32)
    p = np.stack([logmel_np.sample_mask_params(r, 64, 94) for r in u])
    tw, fw = p[:, 1] - p[:, 0], p[:, 3] - p[:, 2]
    assert tw.max() == 19 and fw.max() == 9 and tw.min() == 0        # widths 0..param-1
    assert (p[:, 1] <= 94).all() and (p[:, 3] <= 64).all()
    assert abs((tw > 0).mean() - 0.5 * 19 / 20) < 0.02               # gate 0.5, width 0 w.p. 1/20


def test_pad_trim_collate():
    a = np.ones((64, 94), np.float32)
    assert logmel_np.pad_or_trim(a).shape == (64, 200) and logmel_np.pad_or_trim(a)[:, 94:].sum() == 0
    assert logmel_np.pad_or_trim(np.ones((64, 313), np.float32)).shape == (64, 200)
    mel, lab = logmel_np.collate([(a, 3), (None, 1), (np.zeros((64, 0), np.float32), 2), (np.ones((64, 250)), 4)])
    assert mel.shape == (2, 64, 200) and lab.tolist() == [3, 4] and lab.dtype == np.int64
    assert logmel_np.collate([(None, 0)]) == (None, None)


def test_classifier_matches_reference():
    g = golden("classifier")
    sd = synth.make_weights(int(g["weight_seed"]))
    y = classifier_np.forward(g["x"], sd)
    assert y.shape == (4, 31)
    assert np.max(np.abs(y - g["logits"])) < 2e-4
    assert np.array_equal(y.argmax(1), g["logits"].argmax(1))
    yv = classifier_np.forward(g["x_var"][None, None], sd)           # 4-D, variable T, no padding
    assert np.max(np.abs(yv - g["logits_var"])) < 2e-4


def test_torch_port_matches_reference():
    import torch
    from oracle.torch_port import ClassifierPort, FeaturePort, load_numpy_state
    g, waves, lengths, _ = golden_waves()
    fp = FeaturePort()
    for i in range(6):
        f = fp.one(torch.from_numpy(waves[i:i + 1, :lengths[i]])).numpy()
        assert rel_to_scale(f, g[f"feat_{i}"]) < 1e-6
    gc = golden("classifier")
    m = load_numpy_state(ClassifierPort(31).eval(), synth.make_weights(int(gc["weight_seed"])))
    with torch.no_grad():
        y = m(torch.from_numpy(gc["x"])).numpy()
    assert np.max(np.abs(y - gc["logits"])) < 1e-4
    assert set(m.state_dict().keys()) - {f"bn{i}.num_batches_tracked" for i in (1, 2, 3)} == \
        {k for k, _ in synth.state_dict_spec()}


def test_weights_are_discriminative():
    """'identical argmax' must not be vacuous (SURVEY.md section 7 hard part 4)."""
    waves = synth.speech_like(7, 48)
    feats = np.stack([logmel_np.dataset_item(w) for w in waves])
    y = classifier_np.forward(feats, synth.make_weights(1234))
    counts = collections.Counter(y.argmax(1).tolist())
    top = np.sort(y, axis=1)
    assert len(counts) >= 12 and max(counts.values()) <= 20
    assert np.median(top[:, -1] - top[:, -2]) > 0.2


def test_mfcc_preemphasis_resample_restatements_match_torchaudio():
    """The section-8(f) rows have no counterpart inside the reference tree; their oracle is pinned by a live differential
    test against the torchaudio the reference builds on (MFCC, preemphasis, Resample)."""
    import torch
    import torchaudio
    w = synth.speech_like(3, 2, 30000)
    for n_mels, n_mfcc in ((64, 40), (80, 13)):
        m = torchaudio.transforms.MFCC(sample_rate=16000, n_mfcc=n_mfcc, melkwargs=dict(n_fft=1024, hop_length=512, n_mels=n_mels))
        want = m(torch.from_numpy(w[0:1]))[0].numpy()
        assert rel_to_scale(logmel_np.mfcc(w[0], n_mfcc=n_mfcc, n_mels=n_mels), want) < 1e-5
    assert np.array_equal(logmel_np.preemphasis(w), torchaudio.functional.preemphasis(torch.from_numpy(w), 0.97).numpy())
    x = synth.speech_like(4, 1, 36001)[0]
    for orig in (24000, 22050, 8000):
        want = torchaudio.transforms.Resample(orig, 16000)(torch.from_numpy(x)).numpy()
        got = logmel_np.resample(x, orig, 16000)
        assert got.shape == want.shape and np.max(np.abs(got - want)) < 1e-5, orig
