"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Run on the B200 box: ``python -m pytest tests -m gpu``.  Bars: features within 1e-4 of the tensor scale,
logits within 1e-3 absolute, identical argmax, integer outputs (mask parameters) bit-exact.
"""
import importlib

import numpy as np
import pytest
import torch

from oracle import classifier_np, logmel_np
from tests.util import FEATURE_REL_TOL, LOGIT_ABS_TOL, golden, golden_waves, rel_to_scale, synth

pytestmark = pytest.mark.gpu
native = importlib.import_module("speech-intent-recognizer_b200._native")


@pytest.fixture(scope="module")
def fe():
    return native.Frontend()


@pytest.fixture(scope="module")
def model():
    m = native.Model(31, 64)
    m.load_weights(torch.from_numpy(synth.flatten_weights(synth.make_weights(1234))))
    return m


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def test_library_is_the_cuda_extension():
    lib = native.load_library()
    assert lib.sir_version() >= 100
    before = native.launch_count()
    native.amplitude_to_db(torch.ones(8, device="cuda"))
    assert native.launch_count() == before + 1


def test_frontend_matches_golden_ragged_batch(fe):
    g, waves, lengths, noise = golden_waves()
    out = fe.forward(dev(waves), lengths=dev(np.asarray(lengths, np.int32)), max_samples=80000, out_frames=200)
    out = out.cpu().numpy()
    for i, n in enumerate(lengths):
        want = g[f"feat_{i}"]
        T = want.shape[1]
        assert rel_to_scale(out[i, :, :T], want) < FEATURE_REL_TOL, i
        assert not out[i, :, T:].any()                      # zero padding, scripts/dataset.py:111-113
    single = fe.forward(dev(noise)).cpu().numpy()           # one utterance, unpadded T = 94
    assert single.shape == (1, 64, 94)
    assert rel_to_scale(single[0], g["feat_noise"]) < FEATURE_REL_TOL


def test_frontend_stages_and_untruncated_twin(fe):
    g, waves, lengths, _ = golden_waves()
    w0 = dev(waves[0:1, :lengths[0]].copy())
    p = fe.forward(w0, mode=native.OUT_MEL_POWER).cpu().numpy()[0]
    assert rel_to_scale(p, g["mel_power_0"]) < 1e-5
    d = fe.forward(w0, mode=native.OUT_MEL_DB).cpu().numpy()[0]
    assert np.max(np.abs(d - g["mel_db_0"])) < 1e-2         # dB; |dB| up to 100 -> 1e-4 of scale
    d2 = native.amplitude_to_db(dev(g["mel_power_0"])).cpu().numpy()
    assert np.max(np.abs(d2 - g["mel_db_0"])) < 1e-4
    full = fe.forward(dev(waves[2:3, :90000].copy())).cpu().numpy()[0]      # scripts/test_model.py twin: no 5 s cut
    assert full.shape == (64, 176) and rel_to_scale(full, g["feat_2_untruncated"]) < FEATURE_REL_TOL


def test_frontend_trim_keeps_statistics_of_all_frames(fe):
    """Long audio: normalise over all T frames, THEN trim (scripts/test_model.py:94,116)."""
    g, waves, lengths, _ = golden_waves()
    out = fe.forward(dev(waves[2:3, :90000].copy()), out_frames=100).cpu().numpy()[0]
    assert rel_to_scale(out, g["feat_2_untruncated"][:, :100]) < FEATURE_REL_TOL


def test_frontend_edge_cases(fe):
    g, waves, lengths, _ = golden_waves()
    sil = fe.forward(torch.zeros(1, 16000, device="cuda")).cpu().numpy()[0]
    assert sil.shape == g["feat_silence"].shape and not sil.any()
    # too-short utterances (<= 512 samples: reflect padding undefined, the reference returns None/zeros)
    w = torch.randn(3, 2048, device="cuda") * 0.1
    lens = torch.tensor([2048, 512, 0], dtype=torch.int32, device="cuda")
    status = torch.full((3,), -1, dtype=torch.int32, device="cuda")
    out = fe.forward(w, lengths=lens, out_frames=8, status=status).cpu().numpy()
    assert status.cpu().tolist() == [0, 1, 1]
    assert out[0].any() and not out[1].any() and not out[2].any()
    want = logmel_np.pad_or_trim(logmel_np.extract_features(w[0].cpu().numpy()), 8)
    assert rel_to_scale(out[0], want) < FEATURE_REL_TOL
    # strided rows + unaligned row starts take the scalar staging path
    big = torch.zeros(2, 48003, device="cuda")
    big[:, 1:48001] = dev(waves[:2, :48000].copy())
    view = big[:, 1:48001]
    out = fe.forward(view).cpu().numpy()
    for i in range(2):
        assert rel_to_scale(out[i], logmel_np.extract_features(waves[i, :48000])) < FEATURE_REL_TOL
    # 80-mel frontend (config 5)
    fe80 = native.Frontend(n_mels=80)
    f80 = fe80.forward(dev(waves[0:1, :32000].copy())).cpu().numpy()[0]
    assert rel_to_scale(f80, g["feat80_0"]) < FEATURE_REL_TOL
    with pytest.raises(native.NativeError):
        native.Frontend(n_fft=512, hop_length=256)


@pytest.mark.parametrize("batch,samples", [(1, 48000), (37, 24000), (1200, 4096)])
def test_frontend_cluster_shapes_against_oracle(fe, batch, samples):
    """Cluster size 8 / 4 / 1 launches (small, medium, large batch) agree with the restatement."""
    w = synth.white_noise(100 + batch, batch, samples) * np.linspace(0.01, 1.0, batch, dtype=np.float32)[:, None]
    out = fe.forward(dev(w), out_frames=64).cpu().numpy()
    for i in sorted({0, batch // 2, batch - 1}):
        want = logmel_np.pad_or_trim(logmel_np.extract_features(w[i]), 64)
        assert rel_to_scale(out[i], want) < FEATURE_REL_TOL, (batch, i)


def test_specaugment_sampler_bit_exact():
    frames = np.asarray([94, 40, 157, 10, 200, 94, 94, 33], np.int32)
    for prob in (1.0, 0.7):
        got = native.specaugment_sample(1234567890123, 1000, 8, 64, 0, frames=dev(frames), augment_prob=prob)
        want = logmel_np.sample_masks(1234567890123, 1000, 8, 64, frames, augment_prob=prob)
        assert np.array_equal(got.cpu().numpy(), want)
    got = native.specaugment_sample(7, 0, 4096, 64, 94).cpu().numpy()
    assert np.array_equal(got, logmel_np.sample_masks(7, 0, 4096, 64, 94))
    tw, fw = got[:, 1] - got[:, 0], got[:, 3] - got[:, 2]
    assert tw.max() == 19 and fw.max() == 9 and 0.4 < (tw > 0).mean() < 0.55


def test_fused_masks_match_reference_semantics(fe):
    g, waves, lengths, _ = golden_waves()
    ga = golden("augment")
    w0 = dev(waves[0:1, :lengths[0]].copy()).repeat(len(ga["uniforms"]), 1)
    params = np.stack([logmel_np.sample_mask_params(u, 64, 94) for u in ga["uniforms"]])
    out = fe.forward(w0, masks=dev(params)).cpu().numpy()
    for k in range(len(params)):
        want = ga["outputs"][k]
        assert np.array_equal(out[k] == 0, want == 0) or rel_to_scale(out[k], want) < FEATURE_REL_TOL
        assert rel_to_scale(out[k], want) < FEATURE_REL_TOL
    # the same masks applied to cached features + pad to 200 (dataset.__getitem__ tail on a cache hit)
    base = dev(ga["base"][None].repeat(len(params), 0))
    fin = native.features_finalize(base, 200, masks=dev(params)).cpu().numpy()
    for k in range(len(params)):
        assert np.array_equal(fin[k], logmel_np.pad_or_trim(ga["outputs"][k], 200))


def test_classifier_matches_golden(model):
    g = golden("classifier")
    y = model.forward(dev(g["x"])).cpu().numpy()
    assert np.max(np.abs(y - g["logits"])) < LOGIT_ABS_TOL
    assert np.array_equal(y.argmax(1), g["logits"].argmax(1))
    yv = model.forward(dev(g["x_var"][None])).cpu().numpy()            # variable T = 40, no padding (config 1)
    assert np.max(np.abs(yv - g["logits_var"])) < LOGIT_ABS_TOL


@pytest.mark.parametrize("B,T", [(1, 16), (3, 24), (5, 72), (2, 48), (33, 136), (7, 88), (150, 56)])
def test_classifier_shapes_against_oracle(model, B, T):
    """Frame counts and batch sizes around the kernels' tile shapes: conv1's pooled patches (16 x 8 pooled pixels: T / 2 on and
    off multiples of 8), conv2 / conv3 tiles with ragged edges, GRU slices of 16 and of 64 utterances, 2 to 17 time steps."""
    rng = np.random.default_rng(100 * B + T)
    x = rng.standard_normal((B, 64, T)).astype(np.float32)
    x[:, :, T - T // 4:] = 0.0                                          # a padded tail, as the dataset pads
    got = model.forward(dev(x)).cpu().numpy()
    n = min(B, 6)                                                       # the numpy oracle on a few rows, first and last
    rows = np.r_[0:n // 2 + n % 2, B - n // 2:B]
    want = classifier_np.forward(x[rows], synth.make_weights(1234))
    assert np.max(np.abs(got[rows] - want)) < LOGIT_ABS_TOL
    assert np.isfinite(got).all()


def test_pipeline_matches_oracle_and_argmax(fe, model):
    """waveform -> logits: config-2 shape at a size the oracle finishes in seconds."""
    B = 24
    w = synth.speech_like(21, B)
    logits, feats = model.pipeline(fe, dev(w), max_samples=80000, out_frames=200)
    logits, feats = logits.cpu().numpy(), feats.cpu().numpy()
    want_f = np.stack([logmel_np.dataset_item(x) for x in w])
    assert rel_to_scale(feats, want_f) < FEATURE_REL_TOL
    want = classifier_np.forward(want_f, synth.make_weights(1234))
    assert np.max(np.abs(logits - want)) < LOGIT_ABS_TOL
    top = np.sort(want, axis=1)
    safe = (top[:, -1] - top[:, -2]) > 4 * LOGIT_ABS_TOL
    assert safe.sum() >= B - 2
    assert np.array_equal(logits.argmax(1)[safe], want.argmax(1)[safe])
    assert len(set(want.argmax(1).tolist())) >= 6                       # not vacuous


def test_full_size_properties(fe, model):
    """Config-2 size (256 x 3 s): size-independent properties instead of a full oracle run."""
    B = 256
    w = dev(synth.white_noise(5, 64, 48000)).repeat(4, 1) * 1.0
    logits, feats = model.pipeline(fe, w, out_frames=200)
    f = feats[:, :, :94]
    mean = f.reshape(B, -1).mean(1)
    std = f.reshape(B, -1).std(1)
    assert mean.abs().max().item() < 1e-4 and (std - 1).abs().max().item() < 1e-3     # normalised per utterance
    assert not feats[:, :, 94:].any().item()
    assert torch.equal(feats[:64], feats[192:]) and torch.equal(logits[:64], logits[192:])   # batch-position invariant
    # gain invariance: log-mel of a*x is a dB shift, removed by the normalisation
    f2 = fe.forward(w[:8] * 0.25, out_frames=200)
    assert rel_to_scale(f2.cpu().numpy(), feats[:8].cpu().numpy()) < FEATURE_REL_TOL


def test_host_pipeline_equals_device_path():
    """IntentPipeline.infer_host (H2D overlapped with frontend + conv stack, staged forward) == the plain path."""
    pre = importlib.import_module("speech-intent-recognizer_b200.scripts.precompute_features")
    models = importlib.import_module("speech-intent-recognizer_b200.models.models")
    pipeline = importlib.import_module("speech-intent-recognizer_b200.pipeline")
    ex = pre.AudioFeatureExtractor()
    m = models.CNNAudioGRU(31)
    sd = synth.make_weights(1234)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    m = m.cuda().eval()
    for B, subs in ((37, 4), (256, 4), (5, 8), (400, 4)):
        w = torch.from_numpy(synth.white_noise(B, B, 16000) * np.linspace(0.05, 1, B, dtype=np.float32)[:, None]).pin_memory()
        want = m(ex.extract_batch(w.cuda(), out_frames=200)).cpu()
        got = pipeline.IntentPipeline(ex, m, sub_batches=subs).infer_host(w)
        assert got.is_pinned() and torch.equal(got, want), (B, float((got - want).abs().max()))
    with pytest.raises(native.NativeError):
        pipeline.IntentPipeline(ex, m).infer_host(w.cuda())
    # streaming: several batches in flight on rotating slots, results in order
    batches = [torch.from_numpy(synth.white_noise(50 + i, 24, 16000)).pin_memory() for i in range(5)]
    pipe = pipeline.IntentPipeline(ex, m, sub_batches=3, depth=2)
    for i, got in enumerate(pipe.infer_stream(batches)):
        assert torch.equal(got, m(ex.extract_batch(batches[i].cuda(), out_frames=200)).cpu()), i


def test_pcm16_ingest_is_bit_identical_to_float_path(fe, model):
    """16-bit PCM rows scaled by 1/32768 on load == torchaudio.load's normalisation followed by the fp32 path."""
    g, waves, lengths, _ = golden_waves()
    pcm = np.round(waves * 32767.0).astype(np.int16)
    as_float = (pcm.astype(np.float32) / 32768.0).astype(np.float32)          # what the reference's loader hands on
    lens = dev(np.asarray(lengths, np.int32))
    a = fe.forward(dev(pcm), lengths=lens, max_samples=80000, out_frames=200)
    b = fe.forward(dev(as_float), lengths=lens, max_samples=80000, out_frames=200)
    assert torch.equal(a, b)
    want = np.stack([logmel_np.dataset_item(as_float[i, :n]) for i, n in enumerate(lengths)])
    assert rel_to_scale(a.cpu().numpy(), want) < FEATURE_REL_TOL
    odd = torch.zeros(2, 30001, dtype=torch.int16, device="cuda")              # odd stride -> unaligned rows (scalar loads)
    odd[:, :30000] = dev(pcm[:2, :30000].copy())
    got_odd = fe.forward(odd[:, :30000]).cpu().numpy()          # the scalar-load path contracts FMAs differently: ulp-level
    assert rel_to_scale(got_odd, fe.forward(dev(as_float[:2, :30000].copy())).cpu().numpy()) < 1e-5
    pre = importlib.import_module("speech-intent-recognizer_b200.scripts.precompute_features")
    models = importlib.import_module("speech-intent-recognizer_b200.models.models")
    pipeline = importlib.import_module("speech-intent-recognizer_b200.pipeline")
    ex = pre.AudioFeatureExtractor()
    m = models.CNNAudioGRU(31)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(1234).items()}, strict=False)
    m = m.cuda().eval()
    pipe = pipeline.IntentPipeline(ex, m)
    host_pcm = torch.from_numpy(pcm[:, :48000].copy()).pin_memory()
    got = pipe.infer_host(host_pcm).clone()
    ref = pipe.infer_host(torch.from_numpy(as_float[:, :48000].copy()).pin_memory())
    assert torch.equal(got, ref)


def test_evaluation_head_matches_reference_semantics(model):
    """softmax / argmax / confidence / top-3 / accuracy / confusion counts (scripts/test_model.py:121-156, evaluate.py:79-98)."""
    from oracle import eval_np
    rng = np.random.default_rng(11)
    logits = (rng.standard_normal((300, 31)) * 4).astype(np.float32)
    logits[5] = 0.0                                   # all-equal row: argmax -> 0, top-k -> highest indices first
    logits[6, :] = -200.0
    logits[6, 7] = 30.0                               # saturated softmax: the zero-probability tail ties
    logits[7, 3] = logits[7, 9] = 12.5                # tied maximum
    labels = rng.integers(0, 31, 300).astype(np.int64)
    conf_m = torch.zeros((31, 31), dtype=torch.int64, device="cuda")
    correct = torch.zeros(1, dtype=torch.int64, device="cuda")
    pred, conf, ti, tp = native.predict(dev(logits), k=3, labels=dev(labels), confusion=conf_m, correct=correct)
    w_pred, w_conf, w_idx, w_prob = eval_np.predict(logits, 3)
    assert np.array_equal(pred.cpu().numpy(), w_pred)
    # indices must agree wherever the top-(k+1) probabilities are distinct; among exactly tied probabilities numpy's
    # default (unstable) argsort order is implementation noise - there the kernel's rule is "higher index first"
    probs = eval_np.softmax(logits)
    top4 = -np.sort(-probs, axis=1)[:, :4]
    distinct = np.all(np.diff(top4, axis=1) < 0, axis=1)
    assert distinct.sum() >= 295
    got_idx = ti.cpu().numpy()
    assert np.array_equal(got_idx[distinct], w_idx[distinct])
    assert got_idx[5].tolist() == [30, 29, 28] and got_idx[6].tolist() == [7, 30, 29] and got_idx[7].tolist()[:2] == [9, 3]
    assert np.max(np.abs(conf.cpu().numpy() - w_conf)) < 1e-6 and np.max(np.abs(tp.cpu().numpy() - w_prob)) < 1e-6
    acc, cm = eval_np.accuracy_and_confusion(w_pred, labels, 31)
    assert np.array_equal(conf_m.cpu().numpy(), cm) and int(correct) == int(round(acc * 300))
    # the mirrors of the callers
    tm = importlib.import_module("speech-intent-recognizer_b200.scripts.test_model")
    ev = importlib.import_module("speech-intent-recognizer_b200.scripts.evaluate")
    models = importlib.import_module("speech-intent-recognizer_b200.models.models")
    m = models.CNNAudioGRU(31)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(1234).items()}, strict=False)
    m = m.cuda().eval()
    g = golden("classifier")
    label_map = {f"intent_{i}": i for i in range(31)}
    res = tm.predict_batch(m, dev(g["x"]), label_map)
    want_pred, want_conf, want_idx, _ = eval_np.predict(g["logits"], 3)
    assert [r["predicted_label"] for r in res] == [f"intent_{i}" for i in want_pred]
    assert max(abs(r["confidence"] - c) for r, c in zip(res, want_conf)) < 1e-3
    assert [r["top_predictions"][0]["label"] for r in res] == [f"intent_{i}" for i in want_idx[:, 0]]
    y = torch.from_numpy(want_pred.astype(np.int64))
    y[0] = (y[0] + 1) % 31
    acc, cm = ev.evaluate_loader(m, [(torch.from_numpy(g["x"]), y), (None, None)], 31)
    assert abs(acc - (len(y) - 1) / len(y)) < 1e-9 and int(cm.sum()) == len(y) and int(cm.trace()) == len(y) - 1


def test_mfcc_preemphasis_resample_match_oracle(fe):
    """MFCC epilogue (top_db floor at the per-utterance max, ortho DCT-II), pre-emphasis and the
    sinc-Hann polyphase resampler, against the restatements pinned to torchaudio."""
    g, waves, lengths, _ = golden_waves()
    lens = dev(np.asarray(lengths, np.int32))
    out = fe.mfcc(dev(waves), n_mfcc=40, top_db=80.0, lengths=lens, max_samples=80000, out_frames=160).cpu().numpy()
    for i, n in enumerate(lengths):
        want = logmel_np.mfcc(waves[i, :min(n, 80000)])
        T = want.shape[1]
        assert rel_to_scale(out[i, :, :T], want) < FEATURE_REL_TOL, i
        assert not out[i, :, T:].any()
    one = fe.mfcc(dev(waves[1:2, :lengths[1]].copy()), n_mfcc=13, top_db=0.0).cpu().numpy()[0]      # no floor, 13 coefficients
    assert rel_to_scale(one, logmel_np.mfcc(waves[1, :lengths[1]], n_mfcc=13, top_db=None)) < FEATURE_REL_TOL
    pe = native.preemphasis(dev(waves[:3])).cpu().numpy()
    assert np.max(np.abs(pe - logmel_np.preemphasis(waves[:3]))) < 1e-7
    x = synth.speech_like(4, 3, 36001)
    lens_in = np.asarray([36001, 24000, 513], np.int32)
    for orig in (24000, 22050):
        rs = native.Resampler(orig, 16000)
        got, out_len = rs(dev(x), lengths=dev(lens_in), return_lengths=True)
        got, out_len = got.cpu().numpy(), out_len.cpu().numpy()
        for i, n in enumerate(lens_in):
            want = logmel_np.resample(x[i, :n], orig, 16000)
            assert out_len[i] == want.shape[0]
            assert np.max(np.abs(got[i, :want.shape[0]] - want)) < 2e-6 and not got[i, want.shape[0]:].any()
    # the 24 kHz TTS-clip path end to end: resample -> features (scripts/test_model.py:69-94, no 5 s truncation)
    feats = fe.forward(native.Resampler(24000, 16000)(dev(x[0:1]))).cpu().numpy()[0]
    want = logmel_np.extract_features(logmel_np.resample(x[0], 24000, 16000), max_duration=None)
    assert rel_to_scale(feats, want) < FEATURE_REL_TOL


def test_work_item_scheduling_is_deterministic_and_shape_safe(fe):
    """The frontend draws (utterance, group) items from a ticket counter and lets whichever CTA completes an utterance
    normalise it: results must not depend on the draw order (bit-identical across repeats), and one handle must serve
    launches of changing batch / length / mode back to back (ticket base, partial and counter buffers are reused)."""
    rng = np.random.default_rng(7)
    cases = []
    for B, L, F in ((3, 48000, 200), (300, 16000, 200), (1, 700, 8), (64, 80000, 157), (17, 33333, 66), (300, 16000, 200)):
        w = dev((rng.standard_normal((B, L)) * 0.1).astype(np.float32))
        lens = dev(rng.integers(600, L + 1, size=B).astype(np.int32))
        lens[0] = 100                                                    # one invalid utterance (L <= 512): zeros, status 1
        cases.append((w, lens, F))
    first = [fe.forward(w, lengths=lens, out_frames=F).clone() for w, lens, F in cases]
    for rep in range(3):
        for (w, lens, F), want in zip(cases, first):
            got = fe.forward(w, lengths=lens, out_frames=F)
            assert torch.equal(got, want), (rep, tuple(w.shape))
    for (w, lens, F), got in zip(cases, first):                          # and they are the right values
        g = got.cpu().numpy()
        assert not g[0].any()
        for i in sorted({1, w.shape[0] - 1} - {0}):
            if i >= w.shape[0]:
                continue
            n = int(lens[i])
            want = logmel_np.dataset_item(w[i, :n].cpu().numpy(), target=F)
            assert rel_to_scale(g[i], want) < FEATURE_REL_TOL, (tuple(w.shape), i)
    # MFCC goes through the same finisher (max merge + DCT): repeatable too
    a = fe.mfcc(cases[1][0], lengths=cases[1][1], out_frames=40).clone()
    assert torch.equal(a, fe.mfcc(cases[1][0], lengths=cases[1][1], out_frames=40))


def test_concurrent_streams_match_serial_results(fe, model):
    """One frontend / model handle used from three CUDA streams at once (per-stream workspaces, tile-ticket counters and
    item counters): every stream's results equal the single-stream results bit for bit."""
    rng = np.random.default_rng(11)
    waves = [dev((rng.standard_normal((B, 48000)) * 0.1).astype(np.float32)) for B in (256, 96, 256, 40, 256, 130)]
    want = []
    for w in waves:
        f = fe.forward(w, out_frames=200)
        want.append((f.clone(), model.forward(f).clone()))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    for rep in range(4):
        got = [None] * len(waves)
        for i, w in enumerate(waves):
            s = streams[(i + rep) % 3]
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                f = fe.forward(w, out_frames=200)
                got[i] = (f, model.forward(f))
        torch.cuda.synchronize()
        for i in range(len(waves)):
            assert torch.equal(got[i][0], want[i][0]), ("features", rep, i)
            assert torch.equal(got[i][1], want[i][1]), ("logits", rep, i)


def test_config5_shard_80_mels_long_audio():
    """BASELINE config 5, one GPU's shard: 512 utterances x 10 s @16 kHz, 80-mel frontend + the classifier built for 80
    mels (models/models.py:23 hard-codes 1024 = 128 * 64 / 8; the 80-mel classifier is that class with
    gru_input_size = 1280).  Full size through size-independent properties, a few utterances against the oracle."""
    B, L, n_mels = 512, 160000, 80
    base = synth.speech_like(21, 8, L)
    gains = np.linspace(0.2, 1.0, B // 8, dtype=np.float32)
    gains[-1] = gains[1]                                                # two identical groups at different batch positions
    w = dev((base[None, :, :] * gains[:, None, None]).reshape(B, L))
    fe80 = native.Frontend(n_mels=n_mels)
    feats = fe80.forward(w, out_frames=200)                             # 313 frames normalised, then trimmed to 200
    assert feats.shape == (B, n_mels, 200)
    full = fe80.forward(w[:16])                                          # untrimmed: [16, 80, 313]
    assert full.shape == (16, n_mels, 313)
    flat = full.reshape(16, -1)
    assert flat.mean(1).abs().max().item() < 1e-4 and (flat.std(1) - 1).abs().max().item() < 1e-3
    # trimming does not change the statistics (the two output widths sum their partials in different orders: ulp-level)
    assert rel_to_scale(feats[:16].cpu().numpy(), full[:, :, :200].cpu().numpy()) < 1e-6
    assert torch.equal(feats[8:16], feats[504:512])                     # position in the batch does not matter
    sd = synth.make_weights(77, 31, n_mels)
    model80 = native.Model(31, n_mels)
    model80.load_weights(torch.from_numpy(synth.flatten_weights(sd)))
    logits = model80.forward(feats)
    assert logits.shape == (B, 31) and torch.isfinite(logits).all().item()
    for i in (0, 3, 509):
        want_f = logmel_np.dataset_item(w[i].cpu().numpy(), target=200, n_mels=n_mels, max_duration=None)
        assert rel_to_scale(feats[i].cpu().numpy(), want_f) < FEATURE_REL_TOL, i
        want = classifier_np.forward(want_f[None], sd)[0]
        got = logits[i].cpu().numpy()
        assert np.max(np.abs(got - want)) < LOGIT_ABS_TOL, (i, float(np.max(np.abs(got - want))))


def test_config3_shard_precompute_properties(fe):
    """BASELINE config 3, one of eight shards: 3,756 utterances x 3 s through the frontend alone (feature precompute)."""
    B, L = 3756, 48000
    base = dev(synth.speech_like(31, 12, L))
    w = base.repeat(B // 12, 1)
    feats = fe.forward(w)                                               # [B, 64, 94], no padding
    assert feats.shape == (B, 64, 94)
    flat = feats.reshape(B, -1)
    assert flat.mean(1).abs().max().item() < 1e-4 and (flat.std(1) - 1).abs().max().item() < 1e-3
    assert torch.equal(feats[:12], feats[B - 12:])                      # position in the batch does not matter
    want = logmel_np.extract_features(base[5].cpu().numpy())
    assert rel_to_scale(feats[5].cpu().numpy(), want) < FEATURE_REL_TOL


def test_config1_mic_recording_durations_batch1_and_ragged(fe, model):
    """BASELINE config 1: 95 clips with the durations of the reference's mic_recordings (1.27-3.36 s at 16 kHz), one
    utterance per call with its own T (scripts/test_model.py:106-139: features -> pad to 200 -> model -> arg-max),
    against the same clips as ONE ragged batch and, for a sample of them, against the oracle."""
    lens = synth.config1_lengths()
    assert len(lens) == 95 and lens.min() == 20352 and lens.max() == 53760
    waves = synth.speech_like(41, len(lens), int(lens.max()), lengths=lens)
    d_w = dev(waves)
    feats = fe.forward(d_w, lengths=dev(lens.astype(np.int32)), out_frames=200)        # ragged batch
    logits = model.forward(feats)
    sd = synth.make_weights(1234)
    checked = 0
    for i in range(len(lens)):
        n = int(lens[i])
        f1 = fe.forward(d_w[i:i + 1, :n].contiguous())                                 # batch 1, unpadded [1, 64, T_i]
        T = 1 + n // 512
        assert f1.shape == (1, 64, T)
        # a batch of one and the ragged batch agree to rounding (different output widths sum partials differently)
        assert rel_to_scale(f1[0].cpu().numpy(), feats[i, :, :T].cpu().numpy()) < 1e-6, i
        assert not feats[i, :, T:].any().item()
        l1 = model.forward(torch.nn.functional.pad(f1, (0, 200 - T)))
        assert float((l1[0] - logits[i]).abs().max()) < 1e-4, i
        if i % 12 == 0:                                                                 # 8 clips against the oracle
            want_f = logmel_np.dataset_item(waves[i, :n], target=200, max_duration=None)
            assert rel_to_scale(feats[i].cpu().numpy(), want_f) < FEATURE_REL_TOL, i
            want = classifier_np.forward(want_f[None], sd)[0]
            got = logits[i].cpu().numpy()
            assert np.max(np.abs(got - want)) < LOGIT_ABS_TOL, (i, float(np.max(np.abs(got - want))))
            top = np.sort(want)
            if top[-1] - top[-2] > 4 * LOGIT_ABS_TOL:
                assert int(got.argmax()) == int(want.argmax()), i
            checked += 1
    assert checked == 8


def test_frontend_work_item_boundaries(fe):
    """Utterance lengths around the frontend's 14-frame work items and their two 7-frame stage-2 tiles: T = 1 + L // 512 of
    2, 7, 8, 13, 14, 15, 21, 27, 28, 29, 43 frames, with L on, just below and just above a hop boundary (the last block of an
    utterance goes through the reflect-padding load path), one batch of ragged lengths, fp32 and PCM16, against the oracle;
    plus two pathological utterances whose loud burst ends a few samples into an otherwise near-silent frame (the case that
    broke an fp16 split scaled by the raw maximum): a broadband burst (bar 1e-4) and a PURE TONE 110 dB above the noise floor
    of its own frames - there the bands far from the tone sit at the rounding floor of any fp32 transform (the fp32 oracle is
    5e-5 of the scale away from a float64 evaluation, the three-pass split 3e-4: tools/tc_dft_emulate.py), bar 1e-3."""
    rng = np.random.default_rng(314)
    frames = (2, 7, 8, 13, 14, 15, 21, 27, 28, 29, 43)
    lengths = []
    for T in frames:
        lengths += [max(512 * (T - 1), 513), 512 * (T - 1) + 511, 512 * (T - 1) + int(rng.integers(1, 511))]
    lengths = np.asarray(lengths, np.int32)
    Lmax = int(lengths.max())
    waves = synth.speech_like(77, len(lengths), Lmax)
    # loud bursts that end a few samples into a frame whose remainder is near-silent
    tone = 8
    for i in (5, tone):
        waves[i, :] = 1e-5 * rng.standard_normal(Lmax).astype(np.float32)
    waves[5, 1000:2055] += 0.3 * rng.standard_normal(1055).astype(np.float32)
    waves[tone, 1000:2055] += 0.4 * np.sin(0.3 * np.arange(1055, dtype=np.float32))
    out = fe.forward(dev(waves), lengths=dev(lengths), out_frames=48).cpu().numpy()
    pcm = np.clip(np.round(waves * 32767.0), -32768, 32767).astype(np.int16)
    out16 = fe.forward(dev(pcm), lengths=dev(lengths), out_frames=48).cpu().numpy()
    for i, n in enumerate(lengths):
        T = 1 + int(n) // 512
        want = logmel_np.extract_features(waves[i, :n])
        assert want.shape == (64, T)
        bar = 1e-3 if i == tone else FEATURE_REL_TOL
        assert rel_to_scale(out[i, :, :T], want) < bar, (i, int(n), T)
        assert not out[i, :, T:].any()
        want16 = logmel_np.extract_features(pcm[i, :n].astype(np.float32) / 32768.0)
        assert rel_to_scale(out16[i, :, :T], want16) < bar, (i, int(n), T, "pcm16")
