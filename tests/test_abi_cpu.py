"""CPU tests of the boundary: the C-ABI library builds, loads and exports exactly what include/sir_b200.h
declares; host-side logic that needs no GPU."""
import ctypes
import importlib
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from tests.util import ROOT, synth

native = importlib.import_module("speech-intent-recognizer_b200._native")
builder = importlib.import_module("speech-intent-recognizer_b200.build")


@pytest.fixture(scope="module")
def lib():
    builder.build_library()
    return native.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sir_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sir_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sir_b200.h but not exported"
    assert sorted(native.SIGNATURES) == declared, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (sir_[a-z0-9_]+)", out)))
    assert exported == declared


def test_library_targets_sm100a(lib):
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.sir_version() >= 100 and lib.sir_launch_count() == 0
    with pytest.raises(native.NativeError):
        native.Frontend()
    with pytest.raises(native.NativeError):
        native.Model(31, 64)
    with pytest.raises(native.NativeError):
        native.amplitude_to_db(torch.ones(4))
    h = ctypes.c_void_p()
    assert lib.sir_frontend_create(ctypes.byref(h), 16000, 64, 512, 256) == -3        # unsupported n_fft/hop
    assert b"n_fft" in lib.sir_last_error()
    assert lib.sir_model_create(ctypes.byref(h), 31, 60) == -3


def test_gemm_tile_width_fills_whole_waves(lib):
    """The persistent GEMM's tile width per shape: least rounds x width over 148 SMs (the GRU projections: 12 weight tiles)."""
    assert native.gemm_tile_width(6400, 1536) == 176        # 256 utt x 200 frames: 37 x 12 = 444 tiles = exactly 3 rounds
    assert native.gemm_tile_width(9472, 1536) == 256        # 256 utt x 296 frames: 37 x 12 = 444 tiles again, 176 would need 5 rounds
    assert native.gemm_tile_width(7690, 1536) == 208
    assert native.gemm_tile_width(592, 1536) == 160         # training batch 16: one round either way, narrowest tile
    assert native.gemm_tile_width(100, 1536) == 0           # too few tiles for the persistent kernel
    for M in (353, 1000, 4100, 12345, 86016):
        w = native.gemm_tile_width(M, 1536)
        rounds = lambda bn: -(-12 * -(-M // bn) // 148) * bn
        assert w in (160, 176, 208, 256) and rounds(w) == min(rounds(b) for b in (160, 176, 208, 256))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "speech-intent-recognizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{fn} imports the oracle"


def test_model_mirror_state_dict_and_flat_order():
    models = importlib.import_module("speech-intent-recognizer_b200.models.models")
    m = models.CNNAudioGRU(31)
    keys = [k for k in m.state_dict() if "num_batches_tracked" not in k]
    assert set(keys) == {k for k, _ in synth.state_dict_spec()}
    assert sum(p.numel() for p in m.parameters()) == 3261184
    sd = synth.make_weights(1234)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert np.array_equal(m._flat_weights().numpy(), synth.flatten_weights(sd))
    with pytest.raises(native.NativeError):
        m.train()(torch.zeros(1, 64, 200))


def test_collate_fn_mirror():
    train = importlib.import_module("speech-intent-recognizer_b200.scripts.train")
    a = torch.ones(64, 94)
    mel, lab = train.collate_fn([(a, 3), (None, 1), (torch.zeros(64, 0), 2), (torch.ones(64, 250), 4)])
    assert mel.shape == (2, 64, 200) and lab.tolist() == [3, 4] and lab.dtype == torch.long
    assert mel[0, :, 94:].sum() == 0
    assert train.collate_fn([(None, 0)]) == (None, None)


def test_host_rng_order_matches_reference_masks():
    """draw_mask_params consumes the host generators in the reference's order (golden made by the reference)."""
    from oracle import logmel_np
    from tests.util import golden
    aug = importlib.import_module("speech-intent-recognizer_b200.scripts.augment")
    g = golden("augment")
    for seed, u in zip(g["seeds"], g["uniforms"]):
        np.random.seed(int(seed))
        torch.manual_seed(int(seed))
        got = aug.draw_mask_params(64, 94, gate=np.random.random)
        assert got == logmel_np.sample_mask_params(u, 64, 94).tolist()


def test_wav_reader_roundtrip(tmp_path):
    io = importlib.import_module("speech-intent-recognizer_b200.utils.audio_io")
    w = synth.white_noise(3, 2, 4000)
    p = str(tmp_path / "x.wav")
    io.write_wav_pcm16(p, w, 22050)
    got, sr = io._read_riff_wav(p)
    assert sr == 22050 and got.shape == (2, 4000)
    assert np.max(np.abs(got.numpy() - w)) <= 1.0 / 32768 + 1e-7
    (tmp_path / "bad.wav").write_bytes(b"\xff\xf3\x84\xc4" + b"\0" * 64)             # MP3 frame header, like the reference clips
    with pytest.raises(ValueError):
        io._read_riff_wav(str(tmp_path / "bad.wav"))


def test_host_binding_helper_is_optional():
    """bind_host_to_gpu is an optimisation for multi-socket hosts: without NVML / a GPU it reports None and leaves the
    process affinity alone (bench.py and IntentPipeline users carry on unbound)."""
    import importlib
    import os
    pipeline = importlib.import_module("speech-intent-recognizer_b200.pipeline")
    before = os.sched_getaffinity(0)
    got = pipeline.bind_host_to_gpu(0)
    assert got is None or (isinstance(got, list) and set(got) <= before)
    if got is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)


def test_weight_version_token_sees_every_kind_of_change():
    """CNNAudioGRU re-uploads its weights to the native handle when `_versions()` changes: in-place writes (optimizer steps,
    load_state_dict), buffer updates (BatchNorm statistics) and replaced parameter objects must all change it; reading must not."""
    models = importlib.import_module("speech-intent-recognizer_b200.models.models")
    m = models.CNNAudioGRU(31)
    v0 = m._versions()
    assert m._versions() == v0 and len(v0) == len(list(m.parameters())) + len(list(m.buffers()))
    _ = m.state_dict()                                                   # reading does not change the token
    assert m._versions() == v0
    with torch.no_grad():
        m.fc.bias.add_(1.0)                                              # in-place write to a parameter
    v1 = m._versions()
    assert v1 != v0
    m.bn2.running_mean.mul_(0.5)                                         # in-place write to a buffer
    v2 = m._versions()
    assert v2 != v1
    m.load_state_dict(m.state_dict())                                    # copy_ into every entry
    v3 = m._versions()
    assert v3 != v2
    m.conv1.weight = torch.nn.Parameter(m.conv1.weight.detach().clone())  # a replaced parameter object
    assert m._versions() != v3

