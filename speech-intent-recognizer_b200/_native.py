"""ctypes binding of libsir_b200.so (include/sir_b200.h) - the only door from Python to the CUDA kernels.

There is deliberately no fallback: if the shared library is missing or no CUDA device is usable, every
compute call raises.  PyTorch is used for device memory and streams only; raw device pointers and the
current stream handle cross the C ABI.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIR_B200_LIB lets kernel experiments A/B a differently built library; the default is the in-tree build.
LIB_PATH = os.environ.get("SIR_B200_LIB") or os.path.join(_HERE, "libsir_b200.so")

OUT_MEL_POWER, OUT_MEL_DB, OUT_LOGMEL_NORM = 0, 1, 2

# symbol -> (restype, argtypes); mirrors include/sir_b200.h one to one (tests/test_abi_cpu.py checks it)
SIGNATURES = {
    "sir_last_error": (c_char_p, []),
    "sir_version": (c_int, []),
    "sir_launch_count": (c_int64, []),
    "sir_profile_enable": (None, [c_int]),
    "sir_profile_read": (c_int, [c_char_p, c_int, POINTER(c_float), POINTER(c_int), c_int]),
    "sir_frontend_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int]),
    "sir_frontend_destroy": (None, [c_void_p]),
    "sir_frontend_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "sir_frontend_mfcc": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                                  c_void_p, c_void_p]),
    "sir_preemphasis": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p]),
    "sir_resampler_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_double]),
    "sir_resampler_destroy": (None, [c_void_p]),
    "sir_resampler_output_length": (c_int64, [c_void_p, c_int64]),
    "sir_resampler_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_void_p,
                                      c_void_p]),
    "sir_frontend_forward_pcm16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "sir_amplitude_to_db": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "sir_specaugment_sample": (c_int, [c_uint64, c_uint64, c_int, c_int, c_int, c_void_p, c_float, c_int, c_int,
                                       c_void_p, c_void_p]),
    "sir_features_finalize": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "sir_model_create": (c_int, [POINTER(c_void_p), c_int, c_int]),
    "sir_model_destroy": (None, [c_void_p]),
    "sir_model_load_weights": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "sir_model_weight_count": (c_int64, [c_void_p]),
    "sir_model_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sir_model_forward_convs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sir_model_forward_head": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sir_predict": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p]),
    "sir_model_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_uint64, c_uint64, c_float,
                                        c_float, c_void_p, c_void_p]),
    "sir_model_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sir_model_backward_part": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sir_model_gru_grad_offset": (c_int, [c_void_p, POINTER(c_int64)]),
    "sir_cross_entropy": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "sir_grad_nonfinite": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "sir_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int, c_float, c_float, c_float,
                              c_float, c_float, c_int, c_float, c_void_p, c_void_p]),
    "sir_train_state_init": (c_int, [c_void_p, c_float, c_int, c_uint64, c_void_p]),
    "sir_train_state_set_scale": (c_int, [c_void_p, c_float, c_void_p]),
    "sir_train_state_begin": (c_int, [c_void_p, c_float, c_float, c_int, c_void_p]),
    "sir_train_state_end": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "sir_model_set_train_state": (c_int, [c_void_p, c_void_p]),
    "sir_cross_entropy_state": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sir_adam_step_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int, c_float, c_float, c_float,
                                    c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "sir_gemm_nt_split_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sir_gemm_tile_width": (c_int, [c_int, c_int, c_int]),
    "sir_conv3x3_nhwc_split_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sir_pipeline_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load_library():
    """dlopen the extension once.  Raises NativeError (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing - build it with `python speech-intent-recognizer_b200/build.py` "
                              "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {load_library().sir_last_error().decode()}")


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise NativeError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise NativeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def stream_ptr():
    """torch's current CUDA stream as the ABI's ``void* stream`` (the raw getter: a tenth of current_stream()'s host time)."""
    if _raw_stream is not None and _raw_device is not None:
        return c_void_p(_raw_stream(_raw_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load_library().sir_launch_count())


def profile_enable(on: bool):
    load_library().sir_profile_enable(1 if on else 0)


def profile_read():
    """-> {stage: (total_ms, calls)} since the last read (synchronises the device)."""
    names = ctypes.create_string_buffer(4096)
    ms = (c_float * 64)()
    calls = (c_int * 64)()
    n = load_library().sir_profile_read(names, 4096, ms, calls, 64)
    keys = names.value.decode().split(";") if n else []
    return {k: (float(ms[i]), int(calls[i])) for i, k in enumerate(keys)}


class Frontend:
    """Owns a ``sir_frontend`` handle on the current CUDA device."""

    def __init__(self, sample_rate=16000, n_mels=64, n_fft=1024, hop_length=512):
        lib = load_library()
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the feature frontend has no CPU path")
        h = c_void_p()
        check(lib.sir_frontend_create(ctypes.byref(h), sample_rate, n_mels, n_fft, hop_length), "sir_frontend_create")
        self._h = h
        self.n_mels, self.hop = n_mels, hop_length

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.sir_frontend_destroy(h)

    def forward(self, wave, lengths=None, max_samples=0, mode=OUT_LOGMEL_NORM, out_frames=None, masks=None,
                status=None, out=None):
        """wave [B, L] fp32 (or int16 PCM, scaled by 1/32768) CUDA, rows may be strided -> [B, n_mels, out_frames] fp32."""
        pcm16 = wave.dtype == torch.int16
        require_cuda(wave, "wave", torch.int16 if pcm16 else torch.float32)
        if wave.dim() != 2 or wave.stride(1) != 1:
            raise NativeError("wave must be [batch, samples] with contiguous rows")
        B, L = wave.shape
        if lengths is not None:
            require_cuda(lengths, "lengths", torch.int32)
        if out_frames is None:
            eff = min(L, max_samples) if max_samples and max_samples > 0 else L
            out_frames = 1 + eff // self.hop
        if out is None:
            out = torch.empty((B, self.n_mels, out_frames), device=wave.device, dtype=torch.float32)
        else:
            require_cuda(out, "out")
            assert out.is_contiguous() and tuple(out.shape) == (B, self.n_mels, out_frames)
        if masks is not None:
            require_cuda(masks, "masks", torch.int32)
            assert masks.is_contiguous() and tuple(masks.shape) == (B, 4)
        if status is not None:
            require_cuda(status, "status", torch.int32)
        stride = wave.stride(0) if B > 1 else max(L, 1)
        entry = load_library().sir_frontend_forward_pcm16 if pcm16 else load_library().sir_frontend_forward
        check(entry(self._h, ptr(wave), stride, ptr(lengths), L, B, int(max_samples or 0), mode, out_frames, ptr(out),
                    ptr(masks), ptr(status), stream_ptr()), "sir_frontend_forward")
        return out


def _frontend_mfcc(self, wave, n_mfcc=40, top_db=80.0, lengths=None, max_samples=0, out_frames=None, status=None):
    """``wave [B, L]`` fp32 CUDA -> MFCC ``[B, n_mfcc, out_frames]`` with torchaudio.transforms.MFCC semantics."""
    require_cuda(wave, "wave")
    if wave.dim() != 2 or wave.stride(1) != 1:
        raise NativeError("wave must be [batch, samples] with contiguous rows")
    B, L = wave.shape
    if lengths is not None:
        require_cuda(lengths, "lengths", torch.int32)
    if out_frames is None:
        eff = min(L, max_samples) if max_samples and max_samples > 0 else L
        out_frames = 1 + eff // self.hop
    out = torch.empty((B, n_mfcc, out_frames), device=wave.device, dtype=torch.float32)
    stride = wave.stride(0) if B > 1 else max(L, 1)
    check(load_library().sir_frontend_mfcc(self._h, ptr(wave), stride, ptr(lengths), L, B, int(max_samples or 0), n_mfcc,
                                           float(top_db or 0.0), out_frames, ptr(out), ptr(status), stream_ptr()),
          "sir_frontend_mfcc")
    return out


Frontend.mfcc = _frontend_mfcc


def preemphasis(wave: torch.Tensor, coeff: float = 0.97) -> torch.Tensor:
    """``y[n] = x[n] - coeff * x[n-1]`` along the last axis of ``wave [B, L]`` (torchaudio.functional.preemphasis)."""
    require_cuda(wave, "wave")
    wave = wave.contiguous()
    B, L = wave.shape
    out = torch.empty_like(wave)
    check(load_library().sir_preemphasis(ptr(wave), ptr(out), L, L, B, float(coeff), stream_ptr()), "sir_preemphasis")
    return out


def amplitude_to_db(x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    check(load_library().sir_amplitude_to_db(ptr(x), ptr(out), x.numel(), stream_ptr()), "sir_amplitude_to_db")
    return out


def specaugment_sample(seed, first_index, batch, n_mels, n_frames, frames=None, augment_prob=1.0, time_mask_param=20,
                       freq_mask_param=10, device="cuda"):
    masks = torch.empty((batch, 4), device=device, dtype=torch.int32)
    if frames is not None:
        require_cuda(frames, "frames", torch.int32)
    check(load_library().sir_specaugment_sample(int(seed), int(first_index), batch, n_mels, n_frames, ptr(frames),
                                                float(augment_prob), time_mask_param, freq_mask_param, ptr(masks),
                                                stream_ptr()), "sir_specaugment_sample")
    return masks


def features_finalize(feat, out_frames, frames=None, masks=None):
    require_cuda(feat, "feat")
    feat = feat.contiguous()
    B, M, T = feat.shape
    out = torch.empty((B, M, out_frames), device=feat.device, dtype=torch.float32)
    check(load_library().sir_features_finalize(ptr(feat), B, M, T, ptr(frames), ptr(masks), out_frames, ptr(out),
                                               stream_ptr()), "sir_features_finalize")
    return out


def gemm_nt_split_f16(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """``a [M,K] @ w[N,K].T + bias`` on tcgen05 with the fp16 hi/lo split (fp32 in/out, CUDA tensors)."""
    for t, n in ((a, "a"), (w, "w"), (bias, "bias")):
        require_cuda(t, n)
    a, w, bias = a.contiguous(), w.contiguous(), bias.contiguous()
    M, K = a.shape
    N = w.shape[0]
    c = torch.empty((M, N), device=a.device, dtype=torch.float32)
    check(load_library().sir_gemm_nt_split_f16(ptr(a), ptr(w), ptr(bias), ptr(c), M, N, K, stream_ptr()),
          "sir_gemm_nt_split_f16")
    return c


def gemm_tile_width(M: int, N: int, sms: int = 148) -> int:
    """Activation-row tile width the persistent GEMM picks for this shape (0: the shape takes the non-persistent kernel)."""
    return int(load_library().sir_gemm_tile_width(M, N, sms))


class Model:
    """Owns a ``sir_model`` handle (repacked weights + activation workspace)."""

    def __init__(self, num_classes, n_mels=64):
        lib = load_library()
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the classifier has no CPU path")
        h = c_void_p()
        check(lib.sir_model_create(ctypes.byref(h), num_classes, n_mels), "sir_model_create")
        self._h = h
        self.num_classes, self.n_mels = num_classes, n_mels

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.sir_model_destroy(h)

    def weight_count(self) -> int:
        return int(load_library().sir_model_weight_count(self._h))

    def load_weights(self, flat: torch.Tensor, bn_eps: float = 1e-5):
        """flat: contiguous fp32 tensor (CPU or CUDA) in utils/synth.py:state_dict_spec order."""
        flat = flat.detach().contiguous()
        assert flat.dtype == torch.float32
        check(load_library().sir_model_load_weights(self._h, ptr(flat), flat.numel(), bn_eps, stream_ptr()),
              "sir_model_load_weights")

    def forward(self, feat: torch.Tensor) -> torch.Tensor:
        require_cuda(feat, "features")
        feat = feat.contiguous()
        B, M, T = feat.shape
        logits = torch.empty((B, self.num_classes), device=feat.device, dtype=torch.float32)
        check(load_library().sir_model_forward(self._h, ptr(feat), B, T, ptr(logits), stream_ptr()), "sir_model_forward")
        return logits

    MAX_STAGED_BATCH = 336

    def forward_convs(self, feat_chunk: torch.Tensor, batch_total: int, first: int):
        """conv stack for utterances ``[first, first + len(feat_chunk))`` of a ``batch_total`` batch (staged forward)."""
        require_cuda(feat_chunk, "features")
        assert feat_chunk.is_contiguous()
        n, M, T = feat_chunk.shape
        check(load_library().sir_model_forward_convs(self._h, ptr(feat_chunk), batch_total, first, n, T, stream_ptr()),
              "sir_model_forward_convs")

    def forward_head(self, batch_total: int, n_frames: int, logits: torch.Tensor = None) -> torch.Tensor:
        if logits is None:
            logits = torch.empty((batch_total, self.num_classes), device="cuda", dtype=torch.float32)
        check(load_library().sir_model_forward_head(self._h, batch_total, n_frames, ptr(logits), stream_ptr()),
              "sir_model_forward_head")
        return logits

    # -- training step pieces (flat fp32 parameter / gradient buffers owned by the caller) ---------------------
    def set_train_state(self, state: "TrainState" = None):
        """Read the dropout stream position from a device-resident ``TrainState`` (None: from the ``offset`` argument)."""
        self._train_state = state                                     # keeps the device buffer alive
        check(load_library().sir_model_set_train_state(self._h, ptr(state.buf) if state is not None else None),
              "sir_model_set_train_state")

    def train_forward(self, flat_params: torch.Tensor, feat: torch.Tensor, dropout_keep: torch.Tensor = None, seed: int = 0,
                      offset: int = 0, bn_momentum: float = 0.1, bn_eps: float = 1e-5, logits: torch.Tensor = None) -> torch.Tensor:
        """Train-mode forward; ``feat`` (contiguous, CUDA) must stay alive until ``backward``."""
        require_cuda(flat_params, "flat_params")
        require_cuda(feat, "features")
        assert flat_params.is_contiguous() and feat.is_contiguous() and flat_params.numel() >= self.weight_count()
        B, M, T = feat.shape
        if dropout_keep is not None:
            require_cuda(dropout_keep, "dropout_keep", torch.uint8)
            assert dropout_keep.is_contiguous() and dropout_keep.numel() == B * (T // 8) * 512
        if logits is None:
            logits = torch.empty((B, self.num_classes), device=feat.device, dtype=torch.float32)
        check(load_library().sir_model_train_forward(self._h, ptr(flat_params), ptr(feat), B, T, ptr(dropout_keep), int(seed),
                                                     int(offset), float(bn_momentum), float(bn_eps), ptr(logits),
                                                     stream_ptr()), "sir_model_train_forward")
        return logits

    def backward(self, flat_params: torch.Tensor, dlogits: torch.Tensor, flat_grads: torch.Tensor):
        """``dlogits [B, C]`` -> every parameter gradient, written into ``flat_grads[:weight_count]``."""
        require_cuda(dlogits, "dlogits")
        require_cuda(flat_grads, "flat_grads")
        assert dlogits.is_contiguous() and flat_grads.is_contiguous() and flat_grads.numel() >= self.weight_count()
        check(load_library().sir_model_backward(self._h, ptr(flat_params), ptr(dlogits), ptr(flat_grads), stream_ptr()),
              "sir_model_backward")
        return flat_grads

    def backward_part(self, flat_params: torch.Tensor, dlogits, flat_grads: torch.Tensor, part: int):
        """Half of :meth:`backward`: part 1 = head + GRU layers (zeroes ``flat_grads`` first), part 2 = conv stack."""
        require_cuda(flat_grads, "flat_grads")
        assert flat_grads.is_contiguous() and flat_grads.numel() >= self.weight_count()
        check(load_library().sir_model_backward_part(self._h, ptr(flat_params), ptr(dlogits) if dlogits is not None else None,
                                                     ptr(flat_grads), int(part), stream_ptr()), "sir_model_backward_part")
        return flat_grads

    def gru_grad_offset(self) -> int:
        """Index of the first GRU parameter in the flat buffer: conv / BatchNorm gradients lie in front of it."""
        out = c_int64(0)
        check(load_library().sir_model_gru_grad_offset(self._h, ctypes.byref(out)), "sir_model_gru_grad_offset")
        return int(out.value)

    def pipeline(self, fe: Frontend, wave, lengths=None, max_samples=0, out_frames=200, features=None):
        require_cuda(wave, "wave")
        B, L = wave.shape
        if features is None:
            features = torch.empty((B, self.n_mels, out_frames), device=wave.device, dtype=torch.float32)
        logits = torch.empty((B, self.num_classes), device=wave.device, dtype=torch.float32)
        stride = wave.stride(0) if B > 1 else max(L, 1)
        check(load_library().sir_pipeline_forward(fe._h, self._h, ptr(wave), stride, ptr(lengths), L, B,
                                                  int(max_samples or 0), out_frames, ptr(features), ptr(logits),
                                                  stream_ptr()), "sir_pipeline_forward")
        return logits, features


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, scale: float = 1.0, loss_out: torch.Tensor = None,
                  need_grad: bool = True):
    """Mean cross-entropy and ``scale * d loss / d logits`` in one launch -> ``(loss[1], dlogits | None)``."""
    require_cuda(logits, "logits")
    require_cuda(labels, "labels", torch.int64)
    logits, labels = logits.contiguous(), labels.contiguous()
    B, C = logits.shape
    loss = loss_out if loss_out is not None else torch.empty(1, device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits) if need_grad else None
    check(load_library().sir_cross_entropy(ptr(logits), ptr(labels), B, C, float(scale), ptr(loss), ptr(dlogits),
                                           stream_ptr()), "sir_cross_entropy")
    return loss, dlogits


def grad_nonfinite(grads: torch.Tensor, count: int, flag: torch.Tensor):
    """``flag[0] = 1.0`` if any of ``grads[:count]`` is inf/nan (flag must be zeroed by the caller)."""
    require_cuda(grads, "grads")
    require_cuda(flag, "flag")
    check(load_library().sir_grad_nonfinite(ptr(grads), int(count), ptr(flag), stream_ptr()), "sir_grad_nonfinite")


def adam_step(params, grads, exp_avg, exp_avg_sq, segments, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step=1,
              inv_scale=1.0, found_inf: torch.Tensor = None):
    """Fused unscale + torch.optim.Adam update over ``segments`` = [(offset, count), ...] of the flat buffers."""
    for t, n in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        require_cuda(t, n)
    seg = (c_int64 * (2 * len(segments)))(*[int(v) for pair in segments for v in pair])
    check(load_library().sir_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), seg, len(segments), float(lr),
                                       float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                       float(inv_scale), ptr(found_inf), stream_ptr()), "sir_adam_step")


TRAIN_STATE_BYTES = 32


class TrainState:
    """Device-resident per-step scalars of the training step (include/sir_b200.h "device-resident step state"): Adam step
    count and bias corrections, GradScaler scale, dropout stream position.  With it the step takes no per-step host
    argument, i.e. it can be captured in a CUDA graph once and replayed."""

    def __init__(self, loss_scale: float = 65536.0, step: int = 0, dropout_offset: int = 0):
        self.buf = torch.zeros(TRAIN_STATE_BYTES // 4, dtype=torch.int32, device="cuda")
        check(load_library().sir_train_state_init(ptr(self.buf), float(loss_scale), int(step), int(dropout_offset), stream_ptr()),
              "sir_train_state_init")

    def set_scale(self, loss_scale: float):
        check(load_library().sir_train_state_set_scale(ptr(self.buf), float(loss_scale), stream_ptr()), "sir_train_state_set_scale")

    def begin(self, betas, world: int):
        check(load_library().sir_train_state_begin(ptr(self.buf), float(betas[0]), float(betas[1]), int(world), stream_ptr()),
              "sir_train_state_begin")

    def end(self, found_inf: torch.Tensor, dropout_offset_increment: int):
        check(load_library().sir_train_state_end(ptr(self.buf), ptr(found_inf), int(dropout_offset_increment), stream_ptr()),
              "sir_train_state_end")

    def read(self):
        """(step, loss_scale, dropout_offset) - synchronises; for tests and checkpoints."""
        raw = self.buf.cpu().numpy()
        return int(raw[2]), float(raw[4:5].view("float32")[0]), int(raw[0:2].view("uint64")[0])


def cross_entropy_state(logits: torch.Tensor, labels: torch.Tensor, state: TrainState, loss_out: torch.Tensor, dlogits: torch.Tensor):
    require_cuda(logits, "logits")
    require_cuda(labels, "labels", torch.int64)
    B, C = logits.shape
    check(load_library().sir_cross_entropy_state(ptr(logits), ptr(labels), B, C, ptr(state.buf), ptr(loss_out), ptr(dlogits),
                                                 stream_ptr()), "sir_cross_entropy_state")


def adam_step_state(params, grads, exp_avg, exp_avg_sq, segments, lr, betas, eps, weight_decay, state: TrainState, found_inf):
    seg = (c_int64 * (2 * len(segments)))(*[int(v) for pair in segments for v in pair])
    check(load_library().sir_adam_step_state(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), seg, len(segments), float(lr),
                                             float(betas[0]), float(betas[1]), float(eps), float(weight_decay), ptr(state.buf),
                                             ptr(found_inf), stream_ptr()), "sir_adam_step_state")


def conv3x3_nhwc_split_f16(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``x [B,H,W,Cin]``, ``w [9,Cout,Cin]`` (tap-major) -> ``[B,H,W,Cout]``: the tcgen05 implicit-GEMM convolution."""
    require_cuda(x, "x")
    require_cuda(w, "w")
    x, w = x.contiguous(), w.contiguous()
    B, H, W, cin = x.shape
    cout = w.shape[1]
    out = torch.empty((B, H, W, cout), device=x.device, dtype=torch.float32)
    check(load_library().sir_conv3x3_nhwc_split_f16(ptr(x), ptr(w), ptr(out), B, H, W, cin, cout, stream_ptr()),
          "sir_conv3x3_nhwc_split_f16")
    return out


def predict(logits: torch.Tensor, k: int = 3, labels: torch.Tensor = None, confusion: torch.Tensor = None,
            correct: torch.Tensor = None):
    """softmax / argmax / confidence / top-k of ``logits [B, C]`` (+ accuracy and confusion counts when ``labels`` is
    given) in one launch -> ``(pred int32 [B], conf [B], topk_idx int32 [B,k], topk_prob [B,k])``."""
    require_cuda(logits, "logits")
    logits = logits.contiguous()
    B, C = logits.shape
    pred = torch.empty(B, device=logits.device, dtype=torch.int32)
    conf = torch.empty(B, device=logits.device, dtype=torch.float32)
    tk_i = torch.empty((B, k), device=logits.device, dtype=torch.int32) if k else None
    tk_p = torch.empty((B, k), device=logits.device, dtype=torch.float32) if k else None
    if labels is not None:
        require_cuda(labels, "labels", torch.int64)
    for t, n in ((confusion, "confusion"), (correct, "correct")):
        if t is not None:
            require_cuda(t, n, torch.int64)
    check(load_library().sir_predict(ptr(logits), B, C, k, ptr(labels), ptr(pred), ptr(conf), ptr(tk_i), ptr(tk_p),
                                     ptr(confusion), ptr(correct), stream_ptr()), "sir_predict")
    return pred, conf, tk_i, tk_p


class Resampler:
    """``torchaudio.transforms.Resample(orig_freq, new_freq)`` on the GPU (sinc-Hann polyphase FIR)."""

    def __init__(self, orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
        lib = load_library()
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the resampler has no CPU path")
        h = c_void_p()
        check(lib.sir_resampler_create(ctypes.byref(h), int(orig_freq), int(new_freq), int(lowpass_filter_width),
                                       float(rolloff)), "sir_resampler_create")
        self._h = h
        self.orig_freq, self.new_freq = orig_freq, new_freq

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.sir_resampler_destroy(h)

    def output_length(self, n_in: int) -> int:
        return int(load_library().sir_resampler_output_length(self._h, int(n_in)))

    def __call__(self, wave: torch.Tensor, lengths: torch.Tensor = None, return_lengths: bool = False):
        """``wave [..., L]`` fp32 CUDA (+ per-row ``lengths``) -> ``[..., ceil(new * L / orig)]``."""
        require_cuda(wave, "wave")
        lead = wave.shape[:-1]
        w = wave.reshape(-1, wave.shape[-1]).contiguous()
        B, L = w.shape
        n_out = self.output_length(L)
        out = torch.empty((B, n_out), device=w.device, dtype=torch.float32)
        out_len = torch.empty(B, device=w.device, dtype=torch.int32) if return_lengths else None
        if lengths is not None:
            require_cuda(lengths, "lengths", torch.int32)
        check(load_library().sir_resampler_forward(self._h, ptr(w), L, ptr(lengths), L, B, ptr(out), n_out, n_out, ptr(out_len),
                                                   stream_ptr()), "sir_resampler_forward")
        out = out.reshape(*lead, n_out)
        return (out, out_len) if return_lengths else out

    forward = __call__
