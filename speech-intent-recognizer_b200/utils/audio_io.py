"""Audio file decode for the reference-compatible entry points (host side, outside the hot path).

The reference calls ``torchaudio.load`` (scripts/precompute_features.py:47), which in this image needs the
absent torchcodec.  We try it first and fall back to a small RIFF/WAVE reader (PCM 8/16/24/32-bit and IEEE
float) so that real ``.wav`` files work; anything else (e.g. the reference's MP3-in-.wav clips) raises, and
callers translate that into the reference's None / zeros conventions.
"""
from __future__ import annotations

import struct

import numpy as np
import torch


def _read_riff_wav(path: str):
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:          # WAVE_FORMAT_EXTENSIBLE: real tag in the GUID
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, channels, rate, _, _, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(payload, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(payload, "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(payload[: len(payload) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = np.frombuffer(payload, "<i4").astype(np.float32) / 2147483648.0
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3 and bits == 32:
        x = np.frombuffer(payload, "<f4").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}/{bits} bit")
    x = x[: len(x) // channels * channels].reshape(-1, channels).T
    return torch.from_numpy(np.ascontiguousarray(x)), int(rate)


def load_audio(path: str):
    """-> (waveform [channels, samples] fp32 CPU tensor in [-1, 1], sample_rate)."""
    try:
        import torchaudio
        return torchaudio.load(path)
    except (ImportError, RuntimeError, OSError):
        return _read_riff_wav(path)


def write_wav_pcm16(path: str, wave, sample_rate: int = 16000):
    """Tiny PCM16 mono/stereo writer used by tests and examples."""
    x = np.asarray(wave, dtype=np.float32)
    if x.ndim == 1:
        x = x[None]
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype("<i2").tobytes()
    ch = x.shape[0]
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, ch, sample_rate, sample_rate * ch * 2, ch * 2, 16))
        f.write(b"data" + struct.pack("<I", len(pcm)) + pcm)
