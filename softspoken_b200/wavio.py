"""Minimal RIFF/WAVE reader and writer for the formats the batch path meets.

The reference decodes with soundfile/libsndfile (`sf.read(path, dtype='float32')`,
root/code/backend/voice_activity.py:37) and writes with `sf.write(path, audio.T, sr)`
(root/code/frontend/silencer_ui.py:998; WAV default subtype PCM_16).  Neither
library is in this image, so the host side carries its own decoder for
PCM_16 / PCM_24 / PCM_32 / IEEE float32 WAVE files.  Integer PCM is scaled the
way libsndfile's float read does (divide by 2**(bits-1)).
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

WAVE_FORMAT_PCM = 1
WAVE_FORMAT_IEEE_FLOAT = 3
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class WavError(ValueError):
    pass


def _chunks(buf: memoryview):
    pos = 12
    n = len(buf)
    while pos + 8 <= n:
        cid = bytes(buf[pos:pos + 4])
        size = struct.unpack_from("<I", buf, pos + 4)[0]
        yield cid, pos + 8, min(size, n - pos - 8)
        pos += 8 + size + (size & 1)


def wav_info(path: str) -> Tuple[int, int, int, int, int]:
    """-> (frames, sample_rate, channels, bits, format_tag) from the header only."""
    with open(path, "rb") as f:
        head = f.read(1 << 16)
    return _parse_header(memoryview(head), None)[:5]


def _parse_header(buf: memoryview, total_len):
    if len(buf) < 12 or bytes(buf[0:4]) != b"RIFF" or bytes(buf[8:12]) != b"WAVE":
        raise WavError("not a RIFF/WAVE file")
    fmt = None
    pos = 12
    n = len(buf)
    while pos + 8 <= n:
        cid = bytes(buf[pos:pos + 4])
        size = struct.unpack_from("<I", buf, pos + 4)[0]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack_from("<HHIIHH", buf, pos + 8)
            if tag == WAVE_FORMAT_EXTENSIBLE and size >= 26:
                tag = struct.unpack_from("<H", buf, pos + 8 + 24)[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            if fmt is None:
                raise WavError("data chunk before fmt chunk")
            tag, ch, sr, bits = fmt
            frames = size // (ch * bits // 8)
            return frames, sr, ch, bits, tag, pos + 8, size
        pos += 8 + size + (size & 1)
    raise WavError("no data chunk")


def duration_and_rate(path: str) -> Tuple[float, int]:
    """`get_audio_data` (voice_activity.py:23-30): (frames / sr, sr) without decoding."""
    frames, sr, _, _, _ = wav_info(path)
    return frames / sr, sr


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """-> (float32 `(n,)` mono or `(C, n)` multi-channel, sample_rate).

    Same orientation as `sf.read(...)[0].T` (voice_activity.py:37-38)."""
    raw = np.fromfile(path, dtype=np.uint8)
    buf = memoryview(raw)
    frames, sr, ch, bits, tag, off, size = _parse_header(buf, len(raw))
    size = min(size, len(raw) - off)
    frames = size // (ch * bits // 8)
    body = raw[off:off + frames * ch * (bits // 8)]
    if tag == WAVE_FORMAT_PCM and bits == 16:
        x = body.view("<i2").astype(np.float32) / np.float32(32768.0)
    elif tag == WAVE_FORMAT_PCM and bits == 32:
        x = (body.view("<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == WAVE_FORMAT_PCM and bits == 24:
        b = body.reshape(-1, 3).astype(np.int32)
        v = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = (v.astype(np.float64) / 8388608.0).astype(np.float32)
    elif tag == WAVE_FORMAT_IEEE_FLOAT and bits == 32:
        x = body.view("<f4").astype(np.float32)
    else:
        raise WavError(f"unsupported WAVE format tag={tag} bits={bits}")
    if ch > 1:
        x = x.reshape(frames, ch).T.copy()
    return x, sr


def read_wav_pcm16(path: str):
    """PCM_16 file -> (int16 `(n,)` mono or `(n, C)` interleaved frames as stored, sample_rate); `None` for any
    other sample format.  The device decodes these samples itself (`ss_detect_*_pcm16`, `ss_decode_pcm16`): the
    float32 copy that `sf.read(dtype='float32')` would build on the host is never made."""
    raw = np.fromfile(path, dtype=np.uint8)
    frames, sr, ch, bits, tag, off, size = _parse_header(memoryview(raw), len(raw))
    if tag != WAVE_FORMAT_PCM or bits != 16:
        return None
    size = min(size, len(raw) - off)
    frames = size // (ch * 2)
    x = raw[off:off + frames * ch * 2].view("<i2")
    return (x if ch == 1 else x.reshape(frames, ch)), sr


def encode_pcm16(x: np.ndarray) -> np.ndarray:
    """float32 -> int16 as `sf.write(..., subtype='PCM_16')` (libsndfile pcm.c:f2les_array, normalisation on):
    lrintf(x * 32767.0f) with the product rounded to float32 first; out-of-range samples saturate (the C code
    would wrap).  Host twin of the `ss_encode_pcm16` kernel."""
    y = np.rint(np.asarray(x, np.float32) * np.float32(32767.0))
    return np.clip(y, -32768, 32767).astype(np.int16)


def write_wav_pcm16(path: str, pcm: np.ndarray, sr: int) -> None:
    """int16 `(n,)` or `(n, C)` -> PCM_16 WAVE."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    ch = 1 if pcm.ndim == 1 else pcm.shape[1]
    data = pcm.tobytes()
    _write(path, WAVE_FORMAT_PCM, ch, sr, 16, data)


def write_wav_float32(path: str, x: np.ndarray, sr: int) -> None:
    x = np.ascontiguousarray(x, dtype="<f4")
    ch = 1 if x.ndim == 1 else x.shape[1]
    _write(path, WAVE_FORMAT_IEEE_FLOAT, ch, sr, 32, x.tobytes())


def _write(path, tag, ch, sr, bits, data: bytes) -> None:
    block = ch * bits // 8
    fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * block, block, bits)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)
