"""Oracle: interval silencing.  TEST INFRASTRUCTURE.

Restates `SilenceWorker.run` (root/code/frontend/silencer_ui.py:931-1009) and the
`erase` coercion of `SilenceVoicesScreen.load_review_data` (:1098-1106) on
in-memory float32 buffers.  The reference's file decode/encode goes through
librosa/soundfile/libsndfile, which are absent here: parity is
asserted on the float32 sample buffers, and the PCM_16 encode is restated from
libsndfile's documented float->short conversion (see `float_to_pcm16`).

Pinned by tests/test_oracle_silence.py against tests/golden/silence_*.npz
(frozen by running the real `SilenceWorker.run` with in-memory load/write stubs).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def coerce_erase(values) -> np.ndarray:
    """`pd.to_numeric(erase, errors='coerce').fillna(0).astype(int)` (silencer_ui.py:1100)."""
    import pandas as pd
    return pd.to_numeric(pd.Series(list(values)), errors="coerce").fillna(0).astype(int).to_numpy()


def interval_to_samples(start_time, end_time, sr: int, n: int) -> Tuple[int, int]:
    """silencer_ui.py:975-982: Python round (half-to-even) of the double product, clamped to [0, n]."""
    st = float(start_time)
    et = float(end_time)
    start_index = int(round(st * sr))
    end_index = int(round(et * sr))
    start_index = max(0, min(start_index, n))
    end_index = max(0, min(end_index, n))
    return start_index, end_index


def silence_buffer(audio: np.ndarray, sr: int, rows: Sequence[Tuple[float, float]]) -> np.ndarray:
    """One file: audio `(n,)` or `(C,n)` float32 -> `(C,n)` with every row zeroed (silencer_ui.py:969-985)."""
    a = np.array(audio, dtype=np.float32, copy=True)
    if a.ndim == 1:
        a = np.expand_dims(a, axis=0)
    for st, et in rows:
        s, e = interval_to_samples(st, et, sr, a.shape[1])
        a[:, s:e] = 0.0
    return a


def group_rows(file_path: Sequence[str], file_name: Sequence[str], start: Sequence[float],
               end: Sequence[float], erase: Sequence[int]) -> "Dict[Tuple[str,str], List[Tuple[float,float]]]":
    """`df[df.erase == 1].groupby(['file_path','file_name'])`: sorted keys, original row order inside
    a group (silencer_ui.py:938-945)."""
    groups: Dict[Tuple[str, str], List[Tuple[float, float]]] = {}
    for fp, fn, s, e, er in zip(file_path, file_name, start, end, erase):
        if int(er) == 1:
            groups.setdefault((fp, fn), []).append((float(s), float(e)))
    return dict(sorted(groups.items()))


def float_to_pcm16(a: np.ndarray) -> np.ndarray:
    """float32 -> int16 as libsndfile writes a PCM_16 WAV from float input with its default
    (non-clipping) normalisation, pcm.c:f2les_array: `lrintf(src[i] * normfact)` with `normfact = 1.0 * 0x7FFF`
    held in a float (so the product is a float32 product), round-half-even.  Saturated here where C would wrap.
    UNPINNED: libsndfile is absent from this image, so byte parity of the written wav is NOT claimed
    (SURVEY §8c); float buffers are."""
    y = np.rint(np.asarray(a, np.float32) * np.float32(32767.0))
    return np.clip(y, -32768, 32767).astype(np.int16)


def pcm16_to_float(pcm: np.ndarray) -> np.ndarray:
    """int16 -> float32 as `sf.read(dtype='float32')` (libsndfile pcm.c:s2f_array, normfact 1/0x8000)."""
    return pcm.astype(np.float32) / np.float32(32768.0)


def load_audio_pcm16(frames: np.ndarray) -> np.ndarray:
    """`voice_activity.load_audio` for PCM_16 frames `(n,)` or `(n, C)` at 22,050 Hz
    (root/code/backend/voice_activity.py:37,61-62): float32 read, `.T`, `librosa.to_mono` = `np.mean(y, axis=0)`."""
    f1 = pcm16_to_float(np.asarray(frames))
    data = f1.T
    if data.ndim > 1:
        data = np.mean(data, axis=tuple(range(data.ndim - 1)))      # librosa.to_mono
    return data


def silence_pcm16(frames: np.ndarray, sr: int, rows: Sequence[Tuple[float, float]], requantize: bool = True) -> np.ndarray:
    """PCM_16 file in, PCM_16 file out through the reference's float path (silencer_ui.py:959-998):
    `librosa.load(sr=None, mono=False)` -> zero the rows -> `sf.write(audio.T)`.  `(n,)` or `(n, C)` int16 frames
    -> same shape.  With `requantize=False` the untouched samples are kept as stored (identity round trip)."""
    fr = np.asarray(frames)
    audio = pcm16_to_float(fr).T                                   # (C, n) or (n,)
    out = silence_buffer(audio, sr, rows)                          # (C, n)
    enc = float_to_pcm16(out.T) if requantize else np.where(out.T == 0.0, 0, fr.reshape(out.T.shape)).astype(np.int16)
    return enc.reshape(fr.shape)
