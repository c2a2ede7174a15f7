"""Seeded synthetic inputs of the shapes BASELINE.json names.

There is no sample audio in the reference tree and no network, so every test
and benchmark input is generated: Gaussian noise plus "speech-like" bursts
(amplitude-modulated harmonic stacks with a 120-240 Hz fundamental), quantised
to PCM_16 so that the float32 the detector sees is exactly what a PCM_16 wav
decodes to (`x / 32768`, libsndfile's float read used by the reference at
root/code/backend/voice_activity.py:37).  numpy's PCG64 stream keeps the bytes
identical on every host.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import spec


def burst_plan(duration_s: float, seed: int, n_bursts: int | None = None) -> List[Tuple[float, float, float]]:
    """(start_s, length_s, f0_hz) of the speech-like bursts of clip `seed`."""
    rng = np.random.default_rng([seed, 0xB0057])
    if n_bursts is None:
        n_bursts = max(1, int(round(duration_s / 10.0)))
    out = []
    for _ in range(n_bursts):
        length = float(rng.uniform(0.5, 2.0))
        start = float(rng.uniform(0.0, max(duration_s - length, 0.0)))
        f0 = float(rng.uniform(120.0, 240.0))
        out.append((start, length, f0))
    return sorted(out)


def synth_pcm16(duration_s: float, seed: int = 0, sr: int = spec.SAMPLE_RATE,
                noise_sigma: float = 0.05, n_bursts: int | None = None) -> np.ndarray:
    """int16 mono clip: noise + bursts (SURVEY §8d config 1 recipe)."""
    n = int(round(duration_s * sr))
    rng = np.random.default_rng([seed, 0xA0D10])
    x = rng.normal(0.0, noise_sigma, n)
    for start, length, f0 in burst_plan(duration_s, seed, n_bursts):
        s = int(round(start * sr))
        e = min(n, s + int(round(length * sr)))
        if e <= s:
            continue
        t = np.arange(e - s) / sr
        env = np.sin(np.pi * np.arange(e - s) / (e - s)) ** 2          # smooth on/off
        am = 0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * t)                     # 4 Hz syllabic AM
        vib = f0 * (1.0 + 0.02 * np.sin(2 * np.pi * 5.5 * t))            # slight vibrato
        phase = 2 * np.pi * np.cumsum(vib) / sr
        h = np.zeros(e - s)
        for k in range(1, 13):
            h += np.sin(k * phase) / k
        x[s:e] += 0.25 * env * am * h
    return np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)


def pcm16_to_float32(pcm: np.ndarray) -> np.ndarray:
    """PCM_16 -> float32 exactly as libsndfile's float read does (x / 32768)."""
    return (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def synth_audio(duration_s: float, seed: int = 0, sr: int = spec.SAMPLE_RATE, **kw) -> np.ndarray:
    """float32 mono clip, bit-identical to reading the PCM_16 wav of the same seed."""
    return pcm16_to_float32(synth_pcm16(duration_s, seed, sr, **kw))


def synth_review_rows(n_rows: int, n_files: int, clip_s: float = 600.0, seed: int = 0):
    """Synthetic `erase=1` intervals: uniform start in
    [0, clip_s - 2], length U(0.2, 4.0) s, times rounded to 3 decimals as the
    review screen does (reference review_detections.py:979)."""
    rng = np.random.default_rng([seed, 0x5117])
    files = rng.integers(0, n_files, n_rows)
    start = np.round(rng.uniform(0.0, clip_s - 2.0, n_rows), 3)
    end = np.round(start + rng.uniform(0.2, 4.0, n_rows), 3)
    return files.astype(np.int64), start, end


def stream_hour_pcm16(stream_seed: int, hour: int, sr: int = spec.SAMPLE_RATE) -> np.ndarray:
    """Hour `hour` of the long synthetic recording `stream_seed`: the recording is the
    concatenation of independently seeded one-hour clips, so it never has to exist whole and any prefix can be
    regenerated on another host (360 bursts per hour)."""
    return synth_pcm16(3600.0, 100_000 + 1_000 * int(stream_seed) + int(hour), sr)


def stream_hour(stream_seed: int, hour: int, sr: int = spec.SAMPLE_RATE) -> np.ndarray:
    return pcm16_to_float32(stream_hour_pcm16(stream_seed, hour, sr))
