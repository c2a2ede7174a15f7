"""Sample-rate conversion to the detector's 22,050 Hz as a polyphase windowed-sinc FIR on the GPU.

The reference resamples in `voice_activity.load_audio` (root/code/backend/voice_activity.py:44-66:
`librosa.resample(..., res_type=default)` = soxr "HQ").  soxr is not in the build image and the reference pins no
version, so this is **not** a restatement of soxr: it is a documented filter of comparable quality whose output has
the same length librosa produces (`ceil(n * 22050 / sr)`), and **parity with the reference is unpinned for resampled
files** — the bit-exact claims of this package are for files already at 22,050 Hz, which is what every BASELINE config
uses.  Without this step files at 44.1 / 48 kHz (what field recorders write) could not be processed at all.

Definition (everything the kernel `resample_kernel` and the oracle share):
  L / M = 22050 / sr reduced;  output m sits at input time t_m = m M / L = n0 + p / L  (n0 = mM div L, p = mM mod L);
  y[m] = sum_{j=-T..T} x[n0 - j] g[j][p],   g[j][p] = h(j + p / L),   x = 0 outside the clip
  (the device table stores g's columns in visit order: column m mod L holds phase p = (m M) mod L);
  h(t) = 2c sinc(2c t) kaiser(t / T_half; beta),  |t| <= T_half,  c = 0.5 min(1, L / M) ROLLOFF cycles per input sample,
  T_half = ZEROS / (2c),  T = ceil(T_half);  each phase row is normalised to unit DC gain.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np

from . import spec

ROLLOFF = 0.945        # pass band edge as a fraction of the narrower Nyquist band
ZEROS = 64             # zero crossings of the sinc kept each side
BETA = 12.0            # Kaiser window: ~ -118 dB side lobes
MAX_TAPS = 4096        # 2 T + 1 must stay below this (sr up to ~ 700 kHz)


def ratio(sr_in: int) -> Tuple[int, int]:
    g = math.gcd(spec.SAMPLE_RATE, int(sr_in))
    return spec.SAMPLE_RATE // g, int(sr_in) // g


def out_len(n_in: int, sr_in: int) -> int:
    """librosa.resample's `int(np.ceil(n * ratio))`, in integers."""
    L, M = ratio(sr_in)
    return (int(n_in) * L + M - 1) // M


def kernel_value(t: np.ndarray, c: float, t_half: float) -> np.ndarray:
    """h(t) in float64."""
    t = np.asarray(t, dtype=np.float64)
    inside = np.abs(t) <= t_half
    r = np.clip(t / t_half, -1.0, 1.0)
    win = np.i0(BETA * np.sqrt(1.0 - r * r)) / np.i0(BETA)
    return np.where(inside, 2.0 * c * np.sinc(2.0 * c * t) * win, 0.0)


@lru_cache(maxsize=16)
def design(sr_in: int):
    """-> (L, M, T, table float32 `[2T+1][L]` with table[j + T][q] = g[j][(q M) mod L], q = m mod L)."""
    if sr_in <= 0:
        raise ValueError(f"sample rate {sr_in}")
    L, M = ratio(sr_in)
    c = 0.5 * min(1.0, L / M) * ROLLOFF
    t_half = ZEROS / (2.0 * c)
    T = int(math.ceil(t_half))
    if 2 * T + 1 > MAX_TAPS:
        raise ValueError(f"sample rate {sr_in} Hz needs {2 * T + 1} taps (> {MAX_TAPS})")
    j = np.arange(-T, T + 1, dtype=np.float64)[:, None]
    p = np.arange(L, dtype=np.float64)[None, :]
    g = kernel_value(j + p / L, c, t_half)
    g /= g.sum(axis=0, keepdims=True)                 # unit DC gain in every phase
    # Columns in VISIT order: output m uses phase (m M) mod L, which depends on m mod L only, so column q = m mod L
    # holds phase (q M) mod L and the 32 outputs of a warp read 32 neighbouring columns (one or two cache lines per
    # tap instead of five with columns in phase order: the kernel was bound by exactly those L1 wavefronts).
    visit = (np.arange(L, dtype=np.int64) * M) % L
    return L, M, T, np.ascontiguousarray(g[:, visit], dtype=np.float32)
