"""CPU restatement of the review-screen spectrogram.  TEST INFRASTRUCTURE: only tests/ may import it.

**Parity unpinned.**  The reference computes `np.abs(librosa.stft(data, n_fft=512, win_length=512, hop_length=256))`
(root/code/backend/voice_activity.py:148-154, settings.py:4-6) and displays
`np.abs(librosa.amplitude_to_db(spectrogram ** 2, ref=np.max))` (root/code/frontend/review_detections.py:880-881).
librosa is not in this image and the reference pins no version (requirements.txt), so what follows restates the
published algorithm of librosa >= 0.10 (the versions that run on the Python 3.12 / 3.13 the README names):

  stft: `center=True` pads n_fft // 2 = 256 samples each side with `pad_mode="constant"` (zeros); frames of 512 at hop
        256 -> 1 + n // 256 frames; window = `scipy.signal.get_window("hann", 512, fftbins=True)` (periodic Hann,
        float64) multiplied into the frames (float64 * float32 -> float64); `numpy.fft.rfft` over the frame axis; the
        result is stored as complex64 for float32 input -> |.| is float32 `[257, T]`.
  amplitude_to_db(S, ref=np.max, amin=1e-5, top_db=80.0): magnitude = |S|; ref_value = max(magnitude);
        power = magnitude ** 2; power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db):
        10 log10(max(amin^2, power)) - 10 log10(max(amin^2, ref^2)), then max(., that.max() - top_db).
        All of it in the array's dtype (float32 here: python-float scalars are weak under NEP 50).

`stft_magnitude` is cross-checked against `scipy.signal.stft` (an independent implementation of the same transform)
and against a direct DFT in tests/test_oracle_spectrogram.py.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import get_window

N_FFT = 512
HOP = 256


def n_frames(n_samples: int) -> int:
    return 1 + n_samples // HOP


def stft_magnitude(x: np.ndarray) -> np.ndarray:
    """float32 `(n,)` -> float32 `[257, 1 + n // 256]`."""
    x = np.asarray(x, dtype=np.float32)
    padded = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="constant")
    T = n_frames(len(x))
    idx = np.arange(N_FFT)[:, None] + HOP * np.arange(T)[None, :]
    frames = padded[idx]                                             # [512, T] float32
    window = get_window("hann", N_FFT, fftbins=True)                 # float64
    spec = np.fft.rfft(window[:, None] * frames, axis=0)             # complex128 [257, T]
    return np.abs(spec.astype(np.complex64))


def wav_to_spec(data: np.ndarray, trim_edges: bool = True) -> np.ndarray:
    """voice_activity.py:148-154."""
    D = stft_magnitude(data)
    if trim_edges:
        D = D[..., 0:256, 0:256]
    return D


def display_db(spectrogram: np.ndarray) -> np.ndarray:
    """review_detections.py:880-881: np.abs(librosa.amplitude_to_db(spectrogram ** 2, ref=np.max)), float32."""
    S = np.asarray(spectrogram, dtype=np.float32) ** 2
    magnitude = np.abs(S)
    ref_value = np.max(magnitude) if magnitude.size else np.float32(0)
    power = np.square(magnitude)
    amin = 1e-5 ** 2
    log_spec = 10.0 * np.log10(np.maximum(amin, power))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value ** 2))
    if log_spec.size:
        log_spec = np.maximum(log_spec, log_spec.max() - 80.0)
    return np.abs(log_spec)
