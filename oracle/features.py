"""Oracle: log-mel front end of the detector.  TEST INFRASTRUCTURE.

Restates, without torchaudio, what `SpecUNet_2D.forward` does before the first
convolution (root/code/backend/pytorch_neural_nets.py:92-99,144-153):

    torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=2048,
        win_length=512, hop_length=256, n_mels=128, f_max=8000)   # power=2, center,
    -> sqrt(log10(x + 1)) -> [:, :, :256]                          # reflect, htk

torchaudio 2.11.0 (pinned by the image, unpinned in the reference's
requirements.txt:1-12) implements that as `torch.stft(center=True,
pad_mode='reflect', window zero-padded to n_fft around its centre,
onesided)` -> `abs().pow(2)` -> `matmul(spec^T, fb)^T`.  Two restatements live
here:

* `log_mel`      — torch float32, same library FFT; the reference's arithmetic.
* `log_mel_f64`  — numpy float64 straight from the definition (direct frames,
                   rfft in double); independent of any float32 rounding, used
                   to put error bars on both the reference and the CUDA kernel.

Pinned by tests/test_oracle_features.py against tests/golden/frontend.npz.
"""
from __future__ import annotations

import numpy as np
import torch

N_FFT = 2048
WIN = 512
HOP = 256
N_FRAMES = 256
PAD = N_FFT // 2            # torch.stft(center=True) reflect-pads n_fft // 2 per side
WIN_OFFSET = (N_FFT - WIN) // 2   # torch.stft centres a short window inside n_fft


def frames_f32(x: torch.Tensor) -> torch.Tensor:
    """x `[B, n>=65536]` -> windowless frames `[B, 256, 512]`.

    Frame t holds x[256 t - 256 .. 256 t + 255]; for t = 0 the first 256 samples
    are the reflection x[256], x[255], ..., x[1] (torch 'reflect': no edge repeat).
    """
    xp = torch.nn.functional.pad(x.unsqueeze(1), (PAD, PAD), mode="reflect").squeeze(1)
    fr = xp.unfold(-1, N_FFT, HOP)[:, :N_FRAMES, WIN_OFFSET:WIN_OFFSET + WIN]
    return fr


@torch.no_grad()
def power_spectrum(x: torch.Tensor, window: torch.Tensor) -> torch.Tensor:
    """`[B,n]` -> `[B, 256 frames, 1025 bins]` float32 power, torch.stft-style."""
    fr = frames_f32(x) * window
    buf = torch.zeros(fr.shape[0], N_FRAMES, N_FFT, dtype=torch.float32)
    buf[:, :, WIN_OFFSET:WIN_OFFSET + WIN] = fr
    return torch.fft.rfft(buf, dim=-1).abs().pow(2.0)


@torch.no_grad()
def log_mel(x: torch.Tensor, window: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """`[B, 66150]` float32 -> `[B, 128, 256]` float32 = sqrt(log10(mel + 1))."""
    p = power_spectrum(x, window)                       # [B, T, F]
    mel = torch.matmul(p, fb).transpose(-1, -2)         # MelScale.forward
    return torch.sqrt(torch.log10(mel + 1))


def log_mel_f64(x: np.ndarray, window: np.ndarray, fb: np.ndarray) -> np.ndarray:
    """float64 definition-level restatement, `[n]` -> `[128, 256]`."""
    x = np.asarray(x, np.float64)
    idx = np.arange(N_FRAMES)[:, None] * HOP - HOP + np.arange(WIN)[None, :]
    idx = np.abs(idx)                                   # reflect about sample 0
    fr = x[idx] * np.asarray(window, np.float64)[None, :]
    p = np.abs(np.fft.rfft(fr, n=N_FFT, axis=-1)) ** 2  # zero-padding position only changes phase
    mel = p @ np.asarray(fb, np.float64)
    return np.sqrt(np.log10(mel + 1.0)).T
