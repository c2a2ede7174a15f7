"""The review screen's spectrogram on the GPU — mirror of the two reference functions involved.

`wav_to_spec(data, trim_edges=True)` is root/code/backend/voice_activity.py:148-154:
`np.abs(librosa.stft(data, n_fft=512, win_length=512, hop_length=256))`, optionally trimmed to `[..., 0:256, 0:256]`.
`spectrogram_db(spectrogram)` is the display transform of `ReviewDetectionsScreen.display_spectrogram`
(root/code/frontend/review_detections.py:880-881): `np.abs(librosa.amplitude_to_db(spectrogram ** 2, ref=np.max))`.
Both run as hand-written kernels (`csrc/features.cu`: `stft512_kernel`, `spec_db_kernel`) through `ss_spectrogram*`;
there is no CPU fallback.  Parity is against a restatement of librosa's published algorithm (`oracle/spectrogram.py`;
librosa itself is not in the build image), tolerance 1e-4 of the largest magnitude.
"""
from __future__ import annotations

import numpy as np
import torch

from . import settings

assert (settings.n_fft, settings.win_length, settings.hop_length) == (512, 512, 256), "the kernel bakes these in"


def wav_to_spec(data, trim_edges: bool = True, engine=None) -> np.ndarray:
    """`data`: mono float32 samples (numpy array, or a CUDA tensor to skip the upload); `engine`: a
    `softspoken_b200.engine.Engine`.  -> float32 `[257, 1 + n // 256]`, or `[256, 256]` with `trim_edges`
    (the reference's slice keeps whatever exists when the clip is shorter)."""
    if engine is None:
        raise RuntimeError("softspoken_b200.voice_activity.wav_to_spec needs an Engine (there is no CPU fallback)")
    D = _magnitudes(data, engine, db=False)
    if trim_edges:
        D = D[..., 0:256, 0:256]
    return D.cpu().numpy()


def spectrogram_db(data, engine=None) -> np.ndarray:
    """Samples -> what the review screen draws: `np.abs(amplitude_to_db(wav_to_spec(data, False) ** 2, ref=np.max))`
    (0 at the loudest cell, 80 at the floor), `[257, T]` float32, in two launches."""
    if engine is None:
        raise RuntimeError("softspoken_b200.voice_activity.spectrogram_db needs an Engine (there is no CPU fallback)")
    return _magnitudes(data, engine, db=True).cpu().numpy()


def _magnitudes(data, engine, db: bool) -> torch.Tensor:
    if isinstance(data, torch.Tensor):
        x = data
    else:
        a = np.asarray(data)
        if a.dtype != np.int16:
            a = np.ascontiguousarray(a, dtype=np.float32)     # the review screen may hand over a float64 zero-padded copy
        x = torch.from_numpy(np.ascontiguousarray(a))
    if x.dim() != 1:
        raise ValueError("wav_to_spec expects a mono clip")
    return engine.spectrogram(x.to(engine.device).contiguous(), db=db)
