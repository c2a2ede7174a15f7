"""CPU statement of the polyphase resampler (softspoken_b200/resample.py).  TEST INFRASTRUCTURE.

**Parity unpinned against the reference**: the reference resamples with soxr through librosa
(root/code/backend/voice_activity.py:44-66), neither of which is in this image.  This oracle pins the KERNEL to the
filter definition written in softspoken_b200/resample.py (float64 evaluation, one output at a time from the
definition, no polyphase table), and tests/test_oracle_resample.py checks the definition's properties: librosa's
output length, unit DC gain, tones kept within 1e-4 in the pass band, > 100 dB rejection of what would alias."""
from __future__ import annotations

import numpy as np

from softspoken_b200 import resample as rs


def resample(x: np.ndarray, sr_in: int) -> np.ndarray:
    """float `(n,)` at sr_in -> float64 `(ceil(n * 22050 / sr_in),)` at 22,050 Hz, straight from the definition."""
    x = np.asarray(x, dtype=np.float64)
    L, M = rs.ratio(sr_in)
    c = 0.5 * min(1.0, L / M) * rs.ROLLOFF
    t_half = rs.ZEROS / (2.0 * c)
    T = int(np.ceil(t_half))
    n_out = rs.out_len(len(x), sr_in)
    xp = np.concatenate([np.zeros(T + 1), x, np.zeros(T + 2)])
    m = np.arange(n_out, dtype=np.int64)
    n0 = (m * M) // L
    frac = ((m * M) % L) / L
    j = np.arange(-T, T + 1)
    y = np.zeros(n_out)
    norm = np.zeros(n_out)
    for jj in j:                                   # one tap at a time over all outputs
        w = rs.kernel_value(jj + frac, c, t_half)
        y += xp[n0 - jj + T + 1] * w
        norm += w
    return y / norm
