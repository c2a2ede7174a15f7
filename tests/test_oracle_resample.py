"""Properties of the resampling filter (softspoken_b200/resample.py) through its float64 oracle."""
import numpy as np
import pytest

from oracle import resample as orr
from softspoken_b200 import resample as rs


@pytest.mark.parametrize("sr", [44100, 48000, 16000, 32000, 96000, 8000, 11025, 22050])
def test_length_dc_gain_and_table(sr):
    for n in (0, 1, 999, 48000):
        assert rs.out_len(n, sr) == int(np.ceil(n * 22050 / sr))
    L, M, T, g = rs.design(sr)
    assert g.shape == (2 * T + 1, L) and g.dtype == np.float32 and L * sr == M * 22050
    assert np.allclose(g.sum(axis=0), 1.0, atol=2e-6)
    y = orr.resample(np.ones(6 * T + 50), sr)
    mid = slice(len(y) // 3, 2 * len(y) // 3)
    assert np.allclose(y[mid], 1.0, atol=1e-12)


@pytest.mark.parametrize("sr", [44100, 48000, 16000])
def test_tones_pass_and_aliases_are_rejected(sr):
    n = sr                                     # one second
    t = np.arange(n) / sr
    band = 0.5 * min(sr, 22050)
    for f in (440.0, 3000.0, 0.85 * band):     # inside the pass band: reproduced at the new rate
        y = orr.resample(np.sin(2 * np.pi * f * t), sr)
        tt = np.arange(len(y)) / 22050
        core = slice(2000, len(y) - 2000)
        assert np.max(np.abs(y[core] - np.sin(2 * np.pi * f * tt[core]))) < 1e-4, f
    if sr > 22050:                             # above the new Nyquist (would alias): gone
        for f in (1.08 * 11025, 0.45 * sr):
            y = orr.resample(np.sin(2 * np.pi * f * t), sr)
            assert np.max(np.abs(y[2000:-2000])) < 1e-5, f


def test_identity_rate_keeps_band_limited_signals():
    """L = M = 1 is still a low-pass (h is not a delta): anything below 0.9 Nyquist comes back unchanged and undelayed."""
    rng = np.random.default_rng(0)
    t = np.arange(6000) / 22050
    x = sum(a * np.sin(2 * np.pi * f * t + ph) for a, f, ph in zip(rng.uniform(0.1, 1, 12), rng.uniform(50, 9900, 12),
                                                                   rng.uniform(0, 6, 12)))
    y = orr.resample(x, 22050)
    assert len(y) == len(x) and np.max(np.abs(y[1000:-1000] - x[1000:-1000])) < 1e-4
