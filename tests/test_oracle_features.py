"""Oracle front end vs golden vectors frozen from the reference (tests/golden/frontend.npz)."""
import hashlib

import numpy as np
import torch

from conftest import load_golden
from oracle import features as of
from oracle import postproc as pp
from softspoken_b200 import checkpoint, spec


def test_window_and_filterbank_bit_exact():
    g = load_golden("frontend.npz")
    assert np.array_equal(checkpoint.hann_window().numpy(), g["window"])
    fb = checkpoint.mel_filterbank().numpy()
    assert fb.shape == (1025, 128)
    assert hashlib.sha256(fb.tobytes()).hexdigest() == str(g["fb_sha256"])
    dense = np.zeros_like(fb)
    dense[g["fb_rows"], g["fb_cols"]] = g["fb_vals"]
    assert np.array_equal(dense, fb)


def test_filterbank_structure():
    """rows 0 and 744.. are zero, every band non-empty, 2..31 taps, 1469 non-zeros."""
    fb = checkpoint.mel_filterbank()
    start, count, offs, taps = checkpoint.sparse_filterbank(fb)
    assert int((fb != 0).sum()) == 1469
    assert count.min() >= 2 and count.max() <= 31
    assert start.min() >= 1 and (start + count).max() <= 744
    # packed form reproduces the dense one
    dense = np.zeros((1025, 128), np.float32)
    for m in range(128):
        dense[start[m]:start[m] + count[m], m] = taps[offs[m]:offs[m] + count[m]]
    assert np.array_equal(dense, fb.numpy())


def test_log_mel_matches_reference_golden(clip60):
    g = load_golden("frontend.npz")
    padded = pp.pad_audio(clip60)
    x = torch.stack([torch.from_numpy(padded[i:i + spec.WINDOW_SAMPLES]) for i in g["starts"]])
    mel = of.log_mel(x, checkpoint.hann_window(), checkpoint.mel_filterbank()).numpy()
    ref = g["mel"]
    assert mel.shape == ref.shape == (3, 128, 256)
    # same library FFT as the reference: bit-exact in the build container; a different host CPU may
    # pick another pocketfft/sgemm code path, so allow float32 rounding there.
    assert np.max(np.abs(mel - ref)) <= 2e-6 * np.max(np.abs(ref))


def test_log_mel_f64_definition(clip60):
    """The float64 from-the-definition restatement agrees with the reference to float32 rounding."""
    g = load_golden("frontend.npz")
    padded = pp.pad_audio(clip60)
    for k, s in enumerate(g["starts"]):
        m64 = of.log_mel_f64(padded[s:s + spec.WINDOW_SAMPLES], g["window"], checkpoint.mel_filterbank().numpy())
        assert np.max(np.abs(m64 - g["mel"][k])) <= 2e-6 * np.max(np.abs(g["mel"][k]))


def test_only_first_65536_samples_matter(clip60):
    padded = pp.pad_audio(clip60)
    x = torch.from_numpy(padded[5 * 13230:5 * 13230 + spec.WINDOW_SAMPLES].copy())[None]
    a = of.log_mel(x, checkpoint.hann_window(), checkpoint.mel_filterbank())
    x[:, spec.WINDOW_SAMPLES_USED:] = 123.0
    b = of.log_mel(x, checkpoint.hann_window(), checkpoint.mel_filterbank())
    assert torch.equal(a, b)
