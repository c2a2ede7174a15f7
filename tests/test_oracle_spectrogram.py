"""The spectrogram oracle (oracle/spectrogram.py) against two independent computations.  librosa is not in
the image, so the oracle restates its published algorithm ("parity unpinned"); what CAN be pinned is that the
restatement computes the transform it says it does."""
import numpy as np
from scipy.signal import stft as scipy_stft

from oracle import spectrogram as osp


def test_stft_magnitude_equals_scipy_stft():
    rng = np.random.default_rng(0)
    for n in (600, 1000, 4096, 66150):
        x = rng.normal(size=n).astype(np.float32)
        got = osp.stft_magnitude(x)
        # same frames (zero boundary, no end padding), same periodic Hann; scipy divides by the window sum
        _, _, Z = scipy_stft(x.astype(np.float64), window="hann", nperseg=512, noverlap=256, nfft=512, boundary="zeros",
                             padded=False, return_onesided=True)
        want = np.abs(Z) * 256.0
        assert got.shape == (257, 1 + n // 256) and got.dtype == np.float32
        assert want.shape[1] >= got.shape[1]
        np.testing.assert_allclose(got, want[:, :got.shape[1]], rtol=0, atol=2e-6 * float(want.max()))


def test_stft_magnitude_equals_direct_dft_and_edges():
    rng = np.random.default_rng(1)
    x = rng.normal(size=700).astype(np.float32)
    got = osp.stft_magnitude(x)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(512) / 512)
    padded = np.concatenate([np.zeros(256), x.astype(np.float64), np.zeros(256)])
    k = np.arange(257)[:, None] * np.arange(512)[None, :]
    for t in range(got.shape[1]):
        fr = padded[256 * t:256 * t + 512] * w
        want = np.abs((fr[None, :] * np.exp(-2j * np.pi * k / 512)).sum(1))
        np.testing.assert_allclose(got[:, t], want, rtol=0, atol=1e-5 * want.max())
    assert osp.stft_magnitude(np.zeros(0, np.float32)).shape == (257, 1)
    assert osp.stft_magnitude(np.ones(255, np.float32)).shape == (257, 1)
    assert osp.wav_to_spec(rng.normal(size=70000).astype(np.float32)).shape == (256, 256)
    assert osp.wav_to_spec(rng.normal(size=70000).astype(np.float32), trim_edges=False).shape == (257, 274)


def test_display_db_range_and_floor():
    rng = np.random.default_rng(2)
    S = np.abs(rng.normal(size=(257, 40))).astype(np.float32)
    S[3, 3] = 0.0
    db = osp.display_db(S)
    assert db.dtype == np.float32 and db.min() == 0.0 and db.max() == 80.0
    i = np.unravel_index(np.argmax(S), S.shape)
    assert db[i] == 0.0 and db[3, 3] == 80.0
    mid = S > 0.05 * S.max()
    np.testing.assert_allclose(db[mid], -40.0 * np.log10(S[mid] / S.max()), atol=1e-3)
