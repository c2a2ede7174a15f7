"""K8: the review-screen spectrogram kernels against the oracle's restatement of librosa.stft /
amplitude_to_db.  Tolerance: 1e-4 of the largest magnitude (float32 FFT against a float64 one), 2e-3 dB."""
import numpy as np
import pytest
import torch

from oracle import spectrogram as osp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=1, mode="fp32")
    yield eng
    eng.close()


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 511, 512, 513, 8191, 8192, 66150, 300001])
def test_magnitudes_match_oracle_at_ragged_lengths(engine, n):
    rng = np.random.default_rng(n)
    x = (rng.normal(size=n) * 0.3).astype(np.float32)
    if n > 600:
        x[100:400] += np.sin(2 * np.pi * 440 * np.arange(300) / 22050).astype(np.float32)
    want = osp.stft_magnitude(x)
    got = engine.spectrogram(torch.from_numpy(x).cuda()).cpu().numpy()
    assert got.shape == want.shape == (257, 1 + n // 256) and got.dtype == np.float32
    scale = max(float(want.max()), 1e-6)
    assert float(np.abs(got - want).max()) <= 1e-4 * scale


def test_pcm16_route_and_wav_to_spec_mirror(engine):
    from softspoken_b200 import synth, voice_activity
    pcm = synth.synth_pcm16(5.0, 3)
    x = pcm.astype(np.float32) / np.float32(32768.0)
    a = engine.spectrogram(torch.from_numpy(pcm).cuda())
    b = engine.spectrogram(torch.from_numpy(x).cuda())
    assert torch.equal(a, b)                      # int16 / 32768 is exact: same bits either way
    full = voice_activity.wav_to_spec(x, trim_edges=False, engine=engine)
    trimmed = voice_activity.wav_to_spec(x, engine=engine)
    assert full.shape == (257, 1 + len(x) // 256) and trimmed.shape == (256, 256)
    assert np.array_equal(trimmed, full[:256, :256])
    want = osp.wav_to_spec(x)
    assert float(np.abs(trimmed - want).max()) <= 1e-4 * float(want.max())
    short = voice_activity.wav_to_spec(x[:3000].astype(np.float64), engine=engine)      # float64 in, shorter than the trim
    assert short.shape == (256, 12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        voice_activity.wav_to_spec(x)


def test_display_db_matches_oracle(engine):
    from softspoken_b200 import synth, voice_activity
    x = synth.synth_audio(8.0, 5)
    x[20000:26000] = 0.0                                   # a silent stretch reaches the -80 dB floor
    want = osp.display_db(osp.stft_magnitude(x))
    got = voice_activity.spectrogram_db(x, engine=engine)
    assert got.shape == want.shape and got.min() == 0.0 and got.max() == 80.0
    # away from the floor / threshold kinks the two agree to float32 log10 accuracy
    mid = (want > 0.5) & (want < 79.0)
    assert mid.mean() > 0.5 and float(np.abs(got - want)[mid].max()) <= 2e-3
    assert float(np.abs(got - want).max()) <= 0.05
    zero = engine.spectrogram(torch.zeros(5000, device="cuda"), db=True)
    assert float(zero.abs().max()) == 0.0                  # all-silent clip: every cell is at the reference level


def test_full_size_properties(engine):
    """A 10-minute clip (BASELINE config 2 size): Parseval per frame against the windowed samples, and linearity."""
    from softspoken_b200 import synth
    x = synth.synth_audio(600.0, 1)
    X = engine.spectrogram(torch.from_numpy(x).cuda())
    assert X.shape == (257, 1 + len(x) // 256)
    # Parseval for a real 512-point DFT: sum_n y^2 = (|X0|^2 + |X256|^2 + 2 sum_{0<k<256} |Xk|^2) / 512
    w = torch.from_numpy((0.5 - 0.5 * np.cos(2 * np.pi * np.arange(512) / 512)).astype(np.float32)).cuda()
    xp = torch.nn.functional.pad(torch.from_numpy(x).cuda(), (256, 256))
    frames = xp.unfold(0, 512, 256)[: X.shape[1]] * w
    energy = (frames.double() ** 2).sum(1)
    P = X.double() ** 2
    spec_energy = (P[0] + P[256] + 2 * P[1:256].sum(0)) / 512
    assert float(((energy - spec_energy).abs() / energy.clamp_min(1e-12)).max()) < 1e-4
    Y = engine.spectrogram(torch.from_numpy(x * np.float32(0.25)).cuda())
    assert float((Y - 0.25 * X).abs().max()) <= 1e-5 * float(X.max())
