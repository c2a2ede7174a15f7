"""K9: the polyphase resampling kernel against the float64 statement of the same filter
(oracle/resample.py), and the loaders that use it.  Tolerance 2e-6 of the largest sample (float32 table and
accumulation over 130-560 taps).  Parity with the REFERENCE's soxr resampler is unpinned."""
import numpy as np
import pytest
import torch

from oracle import resample as orr
from softspoken_b200 import resample as rs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=64, mode="f16x3")
    yield eng
    eng.close()


@pytest.mark.parametrize("sr", [44100, 48000, 16000, 96000, 8000, 11025, 32000, 22050])
def test_kernel_matches_filter_definition(engine, sr):
    rng = np.random.default_rng(sr)
    for n in (0, 1, 7, 300, 4097, 50001):
        x = rng.normal(size=n).astype(np.float32)
        want = orr.resample(x, sr)
        got = engine.resample(torch.from_numpy(x).cuda(), sr).cpu().numpy()
        assert got.shape == want.shape == (rs.out_len(n, sr),) and got.dtype == np.float32
        if n:
            assert float(np.abs(got - want).max()) <= 2e-6 * max(1.0, float(np.abs(x).max())), (sr, n)


def test_int16_input_equals_float_input(engine):
    rng = np.random.default_rng(1)
    pcm = rng.integers(-32768, 32768, 30000).astype(np.int16)
    a = engine.resample(torch.from_numpy(pcm).cuda(), 48000)
    b = engine.resample(torch.from_numpy(pcm.astype(np.float32) / np.float32(32768)).cuda(), 48000)
    assert torch.equal(a, b)                    # int16 / 32768 is exact


def test_loaders_resample_files_at_other_rates(engine, tmp_path, capsys):
    """A 48 kHz and a 44.1 kHz stereo file through the corpus loader and the reference-shaped `load_audio`: right
    length, the audible band intact (compared with the same synthetic scene rendered at 22,050 Hz), detection runs."""
    from softspoken_b200 import corpus, synth, wavio
    from softspoken_b200.worker import load_audio
    dur = 12.0
    t48 = np.arange(int(dur * 48000)) / 48000
    t22 = np.arange(int(dur * 22050)) / 22050
    tones = [(0.3, 310.0), (0.2, 1250.0), (0.1, 5200.0)]

    def scene(t):
        return sum(a * np.sin(2 * np.pi * f * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * t)) for a, f in tones)
    p48 = str(tmp_path / "a48.wav")
    wavio.write_wav_pcm16(p48, wavio.encode_pcm16(scene(t48).astype(np.float32)), 48000)
    x = corpus.load_native_22050(p48, engine)
    assert x.dtype == np.float32 and len(x) == rs.out_len(len(t48), 48000) == len(t22)
    assert float(np.abs(x[3000:-3000] - scene(t22)[3000:-3000]).max()) < 2e-4      # PCM_16 quantisation dominates
    t44 = np.arange(int(dur * 44100)) / 44100
    st = np.stack([scene(t44), 0.5 * scene(t44)], axis=1).astype(np.float32)
    p44 = str(tmp_path / "b44.wav")
    wavio.write_wav_pcm16(p44, wavio.encode_pcm16(st), 44100)
    data, sr = load_audio(p44, engine=engine)
    assert sr == 22050 and len(data) == len(t22)
    assert float(np.abs(data[3000:-3000] - 0.75 * scene(t22)[3000:-3000]).max()) < 2e-4
    with pytest.raises(NotImplementedError):
        load_audio(p44)
    with pytest.raises(ValueError):
        corpus.load_mono_22050(p48)
    rows = corpus.detect_corpus([p48, p44], engine.detect_host_batch, load=lambda p: corpus.load_native_22050(p, engine))
    assert isinstance(rows, list)
