"""PCM_16 sample path on the GPU: int16 in, same bits out.

* detection on the int16 samples of a PCM_16 clip == detection on `pcm / 32768` float32 (logits and regions bit for bit);
* `ss_decode_pcm16` == `sf.read(dtype='float32')` + `librosa.to_mono` as restated in oracle/silence.py;
* `ss_encode_pcm16` and the int16-domain silencing == the oracle's float round trip.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=32, mode="f16x3")
    yield eng
    eng.close()


def _pcm16_clip(seconds, seed):
    from softspoken_b200 import synth
    x = synth.synth_audio(seconds, seed)
    return np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)


def test_detect_pcm16_equals_float_route(engine):
    from oracle.silence import pcm16_to_float
    pcm = _pcm16_clip(21.3, 3)
    f32 = pcm16_to_float(pcm)
    reg_f, lg_f = engine.detect_host(f32, want_logits=True)
    reg_i, lg_i = engine.detect_host(pcm, want_logits=True)
    assert np.array_equal(lg_f, lg_i) and np.array_equal(reg_f, reg_i) and len(reg_f) > 0
    # device-resident and batched entry points, odd start alignment (slice at an odd sample offset)
    for off in (0, 1, 3):
        a = engine.detect_device(torch.from_numpy(pcm[off:]).cuda(), want_logits=True)
        b = engine.detect_device(torch.from_numpy(f32[off:]).cuda(), want_logits=True)
        assert torch.equal(a[2], b[2]) and int(a[1].item()) == int(b[1].item())
        assert torch.equal(a[0][: int(a[1].item())], b[0][: int(b[1].item())])
    clips16 = [pcm, pcm[: 22050 * 4], np.zeros(0, np.int16), pcm[7:50000]]
    got16 = engine.detect_host_batch(clips16)
    got32 = engine.detect_host_batch([pcm16_to_float(c) for c in clips16])
    mixed = engine.detect_host_batch([clips16[0], pcm16_to_float(clips16[1]), clips16[2], pcm16_to_float(clips16[3])])
    for a, b, c in zip(got16, got32, mixed):
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_streamed_pcm16_chunks(engine, sd_seed0):
    """A clip longer than one staging chunk (1,024 windows) streams int16 chunks with the same overlap logic."""
    from oracle.silence import pcm16_to_float
    from softspoken_b200.engine import Engine
    pcm = np.tile(_pcm16_clip(30.0, 5), 22)[: 22050 * 640]       # 640 s: 1,072 windows > one chunk
    eng = Engine(sd_seed0, 0, max_batch=64, mode="bf16")
    a = eng.detect_host(pcm, want_logits=True)
    b = eng.detect_host(pcm16_to_float(pcm), want_logits=True)
    assert a[1].shape[0] > 1024 and np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    eng.close()


@pytest.mark.parametrize("channels", [1, 2, 3, 6])
def test_decode_pcm16_matches_load_audio(engine, channels):
    from oracle.silence import load_audio_pcm16
    rng = np.random.default_rng(channels)
    n = 100_003
    fr = rng.integers(-32768, 32768, size=(n, channels) if channels > 1 else n, dtype=np.int16)
    fr.reshape(n, -1)[:4] = [[-32768] * channels, [32767] * channels, [0] * channels, [-1] * channels]
    want = load_audio_pcm16(fr)
    got = engine.decode_pcm16(torch.from_numpy(fr).cuda()).cpu().numpy()
    assert want.dtype == np.float32 and np.array_equal(want, got)
    got_odd = engine.decode_pcm16(torch.from_numpy(fr[1:]).cuda()).cpu().numpy()     # unaligned start
    assert np.array_equal(want[1:], got_odd)


def test_encode_and_silence_pcm16_match_oracle(engine):
    from oracle import silence as osil
    from softspoken_b200 import wavio
    from softspoken_b200.silencer import interval_table
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200_001) * 0.4).astype(np.float32)
    x[:8] = [1.0, -1.0, 0.5, -0.5, 16383.5 / 32767, 1.5, -1.5, 0.0]
    want = osil.float_to_pcm16(x)
    got = engine.encode_pcm16(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(want, got) and np.array_equal(want, wavio.encode_pcm16(x))
    # int16-domain "Silence Voices": (n, C) interleaved frames, rows in seconds
    sr, n, C = 22050, 150_001, 2
    fr = rng.integers(-32768, 32768, size=(n, C), dtype=np.int16)
    rows = [(0.0, 0.25), (1.2345, 2.0), (3.3, 3.3), (5.0, 4.0), (6.5, 99.0)]
    for requant in (True, False):
        want = osil.silence_pcm16(fr, sr, rows, requantize=requant)
        # interleaved frames: element range of a row = [s * C, e * C)
        iv = interval_table(rows, sr, 1, n) * C
        buf = fr.copy()
        engine.silence_pcm16_host(buf, iv, requantize=requant)
        assert np.array_equal(want, buf), requant
        dev = torch.from_numpy(fr.copy()).cuda()
        engine.silence_pcm16(dev, torch.from_numpy(iv), requantize=requant)
        assert np.array_equal(want, dev.cpu().numpy()), requant
    # the reference's read -> write round trip is not the identity: large samples lose one LSB (32767 != 32768)
    rt = osil.float_to_pcm16(osil.pcm16_to_float(np.array([16383, 16384, 16385, -16385, 32767, -32768], np.int16)))
    assert rt.tolist() == [16382, 16384, 16384, -16384, 32766, -32767]


def test_native_loader_and_corpus_route(engine, tmp_path):
    """A mono PCM_16 wav is loaded as int16 and detected without a host float32 copy; same rows as the float route."""
    from softspoken_b200 import corpus, wavio
    pcm = _pcm16_clip(12.0, 9)
    p = os.path.join(tmp_path, "a.wav")
    wavio.write_wav_pcm16(p, pcm, 22050)
    native = corpus.load_native_22050(p)
    assert native.dtype == np.int16 and np.array_equal(native, pcm)
    rows_i = corpus.detect_corpus([p], engine.detect_host_batch, load=corpus.load_native_22050)
    rows_f = corpus.detect_corpus([p], engine.detect_host_batch, load=corpus.load_mono_22050)
    assert rows_i == rows_f and len(rows_i) > 0
