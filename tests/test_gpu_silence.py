"""K7 on the GPU: bit-exact silenced buffers vs the real SilenceWorker goldens and the oracle."""
import hashlib

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import load_golden
from test_oracle_silence import _inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=1, mode="fp32")
    yield eng
    eng.close()


def test_silence_worker_matches_reference(engine):
    from softspoken_b200.silencer import SilenceWorker, coerce_erase
    g = load_golden("silence_cases.npz")
    store = _inputs(g)
    df = pd.DataFrame({"file_path": g["rows_path"], "file_name": g["rows_name"], "start_time": g["rows_start"],
                       "end_time": g["rows_end"], "erase": g["rows_erase"]})
    written = {}
    sw = SilenceWorker(coerce_erase(df), "/out", engine=engine,
                       reader=lambda p: (store[p][0].copy(), store[p][1]),
                       writer=lambda p, a, sr: written.__setitem__(p, (np.array(a), sr)))
    sw.run()
    assert list(written) == list(g["out_paths"])
    for i, (path, (a, sr)) in enumerate(written.items()):
        assert a.shape == tuple(g[f"out{i}_shape"]) and sr == int(g[f"out{i}_sr"])
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == str(g[f"out{i}_sha256"])
    assert len(sw.signals.finished.log) == 1 and len(sw.signals.fileComplete.log) == 3


def test_silence_device_property_random(engine):
    """Random interval tables (unaligned, overlapping, empty, out of range) on a packed corpus buffer."""
    rng = np.random.default_rng(4)
    for n in [1, 3, 4, 5, 17, 4096, 100003]:
        buf = rng.normal(size=n).astype(np.float32)
        buf[buf == 0] = 1.0
        k = int(rng.integers(0, 40))
        b = rng.integers(-5, n + 5, k)
        e = b + rng.integers(-3, max(4, n // 3), k)
        want = buf.copy()
        for bb, ee in zip(b, e):
            lo, hi = max(0, int(bb)), min(n, int(ee))
            if hi > lo:
                want[lo:hi] = 0.0
        # device entry point, buffer at an odd element offset so heads/tails are exercised
        big = torch.zeros(n + 3, device="cuda")
        view = big[1:1 + n]
        view.copy_(torch.from_numpy(buf))
        engine.silence(view, torch.from_numpy(np.stack([b, e], 1).astype(np.int64)) if k else torch.zeros((0, 2), dtype=torch.int64))
        assert np.array_equal(view.cpu().numpy(), want), n
        assert float(big[0]) == 0.0 and float(big[n + 1]) == 0.0
        # host entry point
        h = buf.copy()
        engine.silence_host(h, np.stack([b, e], 1).astype(np.int64) if k else np.zeros((0, 2), np.int64))
        assert np.array_equal(h, want), n


def test_config5_shape_intervals(engine):
    """SURVEY §8d config 5 in miniature: intervals over a packed multi-clip buffer vs the oracle."""
    from oracle import silence as osil
    from softspoken_b200 import synth
    from softspoken_b200.silencer import interval_table
    n_files, clip_n, sr = 6, 22050 * 20, 22050
    files, start, end = synth.synth_review_rows(60, n_files, clip_s=20.0, seed=0)
    rng = np.random.default_rng(9)
    corpus = rng.normal(0, 0.1, (n_files, clip_n)).astype(np.float32)
    want = np.stack([osil.silence_buffer(corpus[f], sr, [(s, e) for ff, s, e in zip(files, start, end) if ff == f])[0]
                     for f in range(n_files)])
    table = np.concatenate([interval_table([(s, e) for ff, s, e in zip(files, start, end) if ff == f], sr, 1, clip_n,
                                           base=f * clip_n) for f in range(n_files)])
    dev = torch.from_numpy(corpus).cuda().reshape(-1)
    engine.silence(dev, torch.from_numpy(table))
    assert np.array_equal(dev.cpu().numpy().reshape(n_files, clip_n), want)


def test_silence_worker_pcm16_files_route_equals_float32_route(engine, tmp_path):
    """PCM_16 wav files take the int16-domain route (no float32 detour); the files it writes must be byte-identical
    to those of the float32 route (decode /32768 -> zero slices -> encode lrintf(x * 32767)), mono and stereo,
    including a file whose only row clamps to nothing (it is still re-encoded) and rows past the end."""
    from softspoken_b200 import wavio
    from softspoken_b200.silencer import SilenceWorker, _load_native, _write_pcm16
    rng = np.random.default_rng(12)
    sr = 8000
    src = tmp_path / "in"
    src.mkdir()
    mono = rng.integers(-32768, 32768, 40001).astype(np.int16)
    mono[:8] = [32767, -32768, 16383, 16384, -16383, 1, -1, 0]
    stereo = rng.integers(-32768, 32768, (30000, 2)).astype(np.int16)
    wavio.write_wav_pcm16(str(src / "m.wav"), mono, sr)
    wavio.write_wav_pcm16(str(src / "s.wav"), stereo, sr)
    wavio.write_wav_pcm16(str(src / "untouched.wav"), mono[:5000], sr)
    wavio.write_wav_float32(str(src / "f.wav"), (mono[:6000] / 32768.0).astype(np.float32), sr)   # not PCM_16: falls back
    rows = [("m.wav", 0.0004, 0.75), ("m.wav", 2.5, 99.0), ("m.wav", 1.0, 1.0), ("s.wav", 0.1234, 0.5678),
            ("s.wav", 3.7, 3.9), ("untouched.wav", 50.0, 60.0), ("f.wav", 0.1, 0.2)]
    df = pd.DataFrame({"file_path": [str(src)] * len(rows), "file_name": [r[0] for r in rows],
                       "start_time": [r[1] for r in rows], "end_time": [r[2] for r in rows], "erase": 1})
    out_a, out_b = tmp_path / "a", tmp_path / "b"
    out_a.mkdir(), out_b.mkdir()
    fast = SilenceWorker(df, str(out_a), engine=engine)
    assert fast._pcm16_route
    fast.run()
    slow = SilenceWorker(df, str(out_b), engine=engine, reader=_load_native, writer=_write_pcm16)
    assert not slow._pcm16_route
    slow.run()
    names = sorted(p.name for p in out_a.iterdir())
    assert names == ["f_silenced.wav", "m_silenced.wav", "s_silenced.wav", "untouched_silenced.wav"]
    for name in names:
        assert (out_a / name).read_bytes() == (out_b / name).read_bytes(), name
    got, _ = wavio.read_wav_pcm16(str(out_a / "m_silenced.wav"))
    assert not got[3:6000].any() and got[int(2.5 * sr):].max() == 0 and got[6001:int(2.5 * sr)].any()
    assert [len(w.signals.fileComplete.log) for w in (fast, slow)] == [4, 4]


def test_silencer_cli(tmp_path, capsys):
    """`python -m softspoken_b200.silencer review.csv out_dir` end to end on two small files."""
    from softspoken_b200 import silencer, wavio
    rng = np.random.default_rng(3)
    src = tmp_path / "in"
    src.mkdir()
    a = rng.integers(-20000, 20000, 30000).astype(np.int16)
    wavio.write_wav_pcm16(str(src / "a.wav"), a, 22050)
    wavio.write_wav_pcm16(str(src / "b.wav"), a[::-1].copy(), 22050)
    pd.DataFrame({"ID": [1, 2, 3], "file_path": [str(src)] * 3, "file_name": ["a.wav", "b.wav", "a.wav"],
                  "start_time": [0.1, 0.2, 0.9], "end_time": [0.3, 0.4, 1.0], "erase": [1, "junk", 1],
                  "user_comment": "", "review_datetime": ""}).to_csv(tmp_path / "r.csv", index=False)
    assert silencer.main([str(tmp_path / "r.csv"), str(tmp_path / "out")]) == 0
    assert "2 intervals silenced in 1 files" in capsys.readouterr().out
    got, _ = wavio.read_wav_pcm16(str(tmp_path / "out" / "a_silenced.wav"))
    want = wavio.encode_pcm16(a.astype(np.float32) / np.float32(32768))
    want[2205:6615] = 0
    want[19845:22050] = 0
    assert np.array_equal(got, want) and not (tmp_path / "out" / "b_silenced.wav").exists()
