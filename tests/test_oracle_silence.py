"""Oracle silencing vs goldens frozen from the real SilenceWorker.run."""
import hashlib

import numpy as np

from conftest import load_golden
from oracle import silence as osil


def tone(n, c=None):
    k = np.arange(n if c is None else n * c, dtype=np.int64)
    a = (((k * 7919) % 2003) / 2003.0 - 0.5 + 1e-3).astype(np.float32)
    return a if c is None else a.reshape(c, n)


def _inputs(g):
    out = {}
    i = 0
    while f"in{i}_path" in g:
        shape = tuple(int(v) for v in g[f"in{i}_shape"])
        a = tone(shape[0]) if len(shape) == 1 else tone(shape[1], shape[0])
        out[str(g[f"in{i}_path"])] = (a, int(g[f"in{i}_sr"]))
        i += 1
    return out


def test_silence_matches_reference():
    g = load_golden("silence_cases.npz")
    store = _inputs(g)
    groups = osil.group_rows(g["rows_path"], g["rows_name"], g["rows_start"], g["rows_end"], g["rows_erase"])
    assert [f"/out/{fn[:-4]}_silenced.wav" for (_, fn) in groups] == list(g["out_paths"])
    for i, ((fp, fn), rows) in enumerate(groups.items()):
        audio, sr = store[f"{fp}/{fn}"]
        out = osil.silence_buffer(audio, sr, rows).T          # (samples, channels) as sf.write receives
        assert out.shape == tuple(g[f"out{i}_shape"])
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g[f"out{i}_sha256"])
        z = (out == 0.0).all(axis=1).astype(np.int8)
        edges = np.flatnonzero(np.diff(np.concatenate([[0], z, [0]]))).reshape(-1, 2)
        assert np.array_equal(edges, g[f"out{i}_zero_runs"])


def test_round_half_even_and_clamp():
    assert osil.interval_to_samples(0.0000227, 0.0000680, 22050, 50000) == (1, 1)
    assert osil.interval_to_samples(0.5 / 8000, 1.5 / 8000, 8000, 100) == (0, 2)      # 0.5 -> 0, 1.5 -> 2
    assert osil.interval_to_samples(2.5 / 8000, 3.5 / 8000, 8000, 100) == (2, 4)
    assert osil.interval_to_samples(-5.0, 1e9, 8000, 100) == (0, 100)


def test_erase_coercion():
    g = load_golden("erase_coercion.npz")
    raw = [None if v == "nan" else v for v in g["raw"]]
    assert np.array_equal(osil.coerce_erase(raw), g["coerced"])
