"""K5 / K6 on the GPU: bit-exact against the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from softspoken_b200 import spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=1, mode="fp32")
    yield eng
    eng.close()


def _gpu_avg(engine, logits, secs):
    from oracle import postproc as pp
    out_len = pp.output_length(secs)
    lg = torch.from_numpy(np.ascontiguousarray(logits.reshape(-1, 256))).cuda() if logits.size else torch.empty(0)
    avg, cnt = engine.average(lg, out_len)
    return avg.cpu().numpy(), cnt.cpu().numpy()


def test_average_bit_exact_seed0(engine):
    from oracle import postproc as pp
    g = load_golden("postproc_seed0.npz")
    logits = load_golden("model_seed0.npz")["logits"]
    secs = int(g["n_padded"]) / 22050
    avg, cnt = _gpu_avg(engine, logits, secs)
    keep = cnt >= 1
    assert np.array_equal(avg[keep], g["avg_values"])            # float64, bit for bit
    assert [pp.bin_time_str(i) for i in np.nonzero(keep)[0]] == list(g["avg_times"])
    assert np.isnan(avg[~keep]).all() and cnt.max() == 5


@pytest.mark.parametrize("name", ["tiny", "short", "mid"])
def test_cases_bit_exact(engine, name):
    from oracle import postproc as pp
    g = load_golden("postproc_cases.npz")
    lg, secs = g[f"{name}_logits"], float(g[f"{name}_secs"])
    avg, cnt = _gpu_avg(engine, lg, secs)
    keep = cnt >= 1
    assert np.array_equal(avg[keep], g[f"{name}_avg"])
    reg = engine.regions(torch.from_numpy(avg).cuda(), torch.from_numpy(cnt).cuda())
    got = [[pp.bin_time_str(s), pp.bin_time_str(e)] for s, e in reg]
    assert got == g[f"{name}_regions"].tolist()


def test_regions_property_random(engine):
    """Random timelines, thresholds and gaps vs the oracle's sequential scan (incl. empty / all-hot)."""
    from oracle import postproc as pp
    rng = np.random.default_rng(0)
    for trial in range(120):
        n = int(rng.choice([1, 2, 43, 255, 1024, 1025, 4096, 4097, 5000, 70001]))
        p_hot = float(rng.choice([0.0, 0.01, 0.2, 0.5, 0.97, 1.0]))
        gap = int(rng.choice([0, 1, 2, 14, 15, 16, 31, 32, 33, 42, 43, 300, 4096]))      # 15 = where the two-test fast path starts
        thr = float(rng.choice([0.1, 0.0, -0.3]))
        avg = np.where(rng.random(n) < p_hot, thr + rng.random(n) + 1e-9, thr - rng.random(n))
        avg[rng.random(n) < 0.05] = thr                       # exactly at threshold: not hot
        n_cov = n if trial % 3 else int(n * 0.8)
        cnt = np.zeros(n, np.int32)
        cnt[:n_cov] = rng.integers(1, 6, n_cov)
        want = pp.find_speech_regions_idx(avg, cnt, gap_bins=gap, threshold=thr)
        got = engine.regions(torch.from_numpy(avg).cuda(), torch.from_numpy(cnt).cuda(), thr, gap, cap=max(16, n))
        assert np.array_equal(got.astype(np.int64), want), (trial, n, p_hot, gap)


def test_average_property_random(engine):
    from oracle import postproc as pp
    rng = np.random.default_rng(1)
    for W in [1, 2, 5, 6, 33, 400]:
        lg = (rng.normal(0, 1, (W, 1, 256)) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
        n_padded = (W - 1) * 13230 + 66150 + int(rng.integers(0, 13230))
        secs = n_padded / 22050
        want_avg, want_cnt = pp.average_idx(lg, secs)
        avg, cnt = _gpu_avg(engine, lg, secs)
        assert np.array_equal(cnt, want_cnt)
        assert np.array_equal(avg[cnt > 0], want_avg[want_cnt > 0])


def test_empty_inputs(engine):
    avg, cnt = engine.average(torch.empty(0), 512)
    assert int(cnt.sum()) == 0 and bool(torch.isnan(avg).all())
    assert engine.regions(avg, cnt).shape == (0, 2)


def test_plan_and_timeline_host_helpers():
    """ss_plan_windows / ss_timeline_bins restate NNDetector.py:72-77,168 on integer sample counts."""
    from oracle import postproc as pp
    from softspoken_b200.engine import plan_windows, timeline_bins
    rng = np.random.default_rng(2)
    for n in list(range(0, 3000)) + [int(v) for v in rng.integers(0, 2_000_000_000, 3000)] + [1323000, 13230000, 1905120000]:
        assert plan_windows(n) == len(pp.plan_windows(n / 22050)), n
        assert timeline_bins(n + 132300) == pp.output_length((n + 132300) / 22050), n
