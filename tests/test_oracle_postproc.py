"""Oracle post-processing vs goldens frozen from NNDetector / ProcessWorker / DetectionProject."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import postproc as pp
from softspoken_b200 import spec


def test_plan_windows_matches_reference():
    g = load_golden("plan.npz")
    for d, n in zip(g["durations"], g["n_windows"]):
        starts = pp.plan_windows(float(d))
        assert len(starts) == int(n), d
        assert np.array_equal(starts, np.arange(n) * 13230)
    assert len(pp.plan_windows(60.0)) == 105 and len(pp.plan_windows(600.0)) == 1005
    assert len(pp.plan_windows(86400.0)) == 144005


def test_all_windows_fit_in_padded_buffer():
    for n in list(range(0, 40000, 997)) + [1323000, 13230000]:
        L = n + 2 * 66150
        W = len(pp.plan_windows(n / 22050))
        assert W == 0 or (W - 1) * 13230 + 66150 <= L + 13229
        # the kernel only reads the first 65536 samples of a window: those always fit
        assert W == 0 or (W - 1) * 13230 + spec.WINDOW_SAMPLES_USED <= L + 13229


def test_window_position_integer_form():
    pos = pp.window_positions(200000)
    i = np.arange(200000, dtype=np.int64)
    assert np.array_equal(pos, (256 * i + 2) // 5)
    assert [spec.window_position(k) for k in range(10)] == [0, 51, 102, 154, 205, 256, 307, 358, 410, 461]


def test_gap_rule_in_bins_equals_string_rule():
    """float(next_start) - float(cur_end) <= 0.5  <=>  next_idx - cur_idx <= 42."""
    rng = np.random.default_rng(0)
    idx = np.concatenate([np.arange(0, 3000), rng.integers(0, 7_400_000, 20000)])
    for a in idx:
        for gap in (41, 42, 43, 44):
            b = a + gap
            lhs = float(pp.bin_time_str(int(b))) - float(pp.bin_time_str(int(a))) <= 0.5
            assert lhs == (gap <= spec.GAP_BINS), (a, gap)


def test_average_and_regions_seed0_match_reference():
    g = load_golden("postproc_seed0.npz")
    logits = load_golden("model_seed0.npz")["logits"]
    secs = int(g["n_padded"]) / 22050
    entries = pp.average_overlapping(logits, secs)
    assert len(entries) == len(g["avg_values"]) == 5581
    assert np.array_equal(np.array([v for v, _ in entries]), g["avg_values"])      # bit-exact float64
    assert [t for _, t in entries] == list(g["avg_times"])
    regions = pp.find_speech_regions(entries)
    assert [list(r) for r in regions] == g["regions"].tolist()
    # index-space form agrees with the string form
    avg, cnt = pp.average_idx(logits, secs)
    ridx = pp.find_speech_regions_idx(avg, cnt)
    assert [(pp.bin_time_str(s), pp.bin_time_str(e)) for s, e in ridx] == [tuple(r) for r in regions]
    assert pp.output_length(secs) == 5632


@pytest.mark.parametrize("name", ["tiny", "short", "mid"])
def test_adversarial_cases_match_reference(name):
    g = load_golden("postproc_cases.npz")
    lg, secs = g[f"{name}_logits"], float(g[f"{name}_secs"])
    entries = pp.average_overlapping(lg, secs)
    assert np.array_equal(np.array([v for v, _ in entries]), g[f"{name}_avg"])
    assert [t for _, t in entries] == list(g[f"{name}_times"])
    regions = pp.find_speech_regions(entries)
    assert [list(r) for r in regions] == g[f"{name}_regions"].tolist()
    avg, cnt = pp.average_idx(lg, secs)
    ridx = pp.find_speech_regions_idx(avg, cnt)
    assert [[pp.bin_time_str(s), pp.bin_time_str(e)] for s, e in ridx] == g[f"{name}_regions"].tolist()


def test_empty_predictions():
    g = load_golden("postproc_cases.npz")
    assert len(pp.average_overlapping(np.array([]), 6.0)) == int(g["empty_n"]) == 0
    assert pp.find_speech_regions([]) == []


def test_csv_rows_match_reference():
    """Rows + CSV text for two files appended to one project (ID continuation)."""
    want = open(os.path.join(GOLDEN, "detections_seed0.csv")).read()
    logits = load_golden("model_seed0.npz")["logits"]
    n_padded = int(load_golden("postproc_seed0.npz")["n_padded"])
    rows = pp.detect_file(logits, n_padded, "/data/clip_seed0.wav", next_id=1)
    text = pp.csv_text(rows)
    n = len(rows)
    assert n == 22
    assert text == "".join(want.splitlines(keepends=True)[:n + 1])
    assert rows[1]["start_time"] == 1.3827999999999996 and rows[0]["start_time"] == -3.0
