"""Oracle: window planning, overlap averaging, region finding, CSV rows.
TEST INFRASTRUCTURE.

Each function restates one reference function statement by statement, keeping
its floating-point evaluation order (the results are compared bit for bit):

  plan_windows            <- NNDetector.plan_detection_job        NNDetector.py:55-82
  pad_audio               <- ProcessWorker.run (padding)           worker.py:57-62
  average_overlapping     <- NNDetector.average_overlapping_detections  NNDetector.py:153-190
  find_speech_regions     <- NNDetector.find_speech_regions       NNDetector.py:103-143
  detection_rows          <- ProcessWorker.run (row building)      worker.py:97-125
  csv_text                <- DetectionProject.save_detections      silencer_ui.py:775-817

plus index-space equivalents (`*_idx`) that return integer bin indices — the
form the CUDA kernels emit — and are proven equal to the string-time form in
tests/test_oracle_postproc.py.

Pinned by tests/test_oracle_postproc.py against tests/golden/postproc_*.npz /
detections_*.csv frozen from the real reference by oracle/make_golden.py.
"""
from __future__ import annotations

import io
import math
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

SAMPLE_RATE = 22050          # settings.py:16
STEP_SIZE = 0.6              # settings.py:9
THRESHOLD = 0.1              # settings.py:13
WINDOW_S = 3                 # NNDetector.py:68
BATCH = 32                   # settings.py:12

CSV_COLUMNS = ["ID", "file_path", "file_name", "start_time", "end_time", "erase",
               "user_comment", "review_datetime"]     # silencer_ui.py:779-788


def plan_windows(audio_len_seconds: float) -> np.ndarray:
    """NNDetector.py:65-80."""
    sample_rate = SAMPLE_RATE
    window_size = WINDOW_S
    audio_data_length = round(audio_len_seconds * sample_rate) + (window_size * 2 * sample_rate)
    samples_per_window = sample_rate * window_size
    samples_per_step = math.floor(sample_rate * STEP_SIZE)
    num_windows = int(np.ceil((audio_data_length - samples_per_window) / samples_per_step))
    return np.arange(num_windows) * samples_per_step


def pad_audio(audio: np.ndarray) -> np.ndarray:
    """worker.py:58-62: three seconds of zeros on both sides."""
    padding_samples = SAMPLE_RATE * 3
    padded = np.zeros(len(audio) + 2 * padding_samples, dtype=audio.dtype)
    padded[padding_samples:padding_samples + len(audio)] = audio
    return padded


def window_positions(n_windows: int) -> np.ndarray:
    """NNDetector.py:172,175: `int(round(i * 0.6 / (3 / 256)))`, evaluated in double."""
    time_resolution = 3 / 256
    return np.array([int(round(i * STEP_SIZE / time_resolution)) for i in range(n_windows)],
                    dtype=np.int64)


def output_length(audio_length_seconds: float) -> int:
    """NNDetector.py:168."""
    return int(round(audio_length_seconds * 256 / 3))


def average_sums(logits: np.ndarray, audio_length_seconds: float) -> Tuple[np.ndarray, np.ndarray]:
    """The two accumulators of NNDetector.py:168-177 (float64 both, as np.zeros gives)."""
    n = output_length(audio_length_seconds)
    s = np.zeros(n)
    c = np.zeros(n)
    logits = np.asarray(logits)
    if logits.size:
        pos = window_positions(len(logits))
        for i, w in enumerate(logits):
            p = int(pos[i])
            s[p:p + 256] += w.reshape(-1)           # float32 -> float64 add, in window order
            c[p:p + 256] += 1
    return s, c


def average_overlapping(logits: np.ndarray, audio_length_seconds: float, min_count: int = 1):
    """-> [(np.float64 avg, 'sss.ssss')] for bins with count >= min_count (NNDetector.py:179-186)."""
    s, c = average_sums(logits, audio_length_seconds)
    out = []
    for idx in np.nonzero(c >= min_count)[0]:
        out.append((s[idx] / c[idx], f"{idx / (256 / 3):.4f}"))
    return out


def average_idx(logits: np.ndarray, audio_length_seconds: float):
    """Index-space form: (avg float64[out_len] with NaN where uncovered, count int32[out_len])."""
    s, c = average_sums(logits, audio_length_seconds)
    with np.errstate(invalid="ignore", divide="ignore"):
        avg = s / c
    return avg, c.astype(np.int32)


def find_speech_regions(entries: Sequence[Tuple[float, str]], break_duration: float = 0.5,
                        threshold: float = THRESHOLD) -> List[Tuple[str, str]]:
    """NNDetector.py:109-141 for one file (string times, inclusive end)."""
    regions = []
    start_time = None
    end_time = None
    for detection, time in entries:
        if detection > threshold:
            if start_time is None:
                start_time = time
            end_time = time
        elif start_time is not None:
            regions.append((start_time, end_time))
            start_time = None
    if start_time is not None:
        regions.append((start_time, end_time))
    if not regions:
        return []
    merged = []
    current = regions[0]
    for nxt in regions[1:]:
        if float(nxt[0]) - float(current[1]) <= break_duration:
            current = (current[0], nxt[1])
        else:
            merged.append(current)
            current = nxt
    merged.append(current)
    return merged


def find_speech_regions_idx(avg: np.ndarray, count: np.ndarray, gap_bins: int = 42,
                            threshold: float = THRESHOLD) -> np.ndarray:
    """Index-space form -> int64 `[R,2]` of (start_bin, end_bin), end inclusive.

    Uncovered bins (count == 0) are skipped, not treated as gaps — the
    reference scans only the emitted entries (NNDetector.py:182-186,117)."""
    idxs = np.nonzero(count >= 1)[0]
    hot = avg[idxs] > threshold
    runs: List[List[int]] = []
    open_run = False
    for j, h in zip(idxs, hot):
        if h:
            if not open_run:
                runs.append([int(j), int(j)])
                open_run = True
            runs[-1][1] = int(j)
        else:
            open_run = False
    merged: List[List[int]] = []
    for r in runs:
        if merged and r[0] - merged[-1][1] <= gap_bins:
            merged[-1][1] = r[1]
        else:
            merged.append(list(r))
    return np.asarray(merged, dtype=np.int64).reshape(-1, 2)


def bin_time_str(idx: int) -> str:
    """NNDetector.py:185."""
    return f"{idx / (256 / 3):.4f}"


def regions_to_times(regions: Sequence[Tuple[str, str]]) -> List[Tuple[float, float]]:
    """worker.py:100: remove the 3 s pad offset."""
    return [(float(s) - 3, float(e) - 3) for (s, e) in regions]


def detection_rows(file: str, times: Sequence[Tuple[float, float]], next_id: int = 1) -> List[dict]:
    """worker.py:103-125."""
    rows = []
    for (st, et) in times:
        rows.append({"ID": next_id, "file_path": os.path.dirname(file), "file_name": os.path.basename(file),
                     "start_time": st, "end_time": et, "erase": 0, "user_comment": "",
                     "review_datetime": ""})
        next_id += 1
    return rows


def csv_text(rows: Sequence[dict]) -> str:
    """Text `DetectionProject.save_detections` writes for these rows
    (silencer_ui.py:779-817; rows appended with `df.loc[len(df)] = row`, worker.py:125)."""
    import pandas as pd
    column_types = {"ID": "int64", "file_path": str, "file_name": str, "start_time": str,
                    "end_time": str, "erase": int, "user_comment": str,
                    "review_datetime": "datetime64[ns]"}
    df = pd.DataFrame(columns=column_types.keys()).astype(column_types)
    for r in rows:
        df.loc[len(df)] = r
    buf = io.StringIO()
    df.to_csv(buf, index=False)
    return buf.getvalue()


def detect_file(logits: np.ndarray, n_padded: int, file: str, next_id: int = 1):
    """Everything after the network for one file (worker.py:89-125)."""
    secs = n_padded / SAMPLE_RATE
    entries = average_overlapping(logits, secs)
    regions = find_speech_regions(entries, break_duration=0.5)
    times = regions_to_times(regions)
    return detection_rows(file, times, next_id)


def min_threshold_margin(avg: np.ndarray, count: np.ndarray, threshold: float = THRESHOLD) -> float:
    """Smallest |avg - threshold| over covered bins: how far the oracle's decisions
    are from flipping under a perturbation of the logits."""
    cov = count >= 1
    if not cov.any():
        return float("inf")
    return float(np.min(np.abs(avg[cov] - threshold)))
