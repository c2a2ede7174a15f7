"""Headless `ProcessWorker` / `DetectionProject`: the detection job of the reference without Qt.

Mirrors root/code/backend/worker.py:21-139 (per-file loop, 3 s zero padding, batches of 32 window
starts, averaging, region finding, -3 s shift, row building with running IDs, CSV save after every
file, cooperative stop) and root/code/frontend/silencer_ui.py:775-817 (`DetectionProject`).  Qt signals
are replaced by a tiny in-process `Signal` with the same `connect` / `emit` surface.

Two execution strategies give identical rows:
  * `fast=True` (default) — `detector.detect_file(audio)`: one call per file, everything on the GPU;
  * `fast=False` — the reference's own call sequence `process_batch` -> `average_overlapping_detections`
    -> `find_speech_regions`, batch by batch (any object with those three methods works, as in the
    reference).
"""
from __future__ import annotations

import os
from os.path import basename, dirname
from typing import Callable, Dict, List

import numpy as np
import pandas as pd

from . import settings, wavio

COLUMN_TYPES = {
    'ID': 'int64', 'file_path': str, 'file_name': str, 'start_time': str, 'end_time': str,
    'erase': int, 'user_comment': str, 'review_datetime': 'datetime64[ns]',
}                                                   # silencer_ui.py:779-788


class Signal:
    def __init__(self, *types):
        self._slots: List[Callable] = []
        self.log: List[tuple] = []

    def connect(self, fn: Callable) -> None:
        self._slots.append(fn)

    def emit(self, *args) -> None:
        self.log.append(args)
        for fn in self._slots:
            fn(*args)


class WorkerSignals:
    def __init__(self):
        self.fileProgressChanged = Signal(float)
        self.overallProgressChanged = Signal(float)
        self.fileStarted = Signal(str)
        self.fileDone = Signal(str)
        self.finished = Signal()
        self.message = Signal(str)


def load_audio(path: str, engine=None):
    """`voice_activity.load_audio(path)` (root/code/backend/voice_activity.py:32-69): float32, mono, 22,050 Hz.

    Decode -> `(n,)` or `(C, n)`; more than one channel is averaged (`librosa.to_mono`); a decode error
    prints and returns `(None, None)` like the reference.  A file at another rate is resampled on the device
    (`engine.resample`, K9) — with this package's documented filter, NOT the reference's soxr (absent here), so for
    such files the detections are not covered by the bit-exactness claims; without an engine it is refused."""
    try:
        data, sr = wavio.read_wav(path)
    except Exception as e:                                     # voice_activity.py:39-41
        print(f'EXCEPTION EXCEPTION EXCEPTION: \n\t{path}\n\t{str({e})}')
        return (None, None)
    if data.ndim > 1:
        data = np.mean(data, axis=0)                           # librosa.to_mono
    if sr != settings.vad_resample:
        if engine is None:
            raise NotImplementedError(
                f"{path}: sample rate {sr} != {settings.vad_resample} and no engine was given to resample it "
                "(there is no CPU resampler)")
        data = resample_on_device(np.ascontiguousarray(data, dtype=np.float32), sr, engine)
        sr = settings.vad_resample                             # the reference returns (data, target rate) too (:66-67)
    return (data, sr)


def resample_on_device(data: np.ndarray, sr: int, engine) -> np.ndarray:
    """Host mono clip at `sr` -> float32 at 22,050 Hz through the K9 kernel."""
    import torch
    x = torch.from_numpy(np.ascontiguousarray(data)).to(engine.device)
    return engine.resample(x, int(sr)).cpu().numpy()


class DetectionProject:
    """silencer_ui.py:775-817.  `project_settings.current_project['detections_file']` names the CSV."""

    def __init__(self, project_settings):
        self.settings = project_settings
        self.columns = COLUMN_TYPES.keys()
        detections_path = self.settings.current_project['detections_file']
        if os.path.exists(detections_path):
            self.df = pd.read_csv(detections_path)
            if 'ID' not in self.df.columns:
                self.df.insert(0, 'ID', range(1, len(self.df) + 1))
            else:
                self.df['ID'] = pd.to_numeric(self.df['ID'], errors='coerce')
                missing_ids = self.df['ID'].isna()
                if missing_ids.any():
                    current_max = self.df['ID'].dropna().max()
                    start_id = int(current_max) if not np.isnan(current_max) else 0
                    for offset, idx in enumerate(self.df.index[missing_ids], start=start_id + 1):
                        self.df.at[idx, 'ID'] = offset
                self.df['ID'] = self.df['ID'].astype('int64')
            if 'review_datetime' in self.df.columns:
                self.df['review_datetime'] = pd.to_datetime(self.df['review_datetime'], errors='coerce')
            self.df = self.df.reindex(columns=self.columns).astype(COLUMN_TYPES)
        else:
            self.df = pd.DataFrame(columns=self.columns).astype(COLUMN_TYPES)

    def save_detections(self):
        self.df.to_csv(self.settings.current_project['detections_file'], index=False)


class ProcessWorker:
    def __init__(self, detector, detection_project, planned_work, parent=None, fast: bool = True):
        self.signals = WorkerSignals()
        self.detector = detector
        self.detection_project = detection_project
        self.planned_work = planned_work
        self.stop_requested = False
        self.fast = fast and hasattr(detector, "detect_file")

    def stop(self):
        self.stop_requested = True

    def _regions_reference_sequence(self, file, audio_data, indexes):
        """worker.py:58-100 with the detector's three reference methods."""
        sample_rate = settings.vad_resample
        padding_samples = sample_rate * 3
        padded = np.zeros(len(audio_data) + 2 * padding_samples, dtype=audio_data.dtype)
        padded[padding_samples:padding_samples + len(audio_data)] = audio_data
        audio_data = padded
        total_work_count = len(indexes)
        batch_predictions = []
        for start_idx in range(0, total_work_count, settings.prediction_batch_size):
            if self.stop_requested:
                return None
            end_idx = min(start_idx + settings.prediction_batch_size, total_work_count)
            speech_pred, mask_pred = self.detector.process_batch(audio_data, indexes[start_idx:end_idx])
            batch_predictions.append(mask_pred)
            self.signals.fileProgressChanged.emit((end_idx / total_work_count) * 100.0)
        audio_length_seconds = len(audio_data) / sample_rate
        if len(batch_predictions) > 0:
            avg = self.detector.average_overlapping_detections({file: np.vstack(batch_predictions)}, audio_length_seconds)
        else:
            avg = self.detector.average_overlapping_detections({file: np.array([])}, audio_length_seconds)
        speech_regions = self.detector.find_speech_regions({file: avg}, break_duration=0.5)
        return [(float(start) - 3, float(end) - 3) for (start, end) in speech_regions[file]]

    def run(self):
        total_files = len(self.planned_work)
        files_done = 0
        for ii, file in enumerate(self.planned_work.keys()):
            if self.stop_requested:
                break
            self.signals.fileStarted.emit(file)
            audio_data, original_sr = load_audio(file, engine=getattr(getattr(self.detector, 'model', None), 'engine', None))
            if audio_data is None:
                # the reference crashes here with a TypeError (worker.py:57-60); keep the
                # exception type but say why
                raise TypeError(f"object of type 'NoneType' has no len(): {file} could not be decoded")
            indexes = self.planned_work[file]
            if self.fast:
                regions = self.detector.detect_file(audio_data)
                self.signals.fileProgressChanged.emit(100.0)
            else:
                regions = self._regions_reference_sequence(file, audio_data, indexes)
            if self.stop_requested or regions is None:
                break
            append_rows(self.detection_project, file, regions)
            self.detection_project.save_detections()
            self.signals.fileDone.emit(file)
            files_done += 1
            self.signals.overallProgressChanged.emit((files_done / total_files) * 100.0)
        self.signals.finished.emit()


def next_detection_id(df: pd.DataFrame) -> int:
    """worker.py:107-112."""
    next_id = 1
    if not df.empty and 'ID' in df.columns:
        existing_max = pd.to_numeric(df['ID'], errors='coerce').max()
        if not np.isnan(existing_max):
            next_id = int(existing_max) + 1
    return next_id


def detection_rows(file: str, regions, next_id: int) -> List[dict]:
    """worker.py:103-124."""
    rows = []
    file_path, file_name = dirname(file), basename(file)
    for (start_time, end_time) in regions:
        rows.append({'ID': next_id, 'file_path': file_path, 'file_name': file_name,
                     'start_time': start_time, 'end_time': end_time, 'erase': 0,
                     'user_comment': '', 'review_datetime': ''})
        next_id += 1
    return rows


def append_rows(detection_project, file: str, regions) -> None:
    """worker.py:103-125: rows appended one by one with `df.loc[len(df)] = row` (so that the frame's dtypes
    evolve exactly as in the reference and `to_csv` prints the same text)."""
    df = detection_project.df
    for row in detection_rows(file, regions, next_detection_id(df)):
        df.loc[len(df)] = row
