"""Headless `SilenceWorker`: the "Silence Voices" job of the reference without Qt.

Mirrors root/code/frontend/silencer_ui.py:918-1015 (`SilenceWorker.run`) and :1098-1106 (erase
coercion): rows with erase == 1, grouped by (file_path, file_name) in sorted key order; per row
`start = int(round(float(start_time) * sr))`, `end = int(round(float(end_time) * sr))` (Python's
round-half-even on the double product), both clamped to [0, n]; `audio[:, start:end] = 0.0`; output
`<stem>_silenced.wav` in one flat output_dir.  The zeroing runs in the K7 kernel; index computation,
grouping and file naming are host logic kept bit-compatible with the reference.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import pandas as pd

from . import wavio
from .worker import Signal


class SilenceWorkerSignals:
    def __init__(self):
        self.fileStarted = Signal(str)
        self.fileProgress = Signal(int)
        self.fileComplete = Signal(str)
        self.overallProgress = Signal(int)
        self.finished = Signal()


def coerce_erase(df: pd.DataFrame) -> pd.DataFrame:
    """`SilenceVoicesScreen.load_review_data` (silencer_ui.py:1098-1100)."""
    if not df.empty and 'erase' in df.columns:
        df['erase'] = pd.to_numeric(df['erase'], errors='coerce').fillna(0).astype(int)
    return df


def row_to_samples(start_time, end_time, sr: int, n: int) -> Tuple[int, int]:
    """silencer_ui.py:975-982."""
    st = float(start_time)
    et = float(end_time)
    start_index = int(round(st * sr))
    end_index = int(round(et * sr))
    start_index = max(0, min(start_index, n))
    end_index = max(0, min(end_index, n))
    return start_index, end_index


def interval_table(rows: Sequence[Tuple[float, float]], sr: int, channels: int, n: int, base: int = 0) -> np.ndarray:
    """Rows of one `(channels, n)` buffer stored at flat offset `base` -> int64 `[K,2]` element ranges
    (one per row and channel; empty and inverted rows dropped, as a Python slice would do nothing)."""
    out = []
    for st, et in rows:
        s, e = row_to_samples(st, et, sr, n)
        if e > s:
            for c in range(channels):
                out.append((base + c * n + s, base + c * n + e))
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


class SilenceWorker:
    """`SilenceWorker(review_df, output_dir).run()`; `engine` is a `softspoken_b200.engine.Engine`."""

    def __init__(self, review_df, output_dir, sr=44100, engine=None, reader=None, writer=None):
        self.signals = SilenceWorkerSignals()
        self.review_df = review_df
        self.output_dir = output_dir
        self.stop_requested = False
        if engine is None:
            raise RuntimeError("softspoken_b200.SilenceWorker needs an Engine (there is no CPU fallback)")
        self.engine = engine
        self._read = reader or _load_native
        self._write = writer or _write_pcm16
        # PCM_16 files with the default reader / writer never leave the int16 domain (see _run_pcm16)
        self._pcm16_route = reader is None and writer is None

    def run(self):
        erase_df = self.review_df[self.review_df['erase'] == 1]
        if erase_df.empty:
            self.signals.finished.emit()
            return
        grouped = erase_df.groupby(['file_path', 'file_name'])
        total_files = len(grouped)
        files_done = 0
        for (fpath, fname), group_rows in grouped:
            if self.stop_requested:
                break
            full_path = os.path.join(fpath, fname)
            self.signals.fileStarted.emit(full_path)
            if self._pcm16_route:
                done = self._run_pcm16(full_path, fname, group_rows)
                if done is not None:
                    if done:
                        self.signals.fileComplete.emit(done)
                    files_done += 1
                    self.signals.overallProgress.emit(int(files_done / total_files * 100))
                    continue
            try:
                audio_data, sr = self._read(full_path)
            except Exception as e:                                  # silencer_ui.py:961-966
                print(f"Error loading {full_path}: {e}")
                files_done += 1
                self.signals.overallProgress.emit(int(files_done / total_files * 100))
                continue
            if audio_data.ndim == 1:
                audio_data = np.expand_dims(audio_data, axis=0)
            audio_data = np.ascontiguousarray(audio_data, dtype=np.float32)
            rows = [(row['start_time'], row['end_time']) for _, row in group_rows.iterrows()]
            table = interval_table(rows, sr, audio_data.shape[0], audio_data.shape[1])
            if len(table):
                self.engine.silence_host(audio_data, table)
            base, ext = os.path.splitext(fname)
            out_fullpath = os.path.join(self.output_dir, f"{base}_silenced.wav")
            try:
                self._write(out_fullpath, audio_data.T, sr)
            except Exception as e:                                  # silencer_ui.py:999-1000
                print(f"Error writing {out_fullpath}: {e}")
            self.signals.fileComplete.emit(out_fullpath)
            files_done += 1
            self.signals.overallProgress.emit(int(files_done / total_files * 100))
        self.signals.finished.emit()

    def _run_pcm16(self, full_path: str, fname: str, group_rows):
        """PCM_16 input -> PCM_16 output without the float32 detour.  The reference decodes to float32
        (`x / 32768`), zeroes slices and lets libsndfile encode again (`lrintf(x * 32767)`), so samples OUTSIDE the
        erase intervals change too (16383 -> 16382); `ss_silence_pcm16_host(requantize=1)` applies exactly that
        round trip per int16 sample on the device and zeroes the intervals: a quarter of the bytes over PCIe and no
        float32 arrays on the host, same file bytes as the float32 route (tests/test_gpu_pcm16.py, the `files`
        workload of tools/bench_aux.py).  -> output path, "" when writing failed, None when the file is not
        PCM_16 or cannot be read here (the caller falls back to the float32 route and its error messages)."""
        try:
            got = wavio.read_wav_pcm16(full_path)
        except Exception:
            return None
        if got is None:
            return None
        frames, sr = got
        frames = np.ascontiguousarray(frames)
        if not frames.flags.writeable:
            frames = frames.copy()
        n = frames.shape[0]
        ch = 1 if frames.ndim == 1 else frames.shape[1]
        table = []
        for _, row in group_rows.iterrows():
            s, e = row_to_samples(row['start_time'], row['end_time'], sr, n)
            if e > s:
                table.append((s * ch, e * ch))          # interleaved frames: all channels of [s, e) are contiguous
        # requantisation touches every sample, so the kernel runs even when no interval survives clamping
        self.engine.silence_pcm16_host(frames, np.asarray(table, dtype=np.int64).reshape(-1, 2), requantize=True)
        base, _ = os.path.splitext(fname)
        out_fullpath = os.path.join(self.output_dir, f"{base}_silenced.wav")
        try:
            wavio.write_wav_pcm16(out_fullpath, frames, sr)
        except Exception as e:                                      # silencer_ui.py:999-1000
            print(f"Error writing {out_fullpath}: {e}")
            return ""
        return out_fullpath

    def stop(self):
        self.stop_requested = True


def _load_native(path: str):
    """`librosa.load(path, sr=None, mono=False)` (silencer_ui.py:959): float32 at the native rate."""
    return wavio.read_wav(path)


def _write_pcm16(path: str, data: np.ndarray, sr: int) -> None:
    """`sf.write(path, data, sr)` (silencer_ui.py:998; WAV default subtype PCM_16).  libsndfile's exact
    float->short conversion could not be pinned in this image: parity is claimed for the
    float32 buffers handed to the writer, not for the encoded bytes."""
    wavio.write_wav_pcm16(path, wavio.encode_pcm16(data), sr)
