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
from concurrent.futures import ThreadPoolExecutor
from typing import Sequence, Tuple

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
        # PCM_16 files with the default reader / writer never leave the int16 domain (see _silence_pcm16)
        self._pcm16_route = reader is None and writer is None

    def run(self):
        erase_df = self.review_df[self.review_df['erase'] == 1]
        if erase_df.empty:
            self.signals.finished.emit()
            return
        groups = [(fpath, fname, rows) for (fpath, fname), rows in erase_df.groupby(['file_path', 'file_name'])]
        total_files = len(groups)
        self._files_done = 0

        def file_done(out_fullpath=None):
            if out_fullpath is not None:
                self.signals.fileComplete.emit(out_fullpath)
            self._files_done += 1
            self.signals.overallProgress.emit(int(self._files_done / total_files * 100))

        # PCM_16 route: the next file is read and the previous one written on two helper threads while this thread
        # drives the GPU (np.fromfile, file writes and the ctypes call all release the GIL).  At most one file waits
        # to be written, so fileComplete(k - 1) may follow fileStarted(k); completions stay in file order.
        pool = ThreadPoolExecutor(max_workers=2) if self._pcm16_route else None
        pending = []                                   # [(write future, output path)], at most one

        def read_ahead(k):
            if pool is None or k >= len(groups):
                return None
            return pool.submit(_read_pcm16_or_none, os.path.join(groups[k][0], groups[k][1]))

        def settle_write():
            while pending:
                fut, outp = pending.pop(0)
                err = fut.result()
                if err is not None:                                 # silencer_ui.py:999-1000
                    print(f"Error writing {outp}: {err}")
                file_done(outp)

        nxt = read_ahead(0)
        try:
            for k, (fpath, fname, group_rows) in enumerate(groups):
                if self.stop_requested:
                    break
                cur, nxt = nxt, read_ahead(k + 1)
                full_path = os.path.join(fpath, fname)
                self.signals.fileStarted.emit(full_path)
                base, ext = os.path.splitext(fname)
                out_fullpath = os.path.join(self.output_dir, f"{base}_silenced.wav")
                rows = list(zip(group_rows['start_time'].tolist(), group_rows['end_time'].tolist()))
                got = cur.result() if cur is not None else None
                if got is not None:
                    frames, sr = self._silence_pcm16(got, rows)
                    settle_write()
                    pending.append((pool.submit(_write_or_error, out_fullpath, frames, sr), out_fullpath))
                    continue
                settle_write()
                try:
                    audio_data, sr = self._read(full_path)
                except Exception as e:                                  # silencer_ui.py:961-966
                    print(f"Error loading {full_path}: {e}")
                    file_done()
                    continue
                if audio_data.ndim == 1:
                    audio_data = np.expand_dims(audio_data, axis=0)
                audio_data = np.ascontiguousarray(audio_data, dtype=np.float32)
                table = interval_table(rows, sr, audio_data.shape[0], audio_data.shape[1])
                if len(table):
                    self.engine.silence_host(audio_data, table)
                try:
                    self._write(out_fullpath, audio_data.T, sr)
                except Exception as e:                                  # silencer_ui.py:999-1000
                    print(f"Error writing {out_fullpath}: {e}")
                file_done(out_fullpath)
            settle_write()
        finally:
            if pool is not None:
                pool.shutdown(wait=True)
        self.signals.finished.emit()

    def _silence_pcm16(self, got, rows):
        """PCM_16 input -> PCM_16 output without the float32 detour.  The reference decodes to float32
        (`x / 32768`), zeroes slices and lets libsndfile encode again (`lrintf(x * 32767)`), so samples OUTSIDE the
        erase intervals change too (16383 -> 16382); `ss_silence_pcm16_host(requantize=1)` applies exactly that
        round trip per int16 sample on the device and zeroes the intervals: a quarter of the bytes over PCIe and no
        float32 arrays on the host, same file bytes as the float32 route (tests/test_gpu_silence.py, the `files`
        workload of tools/bench_aux.py).  `got` = `wavio.read_wav_pcm16(path)`; -> (int16 frames, sample rate)."""
        frames, sr = got
        frames = np.ascontiguousarray(frames)
        if not frames.flags.writeable:
            frames = frames.copy()
        n = frames.shape[0]
        ch = 1 if frames.ndim == 1 else frames.shape[1]
        table = []
        for st, et in rows:
            s, e = row_to_samples(st, et, sr, n)
            if e > s:
                table.append((s * ch, e * ch))          # interleaved frames: all channels of [s, e) are contiguous
        # requantisation touches every sample, so the kernel runs even when no interval survives clamping
        self.engine.silence_pcm16_host(frames, np.asarray(table, dtype=np.int64).reshape(-1, 2), requantize=True)
        return frames, sr

    def stop(self):
        self.stop_requested = True


def _read_pcm16_or_none(path: str):
    """(int16 frames, sr) of a PCM_16 file; None for any other format or an unreadable file (the caller then takes
    the float32 route, which owns the reference's error messages)."""
    try:
        return wavio.read_wav_pcm16(path)
    except Exception:
        return None


def _write_or_error(path: str, frames: np.ndarray, sr: int):
    try:
        wavio.write_wav_pcm16(path, frames, sr)
        return None
    except Exception as e:
        return e


def _load_native(path: str):
    """`librosa.load(path, sr=None, mono=False)` (silencer_ui.py:959): float32 at the native rate."""
    return wavio.read_wav(path)


def _write_pcm16(path: str, data: np.ndarray, sr: int) -> None:
    """`sf.write(path, data, sr)` (silencer_ui.py:998; WAV default subtype PCM_16).  libsndfile's exact
    float->short conversion could not be pinned in this image: parity is claimed for the
    float32 buffers handed to the writer, not for the encoded bytes."""
    wavio.write_wav_pcm16(path, wavio.encode_pcm16(data), sr)


def main(argv=None) -> int:
    """`python -m softspoken_b200.silencer review.csv out_dir`: the "Silence Voices" button without the GUI — the last
    step after `softspoken_b200.corpus` (detections) and `softspoken_b200.review` (review CSV).  Rows are selected and
    coerced as `SilenceVoicesScreen.load_review_data` does (silencer_ui.py:1098-1106)."""
    import argparse
    import time
    ap = argparse.ArgumentParser(description="zero the erase == 1 intervals of a review CSV and write <stem>_silenced.wav files")
    ap.add_argument("review_csv")
    ap.add_argument("output_dir")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    from . import checkpoint
    from .engine import Engine
    df = coerce_erase(pd.read_csv(args.review_csv))
    os.makedirs(args.output_dir, exist_ok=True)
    eng = Engine(checkpoint.synthetic_state_dict(0), args.device, max_batch=1)     # K7 needs no weights: any model will do
    worker = SilenceWorker(df, args.output_dir, engine=eng)
    done = []
    worker.signals.fileComplete.connect(done.append)
    t0 = time.perf_counter()
    worker.run()
    dt = time.perf_counter() - t0
    eng.close()
    n_rows = int((df["erase"] == 1).sum()) if "erase" in df.columns and len(df) else 0
    print(f"{n_rows} intervals silenced in {len(done)} files -> {args.output_dir} ({dt:.2f} s)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
