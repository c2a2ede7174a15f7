"""Headless review step and exporters: detections CSV -> review CSV -> Audacity / Kaleidoscope /
Raven files, byte for byte what the reference's review screen writes.

The reference keeps this state in a Qt table: `ReviewDetectionsScreen` loads `<project>_detections.csv`, drops
detections no longer than `settings.minimum_detection_len` (review_detections.py:764-770), sorts by
(file_name, start_time), rounds the times to 3 decimals and turns every value into cell TEXT
(review_detections.py:970-996); every keep / erase click rewrites two cells and calls `save_review`
(review_detections.py:683-717, 93-172), which rebuilds a DataFrame FROM THE TEXT, writes `<project>_review.csv` and
runs three export transforms (review_exporter.py:129-481).  That text round trip decides the bytes on disk (times
are the repr of the 3-decimal values, `erase` is 1 only where the cell says "Yes", comments are strings and never
NaN), so `ReviewTable` keeps the same representation — a list of rows of strings — without any widget.

Between the detector (`softspoken_b200.corpus`, `worker.ProcessWorker`) and the silencer
(`silencer.SilenceWorker`, which reads the review CSV) this is the piece that makes the batch tool usable with no GUI:

    python -m softspoken_b200.review <project>_detections.csv <project>_review.csv --erase-all \\
        [--export-dir DIR --project NAME]

Pure host code: nothing here touches the GPU.  Parity is pinned by `tests/golden/review_cases.json`, produced by the
reference's own classes (`oracle/make_golden_review.py`).
"""
from __future__ import annotations

import argparse
import csv
import datetime as _dt
import io
import math
import os
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import pandas as pd

from . import settings

REVIEW_COLUMNS = ["ID", "file_path", "file_name", "start_time", "end_time", "erase", "user_comment", "review_datetime"]
STAMP_FORMAT = "%Y-%m-%d %H:%M:%S"        # review_detections.py:698


# ------------------------------------------------------------------------------------------------ cell text
def _is_missing(v) -> bool:
    return v is None or (isinstance(v, (float, np.floating)) and math.isnan(v)) or v is pd.NA or v is pd.NaT


def _cell_text(column: str, v) -> str:
    """What the reference shows for a value (review_detections.py:990-996): erase -> "Yes" or nothing, a missing
    value -> nothing, anything else -> str(value) (for a float that is its shortest repr)."""
    if column == "erase":
        return "Yes" if (not _is_missing(v) and v == 1) else ""
    if _is_missing(v):
        return ""
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    if isinstance(v, (np.integer,)):
        return str(int(v))
    return str(v)


def _id_first(df: pd.DataFrame) -> pd.DataFrame:
    """review_detections.py:62-71: an ID column exists and leads (1..n when it had to be made)."""
    if "ID" not in df.columns:
        out = df.copy()
        out.insert(0, "ID", range(1, len(out) + 1))
        return out
    return df[["ID"] + [c for c in df.columns if c != "ID"]].copy()


def _round_half_even_scaled(x: np.ndarray, decimals: int) -> np.ndarray:
    """DataFrame.round / Series.round on float64 = numpy.round: rint(x * 10^d) / 10^d (NOT Python's decimal-exact
    round: the two differ on a few values, and the files carry numpy's)."""
    return np.round(np.asarray(x, dtype=np.float64), decimals)


# ------------------------------------------------------------------------------------------------ the table
class ReviewTable:
    """The review screen's table, headless: `columns` and `cells[row][col]` hold the TEXT the reference's
    QTableWidget would hold.  Row order is the reference's (sorted by file_name, start_time; stable)."""

    def __init__(self, frame: pd.DataFrame):
        frame = _id_first(frame)
        if len(frame):
            # populate_table (review_detections.py:976-979): sort, renumber rows, round both time columns to 3 places
            frame = frame.sort_values(by=["file_name", "start_time"], ignore_index=True)
            for c in ("start_time", "end_time"):
                frame[c] = _round_half_even_scaled(frame[c].to_numpy(dtype=np.float64), 3)
        self.columns: List[str] = [str(c) for c in frame.columns]
        cols = [frame[c].tolist() for c in frame.columns]
        self.cells: List[List[str]] = [[_cell_text(name, col[r]) for name, col in zip(self.columns, cols)]
                                       for r in range(len(frame))]

    # -- loading (ReviewDetectionsScreen.__init__, review_detections.py:220-237)
    @classmethod
    def open(cls, detections_path: Optional[str], review_path: Optional[str] = None) -> "ReviewTable":
        """An existing review file wins (a review in progress is resumed as it was saved); otherwise the detector's
        CSV is loaded and detections of `end - start <= settings.minimum_detection_len` never reach the table;
        with neither file the table is empty but has the eight columns."""
        if review_path is not None and os.path.exists(review_path):
            return cls(pd.read_csv(review_path))
        if detections_path is not None and os.path.exists(detections_path):
            return cls.from_detections(pd.read_csv(detections_path))
        return cls(pd.DataFrame(columns=REVIEW_COLUMNS))

    @classmethod
    def from_detections(cls, df: pd.DataFrame) -> "ReviewTable":
        keep = (df["end_time"] - df["start_time"]) > settings.minimum_detection_len      # review_detections.py:770
        return cls(df[keep])

    def __len__(self) -> int:
        return len(self.cells)

    def _col(self, name: str) -> int:
        return self.columns.index(name)

    # -- the reviewer's actions
    def label(self, row: int, erase: bool, when: Union[str, _dt.datetime, None] = None) -> None:
        """Keep (erase=False) or erase (True) one detection: the erase cell becomes "Yes" / empty and the row is
        stamped as reviewed (review_detections.py:683-712).  `when` defaults to now."""
        if not 0 <= row < len(self.cells):
            raise IndexError(f"row {row} outside the table of {len(self.cells)} detections")
        if when is None:
            when = _dt.datetime.now()
        stamp = when if isinstance(when, str) else when.strftime(STAMP_FORMAT)
        self.cells[row][self._col("erase")] = "Yes" if erase else ""
        self.cells[row][self._col("review_datetime")] = stamp
        # the cell holds a COPY of the text; nothing else changes (ID, times, comment stay as typed)

    def comment(self, row: int, text: str) -> None:
        self.cells[row][self._col("user_comment")] = str(text)

    def erase_all(self, when: Union[str, _dt.datetime, None] = None) -> None:
        """Mark every detection for erasure — the unattended "Silence Voices" run.  One time stamp for all rows."""
        if when is None:
            when = _dt.datetime.now()
        for r in range(len(self.cells)):
            self.label(r, True, when)

    # -- saving (save_review, review_detections.py:93-172)
    def to_frame(self) -> pd.DataFrame:
        """The DataFrame `save_review` derives from the cell text: ID first, missing IDs numbered on from the
        largest one (review_detections.py:73-87), times parsed back to float64 (unparsable -> NaN), erase = 1 where
        the cell reads "yes" in any case and with any padding, every other column left as text."""
        data = {name: [row[j] for row in self.cells] for j, name in enumerate(self.columns)}
        df = _id_first(pd.DataFrame(data, columns=self.columns))
        ids = pd.to_numeric(df["ID"], errors="coerce").to_numpy(dtype=np.float64, copy=True)
        known = ids[~np.isnan(ids)]
        nxt = int(known.max()) + 1 if known.size else 1
        for i in np.flatnonzero(np.isnan(ids)):
            ids[i] = nxt
            nxt += 1
        df["ID"] = ids.astype(np.int64)
        for c in ("start_time", "end_time"):
            if c in df.columns:
                df[c] = pd.to_numeric(df[c], errors="coerce")
        if "erase" in df.columns:
            df["erase"] = [1 if t.strip().lower() == "yes" else 0 for t in df["erase"]]
        return df

    def save(self, review_path: str) -> pd.DataFrame:
        df = self.to_frame()
        with open(review_path, "w", newline="") as f:
            f.write(frame_text(df))
        return df


# ------------------------------------------------------------------------------------------------ text writers
def _field(v):
    if _is_missing(v):
        return ""
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    if isinstance(v, np.integer):
        return int(v)
    return v


def frame_text(df: pd.DataFrame, sep: str = ",") -> str:
    """`df.to_csv(index=False, sep=sep, lineterminator="\\n")` written directly: ints as ints, float64 with the
    shortest repr, missing values as empty fields, a field quoted only when it holds the separator, a quote or a
    line break (csv.QUOTE_MINIMAL, which is what pandas uses)."""
    buf = io.StringIO()
    w = csv.writer(buf, delimiter=sep, quoting=csv.QUOTE_MINIMAL, lineterminator="\n")
    w.writerow(list(df.columns))
    cols = [df[c].tolist() for c in df.columns]
    for r in range(len(df)):
        w.writerow([_field(col[r]) for col in cols])
    return buf.getvalue()


def wav_seconds(path: str) -> float:
    """frames / sample rate from the RIFF header (review_exporter.py:26-28 asks soundfile for the same two numbers)."""
    from . import wavio
    seconds, _ = wavio.duration_and_rate(path)
    return seconds


# ------------------------------------------------------------------------------------------------ transforms
class Transform:
    """One application-specific export (review_exporter.py:31-50).  Called with a copy of the review DataFrame and
    keyword options; returns a DataFrame (the manager writes it as CSV), str / bytes (written verbatim) or None
    (the transform wrote its own files)."""

    name = "unnamed"
    extension = ".csv"

    def __call__(self, df: pd.DataFrame, **kwargs):
        raise NotImplementedError


def _need(df: pd.DataFrame, who: str, columns: Iterable[str]) -> None:
    missing = set(columns) - set(df.columns)
    if missing:
        raise ValueError(f"{who}: DataFrame missing column(s): {missing}")


def _optional_text(df: pd.DataFrame, column: str) -> list:
    return df[column].tolist() if column in df.columns else [""] * len(df)


class AudacityTxtTransform(Transform):
    """`<base_dir>/Audacity Outputs/<project_name>/<wav stem>.txt`: one label track per wav FILE NAME (same-named
    wavs of different folders share one file, as in the reference), rows `start<TAB>end<TAB>comment`, times with
    `precision` decimals, ordered by start (review_exporter.py:129-213)."""

    name = "audacity"
    extension = ".txt"

    def __call__(self, df, *, base_dir, project_name, comment: str = "Human", precision: int = 6, **kwargs) -> None:
        out_root = Path(base_dir) / "Audacity Outputs" / project_name
        out_root.mkdir(parents=True, exist_ok=True)
        _need(df, "AudacityTxtTransform", ("file_name", "start_time", "end_time"))
        t = pd.DataFrame({"file_name": df["file_name"],
                          "start_time": pd.to_numeric(df["start_time"], errors="coerce"),
                          "end_time": pd.to_numeric(df["end_time"], errors="coerce")})
        t = t.sort_values(["file_name", "start_time"])
        tracks: Dict[str, List[str]] = {}
        for name, s, e in zip(t["file_name"].tolist(), t["start_time"].tolist(), t["end_time"].tolist()):
            if _is_missing(name):
                continue                                   # groupby drops rows without a key
            tracks.setdefault(name, []).append(f"{s:.{precision}f}\t{e:.{precision}f}\t{comment}")
        for name, lines in tracks.items():
            (out_root / f"{Path(name).stem}.txt").write_text("\n".join(lines) + "\n")
        return None


class KaleidoscopeCsvTransform(Transform):
    """`<base_dir>/Kaleidoscope Outputs/<project_name>/<project_name>.csv`: INDIR (common folder of all files),
    FOLDER (each file's folder below it), IN FILE*, OFFSET, DURATION, TOP1MATCH*, MANUAL ID (the reviewer's
    comment) plus end_time / erase / review_datetime for traceability (review_exporter.py:216-337)."""

    name = "kaleidoscope"
    extension = ".csv"

    def __call__(self, df, *, base_dir, project_name, precision: int = 6, human_label: str = "Human", **kwargs) -> None:
        out_root = Path(base_dir) / "Kaleidoscope Outputs" / project_name
        out_root.mkdir(parents=True, exist_ok=True)
        _need(df, "KaleidoscopeCsvTransform", ("file_path", "file_name", "start_time", "end_time"))
        start = pd.to_numeric(df["start_time"], errors="coerce").to_numpy(dtype=np.float64)
        end = pd.to_numeric(df["end_time"], errors="coerce").to_numpy(dtype=np.float64)
        folders_abs = [str(p) for p in df["file_path"].tolist()]
        indir = os.path.commonpath(folders_abs)            # ValueError on an empty table, as in the reference
        if not indir.endswith(os.sep):
            indir += os.sep
        below = [os.path.relpath(p, indir) for p in folders_abs]
        below = ["" if f == "." else f for f in below]
        if indir.endswith("\\"):                           # only a Windows-style separator is trimmed again
            indir = indir[:-1]
        out = pd.DataFrame({
            "INDIR": [indir] * len(df),
            "FOLDER": below,
            "IN FILE*": df["file_name"].tolist(),
            "OFFSET": _round_half_even_scaled(start, precision),
            "DURATION": _round_half_even_scaled(end - start, precision),
            "TOP1MATCH*": [human_label] * len(df),
            "MANUAL ID": _optional_text(df, "user_comment"),
            "end_time": _round_half_even_scaled(end, precision),
            "erase": _optional_text(df, "erase"),
            "review_datetime": _optional_text(df, "review_datetime"),
        })
        (out_root / f"{project_name}.csv").write_text(frame_text(out))
        return None


class RavenTxtTransform(Transform):
    """`<base_dir>/Raven Outputs/<project_name>/<project_name>_listfile.txt` (the wavs, in order of first appearance)
    and `<project_name>.txt` (tab-separated selection table).  Raven lays the listed files end to end, so a
    detection's Begin / End is its time in the file plus the lengths of the files listed before it; a file that
    cannot be opened counts with its largest detection end (review_exporter.py:340-481).  `duration_of(path)`
    returns seconds and raises when the file cannot be read (default: the RIFF header)."""

    name = "raven"
    extension = ".txt"

    def __init__(self, duration_of: Optional[Callable[[str], float]] = None):
        self.duration_of = duration_of or wav_seconds

    def __call__(self, df, *, base_dir, project_name, precision: int = 6, annotation_label: str = "Human",
                 low_freq: int = 0, high_freq: int = 8000, **kwargs) -> None:
        out_root = Path(base_dir) / "Raven Outputs" / project_name
        out_root.mkdir(parents=True, exist_ok=True)
        _need(df, "RavenTxtTransform", ("file_path", "file_name", "start_time", "end_time"))
        paths = [str(Path(fp) / fn) for fp, fn in zip(df["file_path"].tolist(), df["file_name"].tolist())]
        starts = [float(v) for v in df["start_time"].tolist()]
        ends_raw = df["end_time"].tolist()
        ends = [float(v) for v in ends_raw]
        listed = list(dict.fromkeys(paths))
        (out_root / f"{project_name}_listfile.txt").write_text("\n".join(listed) + "\n")
        offset_of: Dict[str, float] = {}
        total = 0.0
        for wav in listed:
            try:
                seconds = self.duration_of(wav)
            except Exception:
                own = [e for p, e in zip(paths, ends) if p == wav and not math.isnan(e)]
                seconds = max(own) if own else float("nan")
            offset_of[wav] = total
            total += seconds
        begin = _round_half_even_scaled(np.array([offset_of[p] + s for p, s in zip(paths, starts)], np.float64), precision)
        finish = _round_half_even_scaled(np.array([offset_of[p] + e for p, e in zip(paths, ends)], np.float64), precision)
        n = len(df)
        table = pd.DataFrame({
            "Selection": list(range(1, n + 1)),
            "View": ["Spectrogram 1"] * n,
            "Channel": [1] * n,
            "Begin Time (s)": begin,
            "End Time (s)": finish,
            "Low Freq (Hz)": [low_freq] * n,
            "High Freq (Hz)": [high_freq] * n,
            "Annotation": [annotation_label] * n,
            "Begin Path": paths,
            "erase": _optional_text(df, "erase"),
            "user_comment": _optional_text(df, "user_comment"),
            "review_datetime": _optional_text(df, "review_datetime"),
        })
        if "confidence" in df.columns:
            table["confidence"] = df["confidence"].tolist()
        (out_root / f"{project_name}.txt").write_text(frame_text(table, sep="\t"))
        return None


class ReviewExportManager:
    """Registry of transforms over one review DataFrame (review_exporter.py:53-126): `export(name, dst, **options)`
    runs one, `export_all(dst_dir, **options)` all of them; a transform's DataFrame / str / bytes result is written
    to `dst` (a directory gets `review<extension>`), None means it wrote its own files."""

    def __init__(self, df: pd.DataFrame):
        self.df = df
        self._registry: Dict[str, Transform] = {}

    def register_transform(self, transform: Transform) -> None:
        if transform.name in self._registry:
            raise KeyError(f"Transform '{transform.name}' already registered")
        self._registry[transform.name] = transform

    def transform(self, cls):
        self.register_transform(cls())
        return cls

    def export(self, name: str, dst, make_dirs: bool = True, **kwargs) -> Path:
        if name not in self._registry:
            raise KeyError(f"No transform named '{name}' registered")
        tr = self._registry[name]
        dst = Path(dst)
        if dst.is_dir():
            dst = dst / f"review{tr.extension}"
        if make_dirs:
            dst.parent.mkdir(parents=True, exist_ok=True)
        result = tr(self.df.copy(), **kwargs)
        if result is None:
            return dst
        if isinstance(result, pd.DataFrame):
            dst.write_text(frame_text(result))
        elif isinstance(result, str):
            dst.write_text(result)
        elif isinstance(result, bytes):
            dst.write_bytes(result)
        else:
            raise TypeError(f"Unsupported return type from transform ({type(result).__name__}).")
        return dst

    def export_all(self, dst_dir, **kwargs) -> Dict[str, Path]:
        return {name: self.export(name, dst_dir, **kwargs) for name in self._registry}


def save_review(table: ReviewTable, review_path: str, base_dir=None, project_name: Optional[str] = None,
                duration_of: Optional[Callable[[str], float]] = None) -> pd.DataFrame:
    """`save_review(persist=True)` (review_detections.py:136-168): write the review CSV, then — when `base_dir` and
    `project_name` are given — the three exports under `<base_dir>/{Audacity,Kaleidoscope,Raven} Outputs/<project_name>/`."""
    df = table.save(review_path)
    if base_dir is not None and project_name is not None:
        mgr = ReviewExportManager(df)
        mgr.register_transform(AudacityTxtTransform())
        mgr.register_transform(KaleidoscopeCsvTransform())
        mgr.register_transform(RavenTxtTransform(duration_of))
        for name in ("audacity", "kaleidoscope", "raven"):
            mgr.export(name, dst=".", base_dir=Path(base_dir), project_name=project_name)
    return df


def main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(description="detections CSV -> review CSV (+ Audacity / Kaleidoscope / Raven exports)")
    ap.add_argument("detections_csv")
    ap.add_argument("review_csv", help="resumed if it exists, as the review screen does")
    ap.add_argument("--erase-all", action="store_true", help="mark every detection for erasure (unattended silencing)")
    ap.add_argument("--export-dir", default=None, help="write the three export trees below this folder")
    ap.add_argument("--project", default=None, help="project name of the export trees (default: review file stem)")
    args = ap.parse_args(argv)
    table = ReviewTable.open(args.detections_csv, args.review_csv)
    if args.erase_all:
        table.erase_all()
    project = args.project or Path(args.review_csv).stem
    df = save_review(table, args.review_csv, args.export_dir, project if args.export_dir else None)
    print(f"{len(df)} detections, {int(sum(df['erase'])) if len(df) else 0} marked for erasure -> {args.review_csv}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
