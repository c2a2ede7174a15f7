"""Freeze golden files for SURVEY.md 8(f2): detections CSV -> review CSV -> Audacity / Kaleidoscope / Raven exports.

TEST INFRASTRUCTURE (build container only: needs /root/reference).  Runs the REAL reference code, unmodified:

  * `ReviewDetectionsScreen` (root/code/frontend/review_detections.py) is lifted out of its module with `ast` (the
    module imports Qt widgets, matplotlib, librosa at top level) and its data methods are driven on an instance made
    with `object.__new__`: `filter_by_minimum_detection_len` (:764-770), `_ensure_id_column_first` (:62-71),
    `populate_table` (:970-1011), `apply_label_to_current_detection` (:683-717) and `save_review` (:93-172).  The
    QTableWidget they talk to is replaced by a 30-line table of strings (`_Table`), `datetime.now()` by a fixed clock.
  * `review_exporter` (root/code/frontend/review_exporter.py) is imported as is; its `soundfile.info` call
    (:26-28) sees a stub that knows the durations of the synthetic corpus and raises for unknown files, which
    exercises the transform's own fallback (:413-421).

Output: tests/golden/review_cases.json  {case: {detections_csv, marks, comments, durations, outputs{relpath: text}}}.

    python -m oracle.make_golden_review
"""
from __future__ import annotations

import ast
import io
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import pandas as pd

from . import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
CLOCK0 = (2026, 1, 2, 3, 4, 5)       # the fixed clock ticks one second per call


class _Item:
    def __init__(self, text=""):
        self._t = str(text)

    def text(self):
        return self._t

    def setBackground(self, *_):
        pass


class _Table:
    """The part of QTableWidget the reference's data path uses."""

    def __init__(self):
        self.cells, self.headers, self.rows, self.cols = {}, [], 0, 0

    def setRowCount(self, n):
        self.rows = n
        self.cells = {k: v for k, v in self.cells.items() if k[0] < n}

    def setColumnCount(self, n):
        self.cols = n

    def setHorizontalHeaderLabels(self, labels):
        self.headers = [str(x) for x in labels]

    def horizontalHeaderItem(self, c):
        return _Item(self.headers[c])

    def insertRow(self, i):
        assert i == self.rows, "the reference appends rows in order"
        self.rows += 1

    def setItem(self, r, c, item):
        self.cells[(r, c)] = item

    def item(self, r, c):
        return self.cells.get((r, c))

    def rowCount(self):
        return self.rows

    def columnCount(self):
        return self.cols

    def blockSignals(self, *_):
        pass


class _Clock:
    """datetime stand-in: datetime.datetime.now() advances one second per call from CLOCK0."""

    def __init__(self):
        import datetime as real
        self._real, self._n = real, 0
        self.datetime = self

    def now(self):
        t = self._real.datetime(*CLOCK0) + self._real.timedelta(seconds=self._n)
        self._n += 1
        return t


def load_screen_class(durations):
    """-> the reference's ReviewDetectionsScreen class (unmodified source), executable headlessly."""
    ref_shim.install_stubs()
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)

    def info(path):
        if path not in durations:
            raise RuntimeError(f"cannot open {path}")
        frames, sr = durations[path]
        return types.SimpleNamespace(frames=frames, samplerate=sr)

    sys.modules["soundfile"].info = info
    from root.code.backend import settings
    path = os.path.join(ref_shim.REFERENCE_ROOT, "root", "code", "frontend", "review_detections.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    ns = {"os": os, "io": io, "pd": pd, "np": np, "time": __import__("time"), "Path": Path, "settings": settings,
          "datetime": _Clock(), "QMainWindow": object, "QTableWidgetItem": _Item, "QColor": lambda *_: None,
          "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "ReviewDetectionsScreen":
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["ReviewDetectionsScreen"]


def run_reference(detections_csv: str, marks, comments, durations, project="golden"):
    """marks: [(row index in the sorted review table, erase_flag)] applied in order through the reference's
    apply_label_to_current_detection; comments: {row: text} typed into the user_comment cell beforehand.
    -> {relative output path: text}"""
    cls = load_screen_class(durations)
    with tempfile.TemporaryDirectory() as tmp:
        det_path = os.path.join(tmp, f"{project}_detections.csv")
        with open(det_path, "w", newline="") as f:
            f.write(detections_csv)
        scr = object.__new__(cls)
        scr.parent_app_screen = None
        scr.project_manager = types.SimpleNamespace(
            projects_folder=tmp,
            current_project={"name": project, "detections_file": det_path,
                             "review_file": os.path.join(tmp, f"{project}_review.csv")})
        # ReviewDetectionsScreen.__init__ (:220-237), first opening of a project
        scr.csv_data = pd.read_csv(det_path)
        scr.filter_by_minimum_detection_len()
        scr.csv_data = scr._ensure_id_column_first(scr.csv_data)
        scr.table = _Table()
        scr.populate_table()
        scr.scroll = lambda direction: None
        # pandas >= 3 refuses the reference's `csv_data.at[i, "review_datetime"] = "<text>"` (:698) on the all-NaN
        # float64 column read_csv makes of the empty field (pandas 2 up-casts it to object silently; the reference
        # pins no pandas version, requirements.txt).  Do that up-cast here; save_review builds its output from the
        # table cells, not from csv_data, so this changes no output byte.
        for col in ("user_comment", "review_datetime"):
            scr.csv_data[col] = scr.csv_data[col].astype(object)
        cc = scr.csv_data.columns.get_loc("user_comment")
        for row, text in sorted(comments.items()):
            scr.table.setItem(int(row), cc, _Item(text))
        if not marks:
            scr.save_review(persist=True)
        for row, flag in marks:
            scr.current_index = int(row)
            scr.apply_label_to_current_detection(int(flag))
        out = {}
        for root, _, files in os.walk(tmp):
            for name in files:
                p = os.path.join(root, name)
                if p == det_path:
                    continue
                with open(p, newline="") as f:
                    out[os.path.relpath(p, tmp)] = f.read()
        return out


def synthetic_detections(seed: int = 3) -> str:
    """A detections CSV as the detector writes it (worker.py:100-128 rows, silencer_ui.py:817 to_csv): four files in
    three folders (two share a file name), starts that dip into the leading pad (negative), detections at or
    below the 0.1 s review minimum, 4-decimal time strings minus 3 (long float reprs)."""
    rng = np.random.default_rng(seed)
    files = [("/corpus/siteA/day1", "rec_001.wav"), ("/corpus/siteA/day2", "rec_002.wav"),
             ("/corpus/siteB", "rec_001.wav"), ("/corpus/siteB", "zz last.wav")]
    rows, ident = [], 1
    for k, (fp, fn) in enumerate(files):
        t = -2.9
        for j in range(int(rng.integers(5, 9))):
            t += float(rng.uniform(0.6, 40.0)) if (j or k != 2) else 1.2      # third file: first detection at -1.7 s
            s_bin = int(round((t + 3) * 256 / 3))
            length = int(rng.choice([0, 4, 8, 9, 30, 200, 900]))
            if k == 2 and j == 0:
                length = 200
            s = float(f"{s_bin / (256 / 3):.4f}") - 3
            e = float(f"{(s_bin + length) / (256 / 3):.4f}") - 3
            rows.append((ident, fp, fn, s, e, 0, "", ""))
            ident += 1
            t = e
    df = pd.DataFrame(rows, columns=["ID", "file_path", "file_name", "start_time", "end_time", "erase",
                                     "user_comment", "review_datetime"])
    return df.to_csv(index=False)


def cases():
    with open(os.path.join(GOLDEN, "detections_seed0.csv"), newline="") as f:
        seed0 = f.read()
    d0 = pd.read_csv(io.StringIO(seed0))
    n0 = int(((d0['end_time'] - d0['start_time']) > 0.1).sum())      # rows that reach the review table
    multi = synthetic_detections()
    durations = {"/corpus/siteA/day1/rec_001.wav": (13230000, 22050), "/corpus/siteA/day2/rec_002.wav": (4410000, 44100),
                 "/corpus/siteB/rec_001.wav": (1323001, 22050)}          # "zz last.wav" cannot be opened: fallback
    return {
        # every detection of the 60 s golden clip marked for erasure (the headless "Silence Voices" default)
        "seed0_erase_all": dict(detections_csv=seed0, marks=[(i, 1) for i in range(n0)], comments={}, durations={}),
        # nothing reviewed yet: save_review straight after the first load
        "multi_unreviewed": dict(detections_csv=multi, marks=[], comments={}, durations=durations),
        # a mixed review: erase, keep, a re-marked row, comments (one with a comma and quotes), unreviewed rows
        "multi_mixed": dict(detections_csv=multi, marks=[(0, 1), (1, 0), (2, 1), (5, 1), (2, 0), (9, 1), (14, 1)],
                            comments={3: "wind, not voice", 9: 'says "hello"', 5: "loud"}, durations=durations),
    }


def main():
    out = {}
    for name, c in cases().items():
        files = run_reference(c["detections_csv"], c["marks"], c["comments"], c["durations"])
        out[name] = dict(c, comments={str(k): v for k, v in c["comments"].items()},
                         durations={k: list(v) for k, v in c["durations"].items()}, outputs=files)
        print(name, sorted(files))
    with open(os.path.join(GOLDEN, "review_cases.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
