"""headless review step + exporters against files written by the reference's own classes.

`tests/golden/review_cases.json` was produced by `oracle/make_golden_review.py`, which drives the real
`ReviewDetectionsScreen` data methods and the real `review_exporter` transforms.  Every output file must match byte
for byte.  Pure host code: no GPU needed.
"""
import datetime
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

from softspoken_b200 import review

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "review_cases.json")) as _f:
    CASES = json.load(_f)
CLOCK0 = datetime.datetime(2026, 1, 2, 3, 4, 5)


def _durations(case):
    known = {k: v[0] / v[1] for k, v in case["durations"].items()}

    def duration_of(path):
        if path not in known:
            raise OSError(path)
        return known[path]
    return duration_of


def _run(case, tmp_path, project="golden"):
    det = tmp_path / f"{project}_detections.csv"
    det.write_text(case["detections_csv"])
    rev = tmp_path / f"{project}_review.csv"
    table = review.ReviewTable.open(str(det), str(rev))
    for row, text in sorted((int(k), v) for k, v in case["comments"].items()):
        table.comment(row, text)
    for tick, (row, flag) in enumerate(case["marks"]):
        table.label(row, bool(flag), CLOCK0 + datetime.timedelta(seconds=tick))
        review.save_review(table, str(rev), tmp_path, project, _durations(case))     # the reference saves per click
    if not case["marks"]:
        review.save_review(table, str(rev), tmp_path, project, _durations(case))
    out = {}
    for root, _, files in os.walk(tmp_path):
        for name in files:
            p = os.path.join(root, name)
            if p != str(det):
                with open(p, newline="") as f:
                    out[os.path.relpath(p, tmp_path)] = f.read()
    return out, table


@pytest.mark.parametrize("name", sorted(CASES))
def test_files_match_reference_bytes(name, tmp_path):
    case = CASES[name]
    got, _ = _run(case, tmp_path)
    assert sorted(got) == sorted(case["outputs"])
    for rel in sorted(case["outputs"]):
        assert got[rel] == case["outputs"][rel], rel


def test_minimum_length_filter_and_order():
    case = CASES["multi_mixed"]
    det = pd.read_csv(io.StringIO(case["detections_csv"]))
    table = review.ReviewTable.from_detections(det)
    long_enough = det[(det["end_time"] - det["start_time"]) > 0.1]
    assert 0 < len(table) == len(long_enough) < len(det)
    names = [r[table.columns.index("file_name")] for r in table.cells]
    starts = [float(r[table.columns.index("start_time")]) for r in table.cells]
    assert names == sorted(names)
    for a, b, na, nb in zip(starts, starts[1:], names, names[1:]):
        assert na != nb or a <= b
    assert any(s < 0 for s in starts), "a detection inside the leading pad keeps its negative start"
    assert all(len(r[table.columns.index("start_time")].split(".")[-1]) <= 3 for r in table.cells)


def test_resume_from_saved_review_is_a_fixed_point(tmp_path):
    """Opening a saved review file and saving again reproduces it (the reference reloads review_file on start)."""
    case = CASES["multi_mixed"]
    got, _ = _run(case, tmp_path)
    rev = tmp_path / "golden_review.csv"
    again = review.ReviewTable.open(None, str(rev))
    df = again.save(str(tmp_path / "second.csv"))
    assert (tmp_path / "second.csv").read_text() == got["golden_review.csv"]
    assert df["erase"].tolist() == pd.read_csv(rev)["erase"].tolist()


def test_erase_all_feeds_the_silencer(tmp_path):
    """detections -> review (all erase) -> the rows SilenceWorker selects (silencer.coerce_erase)."""
    from softspoken_b200 import silencer
    case = CASES["seed0_erase_all"]
    det = tmp_path / "d.csv"
    det.write_text(case["detections_csv"])
    rev = tmp_path / "r.csv"
    assert review.main([str(det), str(rev), "--erase-all"]) == 0
    df = silencer.coerce_erase(pd.read_csv(rev))
    assert len(df) and (df["erase"] == 1).all()
    ref = pd.read_csv(io.StringIO(case["outputs"]["golden_review.csv"]))
    for c in ("ID", "file_name", "start_time", "end_time", "erase"):
        assert df[c].tolist() == ref[c].tolist()


def test_table_edge_cases(tmp_path):
    empty = review.ReviewTable.open(None, None)
    assert len(empty) == 0 and empty.columns == review.REVIEW_COLUMNS
    df = empty.save(str(tmp_path / "e.csv"))
    assert (tmp_path / "e.csv").read_text() == ",".join(review.REVIEW_COLUMNS) + "\n" and len(df) == 0
    # a table without IDs gets 1..n; rows whose ID text is unusable are numbered on from the largest ID
    t = review.ReviewTable(pd.DataFrame({"file_path": ["/a", "/a"], "file_name": ["x.wav", "x.wav"],
                                        "start_time": [2.00049, 1.0], "end_time": [3.0, 1.5], "erase": [0, 1],
                                        "user_comment": [np.nan, "c"], "review_datetime": [np.nan, np.nan]}))
    assert t.columns[0] == "ID" and [r[0] for r in t.cells] == ["2", "1"]      # sorted by start; IDs follow their rows
    assert t.cells[0][t.columns.index("erase")] == "Yes" and t.cells[1][t.columns.index("start_time")] == "2.0"
    t.cells[1][0] = "oops"
    assert t.to_frame()["ID"].tolist() == [2, 3]
    with pytest.raises(IndexError):
        t.label(5, True)
    t.cells[0][t.columns.index("erase")] = "  yEs "
    assert t.to_frame()["erase"].tolist() == [1, 0]


def test_transform_contracts(tmp_path):
    df = pd.DataFrame({"file_name": ["a.wav"], "start_time": [1.0]})
    for tr in (review.AudacityTxtTransform(), review.KaleidoscopeCsvTransform(), review.RavenTxtTransform()):
        with pytest.raises(ValueError, match="missing column"):
            tr(df, base_dir=tmp_path, project_name="p")
    mgr = review.ReviewExportManager(df)
    mgr.register_transform(review.AudacityTxtTransform())
    with pytest.raises(KeyError):
        mgr.register_transform(review.AudacityTxtTransform())
    with pytest.raises(KeyError):
        mgr.export("nope", tmp_path)

    class Text(review.Transform):
        name, extension = "text", ".md"

        def __call__(self, df, **kw):
            return f"{len(df)} rows"
    mgr.register_transform(Text())
    assert mgr.export("text", tmp_path).read_text() == "1 rows" and (tmp_path / "review.md").exists()


def test_raven_reads_wav_headers(tmp_path):
    """Default duration source = RIFF header (frames / rate); unreadable files fall back to their last detection end."""
    from softspoken_b200 import wavio
    a, b = tmp_path / "a.wav", tmp_path / "b.wav"
    wavio.write_wav_pcm16(str(a), np.zeros(22050 * 2, np.int16), 22050)
    df = pd.DataFrame({"file_path": [str(tmp_path)] * 3, "file_name": ["a.wav", "b.wav", "a.wav"],
                       "start_time": [0.5, 0.25, 1.0], "end_time": [0.75, 4.5, 1.5]})
    review.RavenTxtTransform()(df, base_dir=tmp_path, project_name="p")
    rows = (tmp_path / "Raven Outputs" / "p" / "p.txt").read_text().splitlines()
    assert [r.split("\t")[3:5] for r in rows[1:]] == [["0.5", "0.75"], ["2.25", "6.5"], ["1.0", "1.5"]]
    assert (tmp_path / "Raven Outputs" / "p" / "p_listfile.txt").read_text() == f"{a}\n{b}\n"


needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/root/code"), reason="reference tree not present")


@needs_reference
def test_live_reference_agrees_on_a_fresh_case(tmp_path):
    """Not only the frozen cases: a differently seeded table through the real reference right now."""
    from oracle import make_golden_review as mg
    case = dict(detections_csv=mg.synthetic_detections(seed=11), marks=[(1, 1), (4, 0), (0, 1)],
                comments={"2": "x\ty"}, durations={"/corpus/siteB/rec_001.wav": [44100, 44100]})
    case["outputs"] = mg.run_reference(case["detections_csv"], case["marks"], {2: "x\ty"},
                                       {k: tuple(v) for k, v in case["durations"].items()})
    got, _ = _run(case, tmp_path)
    assert got == case["outputs"]


def test_config5_sized_review_table(tmp_path):
    """BASELINE config 5 size: 10,000 detections over 1,000 files through the review step and all three exporters
    (host code must stay linear: the reference appends rows one `df.loc[len(df)]` at a time)."""
    import time
    from softspoken_b200 import synth
    files, start, end = synth.synth_review_rows(10000, 1000, 600.0, 0)
    det = pd.DataFrame({"ID": np.arange(1, 10001), "file_path": [f"/corpus/site{int(f) % 7}" for f in files],
                        "file_name": [f"clip_{int(f):04d}.wav" for f in files], "start_time": start, "end_time": end,
                        "erase": 0, "user_comment": "", "review_datetime": ""})
    t0 = time.perf_counter()
    table = review.ReviewTable.from_detections(det)
    table.erase_all("2026-01-01 00:00:00")
    df = review.save_review(table, str(tmp_path / "r.csv"), tmp_path, "p", duration_of=lambda p: 600.0)
    dt = time.perf_counter() - t0
    assert len(df) == int(((end - start) > 0.1).sum()) and (df["erase"] == 1).all()
    names = df["file_name"].tolist()
    assert names == sorted(names) and dt < 20.0
    back = pd.read_csv(tmp_path / "r.csv")
    assert np.allclose(back["start_time"], df["start_time"]) and len(back) == len(df)
    raven = (tmp_path / "Raven Outputs" / "p" / "p.txt").read_text().splitlines()
    assert len(raven) == len(df) + 1
    listed = (tmp_path / "Raven Outputs" / "p" / "p_listfile.txt").read_text().splitlines()
    # Begin Time = position in the file + 600 s for every file listed before it
    first = raven[1].split("\t")
    assert abs(float(first[3]) - df["start_time"].iloc[0]) < 1e-6 and first[8] == listed[0]
    last = raven[-1].split("\t")
    assert abs(float(last[3]) - (600.0 * (len(listed) - 1) + df["start_time"].iloc[-1])) < 1e-5


def test_voice_activity_mirror_fails_loudly_without_an_engine():
    from softspoken_b200 import voice_activity
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        voice_activity.wav_to_spec(np.zeros(1000, np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        voice_activity.spectrogram_db(np.zeros(1000, np.float32))
