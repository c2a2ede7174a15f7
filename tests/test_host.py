"""Host-side logic (no GPU): checkpoint packing, wav I/O, row/CSV building, silence indices, sharding."""
import os
import struct
import types

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN, load_golden
from softspoken_b200 import checkpoint, spec, synth, wavio


def test_blob_layout_roundtrip(sd_seed0):
    blob = checkpoint.pack_blob(sd_seed0)
    magic, version, n, _ = struct.unpack_from("<IIII", blob, 0)
    assert magic == checkpoint.BLOB_MAGIC and version == checkpoint.BLOB_VERSION
    table = {}
    for i in range(n):
        name, off, cnt = struct.unpack_from("<48sQQ", blob, 16 + 64 * i)
        table[name.rstrip(b"\0").decode()] = (off, cnt)
    payload = np.frombuffer(blob, dtype=np.float32, offset=16 + 64 * n)
    assert all(off % 4 == 0 for off, _ in table.values())        # 16-byte aligned entries
    folded = checkpoint.fold_state_dict(sd_seed0)
    off, cnt = table["conv6.c1.w"]
    w = payload[off:off + cnt].reshape(9, 256, 96)               # [tap][C_in][C_out]
    ref = folded["conv6.c1.w"].numpy()                           # [C_out, C_in, 3, 3]
    assert np.array_equal(w[4], ref[:, :, 1, 1].T) and np.array_equal(w[2], ref[:, :, 0, 2].T)
    off, cnt = table["conv_flatten.w"]
    assert np.array_equal(payload[off:off + cnt].reshape(128, 32, 4)[5, 7], sd_seed0["conv_flatten.weight"][:, 7, 5, 0].numpy())
    off, cnt = table["mel_taps"]
    assert cnt == 1469


def test_strict_state_dict_validation(sd_seed0):
    bad = dict(sd_seed0)
    bad.pop("conv8.conv1.0.weight")
    with pytest.raises(RuntimeError, match="Missing key"):
        checkpoint.validate_state_dict(bad)
    bad = dict(sd_seed0)
    bad["conv8.conv1.0.weight"] = torch.zeros(3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        checkpoint.validate_state_dict(bad)
    bad = dict(sd_seed0)
    bad["mel_spectrogram.mel_scale.fb"] = torch.ones(1025, 128)   # dense bank: band wider than the kernel walks
    with pytest.raises(ValueError):
        checkpoint.pack_blob(bad)


def test_checkpoint_file_format(sd_seed0, tmp_path):
    p = str(tmp_path / "model_checkpoint.pth")
    checkpoint.save_checkpoint(sd_seed0, p, epoch=3)
    ck = torch.load(p, map_location="cpu", weights_only=True)      # exactly how the reference loads it
    assert set(ck) == {"model_state_dict", "epoch"} and ck["epoch"] == 3
    assert list(ck["model_state_dict"]) == [k for k, _, _ in checkpoint.state_dict_spec()]
    assert checkpoint.normalise_model_path(".\\root\\models\\x\\model.pth") == "./root/models/x/model.pth"


def test_wav_roundtrip_and_header(tmp_path):
    pcm = synth.synth_pcm16(1.5, 2)
    p = str(tmp_path / "a.wav")
    wavio.write_wav_pcm16(p, pcm, 22050)
    x, sr = wavio.read_wav(p)
    assert sr == 22050 and x.dtype == np.float32 and np.array_equal(x, synth.synth_audio(1.5, 2))
    assert wavio.duration_and_rate(p) == (len(pcm) / 22050, 22050)
    st = np.stack([pcm, -pcm], axis=1)
    p2 = str(tmp_path / "st.wav")
    wavio.write_wav_pcm16(p2, st, 44100)
    y, sr2 = wavio.read_wav(p2)
    assert y.shape == (2, len(pcm)) and sr2 == 44100 and np.array_equal(y[0], x)
    p3 = str(tmp_path / "f.wav")
    wavio.write_wav_float32(p3, x, 8000)
    z, _ = wavio.read_wav(p3)
    assert np.array_equal(z, x)
    with open(str(tmp_path / "bad.wav"), "wb") as f:
        f.write(b"not a wav")
    with pytest.raises(wavio.WavError):
        wavio.read_wav(str(tmp_path / "bad.wav"))
    from softspoken_b200.worker import load_audio
    assert load_audio(str(tmp_path / "bad.wav")) == (None, None)      # voice_activity.py:39-41
    mono, _ = load_audio(p)
    assert np.array_equal(mono, x)
    with pytest.raises(NotImplementedError):
        load_audio(p2)                                                 # 44,100 Hz needs the resampler (f1)


def test_rows_and_csv_from_golden_regions(tmp_path):
    """Host row building + DetectionProject CSV == the text the reference wrote, given the reference's regions."""
    from softspoken_b200.detector import bin_time_str, plan_windows_from_duration, region_bins_to_times
    from softspoken_b200.worker import DetectionProject, append_rows
    g = load_golden("postproc_seed0.npz")
    want = open(os.path.join(GOLDEN, "detections_seed0.csv")).read().splitlines(keepends=True)
    bins = np.array([[round(float(s) * 256 / 3), round(float(e) * 256 / 3)] for s, e in g["regions"]])
    assert [[bin_time_str(a), bin_time_str(b)] for a, b in bins] == g["regions"].tolist()
    proj = DetectionProject(types.SimpleNamespace(current_project={"detections_file": str(tmp_path / "d.csv")}))
    append_rows(proj, "/data/clip_seed0.wav", region_bins_to_times(bins))
    proj.save_detections()
    assert open(str(tmp_path / "d.csv")).read() == "".join(want[:23])
    p = load_golden("plan.npz")
    for d, n in zip(p["durations"], p["n_windows"]):
        assert len(plan_windows_from_duration(float(d))) == int(n)


def test_silence_index_rule_matches_oracle():
    from oracle import silence as osil
    from softspoken_b200.silencer import coerce_erase, interval_table, row_to_samples
    rng = np.random.default_rng(0)
    for _ in range(2000):
        sr = int(rng.choice([8000, 22050, 44100, 48000]))
        n = int(rng.integers(1, 10 * sr))
        st = float(np.round(rng.uniform(-1, 11), int(rng.integers(0, 8))))
        et = st + float(rng.uniform(-0.5, 4))
        assert row_to_samples(st, et, sr, n) == osil.interval_to_samples(st, et, sr, n)
    # rows: normal, inverted (dropped), clamped to the end, fully beyond the end (dropped); 2 channels at base 100
    t = interval_table([(0.5, 1.0), (2.0, 1.0), (4.0, 99.0), (9.0, 99.0)], 1000, 2, 5000, base=100)
    assert t.tolist() == [[600, 1100], [5600, 6100], [4100, 5100], [9100, 10100]]
    g = load_golden("erase_coercion.npz")
    df = coerce_erase(pd.DataFrame({"erase": [None if v == "nan" else v for v in g["raw"]]}))
    assert np.array_equal(df["erase"].to_numpy(), g["coerced"])


def test_shard_files_lpt():
    from softspoken_b200.dist import shard_files
    assert shard_files([600.0] * 8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]      # equal durations: round-robin
    sh = shard_files([10, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    assert sorted(sum(sh, [])) == list(range(11)) and sh[0] == [0]
    assert shard_files([], 3) == [[], [], []]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from softspoken_b200.dist import gather_detections, shard_files
    files = [f"/data/f{i}.wav" for i in range(7)]
    mine = shard_files([600.0 - i for i in range(7)], world)[rank]
    rng = np.random.default_rng(0)
    per_file = {i: np.sort(rng.integers(0, 5000, (i % 4, 2)), axis=1) for i in range(7)}   # file 0 and 4: no regions
    local = np.concatenate([np.concatenate([np.full((len(per_file[i]), 1), i), per_file[i]], 1) for i in mine]
                           + [np.zeros((0, 3), np.int64)]).astype(np.int32)
    out = gather_detections(local)
    if rank == 0:
        want = np.concatenate([np.concatenate([np.full((len(per_file[i]), 1), i), per_file[i]], 1)
                               for i in range(7)]).astype(np.int32)
        q.put(bool(np.array_equal(out, want)))
    else:
        q.put(out is None)
    dist.destroy_process_group()


def test_gather_detections_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert res == [True, True]


def test_rows_from_triplets_assigns_ids_in_file_order():
    from softspoken_b200.dist import rows_from_triplets
    trip = np.array([[0, 10, 20], [2, 5, 6], [2, 100, 300]], np.int32)
    rows = rows_from_triplets(["/d/a.wav", "/d/b.wav", "/e/c.wav"], trip)
    assert [r["ID"] for r in rows] == [1, 2, 3]
    assert [r["file_name"] for r in rows] == ["a.wav", "c.wav", "c.wav"]
    assert rows[0]["start_time"] == float("0.1172") - 3


def _corpus_worker(rank, world, port, q, files, durations):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from softspoken_b200 import corpus

    def fake_load(path):                      # "audio" = the file's index, so that detections depend on the file
        return np.full(4, float(files.index(path)), np.float32)

    def fake_detect(clips):                   # deterministic regions per file, some files without any
        out = []
        for c in clips:
            i = int(c[0])
            rng = np.random.default_rng(i)
            out.append(np.sort(rng.integers(0, 51000, (i % 5, 2)), axis=1).astype(np.int32))
        return out

    rows = corpus.detect_corpus(files, fake_detect, load=fake_load, durations=durations, group_size=3)
    if rank == 0:
        # the direct text writer (no row dicts) prints the same bytes
        text = corpus.detect_corpus(files, fake_detect, load=fake_load, durations=durations, group_size=3, as_csv=True)
        assert text == corpus.csv_text(rows)
    elif world > 1:
        corpus.detect_corpus(files, fake_detect, load=fake_load, durations=durations, group_size=3, as_csv=True)
    q.put(corpus.csv_text(rows) if rank == 0 else (rows is None))
    if world > 1:
        dist.destroy_process_group()


def test_corpus_driver_same_csv_for_one_and_two_ranks():
    """BASELINE config 3 host logic: per-file sharding + gather + ID assignment give byte-identical CSV text for
    world sizes 1 and 2, and the direct CSV writer prints what DetectionProject's DataFrame prints."""
    import types
    import torch.multiprocessing as mp
    from softspoken_b200 import corpus
    from softspoken_b200.worker import DetectionProject, append_rows
    files = [f"/data/dir{i % 3}/clip, {i}.wav" if i == 4 else f"/data/dir{i % 3}/clip{i}.wav" for i in range(11)]
    durations = [600.0 - 7 * (i % 4) for i in range(11)]
    ctx = mp.get_context("spawn")
    texts = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = 29500 + (os.getpid() + 17 * world) % 2000
        procs = [ctx.Process(target=_corpus_worker, args=(r, world, port, q, files, durations)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(60)
        texts[world] = [r for r in res if isinstance(r, str)][0]
        assert all(r is True for r in res if not isinstance(r, str))
    assert texts[1] == texts[2]
    assert texts[1].splitlines()[0] == "ID,file_path,file_name,start_time,end_time,erase,user_comment,review_datetime"
    # the DataFrame route of the reference prints the same text
    import tempfile
    from softspoken_b200.detector import region_bins_to_times
    with tempfile.TemporaryDirectory() as d:
        proj = DetectionProject(types.SimpleNamespace(current_project={"detections_file": os.path.join(d, "x.csv")}))
        for i, f in enumerate(files):
            rng = np.random.default_rng(i)
            bins = np.sort(rng.integers(0, 51000, (i % 5, 2)), axis=1).astype(np.int32)
            append_rows(proj, f, region_bins_to_times(bins))
        proj.save_detections()
        assert open(os.path.join(d, "x.csv")).read() == texts[1]


def test_pcm16_host_helpers(tmp_path):
    """PCM_16 sample path, host side: the raw int16 reader returns the stored samples, the float decode of the oracle
    equals `wavio.read_wav` (+ channel mean), and the host twin of the encode kernel equals the oracle's restatement
    of libsndfile's float -> short conversion."""
    from oracle import silence as osil
    from softspoken_b200 import corpus, wavio
    rng = np.random.default_rng(1)
    for ch in (1, 2):
        fr = rng.integers(-32768, 32768, size=(5000, ch) if ch > 1 else 5000, dtype=np.int16)
        p = os.path.join(tmp_path, f"c{ch}.wav")
        wavio.write_wav_pcm16(p, fr, 22050)
        raw, sr = wavio.read_wav_pcm16(p)
        assert sr == 22050 and raw.dtype == np.int16 and np.array_equal(raw, fr)
        f32, _ = wavio.read_wav(p)
        want = f32 if ch == 1 else np.mean(f32, axis=0)
        assert np.array_equal(osil.load_audio_pcm16(fr), want)
        got = corpus.load_native_22050(p)
        assert (got.dtype == np.int16) == (ch == 1)
    pf = os.path.join(tmp_path, "f.wav")
    wavio.write_wav_float32(pf, rng.standard_normal(100).astype(np.float32), 22050)
    assert wavio.read_wav_pcm16(pf) is None and corpus.load_native_22050(pf).dtype == np.float32
    x = (rng.standard_normal(10000) * 0.5).astype(np.float32)
    assert np.array_equal(wavio.encode_pcm16(x), osil.float_to_pcm16(x))


def test_corpus_journal_resume(tmp_path):
    """an interrupted corpus run resumed from its per-rank journal gives the CSV of an uninterrupted
    run, skips the finished files (including those without detections), ignores a torn last line and the lines of
    another file list."""
    from softspoken_b200 import corpus
    files = [f"/data/clip{i}.wav" for i in range(13)]
    durations = [600.0] * len(files)
    seen = []

    def fake_load(path):
        return np.full(4, float(files.index(path)), np.float32)

    def make_detect(fail_after=None):
        calls = {"n": 0}

        def detect(clips):
            calls["n"] += 1
            if fail_after is not None and calls["n"] > fail_after:
                raise RuntimeError("simulated crash")
            out = []
            for c in clips:
                i = int(c[0])
                seen.append(i)
                out.append(np.sort(np.random.default_rng(i).integers(0, 51000, (i % 4, 2)), axis=1).astype(np.int32))
            return out
        return detect

    clean = corpus.csv_text(corpus.detect_corpus(files, make_detect(), load=fake_load, durations=durations, group_size=3))
    jp = os.path.join(tmp_path, "run.journal")
    with pytest.raises(RuntimeError):
        corpus.detect_corpus(files, make_detect(fail_after=2), load=fake_load, durations=durations, group_size=3, journal=jp)
    first = corpus.Journal(jp, files, 0).load()
    assert sorted(first) == [0, 1, 2, 3, 4, 5]                       # two groups of three files reached the disk
    with open(jp + ".rank0", "a") as f:
        f.write("deadbeef\t7\t1,2\n")                                # another file list
        f.write(f"{corpus.Journal(jp, files, 0).key}\t8\t5,")         # torn line (no newline, half a pair)
    del seen[:]
    resumed = corpus.csv_text(corpus.detect_corpus(files, make_detect(), load=fake_load, durations=durations,
                                                   group_size=3, journal=jp))
    assert resumed == clean
    assert sorted(seen) == list(range(6, 13))                        # nothing was recomputed
    # a third run finds everything done and still writes the same CSV
    del seen[:]
    again = corpus.csv_text(corpus.detect_corpus(files, make_detect(), load=fake_load, durations=durations,
                                                 group_size=3, journal=jp))
    assert again == clean and seen == []


def test_corpus_prefetch_order_errors_and_early_exit():
    """The reader thread hands groups over in order, a load failure surfaces at its own group (earlier groups are
    still delivered), and abandoning the loop stops the reader instead of leaving it blocked on a full queue."""
    import threading
    import time
    from softspoken_b200 import corpus
    groups = [[0, 1, 2], [3, 4], [5], [6, 7, 8]]
    loaded = []

    def load(i):
        loaded.append(i)
        return np.full(4, i, np.int16)
    for depth in (0, 1, 3):
        loaded.clear()
        got = list(corpus._prefetched(groups, load, depth=depth))
        assert [g for g, _ in got] == groups and loaded == list(range(9))
        assert all(int(c[0]) == i for g, clips in got for i, c in zip(g, clips))

    def bad(i):
        if i == 5:
            raise OSError("unreadable file")
        return np.zeros(1, np.int16)
    seen = []
    with pytest.raises(OSError, match="unreadable"):
        for g, _ in corpus._prefetched(groups, bad, depth=2):
            seen.append(g)
    assert seen == groups[:2]
    before = threading.active_count()
    it = corpus._prefetched([[i] for i in range(100)], load, depth=1)
    next(it)
    it.close()                       # consumer gives up: the reader must notice and end
    deadline = time.time() + 5
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before


def test_silence_worker_pcm16_pipeline_host_logic(tmp_path, capsys):
    """SilenceWorker's file pipeline on CPU (reader / writer helper threads around the device call): a stand-in
    engine applies the oracle's int16 round trip, so what is checked here is the host logic — every file written
    once, in group order, bytes equal to the oracle's, an unreadable file reported with the reference's message
    (silencer_ui.py:961-966) without stalling the files behind it, and a non-PCM_16 file taking the float32 route."""
    from oracle import silence as osil
    from softspoken_b200 import wavio
    from softspoken_b200.silencer import SilenceWorker, interval_table

    class FakeEngine:
        def __init__(self):
            self.calls = []

        def silence_pcm16_host(self, pcm, intervals, requantize=True):
            assert requantize and pcm.dtype == np.int16 and pcm.flags.writeable
            self.calls.append("s16")
            flat = pcm.reshape(-1)
            req = osil.float_to_pcm16(osil.pcm16_to_float(flat))
            for s, e in np.asarray(intervals).reshape(-1, 2):
                req[s:e] = 0
            flat[:] = req

        def silence_host(self, audio, table):
            self.calls.append("f32")
            flat = audio.reshape(-1)
            for s, e in table:
                flat[s:e] = 0.0

    rng = np.random.default_rng(5)
    sr = 4000
    src = tmp_path / "in"
    src.mkdir()
    clips = {}
    for i in range(7):
        shape = (9000 + 37 * i,) if i % 3 else (5000 + i, 2)
        clips[f"c{i}.wav"] = rng.integers(-32768, 32768, shape).astype(np.int16)
        wavio.write_wav_pcm16(str(src / f"c{i}.wav"), clips[f"c{i}.wav"], sr)
    wavio.write_wav_float32(str(src / "f.wav"), rng.uniform(-1, 1, 6000).astype(np.float32), sr)
    rows = [(f"c{i}.wav", 0.1 * i, 0.1 * i + 0.7) for i in range(7)] + [("c2.wav", 1.9, 5.0), ("gone.wav", 0.0, 1.0),
                                                                        ("f.wav", 0.25, 0.5)]
    df = pd.DataFrame({"file_path": [str(src)] * len(rows), "file_name": [r[0] for r in rows],
                       "start_time": [r[1] for r in rows], "end_time": [r[2] for r in rows], "erase": 1})
    out = tmp_path / "out"
    out.mkdir()
    eng = FakeEngine()
    w = SilenceWorker(df, str(out), engine=eng)
    w.run()
    names = sorted(clips) + ["f.wav"]
    assert [a[0] for a in w.signals.fileComplete.log] == [str(out / (n[:-4] + "_silenced.wav")) for n in sorted(names)]
    assert [a[0] for a in w.signals.fileStarted.log] == [str(src / n) for n in sorted(names + ["gone.wav"])]
    assert w.signals.overallProgress.log[-1] == (100,) and len(w.signals.finished.log) == 1
    assert "Error loading" in capsys.readouterr().out and eng.calls.count("s16") == 7 and eng.calls.count("f32") == 1
    for n, fr in clips.items():
        want = osil.silence_pcm16(fr, sr, [(r[1], r[2]) for r in rows if r[0] == n])
        got, got_sr = wavio.read_wav_pcm16(str(out / (n[:-4] + "_silenced.wav")))
        assert got_sr == sr and np.array_equal(got, want), n


def _riff(chunks):
    body = b"WAVE" + b"".join(cid + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) & 1 else b"")
                              for cid, data in chunks)
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_wav_parser_field_recorder_layouts(tmp_path):
    """RIFF layouts field recorders really write: metadata chunks (LIST / ICMT-style, odd-sized, padded) between
    `fmt ` and `data`, WAVE_FORMAT_EXTENSIBLE headers, 24- and 32-bit PCM, a `data` size field that overstates the
    bytes present (recording cut off) and trailing chunks after `data`."""
    rng = np.random.default_rng(8)
    pcm = rng.integers(-32768, 32768, 1001).astype("<i2")
    fmt16 = struct.pack("<HHIIHH", 1, 1, 22050, 44100, 2, 16)
    p = str(tmp_path / "meta.wav")
    with open(p, "wb") as f:
        f.write(_riff([(b"fmt ", fmt16), (b"LIST", b"INFOICMT\x07\x00\x00\x00AudioM\x00"), (b"junk", b"abc"),
                       (b"data", pcm.tobytes()), (b"bext", b"x" * 10)]))
    x, sr = wavio.read_wav(p)
    assert sr == 22050 and np.array_equal(x, pcm.astype(np.float32) / np.float32(32768))
    got, sr2 = wavio.read_wav_pcm16(p)
    assert sr2 == 22050 and np.array_equal(got, pcm)
    assert wavio.duration_and_rate(p) == (1001 / 22050, 22050)
    # WAVE_FORMAT_EXTENSIBLE wrapping PCM_16 stereo
    st = rng.integers(-32768, 32768, (500, 2)).astype("<i2")
    ext = struct.pack("<HHIIHH", 0xFFFE, 2, 48000, 192000, 4, 16) + struct.pack("<HHI", 22, 16, 3) + \
        struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    p2 = str(tmp_path / "ext.wav")
    with open(p2, "wb") as f:
        f.write(_riff([(b"fmt ", ext), (b"data", st.tobytes())]))
    y, sr = wavio.read_wav(p2)
    assert sr == 48000 and y.shape == (2, 500) and np.array_equal(y, (st.astype(np.float32) / np.float32(32768)).T)
    fr, _ = wavio.read_wav_pcm16(p2)
    assert fr.shape == (500, 2) and np.array_equal(fr, st)
    # 24-bit and 32-bit PCM: value / 2^23 and value / 2^31 (libsndfile's float read)
    v24 = np.array([0, 1, -1, 8388607, -8388608, 123456, -654321], np.int32)
    b24 = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in v24)
    p3 = str(tmp_path / "p24.wav")
    with open(p3, "wb") as f:
        f.write(_riff([(b"fmt ", struct.pack("<HHIIHH", 1, 1, 22050, 66150, 3, 24)), (b"data", b24)]))
    z, _ = wavio.read_wav(p3)
    assert np.array_equal(z, (v24.astype(np.float64) / 8388608.0).astype(np.float32))
    assert wavio.read_wav_pcm16(p3) is None
    v32 = np.array([0, 1, -1, 2147483647, -2147483648], "<i4")
    p4 = str(tmp_path / "p32.wav")
    with open(p4, "wb") as f:
        f.write(_riff([(b"fmt ", struct.pack("<HHIIHH", 1, 1, 22050, 88200, 4, 32)), (b"data", v32.tobytes())]))
    z, _ = wavio.read_wav(p4)
    assert np.array_equal(z, (v32.astype(np.float64) / 2147483648.0).astype(np.float32))
    # truncated recording: the data chunk claims more than the file holds -> the frames that are there
    whole = _riff([(b"fmt ", fmt16), (b"data", pcm.tobytes())])
    p5 = str(tmp_path / "cut.wav")
    with open(p5, "wb") as f:
        f.write(whole[:-501])
    x5, _ = wavio.read_wav(p5)
    assert len(x5) == (len(pcm) * 2 - 501) // 2 and np.array_equal(x5, (pcm[:len(x5)].astype(np.float32) / np.float32(32768)))
    for bad in (b"RIFF\x04\x00\x00\x00WAVE", _riff([(b"data", b"\x00\x00")]), _riff([(b"fmt ", struct.pack("<HHIIHH", 85, 1, 8000, 8000, 1, 8)), (b"data", b"\x00")])):
        pb = str(tmp_path / "bad2.wav")
        with open(pb, "wb") as f:
            f.write(bad)
        with pytest.raises(wavio.WavError):
            wavio.read_wav(pb)


def _resume_worker(rank, world, port, q, files, durations, jp, delay):
    import time
    import torch.distributed as dist
    from softspoken_b200 import corpus
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def fake_load(path):
        return np.full(4, float(files.index(path)), np.float32)

    def detect(clips):
        out = []
        for c in clips:
            i = int(c[0])
            seen.append(i)
            out.append(np.sort(np.random.default_rng(i).integers(0, 51000, (i % 4, 2)), axis=1).astype(np.int32))
        return out
    time.sleep(delay[rank])          # a rank that reaches the journal late must not see what the others just wrote
    rows = corpus.detect_corpus(files, detect, load=fake_load, durations=durations, group_size=2, journal=jp, prefetch=0)
    q.put((rank, seen, corpus.csv_text(rows) if rank == 0 else None))
    dist.destroy_process_group()


def test_corpus_journal_resume_two_ranks_one_view(tmp_path):
    """with --resume every rank used to read the progress files on its own; a rank arriving late saw the
    lines another rank had just appended, sharded a different remainder, and files were detected twice or not at all.
    Now rank 0 reads and broadcasts ONE view: each unfinished file is detected exactly once, the CSV is the clean one."""
    import torch.multiprocessing as mp
    from softspoken_b200 import corpus
    files = [f"/data/clip{i}.wav" for i in range(17)]
    durations = [600.0] * len(files)

    def fake_load(path):
        return np.full(4, float(files.index(path)), np.float32)

    def detect(clips):
        return [np.sort(np.random.default_rng(int(c[0])).integers(0, 51000, (int(c[0]) % 4, 2)), axis=1).astype(np.int32)
                for c in clips]
    clean = corpus.csv_text(corpus.detect_corpus(files, detect, load=fake_load, durations=durations))
    jp = os.path.join(tmp_path, "run.journal")
    jr = corpus.Journal(jp, files, 0)
    for i in (0, 3, 4):                                   # what an earlier, interrupted run had finished
        jr.append(i, detect([fake_load(files[i])])[0])
    jr.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 911) % 2000
    procs = [ctx.Process(target=_resume_worker, args=(r, 2, port, q, files, durations, jp, (0.0, 1.5))) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    seen = sorted(i for _, s, _ in res for i in s)
    assert seen == [i for i in range(17) if i not in (0, 3, 4)]          # exactly once each, none of the finished ones
    assert [t for _, _, t in res if t is not None][0] == clean
