"""End-to-end on the GPU: synthetic wav -> regions -> CSV rows, bit-exact vs the reference goldens, in both
parity-grade classifier modes (f16x3 = tensor cores with split fp16 operands, the default; fp32 = CUDA cores)."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from softspoken_b200 import spec, synth, wavio

pytestmark = pytest.mark.gpu


class _PM:
    def __init__(self, files):
        self.files = list(files)

    def get_unprocessed_list(self):
        return list(self.files)


@pytest.fixture(scope="module", params=["f16x3", "fp32"])
def detector(request, sd_seed0, tmp_path_factory):
    from softspoken_b200 import checkpoint, settings
    from softspoken_b200.detector import NNDetector
    d = tmp_path_factory.mktemp("ckpt")
    path = str(d / "model_checkpoint.pth")
    checkpoint.save_checkpoint(sd_seed0, path, epoch=7)
    old = settings.model_dir, settings.model_name
    settings.model_dir, settings.model_name = str(d), "model_checkpoint.pth"
    det = NNDetector(_PM([]), mode=request.param)
    settings.model_dir, settings.model_name = old
    assert det.load_checkpoint(det.model, path) == 8          # epoch + 1 (NNDetector.py:49-50)
    yield det
    det.model.engine.close()


@pytest.fixture(scope="module")
def wavs(tmp_path_factory):
    d = tmp_path_factory.mktemp("data")
    a, b = str(d / "clip_seed0.wav"), str(d / "clip_seed1.wav")
    wavio.write_wav_pcm16(a, synth.synth_pcm16(60.0, 0), 22050)
    wavio.write_wav_pcm16(b, synth.synth_pcm16(20.0, 1), 22050)
    return str(d), [a, b]


def _margin_report(det, audio):
    from oracle import postproc as pp
    times, lg = det.detect_file(audio, want_logits=True)
    avg, cnt = pp.average_idx(lg.reshape(-1, 1, 256), (len(audio) + 132300) / 22050)
    return times, lg, pp.min_threshold_margin(avg, cnt)


def test_detect_file_matches_golden_rows(detector, clip60):
    from oracle import postproc as pp
    want = open(os.path.join(GOLDEN, "detections_seed0.csv")).read().splitlines()[1:23]
    times, lg, margin = _margin_report(detector, clip60)
    ref_lg = load_golden("model_seed0.npz")["logits"][:, 0]
    err = float(np.max(np.abs(lg - ref_lg)))
    print(f"logit err {err:.3e}; min |avg - 0.1| margin of this clip {margin:.3e}")
    assert err < margin, "logit error exceeds the threshold margin: bit-exact regions are not decidable"
    rows = pp.detection_rows("/data/clip_seed0.wav", times, 1)
    got = pp.csv_text(rows).splitlines()[1:]
    assert got == want


@pytest.mark.parametrize("fast", [True, False])
def test_process_worker_csv_bit_exact(detector, wavs, tmp_path, fast):
    """ProcessWorker + DetectionProject over two files == the CSV the reference wrote (IDs continue)."""
    from softspoken_b200.worker import DetectionProject, ProcessWorker
    d, files = wavs
    csv = str(tmp_path / f"det_{fast}.csv")
    project = DetectionProject(types.SimpleNamespace(current_project={"detections_file": csv}))
    detector.files_to_process = files
    detector.detections_project = {f: [] for f in files}
    planned = detector.plan_detection_job()
    assert [len(v) for v in planned.values()] == [105, 39]
    w = ProcessWorker(detector, project, planned, fast=fast)
    w.run()
    text = open(csv).read().replace(d, "/data")
    assert text == open(os.path.join(GOLDEN, "detections_seed0.csv")).read()
    assert len(w.signals.fileDone.log) == 2 and len(w.signals.finished.log) == 1
    # a re-opened project continues the IDs (silencer_ui.py:794-812; worker.py:107-112)
    again = DetectionProject(types.SimpleNamespace(current_project={"detections_file": csv}))
    assert int(again.df["ID"].max()) == 32


def test_process_batch_api_shapes_and_values(detector, clip60):
    from oracle import postproc as pp
    g = load_golden("model_seed0.npz")
    padded = pp.pad_audio(clip60)
    sp, mk = detector.process_batch(padded, g["starts"][32:64])
    assert sp.shape == (32, 2, 128, 256) and mk.shape == (32, 1, 256)
    assert sp.dtype == np.float32 and mk.dtype == np.float32
    assert np.max(np.abs(mk - g["logits"][32:64])) <= 1e-4 * np.max(np.abs(g["logits"]))
    assert np.max(np.abs(sp[9] - g["spec_w41"])) <= 1e-4 * np.max(np.abs(g["spec_w41"]))


def test_streamed_chunks_equal_resident(detector):
    """Long-recording path (BASELINE config 4 in miniature): > 1 chunk of 1024 windows streamed from the
    host must give bitwise the logits and regions of the device-resident run."""
    eng = detector.model.engine
    audio = synth.synth_audio(13 * 60 + 7.3, 3)                # 1,318 windows -> 2 chunks
    reg_h, lg_h = eng.detect_host(audio, want_logits=True)
    reg_d, n_d, lg_d = eng.detect_device(torch.from_numpy(audio).cuda(), want_logits=True)
    assert lg_h.shape[0] == 1318 > 1024
    assert np.array_equal(lg_h, lg_d.cpu().numpy())
    assert np.array_equal(reg_h, reg_d[: int(n_d.item())].cpu().numpy())
    # and equal to the kernel-by-kernel pipeline on the materialised padded clip
    padded = eng.pad(torch.from_numpy(audio).cuda())
    starts = torch.arange(1318, dtype=torch.int64) * spec.STEP_SAMPLES
    lg_k = eng.classify(eng.features(padded, starts))
    assert torch.equal(lg_k, lg_d)


def test_short_and_empty_clips(detector):
    eng = detector.model.engine
    for n in [0, 1, 100, 13229, 13230, 13231, 66150]:
        audio = synth.synth_audio(n / 22050 + 1e-9, 5)[:n]
        reg, lg = eng.detect_host(audio, want_logits=True)
        assert lg.shape[0] == max(5, int(np.ceil((n + 66150) / 13230)))
        assert reg.ndim == 2 and reg.shape[1] == 2


def test_detect_host_batch_equals_per_clip(detector):
    """ss_detect_host_batch (cross-clip overlap of upload and compute) == one ss_detect_host per clip, for more
    clips than the 8 in-flight slots and for ragged / empty clips."""
    eng = detector.model.engine
    clips = [synth.synth_audio(d, 10 + i) for i, d in enumerate([20.0, 0.7, 33.1, 5.0, 12.4, 20.0, 1.0, 8.8, 15.5, 3.3])]
    clips.insert(3, np.zeros(0, np.float32))
    got = eng.detect_host_batch(clips, cap=512)
    assert len(got) == len(clips)
    for c, g in zip(clips, got):
        assert np.array_equal(g, eng.detect_host(c, cap=512))


def test_config2_clip_device_vs_host_paths(detector):
    """A full BASELINE config-2 unit (10-minute clip, 1,005 windows, 51,712 timeline bins): the device-resident path,
    the streamed host path and the batched host path give bitwise the same regions, and the regions are sorted,
    disjoint and separated by more than the 0.5 s merge gap (size-independent properties of NNDetector.py:109-141)."""
    eng = detector.model.engine
    audio = synth.synth_audio(600.0, 42)
    reg_d, n_d = eng.detect_device(torch.from_numpy(audio).cuda())
    reg_d = reg_d[: int(n_d.item())].cpu().numpy()
    reg_h = eng.detect_host(audio)
    reg_b = eng.detect_host_batch([audio, audio[: 22050 * 30]])[0]
    assert lib_windows(len(audio)) == 1005
    assert np.array_equal(reg_d, reg_h) and np.array_equal(reg_d, reg_b)
    assert len(reg_d) > 0
    assert (reg_d[:, 0] <= reg_d[:, 1]).all()
    assert (reg_d[1:, 0] - reg_d[:-1, 1] > spec.GAP_BINS).all()
    assert reg_d.min() >= 0 and reg_d.max() < 51712


def lib_windows(n):
    from softspoken_b200.engine import plan_windows
    return plan_windows(n)


def test_headless_job_from_wav_files(detector, wavs, tmp_path):
    """The whole GUI-less job on wav files: corpus driver (reader thread, int16 upload) -> detections CSV == the
    golden CSV the reference wrote; review step (all erase) == the golden review CSV; "Silence Voices" back to wav
    files == decode -> zero the rounded sample ranges -> re-encode, sample for sample."""
    import json
    import pandas as pd
    from softspoken_b200 import corpus, review, silencer
    d, files = wavs
    eng = detector.model.engine
    if eng.mode != "f16x3":
        pytest.skip("one classifier mode is enough for the file pipeline")
    rows = corpus.detect_corpus(files, eng.detect_host_batch, load=corpus.load_native_22050, group_size=1, prefetch=2)
    text = corpus.csv_text(rows)
    assert text.replace(d, "/data") == open(os.path.join(GOLDEN, "detections_seed0.csv")).read()
    det_csv, rev_csv = str(tmp_path / "p_detections.csv"), str(tmp_path / "p_review.csv")
    with open(det_csv, "w", newline="") as f:
        f.write(text)
    table = review.ReviewTable.open(det_csv, rev_csv)
    for tick in range(len(table)):                 # the golden case clicked "erase" row by row, one second apart
        table.label(tick, True, pd.Timestamp("2026-01-02 03:04:05") + pd.Timedelta(seconds=tick))
    df = review.save_review(table, rev_csv)
    with open(os.path.join(GOLDEN, "review_cases.json")) as f:
        want_review = json.load(f)["seed0_erase_all"]["outputs"]["golden_review.csv"]
    assert open(rev_csv).read().replace(d, "/data") == want_review
    out = tmp_path / "silenced"
    out.mkdir()
    silencer.SilenceWorker(silencer.coerce_erase(pd.read_csv(rev_csv)), str(out), engine=eng).run()
    for path in files:
        x, sr = wavio.read_wav(path)
        want = wavio.encode_pcm16(x)
        mine = df[df["file_name"] == os.path.basename(path)]
        assert len(mine) > 0
        for s, e in zip(mine["start_time"], mine["end_time"]):
            a, b = silencer.row_to_samples(s, e, sr, len(x))
            want[a:b] = 0
        got, got_sr = wavio.read_wav_pcm16(str(out / (os.path.basename(path)[:-4] + "_silenced.wav")))
        assert got_sr == sr and np.array_equal(got, want)


def test_region_capacity_is_grown_not_fatal(sd_seed0, clip60):
    """a clip with more regions than the caller's capacity used to abort the call (and, in a corpus run, the
    whole job after the work was done).  The kernel reports the true count; the engine runs that clip again with room."""
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=32)
    want = eng.detect_host(clip60)
    assert len(want) > 4
    assert np.array_equal(eng.detect_host(clip60, cap=2), want)
    short = clip60[: 22050 * 5]
    outs = eng.detect_host_batch([clip60, short, clip60], cap=3)
    assert np.array_equal(outs[0], want) and np.array_equal(outs[2], want)
    assert np.array_equal(outs[1], eng.detect_host(short))
    eng.close()
