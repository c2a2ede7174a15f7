"""Freeze golden vectors from the REAL reference into tests/golden/.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference, which is read-only and
absent from the GPU box):

    python -m oracle.make_golden

Every array written here is produced by the reference's own code
(`SpecUNet_2D`, `NNDetector`, `ProcessWorker.run`, `DetectionProject`,
`SilenceWorker.run`) imported headlessly through `oracle/ref_shim.py`; the
oracle restatements are NOT used to produce outputs (only `calibrate_head`
uses the oracle network, to choose the synthetic checkpoint's head biases —
an input, not an output).  tests/test_oracle_*.py then pin the oracle against
these files, and the `-m gpu` tests pin the CUDA path against both.
"""
from __future__ import annotations

import hashlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import calibrate, ref_shim  # noqa: E402
from softspoken_b200 import checkpoint, synth, wavio  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CLIP_S = 60.0
SEED = 0
REF_THREADS = 4   # settings.cpu_threads on the 8-core build container (settings.py:32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_shim.load()
    torch.manual_seed(0)

    # ---------------------------------------------------------------- head calibration (input)
    audio = synth.synth_audio(CLIP_S, SEED)
    head = calibrate.calibrate_head(checkpoint.synthetic_state_dict(SEED), audio)
    with open(os.path.join(GOLDEN, f"head_seed{SEED}.json"), "w") as f:
        json.dump(head, f, indent=1)
    sd = checkpoint.synthetic_state_dict(SEED, head)

    # ---------------------------------------------------------------- state-dict layout + buffers
    model = ref.SpecUNet_2D()
    rsd = model.state_dict()
    layout = [[k, list(v.shape), str(v.dtype)] for k, v in rsd.items()]
    with open(os.path.join(GOLDEN, "state_dict_layout.json"), "w") as f:
        json.dump(layout, f)

    det = ref_shim.make_detector(ref, sd, threads=REF_THREADS)   # strict load_state_dict of OUR dict
    model = det.model

    # ---------------------------------------------------------------- front end
    padded = np.zeros(len(audio) + 2 * 66150, np.float32)          # worker.py:58-62
    padded[66150:66150 + len(audio)] = audio
    fe_starts = np.array([0, 5 * 13230, 41 * 13230], dtype=np.int64)  # window 0 is pure zero-pad + onset
    x = torch.stack([torch.from_numpy(padded[i:i + 66150]) for i in fe_starts])
    with torch.no_grad():
        mel = model.sqrt_log10_nonzero(model.mel_spectrogram(x))[:, :, :256]
    fb = rsd["mel_spectrogram.mel_scale.fb"].numpy()
    nz = np.nonzero(fb)
    np.savez_compressed(
        os.path.join(GOLDEN, "frontend.npz"),
        window=rsd["mel_spectrogram.spectrogram.window"].numpy(),
        fb_sha256=np.array(sha(fb)), fb_rows=nz[0].astype(np.int32), fb_cols=nz[1].astype(np.int32),
        fb_vals=fb[nz], starts=fe_starts, mel=mel.numpy(), clip_seed=np.array(SEED), clip_s=np.array(CLIP_S))

    # ---------------------------------------------------------------- full detector run on the 60 s clip
    tmp = tempfile.mkdtemp(prefix="ss_golden_")
    wav = os.path.join(tmp, "clip_seed0.wav")
    wavio.write_wav_pcm16(wav, synth.synth_pcm16(CLIP_S, SEED), 22050)

    # real plan_detection_job, with get_audio_data answered from the wav header
    import root.code.frontend.NNDetector as nnd_mod
    nnd_mod.get_audio_data = lambda file: wavio.duration_and_rate(file)
    det.files_to_process = [wav]
    det.detections_project = {wav: []}
    planned = det.plan_detection_job()
    starts = np.asarray(planned[wav])

    # real process_batch over the reference's own batching (worker.py:71-79)
    logits, spec0 = [], None
    for s in range(0, len(starts), 32):
        sp, mk = det.process_batch(padded, starts[s:s + 32])
        logits.append(mk)
        if s == 32:
            spec0 = sp[9].copy()                # window 41
    logits = np.vstack(logits)

    # intermediate activations of window 41 (means/abs-max only — layer-level localisation of a mismatch)
    acts = {}
    hooks = []
    for name in ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv_bottleneck", "encoder_out",
                 "conv6", "conv7", "conv8", "conv9_1"]:
        hooks.append(getattr(model, name).register_forward_hook(
            lambda m, i, o, name=name: acts.__setitem__(name, o.detach())))
    with torch.no_grad():
        model(torch.from_numpy(padded[41 * 13230:41 * 13230 + 66150])[None])
    for h in hooks:
        h.remove()
    act_stats = {k: np.array([float(v.mean()), float(v.abs().max()), float(v.std())]) for k, v in acts.items()}
    conv9_w41 = acts["conv9_1"][0].numpy()

    np.savez_compressed(
        os.path.join(GOLDEN, "model_seed0.npz"),
        starts=starts, logits=logits, spec_w41=spec0, conv9_w41_rows=conv9_w41[:, ::16, :],
        threads=np.array(REF_THREADS), **{f"act_{k}": v for k, v in act_stats.items()})

    # real averaging + region finding
    secs = len(padded) / 22050
    avg = det.average_overlapping_detections({wav: logits}, secs)
    regions = det.find_speech_regions({wav: avg}, break_duration=0.5)
    np.savez_compressed(
        os.path.join(GOLDEN, "postproc_seed0.npz"),
        avg_values=np.array([v for v, _ in avg[wav]], dtype=np.float64),
        avg_times=np.array([t for _, t in avg[wav]]),
        regions=np.array(regions[wav]), n_padded=np.array(len(padded)))

    # real ProcessWorker.run + real DetectionProject -> CSV text.  The only stub is load_audio
    # (libsndfile is absent): it returns what sf.read(dtype='float32') gives for this PCM_16 file.
    ui = ref_shim.extract_ui_classes()
    csv_path = os.path.join(tmp, "proj_detections.csv")
    settings_stub = types.SimpleNamespace(current_project={"detections_file": csv_path})
    project = ui["DetectionProject"](settings_stub)
    ref.worker.load_audio = lambda path: wavio.read_wav(path)
    worker = ref.worker.ProcessWorker(det, project, {wav: starts})
    worker.run()
    # second file appended to the same project -> ID continuation (worker.py:107-112)
    wav2 = os.path.join(tmp, "clip_seed1.wav")
    wavio.write_wav_pcm16(wav2, synth.synth_pcm16(20.0, 1), 22050)
    det.files_to_process = [wav2]
    det.detections_project = {wav2: []}
    planned2 = det.plan_detection_job()
    ref.worker.ProcessWorker(det, project, planned2).run()
    text = open(csv_path).read().replace(tmp, "/data")
    with open(os.path.join(GOLDEN, "detections_seed0.csv"), "w") as f:
        f.write(text)

    # ---------------------------------------------------------------- window plan for assorted durations
    durs = [0.0, 0.01, 0.59, 0.6, 1.0, 2.9999, 3.0, 3.3, 59.99, 60.0, 600.0, 601.2345, 3599.5, 86400.0,
            13230 / 22050, 13231 / 22050, 7.0000227]
    counts = []
    for d in durs:
        nnd_mod.get_audio_data = lambda file, d=d: (d, 22050)
        det.detections_project = {"x": []}
        counts.append(len(det.plan_detection_job()["x"]))
    np.savez(os.path.join(GOLDEN, "plan.npz"), durations=np.array(durs), n_windows=np.array(counts))

    # ---------------------------------------------------------------- averaging/regions on adversarial synthetic logits
    rng = np.random.default_rng(7)
    cases = {}

    def logits_from_timeline(g, W):
        """Window i sees g[p_i : p_i + 256] (+ tiny per-window jitter so the mean is a real mean)."""
        pos = [int(round(i * 0.6 / (3 / 256))) for i in range(W)]
        lg = np.stack([g[p:p + 256] for p in pos]).astype(np.float32)
        lg += rng.normal(0, 1e-4, lg.shape).astype(np.float32)
        return lg.reshape(W, 1, 256)

    def run_pattern(n, runs):
        """runs = [(gap_before, hot_len), ...] -> +-0.05 around the threshold."""
        g = np.full(n, 0.05)
        j = 0
        for gap, hot in runs:
            j += gap
            g[j:j + hot] = 0.15
            j += hot
        return g

    specs = {
        "tiny": (1, 66150 * 2 / 22050 + 0.01, [(3, 1), (42, 5), (43, 1), (41, 7)]),
        "short": (7, 10.0, [(0, 10), (43, 2), (42, 1), (60, 30), (44, 1), (1, 1), (100, 200)]),
        "mid": (200, 126.0, [(5, 50), (42, 10), (43, 10), (41, 1), (300, 1), (43, 1), (42, 1), (500, 4000),
                             (43, 3), (1000, 2), (4100, 10)]),
    }
    for name, (W, secs_c, runs) in specs.items():
        n_bins = int(round(secs_c * 256 / 3)) + 512
        lg = logits_from_timeline(run_pattern(n_bins, runs), W)
        if name == "short":
            lg[:, :, 100:108] = np.float32(0.1)            # exact-threshold values: `>` must be strict
        a = det.average_overlapping_detections({"f": lg}, secs_c)
        r = det.find_speech_regions({"f": a}, break_duration=0.5)
        cases[f"{name}_logits"] = lg
        cases[f"{name}_secs"] = np.array(secs_c)
        cases[f"{name}_avg"] = np.array([v for v, _ in a["f"]])
        cases[f"{name}_times"] = np.array([t for _, t in a["f"]])
        cases[f"{name}_regions"] = np.array(r["f"]).reshape(-1, 2)
    # empty prediction list (worker.py:93-94)
    a = det.average_overlapping_detections({"f": np.array([])}, 6.0)
    cases["empty_n"] = np.array(len(a["f"]))
    np.savez_compressed(os.path.join(GOLDEN, "postproc_cases.npz"), **cases)

    # ---------------------------------------------------------------- SilenceWorker on in-memory buffers
    import pandas as pd
    store, written = {}, {}
    librosa_stub = types.SimpleNamespace(load=lambda path, sr=None, mono=False: (store[path][0].copy(), store[path][1]))
    sf_stub = types.SimpleNamespace(write=lambda path, data, sr: written.__setitem__(path, (np.array(data), sr)))
    ui2 = ref_shim.extract_ui_classes(extra_globals={"librosa": librosa_stub, "sf": sf_stub})
    def tone(n, c=None):
        """Deterministic, never-zero float32 pattern (no RNG, so tests can rebuild it)."""
        k = np.arange(n if c is None else n * c, dtype=np.int64)
        a = (((k * 7919) % 2003) / 2003.0 - 0.5 + 1e-3).astype(np.float32)
        return a if c is None else a.reshape(c, n)

    store["/d/a.wav"] = (tone(40000), 8000)            # mono (n,)
    store["/d/b.wav"] = (tone(50000, 2), 22050)        # stereo (2,n)
    store["/e/c.wav"] = (tone(30000), 44100)
    rows = [
        # file_path, file_name, start, end, erase
        ("/d", "a.wav", 0.5, 1.25, 1), ("/d", "a.wav", 4.9, 7.0, 1),      # end beyond file -> clamped
        ("/d", "a.wav", 2.0, 2.5, 0),                                     # not erased
        ("/d", "a.wav", -1.0, 0.01, 1),                                   # negative start -> clamped
        ("/d", "a.wav", 3.0, 2.0, 1),                                     # end < start -> no-op
        ("/d", "b.wav", 0.0000227, 0.0000680, 1),                         # rounding: 0.5005 -> 1, 1.4994 -> 1
        ("/d", "b.wav", 1.00002268, 1.5, 1),                              # 22050.5001 -> 22051
        ("/d", "b.wav", 2.123, 2.2675737, 1),
        ("/e", "c.wav", 0.1, 0.2, 1), ("/e", "c.wav", 0.15, 0.3, 1),      # overlapping rows
        ("/e", "c.wav", 0.5, 0.5, 1),                                     # empty
    ]
    df = pd.DataFrame(rows, columns=["file_path", "file_name", "start_time", "end_time", "erase"])
    sw = ui2["SilenceWorker"](df, "/out")
    sw.run()
    sil = {"rows_path": np.array([r[0] for r in rows]), "rows_name": np.array([r[1] for r in rows]),
           "rows_start": np.array([r[2] for r in rows]), "rows_end": np.array([r[3] for r in rows]),
           "rows_erase": np.array([r[4] for r in rows])}
    for i, (path, (a, sr)) in enumerate(store.items()):
        sil[f"in{i}_path"] = np.array(path)
        sil[f"in{i}_shape"] = np.array(a.shape)
        sil[f"in{i}_sr"] = np.array(sr)
    sil["out_paths"] = np.array(list(written.keys()))
    for i, (path, (a, sr)) in enumerate(written.items()):
        # (samples, channels) as handed to sf.write: keep the zero mask (as run boundaries) + a hash
        z = (a == 0.0).all(axis=1).astype(np.int8)
        edges = np.flatnonzero(np.diff(np.concatenate([[0], z, [0]])))
        sil[f"out{i}_shape"] = np.array(a.shape)
        sil[f"out{i}_zero_runs"] = edges.reshape(-1, 2)
        sil[f"out{i}_sha256"] = np.array(sha(a.astype(np.float32)))
        sil[f"out{i}_sr"] = np.array(sr)
    np.savez_compressed(os.path.join(GOLDEN, "silence_cases.npz"), **sil)

    # erase coercion (silencer_ui.py:1100) on awkward CSV values
    csv = "erase\n1\n0\n\nyes\n1.0\n 1\n2\n"
    rd = pd.read_csv(io.StringIO(csv))
    coerced = pd.to_numeric(rd["erase"], errors="coerce").fillna(0).astype(int)
    np.savez(os.path.join(GOLDEN, "erase_coercion.npz"), raw=np.array([str(v) for v in rd["erase"]], dtype="U16"),
             coerced=coerced.to_numpy())

    sizes = {f: os.path.getsize(os.path.join(GOLDEN, f)) for f in sorted(os.listdir(GOLDEN))}
    print(json.dumps(sizes, indent=1))


if __name__ == "__main__":
    main()
