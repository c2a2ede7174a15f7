"""BASELINE.json configs 4 and 5 on one B200: one JSON line each (same vocabulary as bench.py).

    python tools/bench_aux.py config1                  # config 1: one 60 s clip, CPU reference port next to the GPU path
    python tools/bench_aux.py long    [--hours 24]     # config 4: one long recording streamed in chunks
    python tools/bench_aux.py silence [--files 1000]   # config 5: 10k flagged intervals masked across the corpus
    python tools/bench_aux.py files   [--files 24]     # config 2 from wav FILES: read + decode + upload + detect + CSV,
                                                       # then review (all erase) + exports + "Silence Voices" to wavs

config 4 — a synthetic 24 h mono 22,050 Hz recording (a seeded 10-minute clip tiled with a different gain per
tile, so neighbouring hours differ) is handed to `ss_detect_host` as ONE host buffer; the library streams it in
chunks of 1,024 windows with overlapping analysis windows.  Checked here: the logits of the first 30 minutes are
bitwise those of a separate run on that prefix alone (windows are independent of chunking), and the regions of
the prefix run are a prefix of the long run's regions.
config 5 — 10,000 `erase = 1` intervals (SURVEY §8d recipe) over a 1,000-clip corpus resident in HBM as one
packed float32 buffer, one `ss_silence` launch; a sample of files is compared bit-for-bit with the oracle's
`audio[:, s:e] = 0.0` (oracle/silence.py is used as the checker only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SR = 22050


def load_engine(max_batch=64, mode="f16x3"):
    from softspoken_b200 import checkpoint
    from softspoken_b200.engine import Engine
    with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
        head = json.load(f)
    return Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=max_batch, mode=mode)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def _stream_hour16(args):
    from softspoken_b200 import synth
    return synth.stream_hour_pcm16(*args)


def run_long(args):
    """Config 4: one long recording streamed from a host buffer.  The recording is `hours` independently seeded
    one-hour clips (`synth.stream_hour`: it never has to exist anywhere but in the buffer it is generated into); its
    first hour is the clip whose reference results are frozen in tests/golden/scale_stream_hour0.npz, so the run is
    checked against the REAL reference on every window and detection row that lies inside that hour — and against
    a GPU run of the hour alone, bit for bit (chunking must not change a result)."""
    import multiprocessing as mp
    from tools import scale_parity
    n_hours = int(round(args.hours))
    g = np.load(os.path.join(ROOT, "tests", "golden", "scale_stream_hour0.npz"))
    seed = int(g["stream_seed"])
    n_hour = 3600 * SR
    n = n_hours * n_hour
    t0 = time.perf_counter()
    audio = torch.empty(n, dtype=torch.int16 if args.pcm16 else torch.float32).pin_memory()
    view = audio.numpy()
    with mp.get_context("spawn").Pool(min(8, os.cpu_count() or 1)) as pool:      # spawn: the parent already holds a CUDA context (pinned buffer)
        for h, pcm in enumerate(pool.imap(_stream_hour16, [(seed, h) for h in range(n_hours)])):
            view[h * n_hour:(h + 1) * n_hour] = pcm if args.pcm16 else pcm.astype(np.float32) / np.float32(32768.0)
    gen_s = time.perf_counter() - t0
    eng = load_engine(args.max_batch, args.mode)
    W = (n + 66150 + 13229) // 13230
    eng.reserve(n, 1 << 20)
    # warm-up on the first hour, which is also the prefix reference
    reg_pre, lg_pre = eng.detect_host(audio[:n_hour], want_logits=True, cap=1 << 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reg, lg = eng.detect_host(audio, want_logits=True, cap=1 << 20)
    dt = time.perf_counter() - t0
    # windows that lie entirely inside the first hour (and its leading pad) see identical samples in both runs
    w_same = (n_hour + 66150 - 65536) // 13230
    same_logits = bool(np.array_equal(lg[:w_same], lg_pre[:w_same]))
    safe_bin = int((w_same - 5) * 51.2)
    a = reg[reg[:, 1] < safe_bin]
    b = reg_pre[reg_pre[:, 1] < safe_bin]
    same_regions = bool(np.array_equal(a, b))
    # against the reference: logits of the golden's windows (every 4th) that lie inside the hour, and every
    # reference row that ends before the last bin the hour alone determines
    stride = int(g["logits_stride"])
    ref_lg = g["logits"]
    k = (w_same + stride - 1) // stride
    ref_err = float(np.max(np.abs(lg[:w_same:stride].astype(np.float64) - ref_lg[:k])))
    ref_rows = {tuple(r) for r in g["region_bins"].tolist() if r[1] < safe_bin}
    got_rows = {tuple(r) for r in a.astype(np.int64).tolist()}
    hour_alone = scale_parity.compare("hour0", reg_pre, lg_pre)
    hours = n / SR / 3600.0
    line = {
        "metric": "audio_hours_per_sec", "value": hours / dt, "unit": "audio-hours/s", "n_gpus": 1,
        "higher_is_better": True, "dtype": args.mode, "data": "synthetic",
        "config": {"workload": f"config4: one {hours:.1f} h mono 22.05 kHz recording ({n_hours} independently seeded hours, "
                               f"synth.stream_hour({seed}, h)), host buffer "
                               f"({'int16 samples of a PCM_16 file' if args.pcm16 else 'float32'}) streamed in chunks of "
                               "1024 windows (52,920-sample overlap), K5/K6 once over the whole timeline",
                   "max_batch_windows": args.max_batch,
                   "windows": int(lg.shape[0]), "timeline_bins": int(lg.shape[0] * 51.2) + 256, "regions": int(len(reg)),
                   "planned_windows": int(W), "host_generation_s": round(gen_s, 1)},
        "x_realtime": hours * 3600.0 / dt, "seconds": dt,
        "e2e": {"value": hours / dt, "unit": "audio-hours/s", "h2d_bytes_per_step": int(n * (2 if args.pcm16 else 4)),
                "d2h_bytes_per_step": int(lg.nbytes + reg.nbytes)},
        "checks": {"prefix_logits_bitwise_equal_to_hour_alone": same_logits, "prefix_regions_equal_to_hour_alone": same_regions,
                   "prefix_windows_compared": int(w_same),
                   "vs_reference_inside_hour0": {"max_logit_err": ref_err, "reference_rows": len(ref_rows),
                                                 "differing_rows": len(ref_rows ^ got_rows)},
                   "hour0_alone_vs_reference": hour_alone},
    }
    print(json.dumps(line), flush=True)
    assert same_logits and same_regions
    assert ref_err <= 1e-4 and len(ref_rows ^ got_rows) <= 2 * hour_alone["differing_bins"]
    eng.close()


def run_silence(args):
    from oracle import silence as osil          # checker only
    from softspoken_b200 import synth
    from softspoken_b200.silencer import interval_table
    eng = load_engine(4, "bf16")
    n = 600 * SR
    files, start, end = synth.synth_review_rows(args.intervals, args.files, 600.0, 0)
    dev = torch.device("cuda", 0)
    corpus = torch.empty(args.files * n, dtype=torch.float32, device=dev)
    # cheap deterministic non-zero content: every sample is a function of its flat index
    idx = torch.arange(n, device=dev, dtype=torch.float32)
    for f in range(args.files):
        corpus[f * n:(f + 1) * n] = torch.sin(idx * (0.001 + 1e-6 * f)) * 0.5 + 0.25
    tabs = []
    for f in range(args.files):
        sel = np.nonzero(files == f)[0]
        if len(sel):
            tabs.append(interval_table(list(zip(start[sel], end[sel])), SR, 1, n, base=f * n))
    table = np.concatenate(tabs).astype(np.int64)
    masked_elems = int(np.sum(table[:, 1] - table[:, 0]))           # overlaps counted twice, as the kernel writes them
    iv = torch.from_numpy(table).to(dev)
    before = {f: corpus[f * n:(f + 1) * n].cpu().numpy() for f in range(0, args.files, max(1, args.files // 16))}
    stream = torch.cuda.current_stream()
    eng.silence(corpus[:n].clone(), iv[:1] * 0)                       # warm-up launch on a scratch copy
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.silence(corpus, iv)
    e1.record(stream)
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 1e3
    ok = True
    for f, orig in before.items():
        sel = np.nonzero(files == f)[0]
        want = osil.silence_buffer(orig.reshape(1, -1).copy(), SR, list(zip(start[sel], end[sel])))
        got = corpus[f * n:(f + 1) * n].cpu().numpy()
        ok = ok and np.array_equal(got, want.reshape(-1))
    peak, src = peaks()
    gbs = masked_elems * 4 / dt / 1e9
    line = {
        "metric": "silence_GBps", "value": gbs, "unit": "GB/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"config5: {len(table)} erase=1 intervals over {args.files} x 10-min clips resident in HBM "
                               f"({args.files * n * 4 / 1e9:.1f} GB packed float32), one ss_silence launch",
                   "masked_bytes": masked_elems * 4},
        "ms": dt * 1e3,
        "roofline": {"bound": "hbm", "kernel": "silence_kernel (K7)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                     "frac": gbs / peak, "traffic": None, "peak_source": src,
                     "note": "write-only: algorithmic bytes = sum(e - s) * 4"},
        "checks": {"files_compared_bitwise_with_oracle": len(before), "bit_exact": bool(ok)},
    }
    print(json.dumps(line), flush=True)
    assert ok
    eng.close()


def run_config1(args):
    """BASELINE config 1: one synthetic 60 s mono clip.  The reference's CPU detector (oracle port of its exact
    sequence: batches of 32 windows incl. the discarded spec head, float64 averaging, string-time regions) is timed on
    all host cores and on half of them (the reference's own default, settings.py:32); the GPU path is one
    `ss_detect_host` call on the same samples (median of 20).  The regions must be identical, and the CSV rows must be
    the golden rows the real reference wrote for this clip."""
    from oracle import model as om, postproc as pp          # the CPU side of the comparison
    from softspoken_b200 import synth
    from softspoken_b200.detector import region_bins_to_times
    eng = load_engine(128, args.mode)
    with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
        head = json.load(f)
    from softspoken_b200 import checkpoint
    sd = checkpoint.synthetic_state_dict(0, head)
    audio = synth.synth_audio(60.0, 0)
    padded = pp.pad_audio(audio)
    starts = pp.plan_windows(60.0)
    cpu = {}
    ref_times = None
    n_cores = os.cpu_count() or 1
    for threads in sorted({n_cores, max(1, n_cores // 2)}, reverse=True):
        torch.set_num_threads(threads)
        t0 = time.perf_counter()
        preds = []
        for b0 in range(0, len(starts), 32):
            x = torch.stack([torch.from_numpy(padded[i:i + 66150]) for i in starts[b0:b0 + 32]])
            _, mk = om.forward(sd, x, want_spec=True)
            preds.append(mk.numpy())
        entries = pp.average_overlapping(np.vstack(preds), len(padded) / SR)
        regions = pp.find_speech_regions(entries)
        cpu[threads] = time.perf_counter() - t0
        ref_times = pp.regions_to_times(regions)
    bins, _ = eng.detect_host(audio, want_logits=True)       # warm-up
    lat = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bins = eng.detect_host(audio)
        lat.append(time.perf_counter() - t0)
    lat.sort()
    gpu_s = lat[len(lat) // 2]
    got_times = region_bins_to_times(bins)
    same = got_times == ref_times
    rows = pp.csv_text(pp.detection_rows("/data/clip_seed0.wav", got_times, 1)).splitlines()[1:]
    golden = open(os.path.join(ROOT, "tests", "golden", "detections_seed0.csv")).read().splitlines()[1:1 + len(rows)]
    line = {"metric": "audio_hours_per_sec", "unit": "audio-hours/s", "n_gpus": 1, "higher_is_better": True, "dtype": args.mode,
            "data": "synthetic", "value": 60.0 / 3600.0 / gpu_s,
            "config": {"workload": "config1: one 60 s mono 22.05 kHz clip (105 windows), host float32 buffer -> regions on the host"},
            "gpu": {"median_ms": gpu_s * 1e3, "min_ms": lat[0] * 1e3, "regions": len(got_times)},
            "cpu_reference_port": {str(k): {"seconds": v, "audio_hours_per_s": 60.0 / 3600.0 / v} for k, v in cpu.items()},
            "checks": {"regions_identical_to_cpu": bool(same), "csv_rows_equal_golden_reference_rows": rows == golden}}
    print(json.dumps(line), flush=True)
    assert same and rows == golden
    eng.close()


def run_resample(args):
    """K9 on a 10-minute clip recorded at 48 kHz and at 44.1 kHz (int16 samples, as a PCM_16 file stores them)."""
    from oracle import resample as orr               # checker only
    from softspoken_b200 import resample as rs
    eng = load_engine(4, "bf16")
    dev = torch.device("cuda", 0)
    peak, src = peaks()
    out = {}
    for sr in (48000, 44100):
        n = 600 * sr
        g = torch.Generator(device=dev).manual_seed(sr)
        pcm = (torch.randn(n, device=dev, generator=g) * 3000).clamp(-32768, 32767).to(torch.int16)
        for _ in range(2):
            y = eng.resample(pcm, sr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(torch.cuda.current_stream())
        for _ in range(reps):
            y = eng.resample(pcm, sr)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) / reps / 1e3
        L, M, T, _ = rs.design(sr)
        want = orr.resample(pcm[:200000].cpu().numpy().astype(np.float32) / np.float32(32768), sr)
        got = eng.resample(pcm[:200000].contiguous(), sr).cpu().numpy()
        err = float(np.abs(got[:-2000] - want[:-2000]).max())
        b = 2 * n + 4 * y.numel()
        out[str(sr)] = {"ms": dt * 1e3, "taps": 2 * T + 1, "up": L, "down": M, "algorithmic_bytes": b, "GBps": b / dt / 1e9,
                        "frac": b / dt / 1e9 / peak, "gmacs_per_s": y.numel() * (2 * T + 1) / dt / 1e9,
                        "audio_hours_per_s": 600.0 / 3600.0 / dt, "max_abs_err_vs_definition": err}
    line = {"metric": "resample_audio_hours_per_sec", "unit": "audio-hours/s", "n_gpus": 1, "higher_is_better": True,
            "dtype": "f32", "data": "synthetic", "value": out["48000"]["audio_hours_per_s"],
            "config": {"workload": "K9: 10-minute mono PCM_16 clip at 48 kHz / 44.1 kHz -> float32 at 22,050 Hz on the device"},
            "rates": out, "roofline": {"bound": "hbm", "peak": peak, "peak_source": src, "unit": "GB/s",
                                       "note": "algorithmic bytes = 2 n_in + 4 n_out; at 139-297 taps per output the kernel is bound by FP32 / LSU issue"}}
    print(json.dumps(line), flush=True)
    eng.close()


def run_postproc(args):
    """K5 / K6 at the size where they stop being launch-bound: the timeline of a 24 h recording (144,005 windows,
    7,373,312 bins).  Logits are synthetic (smooth noise around the threshold so that regions start and end all
    along the timeline); timed with CUDA events on the launching stream, inputs larger than L2 (147 MB of logits,
    88 MB of timeline).  Algorithmic bytes: K5 = W*1024 read + out_len*12 written; K6 (ss_regions on a
    caller's timeline) = out_len*12 read once + regions written — its count and emit passes work on one hot bit per
    bin; inside ss_detect_* the averaging kernel votes those bits itself and K6 reads no timeline at all.  Checked
    against the oracle on a prefix."""
    from oracle import postproc as pp               # checker only
    from softspoken_b200 import spec
    from softspoken_b200.engine import plan_windows, timeline_bins
    eng = load_engine(4, "bf16")
    dev = torch.device("cuda", 0)
    n = int(round(args.hours * 3600 * SR))
    W = plan_windows(n)
    out_len = timeline_bins(n + 2 * spec.PAD_SAMPLES)
    g = torch.Generator(device=dev).manual_seed(5)
    slow = torch.nn.functional.interpolate(torch.randn(1, 1, W // 16 + 2, device=dev, generator=g), size=W, mode="linear")[0, 0]
    logits = (0.1 + 0.08 * slow[:, None] + 0.02 * torch.randn(W, 256, device=dev, generator=g)).contiguous()
    stream = torch.cuda.current_stream()
    times = {}
    for _ in range(2):
        avg, cnt = eng.average(logits, out_len)
        reg = eng.regions(avg, cnt, cap=1 << 22)
    reps = 5
    for name, fn in (("K5", lambda: eng.average(logits, out_len)),):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        times[name] = e0.elapsed_time(e1) / reps / 1e3
    # K6 through the device-level entry point (Engine.regions adds a D2H of the result): time it on the stream too
    from softspoken_b200 import _lib
    import ctypes as C
    cap = 1 << 22
    regd = torch.empty((cap, 2), dtype=torch.int32, device=dev)
    nd = torch.zeros(1, dtype=torch.int32, device=dev)
    def k6():
        _lib.check(_lib.lib.ss_regions(eng._ctx, C.c_void_p(avg.data_ptr()), C.c_void_p(cnt.data_ptr()), out_len,
                                       float(spec.THRESHOLD), int(spec.GAP_BINS), C.c_void_p(regd.data_ptr()),
                                       C.c_void_p(nd.data_ptr()), cap, eng._stream()))
    k6()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        k6()
    e1.record(stream)
    torch.cuda.synchronize()
    times["K6"] = e0.elapsed_time(e1) / reps / 1e3
    n_reg = int(nd.item())
    # oracle check on a prefix (windows are independent: the first bins depend on the first windows only)
    Wp = 600
    bins_p = (256 * (Wp - 1) + 2) // 5 + 256
    ref_avg, ref_cnt = pp.average_idx(logits[:Wp].cpu().numpy().reshape(-1, 1, 256), bins_p * 3 / 256)
    m = min(len(ref_avg), bins_p) - 300
    ok = bool(np.array_equal(avg[:m].cpu().numpy(), np.asarray(ref_avg[:m], np.float64)))
    peak, src = peaks()
    b5 = W * 1024 + out_len * 12
    b6 = out_len * 12 + n_reg * 8
    line = {"metric": "postproc_GBps", "unit": "GB/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f64 sums of f32, i32 bins",
            "data": "synthetic", "value": (b5 + b6) / (times["K5"] + times["K6"]) / 1e9,
            "config": {"workload": f"K5 + K6 on the timeline of a {args.hours:g} h recording: {W} windows, {out_len} bins, {n_reg} regions"},
            "K5_average": {"ms": times["K5"] * 1e3, "algorithmic_bytes": b5, "GBps": b5 / times["K5"] / 1e9, "frac": b5 / times["K5"] / 1e9 / peak},
            "K6_regions": {"ms": times["K6"] * 1e3, "algorithmic_bytes": b6, "GBps": b6 / times["K6"] / 1e9, "frac": b6 / times["K6"] / 1e9 / peak,
                           "launches": 4},
            "roofline": {"bound": "hbm", "peak": peak, "peak_source": src, "unit": "GB/s"},
            "checks": {"prefix_bins_bitwise_equal_oracle": m, "ok": ok}}
    print(json.dumps(line), flush=True)
    assert ok
    eng.close()


def run_spectrogram(args):
    """K8 (review-screen spectrogram) at config-2 size: 10-minute clips, magnitudes and the dB display
    transform; inputs rotate through four clips (4 x 106 MB of reads + writes: larger than L2).  Algorithmic bytes:
    4 n read + 257 T 4 written for the STFT; 2 x 257 T 4 for the in-place dB pass."""
    from oracle import spectrogram as osp            # checker only
    from softspoken_b200 import synth
    eng = load_engine(4, "bf16")
    dev = torch.device("cuda", 0)
    clips = [torch.from_numpy(synth.synth_audio(600.0, s)).to(dev) for s in range(4)]
    n = clips[0].numel()
    T = 1 + n // 256
    stream = torch.cuda.current_stream()
    outs = [eng.spectrogram(c) for c in clips]       # warm-up (and the buffers the dB pass runs on)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from softspoken_b200 import _lib
    import ctypes as C
    reps = 10
    mxs = torch.zeros(len(clips), device=dev)

    def stft(i):      # the C-ABI call itself on preallocated buffers (Engine.spectrogram allocates its result)
        _lib.check(_lib.lib.ss_spectrogram(eng._ctx, C.c_void_p(clips[i].data_ptr()), n, C.c_void_p(outs[i].data_ptr()),
                                           C.c_void_p(mxs[i:].data_ptr()), eng._stream()))
    for i in range(len(clips)):
        stft(i)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        for i in range(len(clips)):
            stft(i)
    e1.record(stream)
    torch.cuda.synchronize()
    t_stft = e0.elapsed_time(e1) / (reps * len(clips)) / 1e3
    mx = torch.tensor([float(o.max()) for o in outs], device=dev)
    e0.record(stream)
    for i, o in enumerate(outs):
        _lib.check(_lib.lib.ss_spectrogram_db(eng._ctx, C.c_void_p(o.data_ptr()), o.numel(), C.c_void_p(mx[i:].data_ptr()), eng._stream()))
    e1.record(stream)
    torch.cuda.synchronize()
    t_db = e0.elapsed_time(e1) / len(outs) / 1e3
    x = clips[1][:2000000].cpu().numpy()
    want = osp.stft_magnitude(x)
    got = eng.spectrogram(clips[1][:2000000]).cpu().numpy()
    err = float(np.abs(got - want).max() / want.max())
    peak, src = peaks()
    b1, b2 = 4 * n + 257 * T * 4, 2 * 257 * T * 4
    line = {"metric": "spectrogram_GBps", "unit": "GB/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "value": b1 / t_stft / 1e9,
            "config": {"workload": f"K8: 512/256 STFT magnitudes of 10-min mono clips ({n} samples -> [257, {T}]), pool of 4 clips"},
            "K8_stft": {"ms": t_stft * 1e3, "algorithmic_bytes": b1, "GBps": b1 / t_stft / 1e9, "frac": b1 / t_stft / 1e9 / peak,
                        "audio_hours_per_s": 600.0 / 3600.0 / t_stft},
            "K8_db": {"ms": t_db * 1e3, "algorithmic_bytes": b2, "GBps": b2 / t_db / 1e9, "frac": b2 / t_db / 1e9 / peak},
            "roofline": {"bound": "hbm", "peak": peak, "peak_source": src, "unit": "GB/s"},
            "checks": {"max_rel_err_vs_oracle_2M_samples": err, "ok": err <= 1e-4}}
    print(json.dumps(line), flush=True)
    assert err <= 1e-4
    eng.close()


def run_files(args):
    """The whole headless job on wav files ("a second number including host wav decode + H2D"):
    N synthetic 10-minute PCM_16 clips are written to a scratch folder, then timed by wall clock:
      detect  = softspoken_b200.corpus.detect_corpus (reader thread -> ss_detect_host_batch_pcm16 -> CSV text),
                once with the reader thread and once reading in line as the reference does;
      review  = softspoken_b200.review (minimum-length filter, all erase, review CSV + the three export trees);
      silence = softspoken_b200.silencer.SilenceWorker (read wav, mask on the GPU, encode + write <stem>_silenced.wav).
    Checked: same CSV text with and without the reader thread; every sample of a silenced file is the input sample
    re-encoded, or zero inside an erase interval."""
    import shutil
    import tempfile
    import pandas as pd
    from softspoken_b200 import corpus, review, silencer, synth, wavio
    n_files = min(args.files, 64)
    root = tempfile.mkdtemp(prefix="ss_files_", dir=args.scratch)
    try:
        base = [synth.synth_pcm16(600.0, s) for s in range(4)]
        files = []
        for i in range(n_files):
            p = os.path.join(root, f"site{i % 3}", f"clip_{i:04d}.wav")
            os.makedirs(os.path.dirname(p), exist_ok=True)
            wavio.write_wav_pcm16(p, np.roll(base[i % 4], 997 * i), SR)
            files.append(p)
        eng = load_engine(1005, args.mode)
        hours = n_files * 600.0 / 3600.0
        # warm-up on two files (page cache is warm either way: the files were just written)
        corpus.detect_corpus(files[:2], eng.detect_host_batch, load=corpus.load_native_22050, group_size=2)
        res = {}
        for name, depth in (("inline", 0), ("prefetch", 2)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rows = corpus.detect_corpus(files, eng.detect_host_batch, load=corpus.load_native_22050, group_size=4, prefetch=depth)
            text = corpus.csv_text(rows)
            res[name] = (time.perf_counter() - t0, text)
        assert res["inline"][1] == res["prefetch"][1]
        det_csv = os.path.join(root, "p_detections.csv")
        with open(det_csv, "w", newline="") as f:
            f.write(res["prefetch"][1])
        t0 = time.perf_counter()
        table = review.ReviewTable.open(det_csv, None)
        table.erase_all("2026-01-01 00:00:00")
        df = review.save_review(table, os.path.join(root, "p_review.csv"), root, "p")
        review_s = time.perf_counter() - t0
        out_dir = os.path.join(root, "silenced")
        os.makedirs(out_dir)
        t0 = time.perf_counter()
        silencer.SilenceWorker(pd.read_csv(os.path.join(root, "p_review.csv")), out_dir, engine=eng).run()
        torch.cuda.synchronize()
        silence_s = time.perf_counter() - t0
        # check one file per folder sample by sample
        ok, checked = True, 0
        for p in files[:3]:
            rows_f = df[(df["file_path"] == os.path.dirname(p)) & (df["file_name"] == os.path.basename(p))]
            outp = os.path.join(out_dir, os.path.basename(p)[:-4] + "_silenced.wav")
            if not len(rows_f):
                continue
            x, _ = wavio.read_wav(p)
            want = wavio.encode_pcm16(x)
            for s, e in zip(rows_f["start_time"], rows_f["end_time"]):
                a, b = silencer.row_to_samples(s, e, SR, len(x))
                want[a:b] = 0
            got, _ = wavio.read_wav_pcm16(outp)
            ok = ok and np.array_equal(got, want)
            checked += 1
        silenced = len({(a, b) for a, b in zip(df["file_path"], df["file_name"])})
        line = {
            "metric": "audio_hours_per_sec", "unit": "audio-hours/s", "n_gpus": 1, "higher_is_better": True,
            "value": hours / res["prefetch"][0], "data": "synthetic", "dtype": args.mode,
            "config": {"workload": f"config2 from files: {n_files} x 10-min mono PCM_16 wavs ({n_files * 26.46:.0f} MB) in {args.scratch}, "
                                   "wall clock of read + RIFF parse + upload + detect + CSV rows (softspoken_b200.corpus)"},
            "detect": {"with_reader_thread_s": res["prefetch"][0], "read_in_line_s": res["inline"][0],
                       "audio_hours_per_s_in_line": hours / res["inline"][0], "detections": len(rows),
                       "same_csv_both_ways": True},
            "review": {"rows_after_minimum_length_filter": len(df), "seconds_incl_three_exports": review_s},
            "silence_voices": {"files_written": silenced, "seconds": silence_s,
                               "audio_hours_per_s": silenced * 600.0 / 3600.0 / silence_s if silenced else None,
                               "files_checked_sample_exact": checked, "ok": bool(ok)},
        }
        print(json.dumps(line), flush=True)
        assert ok
        eng.close()
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["long", "silence", "files", "postproc", "spectrogram", "config1", "resample"])
    ap.add_argument("--scratch", default="/dev/shm" if os.path.isdir("/dev/shm") else None,
                    help="folder for the wav files of the `files` workload")
    ap.add_argument("--hours", type=float, default=24.0)
    ap.add_argument("--files", type=int, default=1000)
    ap.add_argument("--intervals", type=int, default=10000)
    ap.add_argument("--max-batch", type=int, default=1024)       # = the streaming chunk (1,024 windows)
    ap.add_argument("--pcm16", action="store_true", help="config 4 from the int16 samples of a PCM_16 recording")
    ap.add_argument("--mode", default="f16x3")
    args = ap.parse_args()
    {"long": run_long, "silence": run_silence, "files": run_files, "postproc": run_postproc, "spectrogram": run_spectrogram, "config1": run_config1, "resample": run_resample}[args.what](args)


if __name__ == "__main__":
    main()
