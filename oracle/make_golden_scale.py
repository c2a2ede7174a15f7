"""Freeze config-scale golden vectors from the REAL reference into tests/golden/.  TEST INFRASTRUCTURE.

Build container only (needs /root/reference):

    python -m oracle.make_golden_scale [--clips 0 1] [--hour] [--threads 4]

What the 60 s goldens of make_golden.py cannot show is how the CUDA path compares with the reference at the
scale BASELINE.json's metric is quoted on.  This script runs the reference's own `NNDetector.process_batch`
(real `SpecUNet_2D`, the reference batching of 32 windows + ragged tail, worker.py:71-79), its own
`average_overlapping_detections` and `find_speech_regions` over

  * whole 10-minute clips of the bench pool (`synth.synth_audio(600, seed)`, config 2: 1,005 windows), and
  * the first hour of the config-4 stream (`synth.stream_hour(STREAM_SEED, 0)`: 6,005 windows),

and stores, per case: the logits (all windows for the clips; every `HOUR_STRIDE`-th window for the hour), the
hot decision of every emitted timeline bin as packed bits, every bin whose average lies within `NEAR` of the
threshold (index + float64 average), the smallest |avg - 0.1|, and the merged regions as bin indices and as the
reference's own time strings.  tests/test_gpu_scale.py compares the CUDA path with these files.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from softspoken_b200 import checkpoint, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NEAR = 1e-3
HOUR_STRIDE = 4
STREAM_SEED = 24


def reference_case(det, audio: np.ndarray, tag: str):
    padded = np.zeros(len(audio) + 2 * 66150, np.float32)            # worker.py:58-62
    padded[66150:66150 + len(audio)] = audio
    secs = len(padded) / 22050                                       # worker.py:89
    L = round(len(audio) / 22050 * 22050) + 6 * 22050                # NNDetector.py:72
    n_win = int(np.ceil((L - 66150) / 13230))
    starts = np.arange(n_win) * 13230
    t0 = time.perf_counter()
    preds = []
    for s in range(0, n_win, 32):                                    # worker.py:71-79
        _, mk = det.process_batch(padded, starts[s:s + 32])
        preds.append(mk)
        if (s // 32) % 16 == 0:
            print(f"  {tag}: {s + len(mk)}/{n_win} windows, {time.perf_counter() - t0:.0f} s", flush=True)
    logits = np.vstack(preds)                                        # [W,1,256]
    dt_model = time.perf_counter() - t0
    avg = det.average_overlapping_detections({tag: logits}, secs)    # NNDetector.py:153-190
    regions = det.find_speech_regions({tag: avg}, break_duration=0.5)[tag]
    vals = np.array([v for v, _ in avg[tag]], dtype=np.float64)
    times = [t for _, t in avg[tag]]
    idx_of_time = {t: i for i, t in enumerate(times)}                # emitted bins are 0..len-1 (count >= 1 prefix)
    reg_bins = np.array([[idx_of_time[a], idx_of_time[b]] for a, b in regions], dtype=np.int64).reshape(-1, 2)
    hot = vals > 0.1
    near = np.flatnonzero(np.abs(vals - 0.1) < NEAR)
    return dict(logits=logits[:, 0, :], n_emitted=np.array(len(vals)), hot_bits=np.packbits(hot),
                near_idx=near.astype(np.int64), near_avg=vals[near], min_margin=np.array(np.abs(vals - 0.1).min()),
                region_bins=reg_bins, region_times=np.array(regions).reshape(-1, 2),
                n_padded=np.array(len(padded)), model_seconds=np.array(dt_model)), vals


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--hour", action="store_true")
    ap.add_argument("--threads", type=int, default=4)    # settings.cpu_threads on the 8-core build container
    args = ap.parse_args()
    import json
    with open(os.path.join(GOLDEN, "head_seed0.json")) as f:
        head = json.load(f)
    sd = checkpoint.synthetic_state_dict(0, head)
    ref = ref_shim.load()
    det = ref_shim.make_detector(ref, sd, threads=args.threads)
    for seed in args.clips:
        out, _ = reference_case(det, synth.synth_audio(600.0, seed), f"clip{seed}")
        np.savez_compressed(os.path.join(GOLDEN, f"scale_clip_seed{seed}.npz"), threads=np.array(args.threads),
                            clip_seed=np.array(seed), **out)
        print(f"clip seed {seed}: {len(out['region_bins'])} regions, min margin {float(out['min_margin']):.3e}, "
              f"{len(out['near_idx'])} bins within {NEAR}", flush=True)
    if args.hour:
        out, _ = reference_case(det, synth.stream_hour(STREAM_SEED, 0), "hour0")
        out["logits"] = out["logits"][::HOUR_STRIDE].copy()
        np.savez_compressed(os.path.join(GOLDEN, "scale_stream_hour0.npz"), threads=np.array(args.threads),
                            stream_seed=np.array(STREAM_SEED), logits_stride=np.array(HOUR_STRIDE), **out)
        print(f"hour 0 of stream {STREAM_SEED}: {len(out['region_bins'])} regions, min margin "
              f"{float(out['min_margin']):.3e}, {len(out['near_idx'])} bins within {NEAR}", flush=True)


if __name__ == "__main__":
    main()
