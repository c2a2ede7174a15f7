"""Parity of the CUDA path with the REAL reference at the scale the metric is quoted on.  Test infrastructure.

Compares `Engine.detect_host` (K1 -> classifier -> K5 -> K6, with or without the margin-guided refinement) with the
golden vectors `oracle/make_golden_scale.py` froze from the reference's own classes:

  clip0, clip1   whole 10-minute clips of the bench pool (config 2: 1,005 windows, 51,661 emitted bins each)
  hour0          the first hour of the config-4 stream (6,005 windows, 307,661 emitted bins)

and reports, per case: the largest logit error, the smallest reference margin |avg - 0.1|, the number of timeline
bins whose hot / not-hot decision differs from the reference's (with the reference margin of each), and the number
of detection rows (merged regions) that differ.  Used by tests/test_gpu_scale.py and, as a script on the GPU box,
to write profiles/r2_scale_parity.json:

    python tools/scale_parity.py > gpurun_out/scale_parity.json
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {"clip0": "scale_clip_seed0.npz", "clip1": "scale_clip_seed1.npz", "hour0": "scale_stream_hour0.npz"}


def case_audio(name: str) -> np.ndarray:
    from softspoken_b200 import synth
    if name.startswith("clip"):
        return synth.synth_audio(600.0, int(name[4:]))
    g = np.load(os.path.join(GOLDEN, CASES[name]))
    return synth.stream_hour(int(g["stream_seed"]), 0)


def compare(name: str, bins: np.ndarray, logits: np.ndarray) -> dict:
    """GPU result of one case (region bins `[R,2]`, logits `[W,256]`) against its golden file."""
    from oracle import postproc as pp
    g = np.load(os.path.join(GOLDEN, CASES[name]))
    stride = int(g["logits_stride"]) if "logits_stride" in g.files else 1
    ref_lg = g["logits"]
    got_lg = logits[::stride]
    assert got_lg.shape == ref_lg.shape, (got_lg.shape, ref_lg.shape)
    n_emit = int(g["n_emitted"])
    hot_ref = np.unpackbits(g["hot_bits"])[:n_emit].astype(bool)
    # the averaged timeline of the GPU logits, by the oracle's float64 arithmetic (K5 is bit-identical to it:
    # tests/test_gpu_postproc.py), gives the GPU's decision per bin
    avg, cnt = pp.average_idx(logits.reshape(-1, 1, 256), int(g["n_padded"]) / 22050)
    assert int((cnt >= 1).sum()) == n_emit
    hot_gpu = avg[:n_emit] > 0.1
    flips = np.flatnonzero(hot_gpu != hot_ref)
    near = dict(zip(g["near_idx"].tolist(), g["near_avg"].tolist()))
    flip_margins = [abs(near[int(j)] - 0.1) if int(j) in near else float("inf") for j in flips]
    ref_rows = {tuple(r) for r in g["region_bins"].tolist()}
    got_rows = {tuple(r) for r in np.asarray(bins, dtype=np.int64).tolist()}
    # regions recomputed from the GPU logits by the oracle must be the regions K6 returned
    k6_consistent = bool(np.array_equal(pp.find_speech_regions_idx(avg, cnt), np.asarray(bins, dtype=np.int64).reshape(-1, 2)))
    return {
        "case": name, "windows": int(logits.shape[0]), "emitted_bins": n_emit, "reference_rows": len(ref_rows),
        "max_logit_err": float(np.max(np.abs(got_lg.astype(np.float64) - ref_lg))),
        "max_abs_ref_logit": float(np.max(np.abs(ref_lg))),
        "min_ref_margin": float(g["min_margin"]),
        "bins_within_1e-5_of_threshold": int(np.sum(np.abs(g["near_avg"] - 0.1) < 1e-5)),
        "differing_bins": int(len(flips)), "differing_bin_ref_margins": [float(m) for m in flip_margins],
        "differing_rows": int(len(ref_rows ^ got_rows)), "k6_matches_oracle_on_gpu_logits": k6_consistent,
    }


def run(eng, names=("clip0", "clip1", "hour0")) -> list:
    out = []
    for name in names:
        audio = case_audio(name)
        eng.refine_stats(reset=True)
        t0 = time.perf_counter()
        bins, lg = eng.detect_host(audio, want_logits=True, cap=1 << 16)
        dt = time.perf_counter() - t0
        r = compare(name, bins, lg)
        r["refine"] = eng.refine_stats()
        r["seconds"] = dt
        out.append(r)
    return out


def main():
    from tools.bench_aux import load_engine
    report = {"what": "CUDA path vs the real reference at config scale (tools/scale_parity.py)", "runs": []}
    configs = [("f16x3, library default (refinement off)", "f16x3", None),
               ("f16x3 + fp32 refinement eps 1e-5", "f16x3", 1e-5), ("fp32 (CUDA cores)", "fp32", 0.0)]
    for label, mode, eps in configs:
        eng = load_engine(1005, mode)
        if eps is not None:
            eng.set_refine(eps)
        eng.detect_host(case_audio("clip0")[:22050 * 30])        # warm-up: workspace, first launches
        report["runs"].append({"config": label, "cases": run(eng)})
        eng.close()
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
