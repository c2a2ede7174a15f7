"""GPU box: logits of the config-scale parity cases (10-minute pool clips, hour 0 of the config-4 stream) in the
requested classifier modes -> gpurun_out/scale_logits_<mode>.npz, for offline comparison with tests/golden/scale_*.npz
(tools/scale_parity_report.py).  Test infrastructure."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import synth  # noqa: E402
from tools.bench_aux import load_engine  # noqa: E402

modes = sys.argv[1:] or ["f16x3", "fp32"]
cases = {"clip0": synth.synth_audio(600.0, 0), "clip1": synth.synth_audio(600.0, 1),
         "hour0": synth.stream_hour(24, 0)}
for mode in modes:
    eng = load_engine(1005, mode)
    out = {}
    for name, audio in cases.items():
        t0 = time.perf_counter()
        bins, lg = eng.detect_host(audio, want_logits=True, cap=1 << 16)
        out[name + "_logits"] = lg
        out[name + "_regions"] = bins
        print(f"{mode} {name}: {lg.shape[0]} windows, {len(bins)} regions, {time.perf_counter() - t0:.2f} s", flush=True)
    np.savez(os.path.join(ROOT, "gpurun_out", f"scale_logits_{mode}.npz"), **out)
    eng.close()
