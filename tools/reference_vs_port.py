"""Build container only: the REAL reference detector timed beside the oracle port on the same windows.

`bench.py`'s CPU arm (`cpu_baseline`, `--impl reference`) times the oracle port (`oracle/model.forward`), because
/root/reference cannot travel to the GPU box.  This script backs `kind: "port"` with a number: the reference's own
`NNDetector.process_batch` (root/code/frontend/NNDetector.py:84-101, incl. its whole-file `torch.tensor(audio_data)`
copy per batch, :90) and the port run over the same 64 windows (two batches of 32) of the same 10-minute clip, same
thread count, interleaved repetitions; logits must agree to float32 rounding.

    python tools/reference_vs_port.py > profiles/r2_reference_vs_port.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model as om, postproc as pp, ref_shim  # noqa: E402
from softspoken_b200 import checkpoint, synth  # noqa: E402

THREADS = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, (os.cpu_count() or 2) // 2)    # settings.py:32
with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    sd = checkpoint.synthetic_state_dict(0, json.load(f))
import contextlib
with contextlib.redirect_stdout(sys.stderr):      # the reference prints "No checkpoint found..." (NNDetector.py:52)
    ref = ref_shim.load()
    det = ref_shim.make_detector(ref, sd, threads=THREADS)
torch.set_num_threads(THREADS)
audio = synth.synth_audio(600.0, 0)
padded = pp.pad_audio(audio)
starts = pp.plan_windows(600.0)[:64]


def run_reference():
    out = []
    for s in range(0, 64, 32):
        _, mk = det.process_batch(padded, starts[s:s + 32])
        out.append(mk)
    return np.vstack(out)[:, 0]


def run_port():
    out = []
    for s in range(0, 64, 32):
        x = torch.stack([torch.from_numpy(padded[i:i + 66150]) for i in starts[s:s + 32]])
        _, mk = om.forward(sd, x, want_spec=True)
        out.append(mk.numpy())
    return np.vstack(out)[:, 0]


run_reference(); run_port()                      # warm-up
t_ref, t_port = [], []
for _ in range(3):
    t0 = time.perf_counter(); a = run_reference(); t_ref.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); b = run_port(); t_port.append(time.perf_counter() - t0)
print(json.dumps({
    "what": "real reference NNDetector.process_batch vs oracle port (oracle/model.forward), 64 windows of the 10-min "
            "clip seed 0, build container", "threads": THREADS, "host_cores": os.cpu_count(),
    "reference_s_per_64_windows": [round(t, 3) for t in t_ref], "port_s_per_64_windows": [round(t, 3) for t in t_port],
    "reference_ms_per_window": 1e3 * min(t_ref) / 64, "port_ms_per_window": 1e3 * min(t_port) / 64,
    "reference_over_port": min(t_ref) / min(t_port),
    "max_abs_logit_difference": float(np.abs(a - b).max()),
    "note": "the port calls the same torch-CPU kernels (bit-identical logits); it builds the batch with torch.stack "
            "of numpy slices where the reference copies the whole padded file into a tensor per batch "
            "(NNDetector.py:90, 53 MB for a 10-minute clip).  The two are within run-to-run noise of each other; "
            "bench.py's CPU numbers can be read as the reference's own to within that ratio"}, indent=1))
