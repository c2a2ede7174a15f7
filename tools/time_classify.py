"""Tuning: classifier time per 10-minute clip (1005 windows) under the current environment (SS_TC_* knobs).

    SS_TC_FUSE=0 python tools/time_classify.py [mode] [max_batch] [ref.pt]

Prints ms per 1005 windows (CUDA events, 3 repetitions after 2 warm-ups).  With a third argument the logits are
saved to / compared bit for bit with that file, so that schedule changes can be shown not to change a single bit.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint  # noqa: E402
from softspoken_b200.engine import Engine  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ref_path = sys.argv[3] if len(sys.argv) > 3 else None
with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=B, mode=mode)
W = 1005
torch.manual_seed(0)
mel = torch.rand(W, 128, 256, device="cuda") * 1.5
for _ in range(2):
    lg = eng.classify(mel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    lg = eng.classify(mel)
e1.record()
torch.cuda.synchronize()
eng.check_health()
ms = e0.elapsed_time(e1) / reps
same = ""
if ref_path:
    if os.path.exists(ref_path):
        same = f"  bit-identical to {os.path.basename(ref_path)}: {bool(torch.equal(torch.load(ref_path), lg.cpu()))}"
    else:
        torch.save(lg.cpu(), ref_path)
        same = f"  (saved {os.path.basename(ref_path)})"
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SS_TC_"))
print(f"[{knobs or 'defaults'}] mode={mode} batch={B}: {ms:8.3f} ms per 1005 windows "
      f"({6.359672832e9 * W / ms / 1e9:7.1f} TFLOP/s algorithmic){same}", flush=True)
