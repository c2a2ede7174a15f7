"""A/B of a per-call knob on the whole classifier: CUDA-event time of ss_classify over 1,005 windows under environment
A and environment B, interleaved in one process (same box, clocks and thermal state), median of `reps`.

    python tools/ab_classify.py SS_TC_FUSE=0 SS_TC_FUSE=1 [reps] [mode]

Only knobs that are read per call can be compared this way (SS_TC_FUSE, SS_TC_LAG, SS_TC_RING, SS_TC_PAIR_STORE,
SS_TC_LAYOUT, SS_TC_EPI, SS_TC_FUSE_HEAD); SS_TC_SUB* are read once per process.  Also checks bit-identity of the logits.
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint  # noqa: E402
from softspoken_b200.engine import Engine  # noqa: E402

envs = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:3]]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
mode = sys.argv[4] if len(sys.argv) > 4 else "f16x3"
with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=1005, mode=mode)
torch.manual_seed(0)
mel = torch.rand(1005, 128, 256, device="cuda") * 1.5
times, outs = [[], []], [None, None]
for r in range(reps + 1):
    for e, env in enumerate(envs):
        for k, v in env.items():
            os.environ[k] = v
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lg = eng.classify(mel); b.record()
        torch.cuda.synchronize()
        eng.check_health()
        if r > 0:
            times[e].append(a.elapsed_time(b))
        outs[e] = lg
        for k in env:
            os.environ.pop(k, None)
        del flush
ma, mb = statistics.median(times[0]), statistics.median(times[1])
print(f"A {envs[0]}: {ma:.3f} ms   B {envs[1]}: {mb:.3f} ms   B/A {mb / ma:.4f}   bit-identical: {bool(torch.equal(outs[0], outs[1]))}"
      f"   (median of {reps}, min A {min(times[0]):.3f} B {min(times[1]):.3f})")
