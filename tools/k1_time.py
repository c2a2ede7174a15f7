"""K1 alone: CUDA-event time of ss_features over one 10-minute clip, one environment variable at 0 against 1 (default
SS_MEL_WALK: band-by-band mel walk against the two-band walk; `python tools/k1_time.py SS_K1_PACKED`: scalar against
packed phase 1), two contexts each, median of 15 (profiles/r2_tuning_experiments.txt, sections 13 and 17)."""
import json, os, sys, torch, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint, synth
from softspoken_b200.engine import Engine
from oracle import postproc as pp
import numpy as np
head = json.load(open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")))
sd = checkpoint.synthetic_state_dict(0, head)
audio = synth.synth_audio(600.0, 0)
padded = torch.from_numpy(pp.pad_audio(audio)).cuda()
starts = torch.from_numpy(pp.plan_windows(600.0))
res = {}
VAR = sys.argv[1] if len(sys.argv) > 1 else "SS_MEL_WALK"
for walk in ("0", "1", "0", "1"):
    os.environ[VAR] = walk
    eng = Engine(sd, 0, max_batch=8, mode="fp32")
    for _ in range(3): eng.features(padded, starts)
    ts = []
    for _ in range(15):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m = eng.features(padded, starts); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res.setdefault(walk, []).append(statistics.median(ts))
    eng.close()
print("K1 ms per 10-min clip (", len(starts), "windows ):", VAR, "= 0", res["0"], " = 1", res["1"])
