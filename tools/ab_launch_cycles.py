"""A/B of a runtime knob, launch by launch: MMA-warp cycles of every tcgen05 conv launch under environment A and
under environment B, interleaved in one process (same box, same clocks, same thermal state).

    python tools/ab_launch_cycles.py SS_TC_PAIR_STORE=0 SS_TC_PAIR_STORE=1 [batch] [reps]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint  # noqa: E402
from softspoken_b200._lib import lib, check  # noqa: E402
from softspoken_b200.engine import Engine  # noqa: E402

envs = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:3]]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1005
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=B, mode="f16x3")
mel = torch.rand(B, 128, 256, device="cuda")
eng.classify(mel)
names = ["conv1_1.c2"]
for rb in ["conv2_1", "conv3_1", "conv4_1", "bottleneck", "encoder_out", "conv6", "conv7", "conv8", "conv9_1"]:
    names += [rb + ".c1", rb + ".c2+res"]
buf = np.zeros((148, 8), np.int64)
res = np.zeros((2, len(names), reps))
for r in range(reps):
    for e, env in enumerate(envs):
        for k, v in env.items():
            os.environ[k] = v
        for i in range(len(names)):
            check(lib.ss_debug_tc_profile(eng._ctx, i, None))
            eng.classify(mel)
            check(lib.ss_debug_tc_profile(eng._ctx, -1, C.c_void_p(buf.ctypes.data)))
            res[e, i, r] = buf[buf[:, 7] > 0][:, 3].max()          # slowest CTA's MMA-warp lifetime = the launch
        for k in env:
            os.environ.pop(k, None)
med = np.median(res, axis=2)
print(f"{'launch':20s} {'A kcyc':>10s} {'B kcyc':>10s}  B/A     A = {envs[0]}  B = {envs[1]}")
for i, n in enumerate(names):
    print(f"{n:20s} {med[0, i] / 1e3:10.1f} {med[1, i] / 1e3:10.1f}  {med[1, i] / med[0, i]:.3f}")
print(f"{'sum':20s} {med[0].sum() / 1e3:10.1f} {med[1].sum() / 1e3:10.1f}  {med[1].sum() / med[0].sum():.3f}")
