"""ResBlock by ResBlock: MMA-warp cycles of the two launches of a block (SS_TC_FUSE=0) against the single fused launch
(SS_TC_FUSE=1, extra knobs from argv: e.g. SS_TC_RING=0 SS_TC_LAG=300), interleaved in one process."""
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

extra = dict(kv.split("=") for kv in sys.argv[1:])
with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=1005, mode="f16x3")
mel = torch.rand(1005, 128, 256, device="cuda")
eng.classify(mel)
blocks = ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "bottleneck", "encoder_out", "conv6", "conv7", "conv8", "conv9_1"]
buf = np.zeros((148, 8), np.int64)


def launches(n):
    out = []
    for i in range(n):
        check(lib.ss_debug_tc_profile(eng._ctx, i, None))
        eng.classify(mel)
        check(lib.ss_debug_tc_profile(eng._ctx, -1, C.c_void_p(buf.ctypes.data)))
        out.append(float(buf[buf[:, 7] > 0][:, 3].max()))
    return out


res = {0: [], 1: []}
for rep in range(2):
    os.environ.pop("SS_TC_FUSE", None)
    for k in extra:
        os.environ.pop(k, None)
    a = launches(19)
    res[0].append([a[0]] + [a[1 + 2 * i] + a[2 + 2 * i] for i in range(9)])
    os.environ["SS_TC_FUSE"] = "1"
    os.environ.update(extra)
    b = launches(10)          # conv1_1 has one tcgen05 launch either way (its first convolution is conv1_direct)
    res[1].append(b)
un = np.median(np.array(res[0]), axis=0)
fu = np.median(np.array(res[1]), axis=0)
print(f"{'block':14s} {'two launches':>13s} {'fused':>10s}  fused/two   (kcyc; fused knobs: {extra})")
for n, x, y in zip(blocks, un, fu):
    print(f"{n:14s} {x / 1e3:13.1f} {y / 1e3:10.1f}  {y / x:.3f}")
print(f"{'sum':14s} {un.sum() / 1e3:13.1f} {fu.sum() / 1e3:10.1f}  {fu.sum() / un.sum():.3f}")
