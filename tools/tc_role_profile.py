"""Per-role cycle breakdown of the persistent tcgen05 conv kernel, one launch at a time (debug tool; the launch
names assume one launch per convolution, i.e. SS_TC_FUSE unset or 0)."""
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

with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
MODE = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=B, mode=MODE)
mel = torch.rand(B, 128, 256, device="cuda")
eng.classify(mel)
torch.cuda.synchronize()
names = ["conv1_1.c2"]           # conv1_1's first convolution is the CUDA-core kernel conv1_direct, not a tcgen05 launch
for rb in ["conv2_1", "conv3_1", "conv4_1", "bottleneck", "encoder_out", "conv6", "conv7", "conv8", "conv9_1"]:
    names += [rb + ".c1", rb + ".c2+res"]
buf = np.zeros((148, 8), np.int64)
for i, name in enumerate(names):
    check(lib.ss_debug_tc_profile(eng._ctx, i, None))
    eng.classify(mel)
    check(lib.ss_debug_tc_profile(eng._ctx, -1, C.c_void_p(buf.ctypes.data)))
    act = buf[buf[:, 7] > 0]
    if len(act) == 0:
        print(name, "no data"); continue
    m = act.mean(axis=0)
    units = act[:, 7].mean()
    print(f"{name:18s} ctas={len(act):3d} units/cta={units:5.1f} | mma total {m[3]:9.0f} cyc ({m[3]/units:7.0f}/unit) "
          f"wait_full {100*m[2]/m[3]:5.1f}% wait_acc_empty {100*m[1]/m[3]:5.1f}% | epi total {m[5]:9.0f} wait_acc_full "
          f"{100*m[4]/max(m[5],1):5.1f}% | producer wait_empty {m[0]:9.0f} ({100*m[0]/m[3]:5.1f}% of mma total)")
