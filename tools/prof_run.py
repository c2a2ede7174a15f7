"""Small driver for ncu: one 10-minute clip's worth of windows through K1 + classifier (+K5/K6)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint, synth  # noqa: E402
from softspoken_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="f16x3")
ap.add_argument("--windows", type=int, default=128)
ap.add_argument("--max-batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
sd = checkpoint.synthetic_state_dict(0, head)
eng = Engine(sd, 0, max_batch=args.max_batch, mode=args.mode)
secs = (args.windows - 5) * 0.6
audio = torch.from_numpy(synth.synth_audio(secs, 0)).cuda()
for _ in range(args.reps):
    reg, n, lg = eng.detect_device(audio, want_logits=True)
torch.cuda.synchronize()
print("windows", lg.shape[0], "regions", int(n.item()))
