"""Print the launch configuration (geometry, ring depth, resident weights, stage program length) of every tcgen05 conv
launch of one classify call (SS_TC_VERBOSE=1)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint
from softspoken_b200.engine import Engine
head = json.load(open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")))
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=8, mode="f16x3")
mel = torch.rand(8, 128, 256, device="cuda")
eng.classify(mel); torch.cuda.synchronize()
os.environ["SS_TC_VERBOSE"] = "1"
eng.classify(mel); torch.cuda.synchronize()
