"""Where does each classifier mode sit relative to the float64 truth?  (GPU box; test infrastructure.)

For 3 windows of the 60 s golden clip: the float64 evaluation of the oracle network is the truth; the
reference's own float32 CPU arithmetic (oracle/model.py, oneDNN) and every GPU mode are compared with it
layer by layer as max |delta| / max |truth|.  The table goes into DESIGN.md: it shows that f16x3 logits are
as close to the truth as the reference's own float32 is, while f16 / bf16 are the documented-tolerance modes.
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model as om  # noqa: E402
from oracle import postproc as pp  # noqa: E402
from softspoken_b200 import checkpoint, synth  # noqa: E402
from softspoken_b200._lib import lib, check  # noqa: E402
from softspoken_b200.engine import Engine  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
sd = checkpoint.synthetic_state_dict(0, head)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
clip = synth.synth_audio(60.0, 0)
padded = torch.from_numpy(pp.pad_audio(clip)).cuda()
starts = torch.tensor([0, 41 * 13230, 77 * 13230], dtype=torch.int64)
NAMES = ["conv1", "conv2", "conv3", "conv4", "bottleneck", "up(encoder_out)", "up(conv6)", "up(conv7)", "up(conv8)",
         "conv9"]
up = lambda t: torch.nn.functional.interpolate(t, scale_factor=2, mode="nearest")


def ref_taps(sdx, mel):
    taps = {}
    _, mk = om.forward_from_mel(sdx, mel, want_spec=False, taps=taps)
    acts = [taps["conv1"], taps["conv2"], taps["conv3"], taps["conv4"], taps["bottleneck"], up(taps["encoder_out"]),
            up(taps["conv6"]), up(taps["conv7"]), up(taps["conv8"]), taps["conv9"]]
    return acts, mk[:, 0]


def dump(eng, which, n):
    c, h, w = C.c_int(), C.c_int(), C.c_int()
    check(lib.ss_debug_activation(eng._ctx, which, n, None, C.byref(c), C.byref(h), C.byref(w), None))
    out = torch.empty((n, c.value, h.value, w.value), dtype=torch.float32, device="cuda")
    check(lib.ss_debug_activation(eng._ctx, which, n, C.c_void_p(out.data_ptr()), C.byref(c), C.byref(h), C.byref(w),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu()


rel = lambda a, t: float((a.double() - t).abs().max() / t.abs().max())
rows = {}
eng = Engine(sd, 0, max_batch=4, mode="f16x3")
mel = eng.features(padded, starts).cpu()
truth_acts, truth_logits = ref_taps(sd64, mel.double().unsqueeze(1))
cpu_acts, cpu_logits = ref_taps(sd, mel.unsqueeze(1))
rows["cpu f32 (reference arithmetic)"] = [rel(a, t) for a, t in zip(cpu_acts, truth_acts)] + [rel(cpu_logits, truth_logits)]
for mode in ["fp32", "f16x3", "f16", "bf16"]:
    lg, _ = eng.classify(mel.cuda(), want_spec=True, mode=mode)      # want_spec: conv9 is materialised only for the spec head
    lg = lg.cpu()
    if mode == "fp32":
        rows["gpu fp32 (CUDA cores)"] = [float("nan")] * 10 + [rel(lg, truth_logits)]
        continue
    acts = [dump(eng, i, 3) for i in range(10)]
    rows[f"gpu {mode} (tcgen05)"] = [rel(a, t) for a, t in zip(acts, truth_acts)] + [rel(lg, truth_logits)]
print("max |x - truth64| / max |truth64|, 3 windows of the seed-0 60 s clip, seed-0 checkpoint")
print(f"{'':34s}" + " ".join(f"{n[:9]:>9s}" for n in NAMES + ["logits"]))
for k, v in rows.items():
    print(f"{k:34s}" + " ".join(f"{x:9.2e}" for x in v))
print(f"max |logit| {float(truth_logits.abs().max()):.4f}")
eng.close()
