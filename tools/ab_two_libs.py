"""A/B of BUILDS of the library in one process — for changes that have no runtime knob.  Default: the current build
against tools/bin/libsoftspoken_b200_prev.so (the previous commit's sources built in a scratch worktree:
`git worktree add /tmp/prev HEAD && make -C /tmp/prev/softspoken_b200/csrc && cp .../libsoftspoken_b200.so tools/bin/...`);
AB_LIBS="name=path,name=path,..." compares any builds (at most three fit in 180 GB at max_batch 1005; the first is the
baseline; an empty path = the current build).  The libraries are loaded side by side (the package is imported once per
build under another name with SOFTSPOKEN_B200_LIB pointing at the file) and their classifiers run alternately on the
same input: whole-classifier CUDA-event time (median of `reps`, L2 flushed in between), MMA-warp kilocycles per launch
(median of `preps`; deterministic to ~0.2 %), bit-identity of the logits.

    [AB_LIBS=...] python tools/ab_two_libs.py [reps] [mode] [preps]
"""
import ctypes as C
import importlib.util
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
mode = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
preps = int(sys.argv[3]) if len(sys.argv) > 3 else 3


def load_package(alias, lib_path):
    """softspoken_b200 imported as `alias`, bound to `lib_path`."""
    if lib_path:
        os.environ["SOFTSPOKEN_B200_LIB"] = lib_path
    else:
        os.environ.pop("SOFTSPOKEN_B200_LIB", None)
    spec = importlib.util.spec_from_file_location(alias, os.path.join(ROOT, "softspoken_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ROOT, "softspoken_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    eng = importlib.import_module(alias + ".engine")
    lib = importlib.import_module(alias + "._lib")
    os.environ.pop("SOFTSPOKEN_B200_LIB", None)
    return mod, eng, lib


with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
    head = json.load(f)
builds = []
# AB_LIBS="name=path,name=path,..." compares any number of builds (first = the baseline); default: prev against current
pairs = [("ss_prev", os.path.join(ROOT, "tools", "bin", "libsoftspoken_b200_prev.so")), ("ss_new", None)]
if os.environ.get("AB_LIBS"):
    pairs = [(kv.split("=")[0], os.path.join(ROOT, kv.split("=")[1]) if kv.split("=")[1] else None)
             for kv in os.environ["AB_LIBS"].split(",")]
for alias, path in pairs:
    mod, eng_mod, lib_mod = load_package(alias, path)
    ck = importlib.import_module(alias + ".checkpoint")
    eng = eng_mod.Engine(ck.synthetic_state_dict(0, head), 0, max_batch=1005, mode=mode)
    builds.append((alias, eng, lib_mod))
    print(alias, "->", lib_mod.LIB_PATH, file=sys.stderr)
torch.manual_seed(0)
mel = torch.rand(1005, 128, 256, device="cuda") * 1.5
NB = len(builds)
times, outs = [[] for _ in builds], [None] * NB
for r in range(reps + 1):
    for e, (alias, eng, lib_mod) in enumerate(builds):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lg = eng.classify(mel); b.record()
        torch.cuda.synchronize()
        eng.check_health()
        if r > 0:
            times[e].append(a.elapsed_time(b))
        outs[e] = lg
        del flush
med = [statistics.median(t) for t in times]
for e, (alias, _, _) in enumerate(builds):
    print(f"{mode}: {alias:12s} {med[e]:.3f} ms   vs {builds[0][0]} {med[e] / med[0]:.4f}   bit-identical to it: "
          f"{bool(torch.equal(outs[0], outs[e]))}   (median of {reps}; min {min(times[e]):.3f})")
if mode == "f16x3" and preps > 0:
    names = ["conv1_1.c2"]
    for rb in ["conv2_1", "conv3_1", "conv4_1", "bottleneck", "encoder_out", "conv6", "conv7", "conv8", "conv9_1"]:
        names += [rb + ".c1", rb + ".c2+res"]
    buf = np.zeros((148, 8), np.int64)
    res = np.zeros((NB, len(names), preps))
    for r in range(preps):
        for e, (alias, eng, lib_mod) in enumerate(builds):
            for i in range(len(names)):
                lib_mod.check(lib_mod.lib.ss_debug_tc_profile(eng._ctx, i, None))
                eng.classify(mel)
                lib_mod.check(lib_mod.lib.ss_debug_tc_profile(eng._ctx, -1, C.c_void_p(buf.ctypes.data)))
                res[e, i, r] = buf[buf[:, 7] > 0][:, 3].max()
    medl = np.median(res, axis=2)
    print(f"{'launch (kcyc)':20s} " + " ".join(f"{b[0]:>10s}" for b in builds) + "   ratios to " + builds[0][0])
    for i, n in enumerate(names):
        print(f"{n:20s} " + " ".join(f"{medl[e, i] / 1e3:10.1f}" for e in range(NB)) + "   " +
              " ".join(f"{medl[e, i] / medl[0, i]:.3f}" for e in range(1, NB)))
    print(f"{'sum':20s} " + " ".join(f"{medl[e].sum() / 1e3:10.1f}" for e in range(NB)) + "   " +
          " ".join(f"{medl[e].sum() / medl[0].sum():.3f}" for e in range(1, NB)))
