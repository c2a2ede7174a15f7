"""Benchmark of the voice-detector batch path (BASELINE.json metric: audio-hours/sec detected).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode fp32|bf16]

Workload (BASELINE.json configs[1]): the 1,000-clip corpus of 10-minute mono 22,050 Hz clips, processed
one batch of clips per step; a step is `--clips-per-step` clips per GPU (weak scaling) through the whole
hot path (virtual 3 s padding -> K1 features -> K2/K3 classifier -> K5 averaging -> K6 regions).  Clips
are synthetic (seeded noise + speech-like bursts), drawn round-robin from a pool resident in HBM that is
larger than L2; weights are the seeded synthetic checkpoint (the shipped one is a missing blob).

`value`  = audio-hours per second with the PCM already resident in HBM (ss_detect_device);
`e2e`    = the same through the reference-facing C-ABI call with pinned HOST buffers (ss_detect_host:
           H2D of the clip and D2H of the region list inside the timed region);
`roofline` = the classifier (dominant kernel family) against the measured bf16 tensor peak, its time taken
           with CUDA events on the launching stream; `roofline_features` = K1 against measured HBM.
`cpu_baseline` = the oracle port of the reference's CPU detector timed on this host's cores on a bounded
           sample of the same clip (rank 0, N=1 only).
`--impl reference` times that CPU path alone (the reference is pure Python and cannot travel; the
oracle is its arithmetic on the same torch-CPU kernels — see oracle/model.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_S = 600.0
SR = 22050
WINDOWS_PER_CLIP = 1005
FLOP_PER_WINDOW_MASK = 6_359_672_832          # convs on the mask path, 2 x MAC, BN folded
FEATURE_BYTES_PER_CLIP = 4 * (13_230_000 + 132_300) + WINDOWS_PER_CLIP * 128 * 256 * 4   # PCM once + mel once
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_traffic.json")


def kernel_source_hash() -> str:
    """sha1 over the CUDA sources: ties an ncu traffic record to the kernels it was taken from."""
    import glob
    import hashlib
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(ROOT, "softspoken_b200", "csrc", "*.cu*"))):
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def measured_traffic(mode: str, kernel: str):
    """DRAM bytes per window of a kernel family from the ncu record tools/collect_profiles.sh wrote for THIS build
    (profiles/r2_traffic.json carries the hash of the CUDA sources it was taken from; a record of other sources is
    stale and reported as null rather than as a number of some earlier kernel)."""
    try:
        with open(TRAFFIC_FILE) as f:
            rec = json.load(f)
    except (OSError, ValueError):
        return None, f"{os.path.relpath(TRAFFIC_FILE, ROOT)} missing"
    if rec.get("kernel_source_sha1") != kernel_source_hash():
        return None, f"{os.path.relpath(TRAFFIC_FILE, ROOT)} is stale (taken from sources {rec.get('kernel_source_sha1')})"
    v = rec.get("dram_bytes_per_window", {}).get(mode, {}).get(kernel)
    return v, (f"ncu dram__bytes_read.sum + dram__bytes_write.sum, {rec.get('how', '')} "
               f"(sources {rec.get('kernel_source_sha1')}, {rec.get('when', '')})")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def load_state_dict():
    from softspoken_b200 import checkpoint
    with open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")) as f:
        head = json.load(f)
    return checkpoint.synthetic_state_dict(0, head)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started before the warm-up (nvidia-smi itself takes
    a few hundred ms to come up); `mark()` brackets the timed region and `stop()` summarises the samples inside it
    (falling back to all samples under load if the region was shorter than a sampling period)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def mark(self, begin: bool):
        if begin:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        time.sleep(0.06)                    # let the sample that covers the end of the region arrive
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) >= 6 and self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        where = "timed region"
        if not rows:
            rows, where = [r for _, r in self.rows if len(r) >= 6], "warm-up + timed region"
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": where}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_detector_sample(sd, seconds_budget: float, threads: int):
    """The reference's CPU detector (oracle port) on a bounded sample of one 10-minute clip: batches of 32
    windows (worker.py:71-79) until `seconds_budget` is spent, then averaging + regions of what was run."""
    from oracle import model as om
    from oracle import postproc as pp
    from softspoken_b200 import synth
    torch.set_num_threads(threads)
    audio = synth.synth_audio(CLIP_S, 0)
    padded = pp.pad_audio(audio)
    starts = pp.plan_windows(CLIP_S)
    preds, t0, done = [], time.perf_counter(), 0
    while done < len(starts) and (done == 0 or time.perf_counter() - t0 < seconds_budget):
        idx = starts[done:done + 32]
        x = torch.stack([torch.from_numpy(padded[i:i + 66150]) for i in idx])
        _, mk = om.forward(sd, x, want_spec=True)          # the reference computes the spec head too
        preds.append(mk.numpy())
        done += len(idx)
    lg = np.vstack(preds)
    secs = ((done - 1) * 13230 + 66150) / SR
    pp.find_speech_regions(pp.average_overlapping(lg, secs))
    dt = time.perf_counter() - t0
    audio_hours = done * 0.6 / 3600.0                      # each window advances the clip by 0.6 s
    return audio_hours / dt, done, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = load_state_dict()
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_detector_sample(sd, 0.0, threads)               # one batch of 32 windows
    vals, windows, t_total = [], 0, 0.0
    for _ in range(args.steps):
        v, n, dt = cpu_detector_sample(sd, args.ref_step_seconds, threads)
        vals.append(v); windows += n; t_total += dt
    value = (windows * 0.6 / 3600.0) / t_total
    line = {
        "impl": "reference", "metric": "audio_hours_per_sec", "value": value, "unit": "audio-hours/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: 10-min mono 22.05 kHz clip of the 1000-clip corpus, CPU detector on a bounded "
                               f"sample (~{args.ref_step_seconds:.0f} s of batches of 32 windows per step)"},
        "cpu_baseline": {"value": value, "unit": "audio-hours/s", "cores": threads, "kind": "port",
                         "sample": f"{windows} windows in {t_total:.1f} s over {args.steps} steps"},
        "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "x_realtime": value * 3600.0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def pool_clips(n_clips: int, seed0: int):
    """The bench pool: 10-minute clips of the seeded generator every test uses (`synth.synth_pcm16(600, seed)`: noise +
    60 speech-like bursts, PCM_16 levels).  Clips 0 and 1 of rank 0 are the clips whose reference results are frozen
    in tests/golden/scale_clip_seed{0,1}.npz (tests/test_gpu_scale.py)."""
    from softspoken_b200 import synth
    return [synth.synth_pcm16(CLIP_S, seed0 + c) for c in range(n_clips)]


def write_corpus(pool16, n_files: int, root: str, distinct: bool):
    """`n_files` PCM_16 wavs under `root` (tmpfs): the pool clips written once each, the rest hard links to them
    (or full copies with `distinct`)."""
    import shutil
    from softspoken_b200 import wavio
    os.makedirs(root, exist_ok=True)
    if distinct and shutil.disk_usage(root).free < (n_files + len(pool16) + 8) * (pool16[0].nbytes + 4096):
        print(f"[bench] {root}: not enough room for {n_files} distinct files, using hard links", file=sys.stderr)
        distinct = False
    base = []
    for c, pcm in enumerate(pool16):
        path = os.path.join(root, f"pool_{c}.wav")
        wavio.write_wav_pcm16(path, pcm, SR)
        base.append(path)
    files = []
    for i in range(n_files):
        path = os.path.join(root, f"clip_{i:05d}.wav")
        if os.path.exists(path):
            os.remove(path)
        if distinct:
            shutil.copyfile(base[i % len(base)], path)
        else:
            os.link(base[i % len(base)], path)
        files.append(path)
    return files, distinct


def run_b200(args):
    import torch.distributed as dist
    from softspoken_b200 import _lib, corpus, dist as ssdist
    from softspoken_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    sd = load_state_dict()
    eng = Engine(sd, local, max_batch=args.max_batch, mode=args.mode)
    if args.refine_eps is not None:
        eng.set_refine(args.refine_eps)
    n = int(CLIP_S * SR)
    cap = 4096
    eng.reserve(n, cap)
    C = args.clips_per_step
    pool16 = pool_clips(args.pool, seed0=args.pool * rank)
    # host side: pinned int16 (what a PCM_16 wav stores) and pinned float32 (what load_audio returns: k / 32768, exact);
    # device side: the float32 clips resident in HBM
    pinned16 = [torch.from_numpy(p).pin_memory() for p in pool16]
    pinned = [(torch.from_numpy(p).to(torch.float32) / 32768.0).pin_memory() for p in pool16]
    pool = [p.to(device) for p in pinned]
    reg_bufs = [(torch.empty((cap, 2), dtype=torch.int32, device=device), torch.zeros(1, dtype=torch.int32, device=device))
                for _ in range(C)]
    stream = torch.cuda.current_stream(device)

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    import ctypes as Cc
    from softspoken_b200._lib import lib, check, MODES

    # Config 3 (N > 1): files shard per rank, every rank keeps its (file, start_bin, end_bin) triplets and ONE gather
    # to rank 0 closes the run (softspoken_b200/dist.py) — inside the timed region, after the last step.
    trip: list = []

    def step_device(s):
        for j in range(C):
            clip = pool[(s * C + j) % len(pool)]
            reg, cnt = reg_bufs[j]
            check(lib.ss_detect_device(eng._ctx, Cc.c_void_p(clip.data_ptr()), n, MODES[args.mode],
                                       Cc.c_void_p(reg.data_ptr()), Cc.c_void_p(cnt.data_ptr()), cap, None,
                                       Cc.c_void_p(stream.cuda_stream)))
            if world > 1:                               # the step's detections leave the reusable device buffers
                k = int(cnt.item())
                fi = np.full((k, 1), (s * C + j) * world + rank, np.int32)
                trip.append(np.concatenate([fi, reg[:k].cpu().numpy()], 1))

    def step_host(s, src=None):
        total_regions = 0
        src = pinned if src is None else src
        clips = [src[(s * C + j) % len(src)] for j in range(C)]
        for j, bins in enumerate(eng.detect_host_batch(clips, cap=cap)):   # one C-ABI call per step
            total_regions += len(bins)
            if world > 1:
                trip.append(np.concatenate([np.full((len(bins), 1), (s * C + j) * world + rank, np.int32), bins], 1))
        return total_regions

    def gather_all():
        rows = ssdist.gather_detections(np.concatenate(trip) if trip else np.zeros((0, 3), np.int32), device)
        trip.clear()
        return rows

    def timed(fn, steps, warmup):
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        for s in range(warmup):
            fn(s)
        if world > 1:
            gather_all()                                 # warm the gather path too
        sync_all()
        if sampler:
            sampler.mark(True)
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        last = None
        for s in range(steps):
            last = fn(warmup + s)
        e1.record(stream)
        if world > 1:
            gather_all()
        sync_all()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.mark(False)
        dev_s = e0.elapsed_time(e1) / 1e3
        launches = _lib.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        # host-driven paths (detect_host) are timed by the wall clock between the two synchronisations;
        # device-resident paths by CUDA events on the launching stream
        return dev_s, wall, launches, clocks, last

    dev_s, wall_s, launches, clocks, _ = timed(step_device, args.steps, args.warmup)
    t_dev = max(dev_s, 1e-9) if world == 1 else wall_s        # multi-rank runs end with the gather to rank 0
    eng.refine_stats(reset=True)
    e_dev_s, e_wall_s, _, _, n_regions = timed(step_host, args.steps, max(1, args.warmup // 2))
    refine = eng.refine_stats()
    _, e16_wall_s, _, _, n_regions16 = timed(lambda s: step_host(s, pinned16), args.steps, max(1, args.warmup // 2))

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_dev = max_over_ranks(t_dev)
    t_e2e = max_over_ranks(e_wall_s)
    t_e2e16 = max_over_ranks(e16_wall_s)
    hours = args.steps * C * world * CLIP_S / 3600.0
    value = hours / t_dev
    e2e = hours / t_e2e
    e2e16 = hours / t_e2e16

    # ---- from wav files (the real config 2 / 3 job): `corpus_files` PCM_16 wavs in tmpfs, sharded per file over the
    # ranks, read + parsed + uploaded + detected by softspoken_b200.corpus.detect_corpus, ONE gather at the end
    files_line = None
    n_files = args.corpus_files if args.corpus_files >= 0 else 8 * world
    if n_files > 0:
        root = os.path.join(args.corpus_dir, f"ss_bench_{os.getuid()}")
        t_w = time.perf_counter()
        if rank == 0:
            files, was_distinct = write_corpus(pool16, n_files, root, args.corpus_distinct)
        sync_all()
        if rank != 0:
            files = [os.path.join(root, f"clip_{i:05d}.wav") for i in range(n_files)]
        t_w = time.perf_counter() - t_w
        durations = [CLIP_S] * n_files
        stats: dict = {}
        load = lambda path: corpus.load_native_22050(path, eng)
        corpus.detect_corpus(files[:2 * world], eng.detect_host_batch, load=load, durations=durations[:2 * world],
                             device=device, group_size=args.corpus_group)      # warm-up: page cache, reader thread
        sync_all()
        t0 = time.perf_counter()
        text = corpus.detect_corpus(files, eng.detect_host_batch, load=load, durations=durations, device=device,
                                    group_size=args.corpus_group, stats=stats, as_csv=True)
        if rank == 0:
            rows = text.splitlines()[1:]
        t_files = max_over_ranks(time.perf_counter() - t0)
        one_rank_sha = None
        if args.corpus_check_1rank and world > 1:
            # the same list through ONE rank (rank 0; the others wait): the N-rank CSV must be byte-identical
            if rank == 0:
                import hashlib
                t1 = time.perf_counter()
                text1 = corpus.detect_corpus(files, eng.detect_host_batch, load=load, durations=durations, device=device,
                                             group_size=args.corpus_group, as_csv=True, local_only=True)
                one_rank_sha = {"csv_sha1": hashlib.sha1(text1.encode()).hexdigest(), "identical": text1 == text,
                                "seconds": time.perf_counter() - t1}
            sync_all()
        if rank == 0:
            import hashlib
            files_line = {"value": n_files * CLIP_S / 3600.0 / t_files, "unit": "audio-hours/s", "files": n_files,
                          "seconds": t_files, "x_realtime": n_files * CLIP_S / t_files,
                          "storage": f"{args.corpus_dir} ({'distinct copies' if was_distinct else 'hard links'} of "
                                     f"{args.pool} PCM_16 clips, 26.5 MB each)",
                          "rank0_split_s": {k: round(v, 4) for k, v in stats.items()},
                          "rows": len(rows), "csv_sha1": hashlib.sha1(text.encode()).hexdigest(),
                          "one_rank_run_of_the_same_list": one_rank_sha,
                          "corpus_write_s": round(t_w, 2), "group_size": args.corpus_group,
                          "what": "wall clock of softspoken_b200.corpus.detect_corpus: file read + RIFF parse + int16 "
                                  "upload + detect + one gather + row building + CSV text; max over ranks"}
        sync_all()
        if rank == 0 and not args.keep_corpus:
            import shutil
            shutil.rmtree(root, ignore_errors=True)

    # ---- per-kernel-family times for the roofline: one clip's 1,005 windows, CUDA events on torch's stream
    pk = peaks()
    clip = pool[0]
    padded = eng.pad(clip)
    starts = torch.arange(WINDOWS_PER_CLIP, device=device, dtype=torch.int64) * 13230

    def ev_time(fn, reps):
        fn(); torch.cuda.synchronize(device)
        ts = []
        for _ in range(reps):
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=device).zero_()   # > L2
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize(device)
            ts.append(a.elapsed_time(b) / 1e3)
            del flush
        return sum(ts) / len(ts)

    t_feat = ev_time(lambda: eng.features(padded, starts), 5)
    mel = eng.features(padded, starts)
    t_cls = ev_time(lambda: eng.classify(mel, mode=args.mode), 3)
    achieved_tf = FLOP_PER_WINDOW_MASK * WINDOWS_PER_CLIP / t_cls / 1e12
    achieved_gbs = FEATURE_BYTES_PER_CLIP / t_feat / 1e9
    other_modes = {}
    if world == 1 and not args.no_other_modes:
        for m in ("f16", "bf16"):
            if m != args.mode:
                t_m = ev_time(lambda: eng.classify(mel, mode=m), 2)
                other_modes[m] = {"ms_per_clip": 1e3 * t_m, "tflops": FLOP_PER_WINDOW_MASK * WINDOWS_PER_CLIP / t_m / 1e12,
                                  "frac": FLOP_PER_WINDOW_MASK * WINDOWS_PER_CLIP / t_m / 1e12 / pk["tflops_sustained"]}
    bad_guard_bytes = eng.check_guards()

    line = None
    if rank == 0:
        tr_cls, tr_cls_src = measured_traffic(args.mode, "classifier")
        tr_feat, tr_feat_src = measured_traffic(args.mode, "features")
        line = {
            "metric": "audio_hours_per_sec", "value": value, "unit": "audio-hours/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "f16": "f16", "f16x3": "f16x3 (fp16 hi/lo split operands, fp32 accumulate)"}[args.mode],
            "data": "synthetic",
            "config": {"workload": f"config2: 1000x10-min mono 22.05 kHz corpus, step = {C} clip(s)/GPU "
                                   f"({C * WINDOWS_PER_CLIP} windows) through pad+features+classifier+average+regions"
                                   + ("; config3: files sharded per rank, one gather of the detections to rank 0 inside "
                                      "the timed region" if world > 1 else ""),
                       "classifier_mode": args.mode, "clips_per_step_per_gpu": C, "pool_clips": len(pool),
                       "pool": "synth.synth_pcm16(600 s, seed): seeds 0.. on rank 0 — clips 0 and 1 are the clips of "
                               "tests/golden/scale_clip_seed{0,1}.npz (reference results frozen)",
                       "l2": f"inputs larger than L2: pool of {len(pool)} clips x 53 MB rotates; "
                             "classifier activations stream through a per-batch workspace",
                       "max_batch_windows": args.max_batch, "parallelism": f"files sharded over {world} GPU(s)"},
            "x_realtime": value * 3600.0,
            "e2e": {"value": e2e, "unit": "audio-hours/s", "h2d_bytes_per_step": C * n * 4,
                    "d2h_bytes_per_step": int(C * 4 + (n_regions or 0) * 8), "x_realtime": e2e * 3600.0},
            # same call with the int16 samples of PCM_16 files as host buffers (ss_detect_host_batch_pcm16: decode fused
            # into K1, bit-identical detections, half the upload)
            "e2e_pcm16": {"value": e2e16, "unit": "audio-hours/s", "h2d_bytes_per_step": C * n * 2,
                          "d2h_bytes_per_step": int(C * 4 + (n_regions16 or 0) * 8), "x_realtime": e2e16 * 3600.0,
                          "same_regions_as_float32": bool(n_regions16 == n_regions)},
            "e2e_files": files_line,
            "refinement": {"eps": args.refine_eps if args.refine_eps is not None else 0.0, "mode": "fp32",
                           "windows": refine["windows"], "windows_refined": refine["windows_refined"],
                           "note": "margin-guided fp32 re-classification (ss_ctx_set_refine); off by default since "
                                   "the split-K sub-accumulation put f16x3 logits in the reference's own noise class "
                                   "(profiles/r2_scale_parity.json)"},
            "guard_bytes_overwritten": int(bad_guard_bytes),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "classifier (conv stack, ss_classify)", "achieved": achieved_tf,
                         "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tflops_sustained"],
                         "frac_of_burst_peak": achieved_tf / pk["tflops_burst"],
                         "traffic": (tr_cls * WINDOWS_PER_CLIP) if tr_cls else None, "traffic_source": tr_cls_src,
                         "peak_source": pk["source"], "ms_per_clip": 1e3 * t_cls,
                         "executed_mma_flops_factor": 3 if args.mode == "f16x3" else 1,
                         "single_pass_modes": other_modes,
                         "algorithmic_flops_per_launch_group": FLOP_PER_WINDOW_MASK * WINDOWS_PER_CLIP},
            "roofline_features": {"bound": "hbm", "kernel": "features_kernel (K1)", "achieved": achieved_gbs,
                                  "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / pk["hbm_gbs"],
                                  "traffic": (tr_feat * WINDOWS_PER_CLIP) if tr_feat else None,
                                  "traffic_source": tr_feat_src, "peak_source": pk["source"], "ms_per_clip": 1e3 * t_feat,
                                  "algorithmic_bytes_per_launch": FEATURE_BYTES_PER_CLIP},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, nwin, dt = cpu_detector_sample(sd, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": v, "unit": "audio-hours/s", "cores": threads, "kind": "port",
                                    "sample": f"{nwin} windows (batches of 32) of one 10-min clip in {dt:.1f} s, "
                                              "oracle port of the reference CPU detector incl. spec head",
                                    "port_vs_real_reference": "profiles/r2_reference_vs_port.json (build container: the "
                                                              "real NNDetector.process_batch timed beside the port)"}
            if threads >= 2:
                # the reference itself runs on half the cores (settings.py:32, NNDetector.py:25): a short second sample
                vh, nh, dth = cpu_detector_sample(sd, args.cpu_seconds / 3.0, threads // 2)
                line["cpu_baseline"]["reference_default_threads"] = {
                    "value": vh, "cores": threads // 2, "sample": f"{nh} windows in {dth:.1f} s"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("SS_BENCH_MODE", "f16x3"), choices=["fp32", "bf16", "f16", "f16x3"])
    ap.add_argument("--clips-per-step", type=int, default=2)
    ap.add_argument("--pool", type=int, default=4)
    ap.add_argument("--max-batch", type=int, default=1005)      # one 10-minute clip per classifier batch (36 GB of f16x3 activations)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-step-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--refine-eps", type=float, default=None, help="margin-guided refinement (ss_ctx_set_refine); default: library default")
    ap.add_argument("--corpus-files", type=int, default=-1,
                    help="wav files of the from-files leg (e2e_files); -1 = 8 per GPU, 0 = skip, 1000 = the real config 3")
    ap.add_argument("--corpus-dir", default="/dev/shm")
    ap.add_argument("--corpus-distinct", action="store_true", help="full copies instead of hard links")
    ap.add_argument("--corpus-group", type=int, default=4)
    ap.add_argument("--keep-corpus", action="store_true")
    ap.add_argument("--corpus-check-1rank", action="store_true",
                    help="N > 1: rank 0 also runs the whole list alone and compares the CSV text")
    ap.add_argument("--no-other-modes", action="store_true")
    args = ap.parse_args()
    # stdout carries the ONE JSON line of the contract: whatever libraries print there meanwhile (NCCL announces its
    # version on stdout) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
