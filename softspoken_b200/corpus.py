"""Headless corpus driver: "Run Voice Detector" over a file list on 1..N GPUs of one box (BASELINE config 3).

The reference processes its file list in one loop on one device (root/code/backend/worker.py:49-136) and saves
the detections CSV after every file.  Here the list is sharded per file across ranks (one process per GPU,
`torchrun`), every rank streams its files through `ss_detect_host_batch` in groups, and the
`(file_index, start_bin, end_bin)` triplets are gathered to rank 0 once at the end (softspoken_b200/dist.py) —
the only collective of the path.  Rank 0 assigns `ID = 1..` in file-list order and writes the same CSV text the
reference's `DetectionProject.save_detections` would (silencer_ui.py:816-817), so the output is byte-identical
for any number of ranks.

    torchrun --nproc-per-node 8 -m softspoken_b200.corpus files.txt detections.csv [--checkpoint model.pth]
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import dist as ssdist
from . import spec, wavio


def load_mono_22050(path: str, engine=None) -> np.ndarray:
    """`voice_activity.load_audio` for a corpus file: wav -> float32 mono at 22,050 Hz.  Multi-channel files are
    averaged as `librosa.to_mono` does (voice_activity.py:61-63).  Other rates go through the K9 resampling kernel
    when an `engine` is given (this package's documented filter, not the reference's soxr: SURVEY §8c — such files
    are outside the bit-exactness claims) and are refused without one."""
    x, sr = wavio.read_wav(path)
    if x.ndim > 1:
        x = np.mean(x, axis=0).astype(np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    if sr != spec.SAMPLE_RATE:
        if engine is None:
            raise ValueError(f"{path}: sample rate {sr} != {spec.SAMPLE_RATE} and no engine to resample it")
        from .worker import resample_on_device
        x = resample_on_device(x, sr, engine)
    return x


def load_native_22050(path: str, engine=None) -> np.ndarray:
    """Like `load_mono_22050`, but a mono PCM_16 file is returned as its int16 samples (the library decodes them on
    the device: `ss_detect_host_batch_pcm16`), so the host never builds the float32 copy and the upload is half the
    size.  Detections are bit-identical to the float32 route (tests/test_gpu_pcm16.py)."""
    got = wavio.read_wav_pcm16(path)
    if got is not None and got[0].ndim == 1 and got[1] == spec.SAMPLE_RATE:
        return np.ascontiguousarray(got[0])
    return load_mono_22050(path, engine)


class Journal:
    """Restartable corpus runs.  The reference keeps no record of finished files —
    `get_unprocessed_list` returns every file (root/code/frontend/silencer_ui.py:668-686), so a restart re-processes the
    corpus and appends duplicate rows.  Here every rank appends one line per finished file to its own
    `<prefix>.rank<r>` (no collective, flushed per group of files):

        <sha1 of the file list>\t<file index>\t<start_bin>,<end_bin>;<start_bin>,<end_bin>;...

    A later run — with ANY number of ranks — reads all `<prefix>.rank*`, skips the files found there (a file with no
    detection has a line with an empty last field), shards only the rest and merges old and new triplets before the
    rows are numbered, so the final CSV is byte-identical to an uninterrupted run.  Lines of another file list (hash
    mismatch) and torn last lines are ignored."""

    def __init__(self, prefix: str, files: Sequence[str], rank: int):
        import hashlib
        self.prefix = prefix
        self.key = hashlib.sha1("\n".join(files).encode("utf-8", "surrogatepass")).hexdigest()
        self.n_files = len(files)
        self.path = f"{prefix}.rank{rank}"
        self._fh = None

    def load(self):
        """-> {file_index: int32 [R,2]} of every complete line written for this file list by any earlier rank."""
        import glob
        done = {}
        for path in sorted(glob.glob(glob.escape(self.prefix) + ".rank*")):
            with open(path, "r") as f:
                text = f.read()
            lines = text.split("\n")[:-1]              # a line counts only once its newline is on disk
            for ln in lines:
                parts = ln.split("\t")
                if len(parts) != 3 or parts[0] != self.key:
                    continue
                try:
                    fi = int(parts[1])
                    pairs = [tuple(int(v) for v in p.split(",")) for p in parts[2].split(";") if p]
                except ValueError:
                    continue
                if 0 <= fi < self.n_files and all(len(p) == 2 for p in pairs):
                    done[fi] = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        return done

    def append(self, file_index: int, bins: np.ndarray) -> None:
        if self._fh is None:
            torn = False
            if os.path.exists(self.path) and os.path.getsize(self.path) > 0:
                with open(self.path, "rb") as f:
                    f.seek(-1, os.SEEK_END)
                    torn = f.read(1) != b"\n"
            self._fh = open(self.path, "a")
            if torn:                      # a crash left half a line: terminate it so that it cannot swallow the next one
                self._fh.write("\n")
        body = ";".join(f"{int(s)},{int(e)}" for s, e in np.asarray(bins).reshape(-1, 2))
        self._fh.write(f"{self.key}\t{int(file_index)}\t{body}\n")

    def sync(self) -> None:
        if self._fh is not None:
            self._fh.flush()
            os.fsync(self._fh.fileno())

    def close(self) -> None:
        if self._fh is not None:
            self.sync()
            self._fh.close()
            self._fh = None


def _prefetched(groups, load_one: Callable[[int], np.ndarray], depth: int = 2):
    """Yield `(group, [clip, ...])` in order while a reader thread stays up to `depth` groups ahead.

    A 10-minute PCM_16 clip is 26 MB to read and 27 ms of GPU work; read in line (the reference's `load_audio`
    per file, worker.py:52) the file system would sit on the critical path for a good part of that time.
    `np.fromfile` and the ctypes call into the library both release the GIL, so one plain thread overlaps them.
    A failure in the reader surfaces at the group it belongs to; leaving the loop early stops the reader."""
    if depth <= 0:
        for g in groups:
            yield g, [load_one(i) for i in g]
        return
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=depth)
    stop = threading.Event()

    def put(item) -> bool:
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                pass
        return False

    def reader():
        try:
            for g in groups:
                if not put((g, [load_one(i) for i in g])):
                    return
            put(None)
        except BaseException as e:      # handed to the consumer, which re-raises it
            put(e)

    t = threading.Thread(target=reader, name="softspoken-prefetch", daemon=True)
    t.start()
    try:
        while True:
            item = q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        t.join(timeout=5.0)


def detect_corpus(files: Sequence[str], detect_batch: Callable[[List[np.ndarray]], List[np.ndarray]],
                  load: Callable[[str], np.ndarray] = load_mono_22050, durations: Optional[Sequence[float]] = None,
                  group_size: int = 8, device: Optional[torch.device] = None, next_id: int = 1,
                  journal: Optional[str] = None, prefetch: int = 2, stats: Optional[dict] = None,
                  as_csv: bool = False, local_only: bool = False):
    """-> list of CSV row dicts on rank 0 (None on the other ranks); with `as_csv` the CSV text itself
    (`csv_text_from_triplets`: the same bytes without building a dict per row).

    `detect_batch(clips) -> [int32 [R,2] region bins per clip]` is `Engine.detect_host_batch` (or any stand-in
    with that contract: the CPU tests drive this function with the oracle).  `journal`: path prefix of the
    per-rank progress files that make the run restartable (`Journal`).  `prefetch`: groups of files a reader
    thread decodes ahead of the GPU (0 = read in line, as the reference does).  `stats`: a dict that receives where
    this rank's wall time went (`wait_files_s`, `detect_s`, `gather_rows_s`)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() and not local_only else 1
    rank = dist.get_rank() if world > 1 else 0          # local_only: this process does the whole list by itself
    if durations is None:
        durations = [wavio.duration_and_rate(f)[0] for f in files]
    jr = Journal(journal, files, rank) if journal else None
    # What earlier runs finished must be ONE view shared by all ranks: a rank that read the progress files a moment
    # later would see lines another rank had just appended, shard a different `todo` list, and files would be
    # detected twice or not at all.  Rank 0 reads, everybody takes its answer; nobody appends before that.
    done = jr.load() if (jr and rank == 0) else {}
    if jr and world > 1:
        box = [done]
        dist.broadcast_object_list(box, src=0, device=device)
        done = box[0]
    todo = [i for i in range(len(files)) if i not in done]
    mine = [todo[k] for k in ssdist.shard_files([durations[i] for i in todo], world)[rank]]
    # triplets of earlier runs enter the gather once, through rank 0
    parts = [np.concatenate([np.full((len(b), 1), i, np.int32), b], axis=1) for i, b in sorted(done.items())] if rank == 0 else []
    groups = [mine[g0:g0 + group_size] for g0 in range(0, len(mine), group_size)]
    import time
    t_wait = t_detect = 0.0
    t_mark = time.perf_counter()
    try:
        for idx, clips in _prefetched(groups, lambda i: load(files[i]), depth=prefetch):
            t_now = time.perf_counter()
            t_wait += t_now - t_mark
            results = detect_batch(clips)
            t_mark = time.perf_counter()
            t_detect += t_mark - t_now
            for i, bins in zip(idx, results):
                b = np.asarray(bins, dtype=np.int32).reshape(-1, 2)
                parts.append(np.concatenate([np.full((len(b), 1), i, np.int32), b], axis=1))
                if jr:
                    jr.append(i, b)
            if jr:
                jr.sync()
    finally:
        if jr:
            jr.close()
    t_mark = time.perf_counter()
    local = np.concatenate(parts) if parts else np.zeros((0, 3), np.int32)
    allrows = ssdist._order(local) if local_only else ssdist.gather_detections(local, device)
    if rank != 0:
        rows = None
    elif as_csv:
        rows = csv_text_from_triplets(list(files), allrows, next_id)
    else:
        rows = ssdist.rows_from_triplets(list(files), allrows, next_id)
    if stats is not None:
        stats.update(wait_files_s=t_wait, detect_s=t_detect, gather_rows_s=time.perf_counter() - t_mark)
    return rows


def csv_text(rows) -> str:
    """The text `DetectionProject.save_detections` writes for these rows (silencer_ui.py:779-788,816-817):
    `DataFrame.to_csv(index=False)` prints ints as ints, floats with the shortest repr, empty strings as empty
    fields and quotes a field only when it must (csv.QUOTE_MINIMAL).  Written directly because appending ten
    thousand rows one `df.loc[len(df)] = row` at a time, as the reference does per file, is quadratic."""
    import csv
    import io
    from .worker import COLUMN_TYPES
    cols = list(COLUMN_TYPES.keys())
    buf = io.StringIO()
    w = csv.writer(buf, quoting=csv.QUOTE_MINIMAL, lineterminator="\n")
    w.writerow(cols)
    float_cols = [c in ("start_time", "end_time") for c in cols]
    w.writerows([repr(float(r[c])) if f else r[c] for c, f in zip(cols, float_cols)] for r in rows)
    return buf.getvalue()


def csv_text_from_triplets(files: Sequence[str], triplets: np.ndarray, next_id: int = 1) -> str:
    """`csv_text(dist.rows_from_triplets(files, triplets, next_id))` without the quarter of a million row dicts a
    1,000-file corpus makes: same bytes (tests/test_host.py), a fifth of the time on rank 0."""
    import csv
    import io
    from .detector import _row_times
    from .worker import COLUMN_TYPES, basename, dirname
    triplets = np.asarray(triplets, dtype=np.int32).reshape(-1, 3)
    _, first = np.unique(triplets, axis=0, return_index=True)
    triplets = ssdist._order(triplets[np.sort(first)])
    bounds = np.searchsorted(triplets[:, 0], np.arange(len(files) + 1))
    starts = _row_times(triplets[:, 1]).tolist()
    ends = _row_times(triplets[:, 2]).tolist()
    out = [",".join(COLUMN_TYPES.keys()) + "\n"]
    quote = io.StringIO()
    for fi, file in enumerate(files):
        lo, hi = int(bounds[fi]), int(bounds[fi + 1])
        if hi == lo:
            continue
        quote.seek(0); quote.truncate()
        csv.writer(quote, quoting=csv.QUOTE_MINIMAL, lineterminator="").writerow([dirname(file), basename(file)])
        mid = quote.getvalue()                                   # the two path fields, quoted only if they must be
        out.extend(f"{next_id + k - lo},{mid},{starts[k]!r},{ends[k]!r},0,,\n" for k in range(lo, hi))
        next_id += hi - lo
    return "".join(out)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("file_list", help="text file with one wav path per line (the reference's <project>_files.txt)")
    ap.add_argument("out_csv")
    ap.add_argument("--checkpoint", default=None, help="reference checkpoint (.pth); seeded init if absent, as the reference")
    ap.add_argument("--mode", default=None)
    ap.add_argument("--max-batch", type=int, default=512)
    ap.add_argument("--group-size", type=int, default=4,
                    help="files per library call (uploads overlap compute inside a call; the reader thread loads the next groups meanwhile)")
    ap.add_argument("--resume", action="store_true",
                    help="keep per-rank progress files next to out_csv (<out_csv>.journal.rank<r>) and skip the files "
                         "an earlier, interrupted run of the same file list already finished")
    args = ap.parse_args(argv)
    from . import checkpoint
    from .engine import Engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    with open(args.file_list) as f:
        files = [ln.strip() for ln in f if ln.strip()]
    if args.checkpoint and os.path.exists(checkpoint.normalise_model_path(args.checkpoint)):
        sd = torch.load(checkpoint.normalise_model_path(args.checkpoint), map_location="cpu", weights_only=True)["model_state_dict"]
    else:
        print("No checkpoint found. Starting training from scratch.")     # NNDetector.py:52
        sd = checkpoint.synthetic_state_dict(0)
    eng = Engine(sd, local, max_batch=args.max_batch, **({"mode": args.mode} if args.mode else {}))
    import time
    durations = [wavio.duration_and_rate(f)[0] for f in files]
    # One-off costs out of the way before the clock starts: workspace allocation for the longest file (tens of GB for
    # a 1,005-window batch), the first launch of every kernel, the NCCL communicator and its gather path.
    t_init = time.perf_counter()
    from . import detector, worker  # noqa: F401  (row building imports them: 0.9 s on first use)
    eng.reserve(int(max(durations, default=0.0) * spec.SAMPLE_RATE) + 1)
    eng.detect_host_batch([np.zeros(spec.SAMPLE_RATE, np.int16)])
    if world > 1:
        ssdist.gather_detections(np.zeros((0, 3), np.int32), device)
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_init = t0 - t_init
    stats: dict = {}
    rows = detect_corpus(files, eng.detect_host_batch, load=lambda path: load_native_22050(path, eng),
                         durations=durations, device=device,
                         group_size=max(1, args.group_size), stats=stats, journal=(args.out_csv + ".journal") if args.resume else None,
                         as_csv=True)
    if rows is not None:
        with open(args.out_csv, "w", newline="") as f:
            f.write(rows)
        rows = rows.splitlines()[1:]
        dt = time.perf_counter() - t0          # rank 0 returns from the gather last: slowest rank + gather + CSV
        hours = sum(durations) / 3600.0
        print(f"{len(rows)} detections in {len(files)} files -> {args.out_csv}")
        print(f"{hours:.3f} audio-hours in {dt:.3f} s on {world} GPU(s): {hours / dt:.2f} audio-hours/s "
              f"({hours * 3600 / dt:,.0f}x real time; file read + upload + detect + gather + CSV; "
              f"{t_init:.1f} s of one-off initialisation before that)")
        print("rank 0: " + ", ".join(f"{k} {v:.3f}" for k, v in stats.items()) + f", csv_s {time.perf_counter() - t0 - sum(stats.values()):.3f}")
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
