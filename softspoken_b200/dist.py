"""Multi-GPU: per-file sharding and the one collective of the path.

The reference is single-process; the file loop of `ProcessWorker.run`
(root/code/backend/worker.py:49) has no cross-file state except the running `ID`, so files shard
across ranks with no data-path exchange.  One process per GPU (`torchrun`), each rank runs K1-K6 on
its files, and the `(file_index, start_bin, end_bin)` triplets are gathered to rank 0 once at the end
of the corpus (count exchange + padded gather; a few KB, latency-bound).  Rank 0 orders by file index
and assigns `ID = 1..` in file-list order (worker.py:107-124), so the CSV is byte-identical to the
single-GPU run.  Works on `nccl` (GPU tensors) and `gloo` (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_files(durations: Sequence[float], world_size: int) -> List[List[int]]:
    """Static longest-processing-time-first assignment of file indices to ranks.

    Deterministic: ties broken by file index, then by rank; with equal durations this is round-robin.
    Each rank's list is returned in ascending file order."""
    order = sorted(range(len(durations)), key=lambda i: (-float(durations[i]), i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(durations[i])
    return [sorted(x) for x in out]


def gather_detections(local: np.ndarray, device: Optional[torch.device] = None, group=None) -> Optional[np.ndarray]:
    """local int32 `[K,3]` (file_index, start_bin, end_bin) -> on rank 0 all ranks' rows ordered by
    file index (stable, so a file's regions keep their ascending order); None on other ranks."""
    local = np.ascontiguousarray(local, dtype=np.int32).reshape(-1, 3)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _order(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros((cap, 3), dtype=torch.int32, device=device)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(local).to(device)
    if rank == 0:
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=0, group=group)
        rows = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)], axis=0)
        return _order(rows)
    dist.gather(buf, None, dst=0, group=group)
    return None


def _order(rows: np.ndarray) -> np.ndarray:
    if rows.shape[0] == 0:
        return rows.reshape(0, 3)
    return rows[np.argsort(rows[:, 0], kind="stable")]


def rows_from_triplets(files: Sequence[str], triplets: np.ndarray, next_id: int = 1) -> List[dict]:
    """Rank-0 row building in file-list order (worker.py:100-124)."""
    from .detector import region_bins_to_times
    from .worker import detection_rows
    rows: List[dict] = []
    # a file detected twice (e.g. by two runs that both journalled it) contributes its regions once
    triplets = np.asarray(triplets, dtype=np.int32).reshape(-1, 3)
    _, first = np.unique(triplets, axis=0, return_index=True)
    triplets = _order(triplets[np.sort(first)])                  # first occurrences, in file order, regions as given
    bounds = np.searchsorted(triplets[:, 0], np.arange(len(files) + 1))
    for fi in range(len(files)):
        sel = triplets[bounds[fi]:bounds[fi + 1], 1:3]
        new = detection_rows(files[fi], region_bins_to_times(sel), next_id)
        next_id += len(new)
        rows += new
    return rows
