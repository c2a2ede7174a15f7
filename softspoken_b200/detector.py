"""`NNDetector`: drop-in for the reference detector (root/code/frontend/NNDetector.py:11-190).

Same constructor, method names, argument meaning, return types and error
behaviour as the reference class, so `ProcessWorker` (reference
root/code/backend/worker.py:78,92,97) and `VoiceDetectorScreen`
(silencer_ui.py:225,228) can use it unchanged.  Every array operation runs in
the CUDA library; the host formats times and strings exactly as the reference
does (`f"{idx / (256 / 3):.4f}"`, NNDetector.py:185).

On top of the reference surface, `detect_file` runs the whole per-file body of
`ProcessWorker.run` (worker.py:57-100) in one call without per-batch host
round trips; that is the throughput path.
"""
from __future__ import annotations

import logging
import math
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import checkpoint, settings, spec, wavio
from ._lib import DEFAULT_MODE
from .model import SpecUNet_2D


def get_audio_data(file):
    """`voice_activity.get_audio_data` (root/code/backend/voice_activity.py:23-30): (duration_s, native_sr)
    from the header only."""
    return wavio.duration_and_rate(file)


class AveragedDetections(list):
    """The reference's `[(np.float64 avg, 'sss.ssss'), ...]` list, plus the arrays it was built from so that
    `find_speech_regions` need not re-parse ~50k strings per 10-minute file."""
    values: np.ndarray
    bins: np.ndarray


def bin_time_str(idx: int) -> str:
    return f"{idx / (256 / 3):.4f}"            # NNDetector.py:185


class NNDetector:
    def __init__(self, project_manager, mode: str = DEFAULT_MODE, device=None, max_batch: int = 32):
        if not torch.cuda.is_available():
            raise RuntimeError("softspoken_b200.NNDetector needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        logging.info(f"Device: {self.device}")
        torch.set_grad_enabled(False)           # NNDetector.py:26 (process-global in the reference too)

        self.project_manager = project_manager
        self.model = SpecUNet_2D(mode=mode, max_batch=max_batch).to(self.device)
        self.load_checkpoint(self.model, os.path.join(settings.model_dir, settings.model_name))
        self.model.eval()

        self.files_to_process = self.project_manager.get_unprocessed_list()
        self.detections_project = {f: [] for f in self.files_to_process}

    # ------------------------------------------------------------------ NNDetector.py:42-53
    def load_checkpoint(self, model, file_path='checkpoint.pth'):
        file_path = checkpoint.normalise_model_path(file_path)
        if os.path.exists(file_path):
            ck = torch.load(file_path, map_location="cpu", weights_only=True)
            model.load_state_dict(ck['model_state_dict'])
            start_epoch = ck['epoch'] + 1
            return start_epoch
        else:
            print("No checkpoint found. Starting training from scratch.")
            return -1

    # ------------------------------------------------------------------ NNDetector.py:55-82
    def plan_detection_job(self):
        p = self.detections_project
        for file in p.keys():
            logging.info(f"Analyzing file: {file}")
            (audio_len_seconds, _) = get_audio_data(file)
            p[file] = plan_windows_from_duration(audio_len_seconds)
        return p

    # ------------------------------------------------------------------ NNDetector.py:84-101
    def process_batch(self, audio_data, batch_indexes):
        """Padded float32 clip + window start indexes -> (speech_pred `[B,2,128,256]`, mask_pred `[B,1,256]`)
        as fresh numpy arrays.  Only the span the batch touches is uploaded (the reference copies the whole
        file for every batch, NNDetector.py:90)."""
        eng = self.model.engine
        idx = np.asarray(batch_indexes, dtype=np.int64).reshape(-1)
        audio = np.asarray(audio_data, dtype=np.float32)
        if idx.size == 0:
            raise RuntimeError("stack expects a non-empty TensorList")       # torch.stack([]) in the reference
        if (idx < 0).any() or (idx + spec.WINDOW_SAMPLES > audio.size).any():
            raise RuntimeError("stack expects each tensor to be equal size")   # ragged slices in the reference
        lo, hi = int(idx.min()), int(idx.max()) + spec.WINDOW_SAMPLES
        span = torch.from_numpy(np.ascontiguousarray(audio[lo:hi])).to(eng.device)
        starts = torch.from_numpy(idx - lo).to(eng.device)
        mel = eng.features(span, starts)
        logits, spec_out = eng.classify(mel, want_spec=True)
        out = spec_out.cpu().numpy(), logits.unsqueeze(1).cpu().numpy()
        # ss_classify only enqueues: a pipeline time-out or an activation beyond the fp16 range (an fp16-operand mode
        # on a checkpoint it does not fit) must fail this call, not end up as garbage rows in the detections CSV
        eng.check_health()
        return out

    # ------------------------------------------------------------------ NNDetector.py:103-143
    def find_speech_regions(self, averaged_detections, break_duration=0.5):
        threshold = settings.threshold
        speech_regions = {}
        for file, file_detections in averaged_detections.items():
            entries = file_detections[file]
            n = len(entries)
            if n == 0:
                speech_regions[file] = []
                continue
            if isinstance(entries, AveragedDetections):
                values, times = entries.values, None
            else:
                values = np.array([v for v, _ in entries], dtype=np.float64)
                times = [t for _, t in entries]
            eng = self.model.engine
            avg = torch.from_numpy(np.ascontiguousarray(values)).to(eng.device)
            cnt = torch.ones(n, dtype=torch.int32, device=eng.device)
            # The 0.5 s rule is exactly "gap <= 42 bins" (spec.GAP_BINS); any other break_duration is
            # applied on the host to the un-merged runs with the reference's own string arithmetic.
            exact = (break_duration == spec.BREAK_DURATION_S and isinstance(entries, AveragedDetections)
                     and entries.contiguous)
            runs = eng.regions(avg, cnt, threshold, spec.GAP_BINS if exact else 0, cap=max(16, n))
            tstr = (lambda k: entries[k][1]) if times is None else (lambda k: times[k])
            regions = [(tstr(int(s)), tstr(int(e))) for s, e in runs]
            if not exact and regions:
                merged, current = [], regions[0]
                for nxt in regions[1:]:
                    if float(nxt[0]) - float(current[1]) <= break_duration:
                        current = (current[0], nxt[1])
                    else:
                        merged.append(current)
                        current = nxt
                merged.append(current)
                regions = merged
            speech_regions[file] = regions
        return speech_regions

    # ------------------------------------------------------------------ NNDetector.py:145-151
    def extract_filename(self, file_path):
        full_filename = os.path.basename(file_path)
        return full_filename.rsplit('.', 1)[0]

    # ------------------------------------------------------------------ NNDetector.py:153-190
    def average_overlapping_detections(self, detections, audio_length_seconds, padding=0, min_count=1):
        averaged_detections = {}
        eng = self.model.engine
        for file, file_detections in detections.items():
            output_length = int(round(audio_length_seconds * 256 / 3))
            arr = np.asarray(file_detections, dtype=np.float32)
            W = 0 if arr.size == 0 else arr.reshape(-1, 256).shape[0]
            if W:
                last = spec.window_position(W - 1) + 256
                if last > output_length:       # numpy would raise a broadcast error in the reference
                    raise ValueError(f"operands could not be broadcast together: window {W - 1} ends at bin "
                                     f"{last} > output_length {output_length}")
                lg = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1, 256))).to(eng.device)
                avg_t, cnt_t = eng.average(lg, output_length)
                avg, cnt = avg_t.cpu().numpy(), cnt_t.cpu().numpy()
            else:
                avg, cnt = np.zeros(output_length), np.zeros(output_length, np.int32)
            keep = np.nonzero(cnt >= min_count)[0] if min_count >= 1 else np.arange(output_length)
            if min_count < 1:                  # count 0 bins: the reference divides 0/0 -> nan (with a warning)
                avg = np.where(cnt > 0, avg, np.nan)
            out = AveragedDetections((np.float64(avg[i]), bin_time_str(int(i) + padding)) for i in keep)
            out.values = avg[keep]
            out.bins = keep + padding
            out.contiguous = bool(keep.size == 0 or (keep[-1] - keep[0] + 1 == keep.size))
            averaged_detections[file] = out
        return averaged_detections

    # ------------------------------------------------------------------ throughput path (new)
    def detect_file(self, audio: np.ndarray, want_logits: bool = False):
        """Unpadded mono clip at 22,050 Hz (float32, or the int16 samples of a PCM_16 file) -> [(start_s, end_s)]
        exactly as `ProcessWorker.run` derives them (pad 3 s, window, classify, average, threshold, merge, subtract 3 s; worker.py:57-100)."""
        eng = self.model.engine
        res = eng.detect_host(audio, want_logits=want_logits)
        bins = res[0] if want_logits else res
        times = region_bins_to_times(bins)
        return (times, res[1]) if want_logits else times


def plan_windows_from_duration(audio_len_seconds: float) -> np.ndarray:
    """NNDetector.py:67-80, statement for statement."""
    sample_rate = settings.vad_resample
    window_size = 3
    step_size = settings.step_size
    audio_data_length = round(audio_len_seconds * sample_rate) + (window_size * 2 * sample_rate)
    samples_per_window = sample_rate * window_size
    samples_per_step = math.floor(sample_rate * step_size)
    num_windows = int(np.ceil((audio_data_length - samples_per_window) / samples_per_step))
    return np.arange(num_windows) * samples_per_step


_ROW_TIMES = np.zeros(0, dtype=np.float64)      # _ROW_TIMES[idx] = float(bin_time_str(idx)) - 3, grown on demand


def _row_times(idx: np.ndarray) -> np.ndarray:
    """`float(f"{idx / (256 / 3):.4f}") - 3` per bin index: the string round trip of NNDetector.py:185 and the shift
    of worker.py:100, evaluated by Python exactly as the reference does — once per distinct bin (a corpus of
    10-minute clips asks for the same 51,661 bins a quarter of a million times)."""
    global _ROW_TIMES
    idx = np.asarray(idx, dtype=np.int64)
    if idx.size == 0:
        return np.zeros(0, dtype=np.float64)
    top = int(idx.max()) + 1
    if top > len(_ROW_TIMES):
        if top - len(_ROW_TIMES) > 8 * idx.size + 65536:          # a few bins of a very long timeline: no table
            return np.array([float(bin_time_str(int(j))) - 3 for j in idx], dtype=np.float64)
        top = max(top, min(2 * len(_ROW_TIMES), 1 << 20))           # grow geometrically while the table is small
        ext = np.array([float(bin_time_str(j)) - 3 for j in range(len(_ROW_TIMES), top)], dtype=np.float64)
        _ROW_TIMES = np.concatenate([_ROW_TIMES, ext])
    return _ROW_TIMES[idx]


def region_bins_to_times(bins: np.ndarray) -> List[Tuple[float, float]]:
    """(start_bin, end_bin) -> (float(start_str) - 3, float(end_str) - 3) (NNDetector.py:185; worker.py:100)."""
    b = np.asarray(bins, dtype=np.int64).reshape(-1, 2)
    return list(zip(_row_times(b[:, 0]).tolist(), _row_times(b[:, 1]).tolist()))
