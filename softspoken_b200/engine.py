"""Thin Python face of the C ABI: one `Engine` = one `ss_ctx` on one GPU.

PyTorch is used for device memory and streams only (tensors in, tensors out,
`torch.cuda.current_stream()` handed to the kernels); every computation is a
call into libsoftspoken_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, checkpoint, spec
from ._lib import lib, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, state_dict, device: int | torch.device = 0, max_batch: int = 32, mode: str = _lib.DEFAULT_MODE):
        if not torch.cuda.is_available() or _lib.device_count() == 0:
            raise _lib.SoftspokenError(_lib.SS_E_NODEVICE,
                                       "no CUDA device: softspoken_b200 has no CPU fallback")
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.mode = mode
        self._resample_tables = {}          # sr_in -> polyphase table on the device (K9)
        self.max_batch = int(max_batch)
        blob = checkpoint.pack_blob(state_dict)
        self._blob = blob
        ctx = C.c_void_p()
        check(lib.ss_ctx_create(self.device.index, blob, len(blob), self.max_batch, C.byref(ctx)))
        self._ctx = ctx
        self._reserved_samples = -1
        self._region_cap = 0
        for name, want in (("sample_rate", spec.SAMPLE_RATE), ("window_samples", spec.WINDOW_SAMPLES),
                           ("step_samples", spec.STEP_SAMPLES), ("pad_samples", spec.PAD_SAMPLES),
                           ("n_frames", spec.N_FRAMES), ("n_mels", spec.N_MELS), ("gap_bins", spec.GAP_BINS),
                           ("threshold", spec.THRESHOLD), ("hop_length", spec.HOP_LENGTH)):
            got = _lib.get_constant(name)
            if got != want:
                raise RuntimeError(f"kernel constant {name}={got} disagrees with settings ({want})")

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            lib.ss_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _mode(self, mode: Optional[str]) -> int:
        m = mode or self.mode
        if m not in _lib.MODES:
            raise ValueError(f"unknown mode {m!r}; expected one of {sorted(_lib.MODES)}")
        return _lib.MODES[m]

    def set_refine(self, eps: float, mode: str = "fp32") -> None:
        """Margin-guided refinement of the detect_* calls (`ss_ctx_set_refine`): windows covering a timeline bin whose
        average lies within `eps` of the 0.1 threshold are classified again in `mode`; 0 switches it off."""
        check(lib.ss_ctx_set_refine(self._ctx, float(eps), _lib.MODES[mode]))

    def refine_stats(self, reset: bool = False) -> dict:
        st = (C.c_uint64 * 4)()
        check(lib.ss_ctx_refine_stats(self._ctx, st, int(reset)))
        return {"windows": int(st[0]), "windows_refined": int(st[1]), "clips": int(st[2]),
                "clips_refined": int(st[3])}

    def check_guards(self) -> int:
        """Bytes of the guard bands around the context's device allocations that a kernel overwrote (0 = none)."""
        bad, n = C.c_uint64(), C.c_int()
        check(lib.ss_debug_check_guards(self._ctx, C.byref(bad), C.byref(n)))
        assert n.value > 0
        return int(bad.value)

    def device_bytes(self) -> int:
        n = C.c_size_t()
        check(lib.ss_ctx_device_bytes(self._ctx, C.byref(n)))
        return n.value

    def reserve(self, max_samples: int, region_cap: int = 1 << 16) -> None:
        if max_samples > self._reserved_samples or region_cap > self._region_cap:
            check(lib.ss_ctx_reserve(self._ctx, max(int(max_samples), self._reserved_samples, 0),
                                     max(int(region_cap), self._region_cap)))
            self._reserved_samples = max(int(max_samples), self._reserved_samples)
            self._region_cap = max(int(region_cap), self._region_cap)

    # ------------------------------------------------------------------ kernel-level entry points
    def pad(self, pcm: torch.Tensor) -> torch.Tensor:
        pcm = self._f32(pcm)
        out = torch.empty(pcm.numel() + 2 * spec.PAD_SAMPLES, dtype=torch.float32, device=self.device)
        check(lib.ss_pad(self._ctx, _ptr(pcm), pcm.numel(), _ptr(out), self._stream()))
        return out

    def features(self, padded: torch.Tensor, starts: torch.Tensor) -> torch.Tensor:
        """padded f32 `[n]`, starts int64 `[W]` -> mel `[W,128,256]` (K1)."""
        padded = self._f32(padded)
        starts = starts.to(device=self.device, dtype=torch.int64).contiguous()
        W = starts.numel()
        mel = torch.empty((W, spec.N_MELS, spec.N_FRAMES), dtype=torch.float32, device=self.device)
        check(lib.ss_features(self._ctx, _ptr(padded), padded.numel(), _ptr(starts), W, _ptr(mel), self._stream()))
        return mel

    def classify(self, mel: torch.Tensor, want_spec: bool = False, mode: Optional[str] = None):
        """mel `[W,128,256]` -> logits `[W,256]` (and spec `[W,2,128,256]`) (K2-K4)."""
        mel = self._f32(mel)
        W = mel.shape[0]
        logits = torch.empty((W, spec.N_FRAMES), dtype=torch.float32, device=self.device)
        spec_out = (torch.empty((W, 2, spec.N_MELS, spec.N_FRAMES), dtype=torch.float32, device=self.device)
                    if want_spec else None)
        check(lib.ss_classify(self._ctx, _ptr(mel), W, _ptr(logits), _ptr(spec_out), self._mode(mode),
                              self._stream()))
        return (logits, spec_out) if want_spec else logits

    def average(self, logits: torch.Tensor, out_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """logits `[W,256]` -> (avg f64 `[out_len]`, count i32 `[out_len]`) (K5)."""
        logits = self._f32(logits).reshape(-1, spec.N_FRAMES) if logits.numel() else logits
        W = logits.shape[0] if logits.numel() else 0
        avg = torch.empty(out_len, dtype=torch.float64, device=self.device)
        cnt = torch.empty(out_len, dtype=torch.int32, device=self.device)
        check(lib.ss_average(self._ctx, _ptr(logits) if W else None, W, out_len, _ptr(avg), _ptr(cnt),
                             self._stream()))
        return avg, cnt

    def regions(self, avg: torch.Tensor, cnt: torch.Tensor, threshold: float = spec.THRESHOLD,
                gap_bins: int = spec.GAP_BINS, cap: int = 1 << 16) -> np.ndarray:
        """-> int32 `[R,2]` (start_bin, end_bin inclusive) on the host (K6)."""
        out_len = avg.numel()
        self.reserve(max(self._reserved_samples, int(out_len * 3 / 256 * spec.SAMPLE_RATE) + 1), cap)
        reg = torch.empty((cap, 2), dtype=torch.int32, device=self.device)
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(lib.ss_regions(self._ctx, _ptr(avg), _ptr(cnt), out_len, float(threshold), int(gap_bins), _ptr(reg),
                             _ptr(n), cap, self._stream()))
        k = int(n.item())
        if k > cap:
            raise _lib.SoftspokenError(_lib.SS_E_CAPACITY, f"{k} regions exceed capacity {cap}")
        return reg[:k].cpu().numpy()

    def silence(self, pcm: torch.Tensor, intervals: torch.Tensor) -> None:
        """In place: zero element ranges `intervals` int64 `[K,2]` of the flat f32 buffer `pcm` (K7)."""
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous()
        iv = intervals.to(device=self.device, dtype=torch.int64).contiguous()
        check(lib.ss_silence(self._ctx, _ptr(pcm), pcm.numel(), _ptr(iv), iv.shape[0], self._stream()))

    # ------------------------------------------------------------------ file-level entry points
    def detect_device(self, pcm: torch.Tensor, mode: Optional[str] = None, cap: int = 1 << 16,
                      want_logits: bool = False):
        """Unpadded device clip -> (regions i32 `[cap,2]` device, n i32 `[1]` device[, logits]).

        `pcm` is float32 (what `load_audio` returns) or int16 (the samples of a mono PCM_16 file as stored; K1
        decodes them as sample / 32768, bit-identical to the float32 route)."""
        pcm16 = isinstance(pcm, torch.Tensor) and pcm.dtype == torch.int16
        pcm = pcm.to(self.device).contiguous() if pcm16 else self._f32(pcm)
        n = pcm.numel()
        self.reserve(n, cap)
        reg = torch.empty((cap, 2), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        W = lib.ss_plan_windows(n)
        lg = torch.empty((W, spec.N_FRAMES), dtype=torch.float32, device=self.device) if want_logits else None
        fn = lib.ss_detect_device_pcm16 if pcm16 else lib.ss_detect_device
        check(fn(self._ctx, _ptr(pcm), n, self._mode(mode), _ptr(reg), _ptr(cnt), cap, _ptr(lg), self._stream()))
        return (reg, cnt, lg) if want_logits else (reg, cnt)

    def detect_host(self, audio: np.ndarray | torch.Tensor, mode: Optional[str] = None, cap: int = 1 << 16,
                    want_logits: bool = False):
        """Host mono clip -> int32 `[R,2]` region bins (host).  float32 (what `load_audio` returns) or int16 (the
        samples of a PCM_16 file as stored: half the upload, same bits out)."""
        n, ptr, pcm16, audio = _host_clip(audio)
        self.reserve(n, cap)
        reg = np.empty((cap, 2), dtype=np.int32)
        k = C.c_int()
        W = lib.ss_plan_windows(n)
        lg = np.empty((W, spec.N_FRAMES), dtype=np.float32) if want_logits else None
        fn = lib.ss_detect_host_pcm16 if pcm16 else lib.ss_detect_host
        check(fn(self._ctx, ptr, n, self._mode(mode), C.c_void_p(reg.ctypes.data), cap, C.byref(k),
                 C.c_void_p(lg.ctypes.data) if want_logits else None))
        if k.value > cap:             # the kernel reports the true count: once more with room for it
            return self.detect_host(audio, mode, cap=k.value, want_logits=want_logits)
        out = reg[:k.value].copy()
        return (out, lg) if want_logits else out

    def detect_host_batch(self, clips, mode: Optional[str] = None, cap: int = 4096):
        """Several host mono clips (numpy arrays or CPU tensors, ideally pinned; all float32 or all int16) -> list of
        int32 `[R,2]` region-bin arrays.  One library call: uploads overlap compute across clips."""
        ptrs, sizes, keep, kinds = [], [], [], set()
        for a in clips:
            k, ptr, pcm16, a = _host_clip(a)
            ptrs.append(ptr.value or 0); sizes.append(k); keep.append(a); kinds.add(pcm16)
        n = len(keep)
        if n == 0:
            return []
        if len(kinds) != 1:      # mixed sample types: one library call per type, results back in input order
            out = [None] * n
            for want in (False, True):
                idx = [i for i, a in enumerate(keep) if _host_clip(a)[2] == want]
                for i, r in zip(idx, self.detect_host_batch([keep[i] for i in idx], mode, cap)):
                    out[i] = r
            return out
        batch_fn = lib.ss_detect_host_batch_pcm16 if kinds.pop() else lib.ss_detect_host_batch
        self.reserve(max(sizes), cap)
        reg = np.empty((n, cap, 2), dtype=np.int32)
        cnt = (C.c_int * n)()
        check(batch_fn(self._ctx, n, (C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*sizes), self._mode(mode),
                       C.c_void_p(reg.ctypes.data), cap, cnt))
        out = []
        for i in range(n):
            if cnt[i] > cap:
                # a long or busy recording: the kernel reported the true count, run that clip again with room for it
                # (one long file must not abort a corpus run that has already done the work for the others)
                out.append(self.detect_host(keep[i], mode, cap=int(cnt[i])))
            else:
                out.append(reg[i, :cnt[i]].copy())
        return out

    def check_health(self) -> None:
        """Synchronise and raise if a kernel flagged a pipeline time-out or an fp16 range overflow (device-level
        calls only enqueue work; the host-level detect calls check by themselves)."""
        check(lib.ss_check_health(self._ctx, self._stream()))

    def silence_host(self, audio: np.ndarray, intervals: np.ndarray) -> None:
        """In place on a host float32 buffer (any shape, C-contiguous); `intervals` int64 `[K,2]` flat offsets."""
        assert audio.dtype == np.float32 and audio.flags.c_contiguous and audio.flags.writeable
        iv = np.ascontiguousarray(intervals, dtype=np.int64).reshape(-1, 2)
        self.reserve(max(self._reserved_samples, 0), max(self._region_cap, 1))
        check(lib.ss_silence_host(self._ctx, C.c_void_p(audio.ctypes.data), audio.size,
                                  C.c_void_p(iv.ctypes.data), iv.shape[0]))

    # ------------------------------------------------------------------ PCM_16 sample path
    def decode_pcm16(self, frames: torch.Tensor) -> torch.Tensor:
        """Interleaved int16 `[n]` or `[n, C]` (device) -> float32 mono `[n]`: `sf.read(dtype='float32')` +
        `librosa.to_mono` (voice_activity.py:37,61-62)."""
        assert frames.dtype == torch.int16
        frames = frames.to(self.device).contiguous()
        n = frames.shape[0]
        ch = 1 if frames.dim() == 1 else frames.shape[1]
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        check(lib.ss_decode_pcm16(self._ctx, _ptr(frames), n, ch, _ptr(out), self._stream()))
        return out

    def encode_pcm16(self, x: torch.Tensor) -> torch.Tensor:
        """float32 (device, any shape) -> int16 of the same shape, as `sf.write(..., subtype='PCM_16')` stores it."""
        x = self._f32(x)
        out = torch.empty(x.shape, dtype=torch.int16, device=self.device)
        check(lib.ss_encode_pcm16(self._ctx, _ptr(x), x.numel(), _ptr(out), self._stream()))
        return out

    def silence_pcm16(self, pcm: torch.Tensor, intervals: torch.Tensor, requantize: bool = True) -> None:
        """In place on a device int16 buffer: optional read->write round trip of every sample, then zero the ranges."""
        assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()
        iv = intervals.to(device=self.device, dtype=torch.int64).contiguous().reshape(-1, 2)
        check(lib.ss_silence_pcm16(self._ctx, _ptr(pcm), pcm.numel(), _ptr(iv), iv.shape[0], int(requantize),
                                   self._stream()))

    def silence_pcm16_host(self, pcm: np.ndarray, intervals: np.ndarray, requantize: bool = True) -> None:
        """In place on a host int16 buffer (any shape, C-contiguous); `intervals` int64 `[K,2]` flat offsets."""
        assert pcm.dtype == np.int16 and pcm.flags.c_contiguous and pcm.flags.writeable
        iv = np.ascontiguousarray(intervals, dtype=np.int64).reshape(-1, 2)
        self.reserve(max(self._reserved_samples, 0), max(self._region_cap, 1))
        check(lib.ss_silence_pcm16_host(self._ctx, C.c_void_p(pcm.ctypes.data), pcm.size,
                                        C.c_void_p(iv.ctypes.data), iv.shape[0], int(requantize)))

    def resample(self, pcm: torch.Tensor, sr_in: int) -> torch.Tensor:
        """Mono clip on the device at `sr_in` (float32, or int16 samples of a PCM_16 file) -> float32 at 22,050 Hz,
        `ceil(n * 22050 / sr_in)` samples (K9; filter of `softspoken_b200.resample.design`).  Not soxr: see there."""
        from . import resample as rs
        assert pcm.is_cuda and pcm.dim() == 1 and pcm.is_contiguous() and pcm.dtype in (torch.float32, torch.int16)
        L, M, T, table = rs.design(int(sr_in))
        key = int(sr_in)
        if key not in self._resample_tables:
            self._resample_tables[key] = torch.from_numpy(table).to(self.device)
        n = pcm.numel()
        n_out = rs.out_len(n, sr_in)
        out = torch.empty(n_out, dtype=torch.float32, device=self.device)
        fn = lib.ss_resample_pcm16 if pcm.dtype == torch.int16 else lib.ss_resample
        check(fn(self._ctx, _ptr(pcm) if n else None, n, _ptr(out) if n_out else None, n_out, L, M, T,
                 _ptr(self._resample_tables[key]), self._stream()))
        return out

    def spectrogram(self, pcm: torch.Tensor, db: bool = False) -> torch.Tensor:
        """Mono clip on the device (float32, or int16 samples of a PCM_16 file) -> `[257, 1 + n // 256]` float32
        magnitudes of the 512 / 256 STFT (K8, `voice_activity.wav_to_spec(trim_edges=False)`); `db=True` applies the
        review screen's display transform in place (`ss_spectrogram_db`)."""
        assert pcm.is_cuda and pcm.dim() == 1 and pcm.is_contiguous() and pcm.dtype in (torch.float32, torch.int16)
        n = pcm.numel()
        T = int(lib.ss_spectrogram_frames(n))
        mag = torch.empty((257, T), dtype=torch.float32, device=self.device)
        mx = torch.zeros(1, dtype=torch.float32, device=self.device)
        fn = lib.ss_spectrogram_pcm16 if pcm.dtype == torch.int16 else lib.ss_spectrogram
        check(fn(self._ctx, _ptr(pcm) if n else None, n, _ptr(mag), _ptr(mx), self._stream()))
        if db:
            check(lib.ss_spectrogram_db(self._ctx, _ptr(mag), mag.numel(), _ptr(mx), self._stream()))
        return mag

    def _f32(self, t: torch.Tensor) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        return t.to(device=self.device, dtype=torch.float32).contiguous()


def _host_clip(audio):
    """-> (n, pointer, is_int16, keep-alive object) of a host clip given as numpy array or CPU tensor."""
    if isinstance(audio, torch.Tensor):
        assert audio.device.type == "cpu" and audio.dtype in (torch.float32, torch.int16) and audio.is_contiguous()
        return audio.numel(), C.c_void_p(audio.data_ptr()), audio.dtype == torch.int16, audio
    audio = np.asarray(audio)
    if audio.dtype != np.int16:
        audio = np.ascontiguousarray(audio, dtype=np.float32)
    audio = np.ascontiguousarray(audio)
    return audio.size, C.c_void_p(audio.ctypes.data), audio.dtype == np.int16, audio


def plan_windows(n_samples: int) -> int:
    return int(lib.ss_plan_windows(int(n_samples)))


def timeline_bins(n_padded: int) -> int:
    return int(lib.ss_timeline_bins(int(n_padded)))
