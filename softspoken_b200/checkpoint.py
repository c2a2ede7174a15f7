"""Checkpoint format of the reference detector and its kernel-side packing.

The reference keeps its weights in a `torch.save`d dict
`{'model_state_dict': <224 tensors>, 'epoch': int}` and loads it with
`torch.load(path, map_location=device, weights_only=True)` followed by a strict
`load_state_dict` (reference `root/code/frontend/NNDetector.py:42-53`).  This
module owns

* the key/shape table of that state dict (`state_dict_spec`), restated from the
  layer list in `root/code/backend/pytorch_neural_nets.py:92-140`;
* the front-end buffers (`hann_window`, `mel_filterbank`) that torchaudio's
  `MelSpectrogram` registers (`pytorch_neural_nets.py:92-99`);
* the one-time "fold BatchNorm, repack to kernel layout" step that produces the
  flat blob `ss_ctx_create` consumes (`include/softspoken_b200.h`).

Nothing here touches CUDA; everything is plain torch-CPU / numpy.
"""
from __future__ import annotations

import os
import re
import struct
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import spec

_BN_GAMMA = re.compile(r"\.(residual|conv1|conv2)\.1\.weight$")
BN_EPS = 1e-5  # torch.nn.BatchNorm{1,2}d default, never overridden by the reference

# (state-dict prefix, C_in, C_out) of the eleven 2-D residual blocks, in
# registration order (pytorch_neural_nets.py:102-123,126-127).
RESBLOCKS_2D: List[Tuple[str, int, int]] = [
    ("conv1_1", 1, 32),
    ("conv2_1", 32, 64),
    ("conv3_1", 64, 96),
    ("conv4_1", 96, 128),
    ("conv_bottleneck", 128, 128),
    ("encoder_out", 128, 128),
    ("conv6", 256, 96),
    ("conv7", 192, 64),
    ("conv8", 128, 32),
    ("conv9_1", 64, 32),
    ("spec_output_conv.0", 32, 32),
]


def _bn_keys(prefix: str, c: int):
    return [
        (f"{prefix}.weight", (c,), torch.float32),
        (f"{prefix}.bias", (c,), torch.float32),
        (f"{prefix}.running_mean", (c,), torch.float32),
        (f"{prefix}.running_var", (c,), torch.float32),
        (f"{prefix}.num_batches_tracked", (), torch.int64),
    ]


def _resblock_keys(prefix: str, cin: int, cout: int, dims: int):
    k1 = (1,) * dims
    k3 = (3,) * dims
    out = [(f"{prefix}.residual.0.weight", (cout, cin) + k1, torch.float32)]
    out += _bn_keys(f"{prefix}.residual.1", cout)
    out += [(f"{prefix}.conv1.0.weight", (cout, cin) + k3, torch.float32)]
    out += _bn_keys(f"{prefix}.conv1.1", cout)
    out += [(f"{prefix}.conv2.0.weight", (cout, cout) + k3, torch.float32)]
    out += _bn_keys(f"{prefix}.conv2.1", cout)
    return out


def state_dict_spec() -> List[Tuple[str, tuple, torch.dtype]]:
    """Ordered (key, shape, dtype) of `SpecUNet_2D().state_dict()` — 224 entries."""
    keys = [
        ("mel_spectrogram.spectrogram.window", (spec.WIN_LENGTH,), torch.float32),
        ("mel_spectrogram.mel_scale.fb", (spec.N_FREQS, spec.N_MELS), torch.float32),
    ]
    for prefix, cin, cout in RESBLOCKS_2D[:-1]:
        keys += _resblock_keys(prefix, cin, cout, 2)
    keys += _resblock_keys("spec_output_conv.0", 32, 32, 2)
    keys += [
        ("spec_output_conv.1.weight", (2, 32, 1, 1), torch.float32),
        ("spec_output_conv.1.bias", (2,), torch.float32),
        ("conv_flatten.weight", (4, 32, spec.N_MELS, 1), torch.float32),
        ("conv_flatten.bias", (4,), torch.float32),
    ]
    keys += _resblock_keys("mask_output_conv.0", 4, 4, 1)
    keys += [
        ("mask_output_conv.1.weight", (1, 4, 1), torch.float32),
        ("mask_output_conv.1.bias", (1,), torch.float32),
    ]
    return keys


# ----------------------------------------------------------------------------
# front-end buffers
# ----------------------------------------------------------------------------

def hann_window() -> torch.Tensor:
    """Periodic Hann window, as `torchaudio.transforms.Spectrogram` registers it
    (`torch.hann_window(win_length)`; pytorch_neural_nets.py:92-95)."""
    return torch.hann_window(spec.WIN_LENGTH, periodic=True, dtype=torch.float32)


def mel_filterbank() -> torch.Tensor:
    """HTK triangular filterbank `[1025, 128]`, norm=None.

    Restates torchaudio 2.11 `functional.melscale_fbanks(n_freqs=1025, f_min=0,
    f_max=8000, n_mels=128, sample_rate=22050, norm=None, mel_scale='htk')`
    (reached from pytorch_neural_nets.py:92-99 via `transforms.MelScale`),
    operation by operation in float32 so the result is bit-identical
    (pinned in tests/test_oracle_features.py against tests/golden/frontend.npz).
    """
    all_freqs = torch.linspace(0, spec.SAMPLE_RATE // 2, spec.N_FREQS)
    m_min = 2595.0 * np.log10(1.0 + (0.0 / 700.0))
    m_max = 2595.0 * np.log10(1.0 + (spec.F_MAX / 700.0))
    m_pts = torch.linspace(m_min, m_max, spec.N_MELS + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(zero, torch.min(down, up)).contiguous()


# ----------------------------------------------------------------------------
# synthetic, seeded checkpoint (the shipped one is a missing blob:
# reference `.MISSING_LARGE_BLOBS:1`)
# ----------------------------------------------------------------------------

def synthetic_state_dict(seed: int = 0, head: Dict[str, list] | None = None) -> "OrderedDict[str, torch.Tensor]":
    """Seeded stand-in for the missing shipped weights, in the reference format.

    Conv weights are He-normal, BatchNorm affine/running statistics are
    non-trivial so that folding is exercised.  `head` optionally overrides a few
    head tensors with frozen calibration values (tests/golden/head_seed*.json) that
    move the logit distribution across the 0.1 threshold on the synthetic clips.  numpy's PCG64 stream is used so that the
    same seed gives the same bytes on any host.
    """
    rng = np.random.default_rng(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape, dtype in state_dict_spec():
        if key == "mel_spectrogram.spectrogram.window":
            t = hann_window()
        elif key == "mel_spectrogram.mel_scale.fb":
            t = mel_filterbank()
        elif key.endswith("num_batches_tracked"):
            t = torch.tensor(1000, dtype=torch.int64)
        elif key.endswith("running_mean"):
            t = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif key.endswith("running_var"):
            t = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif _BN_GAMMA.search(key):
            # the two branches that are summed get ~1/sqrt(2) gain so activations stay O(1)
            lo, hi = (0.7, 1.3) if ".conv1." in key else (0.45, 0.95)
            t = torch.from_numpy(rng.uniform(lo, hi, shape).astype(np.float32))     # BN gamma
        elif len(shape) == 1:
            t = torch.from_numpy(rng.normal(0.0, 0.05, shape).astype(np.float32))   # BN beta / conv bias
        else:
            fan_in = int(np.prod(shape[1:]))
            std = np.sqrt(2.0 / fan_in)
            t = torch.from_numpy(rng.normal(0.0, std, shape).astype(np.float32))
        assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, key
        sd[key] = t
    for key, values in (head or {}).items():       # frozen head calibration (tests/golden/head_seed*.json)
        sd[key] = torch.tensor(values, dtype=torch.float32).reshape(sd[key].shape)
    return sd


def save_checkpoint(sd: Dict[str, torch.Tensor], path: str, epoch: int = 0) -> None:
    """Write `{'model_state_dict', 'epoch'}` exactly as NNDetector.load_checkpoint expects
    (NNDetector.py:47-49)."""
    torch.save({"model_state_dict": OrderedDict(sd), "epoch": int(epoch)}, path)


def read_checkpoint(path: str, map_location="cpu") -> Tuple[Dict[str, torch.Tensor], int]:
    ck = torch.load(path, map_location=map_location, weights_only=True)
    return ck["model_state_dict"], int(ck["epoch"])


def validate_state_dict(sd: Dict[str, torch.Tensor]) -> None:
    """Strict key/shape check — the analogue of `load_state_dict(strict=True)`."""
    want = state_dict_spec()
    missing = [k for k, _, _ in want if k not in sd]
    extra = [k for k in sd if k not in {k for k, _, _ in want}]
    if missing or extra:
        raise RuntimeError(
            "Error(s) in loading state_dict for SpecUNet_2D: "
            f"Missing key(s): {missing}. Unexpected key(s): {extra}.")
    for k, shape, _ in want:
        if tuple(sd[k].shape) != tuple(shape):
            raise RuntimeError(
                f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)} vs model {tuple(shape)}")


# ----------------------------------------------------------------------------
# BatchNorm folding and blob packing
# ----------------------------------------------------------------------------

def fold_bn(sd, conv_w_key: str, bn_prefix: str):
    """Fold eval-mode BatchNorm into the preceding bias-free conv.

    y = gamma (conv(x) - mean) / sqrt(var + eps) + beta
      = conv_{w * s}(x) + (beta - mean * s),   s = gamma / sqrt(var + eps)
    (ResBlock: pytorch_neural_nets.py:12-27).  Done in float64, rounded once.
    """
    w = sd[conv_w_key].detach().cpu().double()
    g = sd[f"{bn_prefix}.weight"].detach().cpu().double()
    b = sd[f"{bn_prefix}.bias"].detach().cpu().double()
    m = sd[f"{bn_prefix}.running_mean"].detach().cpu().double()
    v = sd[f"{bn_prefix}.running_var"].detach().cpu().double()
    s = g / torch.sqrt(v + BN_EPS)
    wf = w * s.reshape(-1, *([1] * (w.dim() - 1)))
    bf = b - m * s
    return wf.float().contiguous(), bf.float().contiguous()


def fold_state_dict(sd) -> Dict[str, torch.Tensor]:
    """All conv weights with BN folded: `<prefix>.{res,c1,c2}.{w,b}` plus heads."""
    out: Dict[str, torch.Tensor] = {}
    for prefix, _, _ in RESBLOCKS_2D + [("mask_output_conv.0", 4, 4)]:
        for short, sub in (("res", "residual"), ("c1", "conv1"), ("c2", "conv2")):
            w, b = fold_bn(sd, f"{prefix}.{sub}.0.weight", f"{prefix}.{sub}.1")
            out[f"{prefix}.{short}.w"] = w
            out[f"{prefix}.{short}.b"] = b
    for k in ("spec_output_conv.1", "conv_flatten", "mask_output_conv.1"):
        out[f"{k}.w"] = sd[f"{k}.weight"].detach().cpu().float().contiguous()
        out[f"{k}.b"] = sd[f"{k}.bias"].detach().cpu().float().contiguous()
    return out


def sparse_filterbank(fb: torch.Tensor):
    """Per-band (first_bin, n_taps) + packed taps of the dense `[1025,128]` filterbank.

    The kernel walks each band's contiguous support only (2..31 taps for the
    reference's HTK bank; 1,469 non-zeros).  A bank whose bands are not
    contiguous, or wider than MAX_TAPS, is rejected at load time.
    """
    fbn = fb.detach().cpu().float().numpy()
    if fbn.shape != (spec.N_FREQS, spec.N_MELS):
        raise ValueError(f"mel filterbank must be {(spec.N_FREQS, spec.N_MELS)}, got {fbn.shape}")
    start = np.zeros(spec.N_MELS, np.int32)
    count = np.zeros(spec.N_MELS, np.int32)
    taps: List[np.ndarray] = []
    offs = np.zeros(spec.N_MELS, np.int32)
    pos = 0
    for m in range(spec.N_MELS):
        nz = np.nonzero(fbn[:, m])[0]
        if nz.size == 0:
            start[m], count[m], offs[m] = 0, 0, pos
            continue
        lo, hi = int(nz[0]), int(nz[-1])
        start[m], count[m], offs[m] = lo, hi - lo + 1, pos
        taps.append(fbn[lo:hi + 1, m].copy())
        pos += hi - lo + 1
    w = np.concatenate(taps) if taps else np.zeros(0, np.float32)
    return start, count, offs, w.astype(np.float32)


BLOB_MAGIC = 0x53534232  # 'SSB2'
BLOB_VERSION = 2
MAX_MEL_TAPS = 32


def _conv2d_to_kernel_layout(w: torch.Tensor) -> np.ndarray:
    """[C_out, C_in, kh, kw] -> [kh*kw, C_in, C_out] (tap-major, C_out fastest)."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw, ci, co).contiguous().numpy()


def pack_blob(sd) -> bytes:
    """Flat little-endian blob consumed by `ss_ctx_create`.

    Layout: header {magic u32, version u32, n_entries u32, reserved u32},
    n_entries x {name char[48], offset u64, count u64} (float32 element
    offsets into the payload), then the float32 payload.  int32 tables are
    stored bit-cast inside the float32 payload.
    """
    validate_state_dict(sd)
    folded = fold_state_dict(sd)
    entries: List[Tuple[str, np.ndarray]] = []

    entries.append(("window", sd["mel_spectrogram.spectrogram.window"].detach().cpu().float().numpy()))
    start, count, offs, taps = sparse_filterbank(sd["mel_spectrogram.mel_scale.fb"])
    if int(count.max()) > MAX_MEL_TAPS:
        raise ValueError(f"mel band wider than {MAX_MEL_TAPS} taps")
    if int(start.min()) < 0 or int((start + count).max()) > spec.N_FREQS:
        raise ValueError("mel band outside the one-sided spectrum")
    entries.append(("mel_start", start.view(np.float32)))
    entries.append(("mel_count", count.view(np.float32)))
    entries.append(("mel_offs", offs.view(np.float32)))
    entries.append(("mel_taps", taps))

    for prefix, _, _ in RESBLOCKS_2D:
        for short in ("res", "c1", "c2"):
            entries.append((f"{prefix}.{short}.w", _conv2d_to_kernel_layout(folded[f"{prefix}.{short}.w"])))
            entries.append((f"{prefix}.{short}.b", folded[f"{prefix}.{short}.b"].numpy()))
    entries.append(("spec_output_conv.1.w", _conv2d_to_kernel_layout(folded["spec_output_conv.1.w"])))
    entries.append(("spec_output_conv.1.b", folded["spec_output_conv.1.b"].numpy()))
    # conv_flatten [4, 32, 128, 1] -> [mel(128), C_in(32), C_out(4)]
    wf = folded["conv_flatten.w"][:, :, :, 0].permute(2, 1, 0).contiguous().numpy()
    entries.append(("conv_flatten.w", wf))
    entries.append(("conv_flatten.b", folded["conv_flatten.b"].numpy()))
    # 1-D head: [C_out, C_in, k] -> [k, C_in, C_out]
    for short in ("res", "c1", "c2"):
        w = folded[f"mask_output_conv.0.{short}.w"].permute(2, 1, 0).contiguous().numpy()
        entries.append((f"mask_output_conv.0.{short}.w", w))
        entries.append((f"mask_output_conv.0.{short}.b", folded[f"mask_output_conv.0.{short}.b"].numpy()))
    entries.append(("mask_output_conv.1.w", folded["mask_output_conv.1.w"].reshape(-1).numpy()))
    entries.append(("mask_output_conv.1.b", folded["mask_output_conv.1.b"].numpy()))

    head = struct.pack("<IIII", BLOB_MAGIC, BLOB_VERSION, len(entries), 0)
    table = b""
    payload = []
    off = 0
    for name, arr in entries:
        a = np.ascontiguousarray(arr).reshape(-1)
        assert a.dtype == np.float32, (name, a.dtype)
        nm = name.encode("ascii")
        assert len(nm) < 48
        table += struct.pack("<48sQQ", nm, off, a.size)
        payload.append(a.tobytes())
        off += a.size
        pad = (-off) % 4            # keep every entry 16-byte aligned
        if pad:
            payload.append(b"\0" * (4 * pad))
            off += pad
    return head + table + b"".join(payload)


def normalise_model_path(path: str) -> str:
    """`settings.model_dir` is a Windows-style relative path (`settings.py:19`);
    on POSIX `os.path.join` would keep the back-slashes.  Accept either."""
    if os.sep == "/" and "\\" in path and not os.path.exists(path):
        return path.replace("\\", "/")
    return path
