"""Head calibration of the synthetic checkpoint.  TEST INFRASTRUCTURE.

A randomly initialised `SpecUNet_2D` leaves the mask head's ReLUs dead and its
logit far from the 0.1 threshold (settings.py:13), which would make every
detection test trivial (no regions at all).  `calibrate_head` runs the oracle
network on a sample clip and returns new values for a handful of head biases /
the last 1x1 weight so that the window-averaged logit crosses 0.1 at the
speech-like bursts.  `oracle/make_golden.py` freezes the result in
`tests/golden/head_seed<k>.json`; `checkpoint.synthetic_state_dict(seed,
head=...)` applies it, so the calibrated checkpoint is reproducible on the GPU
box without running the oracle.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import model as om
from . import postproc as pp


@torch.no_grad()
def calibrate_head(sd, audio: np.ndarray, hot_fraction: float = 0.25, stride: int = 3) -> Dict[str, List[float]]:
    sd = {k: v.clone() for k, v in sd.items()}
    padded = pp.pad_audio(audio)
    starts = pp.plan_windows(len(audio) / pp.SAMPLE_RATE)[::stride]
    x = torch.stack([torch.from_numpy(padded[i:i + 66150]) for i in starts])
    taps: Dict[str, torch.Tensor] = {}
    om.forward(sd, x, want_spec=False, taps=taps)
    conv9 = taps["conv9"]
    # 1. keep ~60 % of conv_flatten's ReLU inputs alive
    sd["conv_flatten.bias"] = torch.zeros(4)
    z = F.conv2d(conv9, sd["conv_flatten.weight"], None).squeeze(2)
    sd["conv_flatten.bias"] = -(z.mean(dim=(0, 2)) - 0.25 * z.std(dim=(0, 2)))
    # 2. scale / shift the last 1x1 so that `hot_fraction` of the logits exceed 0.1
    sd["mask_output_conv.1.bias"] = torch.zeros(1)
    lg = om.mask_head(sd, conv9).reshape(-1)
    scale = 0.08 / float(lg.std())
    sd["mask_output_conv.1.weight"] = sd["mask_output_conv.1.weight"] * scale
    lg = lg * scale
    q = float(torch.quantile(lg, 1.0 - hot_fraction))
    sd["mask_output_conv.1.bias"] = torch.tensor([0.1 - q], dtype=torch.float32)
    keys = ("conv_flatten.bias", "mask_output_conv.1.weight", "mask_output_conv.1.bias")
    return {k: [float(np.float32(v)) for v in sd[k].reshape(-1).tolist()] for k in keys}
