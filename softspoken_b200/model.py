"""`SpecUNet_2D` stand-in: the reference's model-loading API over the CUDA engine.

The reference builds `SpecUNet_2D()` with no arguments, moves it `.to(device)`,
loads a 224-key state dict strictly and calls `.eval()` and
`model(audio_slices) -> (spec_output, mask_output)`
(root/code/backend/pytorch_neural_nets.py:79-197; root/code/frontend/NNDetector.py:32-34,99).
This class keeps exactly that surface; the arithmetic is the kernels'.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from . import checkpoint, spec
from ._lib import DEFAULT_MODE
from .engine import Engine


class SpecUNet_2D:
    input_shape = (66150)
    output_shape = (2, 128, 256)
    n_mels = 128

    def __init__(self, mode: str = DEFAULT_MODE, max_batch: int = 32):
        # The reference initialises randomly (torch defaults); a fixed seeded init keeps runs repeatable.
        self._sd = checkpoint.synthetic_state_dict(0)
        self._device: Optional[torch.device] = None
        self._engine: Optional[Engine] = None
        self._mode = mode
        self._max_batch = max_batch
        self.training = True

    # --- torch.nn.Module look-alikes the reference calls -------------------------------------------
    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("softspoken_b200.SpecUNet_2D runs on CUDA only (no CPU fallback)")
        self._device = dev
        self._engine = None
        return self

    def eval(self):
        self.training = False
        return self

    def state_dict(self):
        return OrderedDict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True):
        checkpoint.validate_state_dict(state_dict)
        self._sd = OrderedDict((k, state_dict[k].detach().cpu().clone()) for k, _, _ in checkpoint.state_dict_spec())
        self._engine = None

    # --- engine ------------------------------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            dev = self._device or torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
            self._engine = Engine(self._sd, dev, self._max_batch, self._mode)
        return self._engine

    def forward(self, x: torch.Tensor):
        """x `[B, 66150]` on the GPU -> (spec `[B,2,128,256]`, mask `[B,1,256]`) on the GPU."""
        eng = self.engine
        x = x.to(device=eng.device, dtype=torch.float32).contiguous()
        B, n = x.shape
        if n < spec.WINDOW_SAMPLES_USED:
            raise ValueError(f"window of {n} samples is shorter than the {spec.WINDOW_SAMPLES_USED} the front end reads")
        starts = torch.arange(B, device=eng.device, dtype=torch.int64) * n
        mel = eng.features(x.reshape(-1), starts)
        logits, spec_out = eng.classify(mel, want_spec=True)
        eng.check_health()       # synchronises: invalid logits (time-out, fp16 overflow) raise here instead of flowing on
        return spec_out, logits.unsqueeze(1)

    __call__ = forward
