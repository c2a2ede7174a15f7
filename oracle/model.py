"""Oracle: float32 CPU restatement of `SpecUNet_2D.forward`.  TEST INFRASTRUCTURE.

Follows the reference op by op (root/code/backend/pytorch_neural_nets.py:142-197)
on the raw state dict — un-folded BatchNorm in eval mode, the same torch CPU
kernels (oneDNN conv, native batch-norm) the reference dispatches to — so that
it is the reference's arithmetic, not an approximation of it.  The mel front
end is restated separately in `oracle/features.py` (no torchaudio needed).

Pinned by tests/test_oracle_model.py against tests/golden/model_*.npz, which
`oracle/make_golden.py` froze from the real reference module.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import features

BN_EPS = 1e-5


def _bn(sd, prefix, x):
    # nn.BatchNorm{1,2}d in eval(): running statistics, eps=1e-5
    return F.batch_norm(x, sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"],
                        sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], False, 0.0, BN_EPS)


def res_block(sd, prefix, x, dims=2):
    """ResBlock / ResBlock1D.forward (pytorch_neural_nets.py:32-41, 68-77)."""
    conv = F.conv2d if dims == 2 else F.conv1d
    identity = _bn(sd, f"{prefix}.residual.1", conv(x, sd[f"{prefix}.residual.0.weight"]))
    out = F.relu(_bn(sd, f"{prefix}.conv1.1", conv(x, sd[f"{prefix}.conv1.0.weight"], padding=1)))
    out = _bn(sd, f"{prefix}.conv2.1", conv(out, sd[f"{prefix}.conv2.0.weight"], padding=1))
    return F.relu(out + identity)          # Dropout is the identity in eval()


def up(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")   # nn.Upsample(scale_factor=2, 'nearest')


def trunk(sd, mel: torch.Tensor, taps: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """mel `[B,1,128,256]` -> conv9 `[B,32,128,256]` (pytorch_neural_nets.py:156-181)."""
    conv1 = res_block(sd, "conv1_1", mel)
    conv2 = res_block(sd, "conv2_1", F.max_pool2d(conv1, 2))
    conv3 = res_block(sd, "conv3_1", F.max_pool2d(conv2, 2))
    conv4 = res_block(sd, "conv4_1", F.max_pool2d(conv3, 2))
    bott = res_block(sd, "conv_bottleneck", F.max_pool2d(conv4, 2))
    enc = res_block(sd, "encoder_out", bott)
    conv6 = res_block(sd, "conv6", torch.cat([conv4, up(enc)], dim=1))
    conv7 = res_block(sd, "conv7", torch.cat([conv3, up(conv6)], dim=1))
    conv8 = res_block(sd, "conv8", torch.cat([conv2, up(conv7)], dim=1))
    conv9 = res_block(sd, "conv9_1", torch.cat([conv1, up(conv8)], dim=1))
    if taps is not None:
        taps.update(conv1=conv1, conv2=conv2, conv3=conv3, conv4=conv4, bottleneck=bott,
                    encoder_out=enc, conv6=conv6, conv7=conv7, conv8=conv8, conv9=conv9)
    return conv9


def spec_head(sd, conv9):
    """pytorch_neural_nets.py:126-130,184-185."""
    y = res_block(sd, "spec_output_conv.0", conv9)
    y = F.conv2d(y, sd["spec_output_conv.1.weight"], sd["spec_output_conv.1.bias"])
    return F.relu(y)


def flatten_head(sd, conv9):
    """conv_flatten + ReLU + squeeze (pytorch_neural_nets.py:133-134,188-192) -> [B,4,256]."""
    return F.relu(F.conv2d(conv9, sd["conv_flatten.weight"], sd["conv_flatten.bias"])).squeeze(2)


def mask_head(sd, conv9):
    """pytorch_neural_nets.py:133-140,188-195 -> raw logits `[B,1,256]` (no sigmoid)."""
    y = res_block(sd, "mask_output_conv.0", flatten_head(sd, conv9), dims=1)
    return F.conv1d(y, sd["mask_output_conv.1.weight"], sd["mask_output_conv.1.bias"])


@torch.no_grad()
def forward_from_mel(sd, mel: torch.Tensor, want_spec: bool = True, taps=None):
    conv9 = trunk(sd, mel, taps)
    spec_out = spec_head(sd, conv9) if want_spec else None
    return spec_out, mask_head(sd, conv9)


@torch.no_grad()
def forward(sd, x: torch.Tensor, want_spec: bool = True, taps=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """`SpecUNet_2D.forward`: x `[B,66150]` -> (spec `[B,2,128,256]`, mask `[B,1,256]`)."""
    mel = features.log_mel(x, sd["mel_spectrogram.spectrogram.window"],
                           sd["mel_spectrogram.mel_scale.fb"]).unsqueeze(1)
    if taps is not None:
        taps["mel"] = mel
    return forward_from_mel(sd, mel, want_spec, taps)
