"""Throughput mode (tcgen05, bf16 operands): layer-by-layer agreement with the fp32 oracle, within bf16 tolerance.

north_star allows a documented tolerance for a bf16 GEMM path; the numbers asserted here are the ones
DESIGN.md quotes.  Bit-exact detections are a property of the fp32 parity mode (test_gpu_detect.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from softspoken_b200 import spec

pytestmark = pytest.mark.gpu

ACT_TOL = 4e-2        # max |delta| / max |ref| per activation tensor (bf16 storage between 25 conv layers)
LOGIT_TOL = 6e-2      # logits: max |delta| / max |ref| (measured 3.4e-2..4.3e-2 on the seed-0 checkpoint)


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=4, mode="bf16")
    yield eng
    eng.close()


def _dump(engine, which, n):
    from softspoken_b200._lib import lib, check
    c, h, w = C.c_int(), C.c_int(), C.c_int()
    check(lib.ss_debug_activation(engine._ctx, which, n, None, C.byref(c), C.byref(h), C.byref(w), None))
    out = torch.empty((n, c.value, h.value, w.value), dtype=torch.float32, device="cuda")
    check(lib.ss_debug_activation(engine._ctx, which, n, C.c_void_p(out.data_ptr()), C.byref(c), C.byref(h),
                                  C.byref(w), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu()


def test_layerwise_against_oracle(engine, sd_seed0, clip60):
    from oracle import model as om
    from oracle import postproc as pp
    g = load_golden("model_seed0.npz")
    sel = g["starts"][[0, 41, 77]]
    padded = torch.from_numpy(pp.pad_audio(clip60)).cuda()
    mel = engine.features(padded, torch.from_numpy(sel))
    logits = engine.classify(mel, mode="bf16")
    torch.cuda.synchronize()
    taps = {}
    _, mk = om.forward_from_mel(sd_seed0, mel.cpu().unsqueeze(1), want_spec=False, taps=taps)
    up = lambda t: torch.nn.functional.interpolate(t, scale_factor=2, mode="nearest")
    ref = {11: None, 10: None, 0: taps["conv1"], 1: taps["conv2"], 2: taps["conv3"], 3: taps["conv4"],
           4: taps["bottleneck"], 5: up(taps["encoder_out"]), 6: up(taps["conv6"]), 7: up(taps["conv7"]),
           8: up(taps["conv8"]), 9: taps["conv9"]}
    x0 = _dump(engine, 11, 3)
    e0 = float((x0[:, 0] - mel.cpu()).abs().max() / mel.cpu().abs().max())
    print(f"x0 (mel as bf16): rel err {e0:.3e}; other channels max {float(x0[:, 1:].abs().max()):.1e}")
    assert e0 < 1e-2 and float(x0[:, 1:].abs().max()) == 0.0
    worst = 0.0
    for which in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]:
        got = _dump(engine, which, 3)
        want = ref[which]
        assert got.shape == want.shape, (which, got.shape, want.shape)
        err = float((got - want).abs().max() / want.abs().max())
        print(f"activation {which}: shape {tuple(got.shape)} rel err {err:.3e}")
        worst = max(worst, err)
    lerr = float((logits.cpu() - mk[:, 0]).abs().max() / mk.abs().max())
    print(f"logits rel err {lerr:.3e} (abs {float((logits.cpu() - mk[:, 0]).abs().max()):.3e})")
    assert worst <= ACT_TOL
    assert lerr <= LOGIT_TOL


def test_bf16_logits_vs_reference_golden_and_interval_agreement(engine, clip60):
    """Whole 60 s clip: logit error against the reference golden, and how many timeline bins change side."""
    from oracle import postproc as pp
    g = load_golden("model_seed0.npz")
    padded = torch.from_numpy(pp.pad_audio(clip60)).cuda()
    lg = engine.classify(engine.features(padded, torch.from_numpy(g["starts"])), mode="bf16").cpu().numpy()
    ref = g["logits"][:, 0]
    err = np.max(np.abs(lg - ref))
    a_ref, c_ref = pp.average_idx(ref.reshape(-1, 1, 256), len(padded) / 22050)
    a_got, _ = pp.average_idx(lg.reshape(-1, 1, 256), len(padded) / 22050)
    cov = c_ref > 0
    flips = int(np.sum((a_ref[cov] > 0.1) != (a_got[cov] > 0.1)))
    print(f"bf16 logits vs golden: max abs {err:.3e} (max |ref| {np.abs(ref).max():.3f}); averaged-bin decision flips "
          f"{flips}/{int(cov.sum())}")
    assert err <= LOGIT_TOL * np.abs(ref).max()
    assert flips <= 0.02 * cov.sum()


def test_bf16_batch_invariance_and_spec_head(engine, clip60):
    from oracle import postproc as pp
    g = load_golden("model_seed0.npz")
    padded = torch.from_numpy(pp.pad_audio(clip60)).cuda()
    mel = engine.features(padded, torch.from_numpy(g["starts"][38:47]))
    full, sp = engine.classify(mel, want_spec=True, mode="bf16")
    parts = torch.cat([engine.classify(mel[:1], mode="bf16"), engine.classify(mel[1:6], mode="bf16"),
                       engine.classify(mel[6:], mode="bf16")])
    assert torch.equal(full, parts)
    ref = torch.from_numpy(g["spec_w41"])
    err = float((sp[3].cpu() - ref).abs().max() / ref.abs().max())
    print(f"bf16 spec head rel err {err:.3e}")
    assert err <= ACT_TOL
