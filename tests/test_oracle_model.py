"""Oracle network vs golden logits frozen from the reference (tests/golden/model_seed0.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import model as om
from oracle import postproc as pp
from softspoken_b200 import checkpoint, spec


def test_state_dict_layout_matches_reference():
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        layout = json.load(f)
    mine = [[k, list(s), str(d)] for k, s, d in checkpoint.state_dict_spec()]
    assert mine == layout
    assert len(mine) == 224


def test_synthetic_state_dict_is_deterministic(head_seed0):
    a = checkpoint.synthetic_state_dict(0, head_seed0)
    b = checkpoint.synthetic_state_dict(0, head_seed0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = checkpoint.synthetic_state_dict(1)
    assert not torch.equal(a["conv2_1.conv1.0.weight"], c["conv2_1.conv1.0.weight"])
    n_params = sum(v.numel() for k, v in a.items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))
                   and not k.startswith("mel_spectrogram"))
    assert n_params == 1_713_555


def _windows(clip, starts):
    padded = pp.pad_audio(clip)
    return torch.stack([torch.from_numpy(padded[i:i + spec.WINDOW_SAMPLES]) for i in starts])


def test_logits_match_reference_golden(sd_seed0, clip60):
    g = load_golden("model_seed0.npz")
    torch.set_num_threads(int(g["threads"]))
    starts = g["starts"]
    assert np.array_equal(starts, pp.plan_windows(60.0))
    sel = np.r_[0:8, 40:44, 97:105]                 # leading pad, middle, ragged tail
    _, mk = om.forward(sd_seed0, _windows(clip60, starts[sel]), want_spec=False)
    ref = g["logits"][sel]
    assert mk.shape == (len(sel), 1, 256)
    assert np.max(np.abs(mk.numpy() - ref)) <= 1e-6   # oracle self-consistency is ~3e-8


def test_spec_head_and_trunk_match_reference_golden(sd_seed0, clip60):
    g = load_golden("model_seed0.npz")
    taps = {}
    sp, _ = om.forward(sd_seed0, _windows(clip60, g["starts"][41:42]), want_spec=True, taps=taps)
    assert np.max(np.abs(sp[0].numpy() - g["spec_w41"])) <= 1e-5 * max(1.0, np.abs(g["spec_w41"]).max())
    c9 = taps["conv9"][0].numpy()[:, ::16, :]
    assert np.max(np.abs(c9 - g["conv9_w41_rows"])) <= 1e-5 * np.abs(g["conv9_w41_rows"]).max()
    names = {"conv1_1": "conv1", "conv2_1": "conv2", "conv3_1": "conv3", "conv4_1": "conv4",
             "conv_bottleneck": "bottleneck", "encoder_out": "encoder_out", "conv6": "conv6",
             "conv7": "conv7", "conv8": "conv8", "conv9_1": "conv9"}
    for rname, oname in names.items():
        v = taps[oname]
        stats = np.array([float(v.mean()), float(v.abs().max()), float(v.std())])
        assert np.allclose(stats, g[f"act_{rname}"], rtol=1e-5), rname


def test_folded_weights_reproduce_unfolded_network(sd_seed0, clip60):
    """BN folding (checkpoint.fold_state_dict) is what the kernels consume: conv(w', b') == BN(conv(w))."""
    import torch.nn.functional as F
    folded = checkpoint.fold_state_dict(sd_seed0)
    x = torch.randn(2, 32, 16, 24, generator=torch.Generator().manual_seed(0))
    want = om.res_block(sd_seed0, "conv2_1", x)
    idn = F.conv2d(x, folded["conv2_1.res.w"], folded["conv2_1.res.b"])
    y = F.relu(F.conv2d(x, folded["conv2_1.c1.w"], folded["conv2_1.c1.b"], padding=1))
    y = F.conv2d(y, folded["conv2_1.c2.w"], folded["conv2_1.c2.b"], padding=1)
    got = F.relu(y + idn)
    assert torch.max(torch.abs(got - want)) <= 2e-5 * want.abs().max()


@pytest.mark.needs_reference
def test_oracle_is_bit_exact_against_live_reference(sd_seed0, clip60):
    from oracle import ref_shim
    ref = ref_shim.load()
    det = ref_shim.make_detector(ref, sd_seed0, threads=4)
    x = _windows(clip60, pp.plan_windows(60.0)[50:53])
    with torch.no_grad():
        sp_r, mk_r = det.model(x)
    sp_o, mk_o = om.forward(sd_seed0, x)
    assert torch.equal(sp_r, sp_o) and torch.equal(mk_r, mk_o)


def test_oracle_reproduces_scale_goldens(sd_seed0):
    """The config-scale goldens (real reference, oracle/make_golden_scale.py): the oracle network on the first
    reference batch of clip 0 and the oracle post-processing on all of its logits give the reference's results."""
    import numpy as np
    import torch
    from oracle import model as om, postproc as pp
    from softspoken_b200 import synth
    g = np.load(os.path.join(GOLDEN, "scale_clip_seed0.npz"))
    audio = synth.synth_audio(600.0, 0)
    padded = pp.pad_audio(audio)
    starts = pp.plan_windows(600.0)
    assert len(starts) == g["logits"].shape[0] == 1005 and len(padded) == int(g["n_padded"])
    x = torch.stack([torch.from_numpy(padded[i:i + 66150]) for i in starts[:32]])
    _, mk = om.forward(sd_seed0, x, want_spec=False)
    assert np.abs(mk[:, 0].numpy() - g["logits"][:32]).max() <= 1e-6      # oneDNN thread-count noise is ~3e-8
    avg, cnt = pp.average_idx(g["logits"].reshape(-1, 1, 256), len(padded) / 22050)
    n = int(g["n_emitted"])
    assert int((cnt >= 1).sum()) == n
    assert np.array_equal(avg[:n] > 0.1, np.unpackbits(g["hot_bits"])[:n].astype(bool))
    assert np.array_equal(pp.find_speech_regions_idx(avg, cnt), g["region_bins"])
    assert abs(float(np.abs(avg[:n] - 0.1).min()) - float(g["min_margin"])) < 1e-18
