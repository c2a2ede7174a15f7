"""K1 on the GPU vs the reference's golden mel features and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from softspoken_b200 import checkpoint, spec

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4     # north_star: mel features within 1e-4 relative (of the tensor's max) in fp32


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=8, mode="fp32")
    yield eng
    eng.close()


def _padded(clip):
    from oracle import postproc as pp
    return pp.pad_audio(clip)


def test_features_match_reference_golden(engine, clip60):
    g = load_golden("frontend.npz")
    padded = torch.from_numpy(_padded(clip60)).cuda()
    mel = engine.features(padded, torch.from_numpy(g["starts"])).cpu().numpy()
    ref = g["mel"]
    err = np.max(np.abs(mel - ref)) / np.max(np.abs(ref))
    print(f"K1 vs reference golden: max|d|/max|ref| = {err:.3e}")
    assert err <= REL_TOL


def test_features_match_oracle_all_windows(engine, clip60):
    from oracle import features as of
    from oracle import postproc as pp
    padded_np = _padded(clip60)
    starts = pp.plan_windows(60.0)
    mel = engine.features(torch.from_numpy(padded_np).cuda(), torch.from_numpy(starts)).cpu().numpy()
    x = torch.stack([torch.from_numpy(padded_np[i:i + spec.WINDOW_SAMPLES]) for i in starts])
    ref = of.log_mel(x, checkpoint.hann_window(), checkpoint.mel_filterbank()).numpy()
    err = np.max(np.abs(mel - ref)) / np.max(np.abs(ref))
    print(f"K1 vs oracle, 105 windows: {err:.3e}")
    assert err <= REL_TOL
    # error against the float64 definition is of the same order as the reference's own float32 error
    m64 = of.log_mel_f64(padded_np[starts[40]:starts[40] + spec.WINDOW_SAMPLES], checkpoint.hann_window().numpy(),
                         checkpoint.mel_filterbank().numpy())
    e_gpu = np.max(np.abs(mel[40] - m64))
    e_ref = np.max(np.abs(ref[40] - m64))
    print(f"vs float64 definition: gpu {e_gpu:.3e}, reference fp32 {e_ref:.3e}")
    assert e_gpu <= 1e-4


def test_features_edge_inputs(engine):
    """Silence, full-scale square wave, impulse at the reflected boundary, window at the very end."""
    from oracle import features as of
    n = spec.WINDOW_SAMPLES * 2
    rng = np.random.default_rng(3)
    x = np.zeros(n, np.float32)
    x[spec.WINDOW_SAMPLES:] = np.sign(rng.normal(size=spec.WINDOW_SAMPLES)).astype(np.float32)
    x[1] = 1.0                                   # inside frame 0's reflected half
    starts = np.array([0, 13230, n - spec.WINDOW_SAMPLES], dtype=np.int64)
    mel = engine.features(torch.from_numpy(x).cuda(), torch.from_numpy(starts)).cpu().numpy()
    xw = torch.stack([torch.from_numpy(x[i:i + spec.WINDOW_SAMPLES]) for i in starts])
    ref = of.log_mel(xw, checkpoint.hann_window(), checkpoint.mel_filterbank()).numpy()
    assert np.isfinite(mel).all()
    assert np.max(np.abs(mel - ref)) <= REL_TOL * np.max(np.abs(ref))
    # all-zero input -> exactly zero features: sqrt(log10(0 + 1))
    z = engine.features(torch.zeros(spec.WINDOW_SAMPLES, device="cuda"), torch.zeros(1, dtype=torch.int64))
    assert float(z.abs().max()) == 0.0


def test_virtual_padding_equals_materialised_padding(engine, clip60):
    """ss_detect_* never builds the padded buffer (worker.py:58-62); ss_pad + ss_features must agree."""
    from oracle import postproc as pp
    clip = clip60[: 22050 * 7]
    padded = engine.pad(torch.from_numpy(clip).cuda())
    assert np.array_equal(padded.cpu().numpy(), pp.pad_audio(clip))
    starts = torch.from_numpy(pp.plan_windows(7.0))
    a = engine.features(padded, starts)
    _, _, lg = engine.detect_device(torch.from_numpy(clip).cuda(), want_logits=True)
    b = engine.classify(a)
    assert torch.equal(lg, b)


def test_two_band_walk_has_the_bits_of_the_band_by_band_walk(sd_seed0, clip60, monkeypatch):
    """K1's mel phase walks the bins of a warp's bands once, feeding the two bands a bin belongs to (default for
    triangular banks, FrontEnd::mel_rec); SS_MEL_WALK=0 (read at context creation) keeps the band-by-band walk of the
    sparse taps.  Same ascending order of the bins within every band, so the same bits."""
    from oracle import postproc as pp
    from softspoken_b200.engine import Engine
    padded = torch.from_numpy(_padded(clip60)).cuda()
    starts = torch.from_numpy(pp.plan_windows(60.0))
    eng = Engine(sd_seed0, 0, max_batch=8, mode="fp32")
    walk = eng.features(padded, starts)
    eng.close()
    monkeypatch.setenv("SS_MEL_WALK", "0")
    eng = Engine(sd_seed0, 0, max_batch=8, mode="fp32")
    monkeypatch.delenv("SS_MEL_WALK")
    plain = eng.features(padded, starts)
    eng.close()
    assert torch.equal(walk, plain)


def test_non_triangular_filterbank_takes_the_general_path(sd_seed0, clip60):
    """A bank whose bands overlap three deep (every band widened by half of its upper neighbour) cannot be walked two
    bands at a time: the context falls back to the band-by-band walk, and the features still match the oracle."""
    from collections import OrderedDict
    from oracle import features as of
    from oracle import postproc as pp
    from softspoken_b200.engine import Engine
    fb = checkpoint.mel_filterbank().clone()
    wide = fb.clone()
    wide[:, :90] += 0.5 * fb[:, 1:91]              # bands 0..89 now reach into band m + 2's support; all stay <= 32 taps
    sd = OrderedDict((k, v.clone()) for k, v in sd_seed0.items())
    sd["mel_spectrogram.mel_scale.fb"] = wide
    padded_np = _padded(clip60)
    starts = pp.plan_windows(60.0)[:16]
    eng = Engine(sd, 0, max_batch=8, mode="fp32")
    mel = eng.features(torch.from_numpy(padded_np).cuda(), torch.from_numpy(starts)).cpu().numpy()
    eng.close()
    x = torch.stack([torch.from_numpy(padded_np[i:i + spec.WINDOW_SAMPLES]) for i in starts])
    ref = of.log_mel(x, checkpoint.hann_window(), wide).numpy()
    err = np.max(np.abs(mel - ref)) / np.max(np.abs(ref))
    print(f"K1 with a three-deep overlapping bank vs oracle: {err:.3e}")
    assert err <= REL_TOL


def test_packed_phase1_against_the_scalar_path(sd_seed0, clip60, monkeypatch):
    """K1's phase 1 runs on packed pairs (FADD2 / FMUL2 / FFMA2: a lane's two transform columns as the halves of a
    64-bit register pair) by default; SS_K1_PACKED=0 (read at context creation) keeps the scalar path.  Each packed
    half is rounded as the scalar instruction would be, but the contraction of products into sums is the compiler's in
    the scalar path and spelled out by hand in the packed one: measured 3.8e-7 of the largest feature apart, not
    bit-identical.  ASSERTED: a bound two orders of magnitude below the parity budget, over whole windows incl. the
    reflected first frame, the virtual zero padding (slow sample path) and int16 samples."""
    from oracle import postproc as pp
    from softspoken_b200.engine import Engine
    padded = torch.from_numpy(_padded(clip60)).cuda()
    starts = torch.from_numpy(pp.plan_windows(60.0))
    clip = torch.from_numpy(clip60[: 22050 * 9]).cuda()
    pcm16 = torch.from_numpy((clip60[: 22050 * 9] * 32767.0).astype(np.int16)).cuda()

    def run():
        eng = Engine(sd_seed0, 0, max_batch=8, mode="fp32")
        mel = eng.features(padded, starts)
        _, _, lg = eng.detect_device(clip, want_logits=True)          # virtual padding: frames that straddle the clip's ends
        _, _, lg16 = eng.detect_device(pcm16, want_logits=True)
        eng.close()
        return mel, lg, lg16

    packed = run()
    monkeypatch.setenv("SS_K1_PACKED", "0")
    scalar = run()
    monkeypatch.delenv("SS_K1_PACKED")
    d = float((packed[0] - scalar[0]).abs().max() / scalar[0].abs().max())
    same = [bool(torch.equal(a, b)) for a, b in zip(packed, scalar)]
    print(f"K1 packed vs scalar: mel max|d|/max = {d:.3e}; bit-identical (mel, logits f32, logits s16): {same}")
    assert d <= 1e-6
    for a, b in zip(packed[1:], scalar[1:]):
        assert float((a - b).abs().max()) <= 1e-5 * max(1.0, float(b.abs().max()))
