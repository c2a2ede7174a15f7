"""Margin-guided refinement (ss_ctx_set_refine) and the guard-band canaries (ss_debug_check_guards)."""
import numpy as np
import pytest
import torch

from softspoken_b200 import spec, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=48, mode="f16x3")
    yield eng
    eng.close()


def _flagged_windows(logits, n_samples, eps):
    """The windows K5 must mark: those covering a bin whose float64 average lies within eps of the threshold."""
    from oracle import postproc as pp
    avg, cnt = pp.average_idx(logits.reshape(-1, 1, 256), (n_samples + 2 * spec.PAD_SAMPLES) / 22050)
    pos = pp.window_positions(len(logits))
    near = np.flatnonzero((cnt >= 1) & (np.abs(avg - spec.THRESHOLD) < eps))
    marked = set()
    for j in near:
        for i in np.flatnonzero((pos <= j) & (j < pos + 256)):
            marked.add(int(i))
    return sorted(marked)


@pytest.mark.parametrize("pcm16", [False, True])
def test_refined_windows_carry_fp32_logits(engine, pcm16):
    from oracle import postproc as pp
    eps = 2e-3                                            # wide band: a few dozen windows of the 60 s clip
    audio = synth.synth_pcm16(60.0, 0) if pcm16 else synth.synth_audio(60.0, 0)
    n = len(audio)
    engine.set_refine(0.0)
    _, first = engine.detect_host(audio, want_logits=True)
    _, fp32 = engine.detect_host(audio, want_logits=True, mode="fp32")
    engine.set_refine(eps, "fp32")
    engine.refine_stats(reset=True)
    bins, got = engine.detect_host(audio, want_logits=True)
    want_marked = _flagged_windows(first, n, eps)
    st = engine.refine_stats()
    assert 0 < len(want_marked) < len(first)
    assert st == {"windows": len(first), "windows_refined": len(want_marked), "clips": 1, "clips_refined": 1}
    marked = np.zeros(len(first), bool)
    marked[want_marked] = True
    assert np.array_equal(got[marked], fp32[marked])          # refined rows: the fp32 classifier's bits
    assert np.array_equal(got[~marked], first[~marked])       # the rest: untouched first-pass logits
    avg, cnt = pp.average_idx(got.reshape(-1, 1, 256), (n + 2 * spec.PAD_SAMPLES) / 22050)
    assert np.array_equal(bins.astype(np.int64), pp.find_speech_regions_idx(avg, cnt))   # K5/K6 ran on the patched logits
    # same through the device-resident and the batched host entry points
    reg, cnt_dev, lg_dev = engine.detect_device(torch.from_numpy(audio).cuda(), want_logits=True)
    assert np.array_equal(lg_dev.cpu().numpy(), got)
    assert np.array_equal(reg[:int(cnt_dev.item())].cpu().numpy(), bins)
    short = audio[: 22050 * 7]
    outs = engine.detect_host_batch([audio, short, audio, audio[:0], short])
    assert np.array_equal(outs[0], bins) and np.array_equal(outs[2], bins)
    assert np.array_equal(outs[3], engine.detect_host(audio[:0]))       # an empty clip is five windows of padding
    assert np.array_equal(outs[1], engine.detect_host(short)) and np.array_equal(outs[4], outs[1])
    engine.set_refine(0.0)
    assert engine.check_guards() == 0


def test_refinement_off_and_same_mode_are_no_ops(engine):
    audio = synth.synth_audio(20.0, 1)
    engine.set_refine(0.0)
    a, la = engine.detect_host(audio, want_logits=True)
    engine.set_refine(1e-2, "f16x3")                       # refine mode == first-pass mode: skipped
    engine.refine_stats(reset=True)
    b, lb = engine.detect_host(audio, want_logits=True)
    assert engine.refine_stats()["windows_refined"] == 0
    assert np.array_equal(a, b) and np.array_equal(la, lb)
    engine.set_refine(0.0)


def test_streamed_recording_refines_from_the_host_buffer(sd_seed0):
    """A clip longer than one staging chunk (1,024 windows) is refined from the caller's host buffer."""
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=64, mode="f16x3")
    audio = np.tile(synth.synth_audio(60.0, 3), 11)[: 22050 * 640]        # 1,072 windows: two chunks
    eng.set_refine(0.0)
    _, first = eng.detect_host(audio, want_logits=True)
    _, fp32 = eng.detect_host(audio, want_logits=True, mode="fp32")
    eps = 5e-4
    eng.set_refine(eps, "fp32")
    _, got = eng.detect_host(audio, want_logits=True)
    marked = np.zeros(len(first), bool)
    marked[_flagged_windows(first, len(audio), eps)] = True
    assert marked.any() and marked[1030:].any() and not marked.all()
    assert np.array_equal(got[marked], fp32[marked]) and np.array_equal(got[~marked], first[~marked])
    assert eng.check_guards() == 0
    eng.close()


@pytest.mark.parametrize("mode", ["fp32", "bf16", "f16", "f16x3"])
def test_no_kernel_writes_outside_its_buffers(sd_seed0, mode):
    """Guard bands around every device allocation stay intact for ragged batch sizes in every classifier mode."""
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=48, mode=mode)
    audio = torch.from_numpy(synth.synth_audio(60.0, 2)).cuda()
    padded = eng.pad(audio)
    for n_win in (1, 5, 48, 53):
        starts = torch.arange(n_win, dtype=torch.int64) * spec.STEP_SAMPLES
        mel = eng.features(padded, starts)
        eng.classify(mel, want_spec=(n_win == 5))
    eng.detect_host(audio.cpu().numpy())
    eng.check_health()
    assert eng.check_guards() == 0
    eng.close()
