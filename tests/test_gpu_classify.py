"""K2-K4 on the GPU vs golden logits frozen from the reference's CPU network."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from softspoken_b200 import spec

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-4      # north_star: logits within 1e-4 relative (fp32 parity mode)


@pytest.fixture(scope="module")
def engine(sd_seed0):
    from softspoken_b200.engine import Engine
    eng = Engine(sd_seed0, 0, max_batch=16, mode="fp32")
    yield eng
    eng.close()


def _mel(engine, clip60, starts):
    from oracle import postproc as pp
    padded = torch.from_numpy(pp.pad_audio(clip60)).cuda()
    return engine.features(padded, torch.from_numpy(np.asarray(starts)))


def test_logits_match_reference_golden_all_windows(engine, clip60):
    g = load_golden("model_seed0.npz")
    logits = engine.classify(_mel(engine, clip60, g["starts"])).cpu().numpy()
    ref = g["logits"][:, 0, :]
    err = np.max(np.abs(logits - ref))
    scale = np.max(np.abs(ref))
    print(f"fp32 logits vs reference golden: max|d| = {err:.3e}, max|ref| = {scale:.3f}, rel = {err / scale:.3e}")
    assert err <= LOGIT_TOL * scale


def test_spec_head_matches_reference_golden(engine, clip60):
    g = load_golden("model_seed0.npz")
    logits, spec_out = engine.classify(_mel(engine, clip60, g["starts"][41:42]), want_spec=True)
    ref = g["spec_w41"]
    err = np.max(np.abs(spec_out[0].cpu().numpy() - ref)) / np.max(np.abs(ref))
    print(f"spec head rel err {err:.3e}")
    assert err <= LOGIT_TOL
    assert np.max(np.abs(logits[0].cpu().numpy() - g["logits"][41, 0])) <= LOGIT_TOL * np.max(np.abs(g["logits"]))


def test_batch_split_is_bitwise_invariant(engine, clip60):
    """Windows are independent: any batching (the reference uses 32 + ragged tail) gives the same bits."""
    g = load_golden("model_seed0.npz")
    mel = _mel(engine, clip60, g["starts"][:37])
    full = engine.classify(mel)
    parts = torch.cat([engine.classify(mel[:5]), engine.classify(mel[5:6]), engine.classify(mel[6:])])
    assert torch.equal(full, parts)


def test_classify_vs_oracle_on_random_mel(engine, sd_seed0):
    """Inputs the reference never sees (random positive 'mel'): the trunk must still agree with the oracle."""
    from oracle import model as om
    gen = torch.Generator().manual_seed(5)
    mel = torch.rand(2, spec.N_MELS, spec.N_FRAMES, generator=gen) * 1.5
    logits, spec_out = engine.classify(mel.cuda(), want_spec=True)
    sp_ref, mk_ref = om.forward_from_mel(sd_seed0, mel.unsqueeze(1))
    assert torch.max(torch.abs(logits.cpu() - mk_ref[:, 0])) <= LOGIT_TOL * mk_ref.abs().max()
    assert torch.max(torch.abs(spec_out.cpu() - sp_ref)) <= LOGIT_TOL * sp_ref.abs().max()
