"""The C-ABI library loads, exports every symbol the header declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "softspoken_b200.h")).read()
    return re.findall(r"^SS_API\s+[\w\s\*]+?\b(ss_\w+)\(", text, flags=re.M)


def test_library_exports_every_declared_symbol():
    from softspoken_b200 import _lib
    names = _declared()
    assert len(names) >= 19 and len(set(names)) == len(names)
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/softspoken_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert _lib.lib.ss_abi_version() == _lib.ABI_VERSION == 2


def test_constants_match_reference_settings():
    from softspoken_b200 import _lib, settings, spec
    assert _lib.get_constant("sample_rate") == settings.vad_resample == 22050
    assert _lib.get_constant("hop_length") == settings.hop_length == 256
    assert _lib.get_constant("win_length") == settings.win_length == 512
    assert _lib.get_constant("n_fft") == settings.n_fft * 4
    assert _lib.get_constant("window_samples") == spec.WINDOW_SAMPLES == 3 * settings.vad_resample
    assert _lib.get_constant("step_samples") == spec.STEP_SAMPLES == int(settings.vad_resample * settings.step_size)
    assert _lib.get_constant("threshold") == settings.threshold
    with pytest.raises(_lib.SoftspokenError):
        _lib.get_constant("no_such_constant")


def test_host_helpers_without_gpu():
    from oracle import postproc as pp
    from softspoken_b200.engine import plan_windows, timeline_bins
    for n in [0, 1, 13229, 13230, 13231, 66150, 1323000, 13230000, 1905120000]:
        assert plan_windows(n) == len(pp.plan_windows(n / 22050))
        assert timeline_bins(n + 132300) == pp.output_length((n + 132300) / 22050)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from softspoken_b200 import _lib, checkpoint
    from softspoken_b200.engine import Engine
    ctx = C.c_void_p()
    blob = checkpoint.pack_blob(checkpoint.synthetic_state_dict(0))
    rc = _lib.lib.ss_ctx_create(0, blob, len(blob), 8, C.byref(ctx))
    assert rc == _lib.SS_E_NODEVICE and not ctx.value
    assert b"no CPU fallback" in _lib.lib.ss_last_error()
    with pytest.raises(_lib.SoftspokenError):
        Engine(checkpoint.synthetic_state_dict(0))
    from softspoken_b200.detector import NNDetector
    with pytest.raises(RuntimeError):
        NNDetector(object())


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU restatement)."""
    pkg = os.path.join(ROOT, "softspoken_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "ref_shim" not in src and "/root/reference" not in src, fn
