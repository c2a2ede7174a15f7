import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_library():
    """A fresh checkout has no libsoftspoken_b200.so (built artefacts are git-ignored): build it once, as
    `__graft_entry__.build()` does, so that the suite does not depend on the order the driver runs things in."""
    so = os.path.join(ROOT, "softspoken_b200", "libsoftspoken_b200.so")
    if not os.path.exists(so):
        import shutil
        import subprocess
        if shutil.which("nvcc") and shutil.which("make"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "softspoken_b200", "csrc"), "-j", str(os.cpu_count() or 4)],
                           check=True)


def pytest_configure(config):
    _ensure_library()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    from oracle import ref_shim
    has_ref = ref_shim.available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "needs_reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def head_seed0():
    with open(os.path.join(GOLDEN, "head_seed0.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def sd_seed0(head_seed0):
    from softspoken_b200 import checkpoint
    return checkpoint.synthetic_state_dict(0, head_seed0)


@pytest.fixture(scope="session")
def clip60():
    from softspoken_b200 import synth
    return synth.synth_audio(60.0, 0)
