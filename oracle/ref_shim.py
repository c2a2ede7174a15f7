"""Import the REAL reference (read-only tree at /root/reference) headlessly.

Only usable in the build container: `/root/reference` does not exist on the GPU
box.  Used by `oracle/make_golden.py` (to freeze golden vectors) and by the
tests marked `needs_reference` (skipped when the tree is absent).

The reference's hot-path modules fail to import only because of top-level
imports of packages absent from this image (sounddevice, matplotlib, librosa,
soundfile, PySide6 — voice_activity.py:3-13, worker.py:2,16).  Stub modules are
installed for exactly those names; all arithmetic runs the reference's own code
on the installed torch / torchaudio / numpy / pandas.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("SOFTSPOKEN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "root", "code"))


class _Signal:
    def __init__(self, *a, **k):
        self.log = []

    def connect(self, fn):
        pass

    def emit(self, *a):
        self.log.append(a)


def install_stubs() -> None:
    for name in ("sounddevice", "matplotlib", "matplotlib.pyplot", "librosa",
                 "librosa.display", "soundfile"):
        sys.modules.setdefault(name, MagicMock())
    if "PySide6" not in sys.modules:
        pyside = types.ModuleType("PySide6")
        qtcore = types.ModuleType("PySide6.QtCore")

        class QObject:
            def __init__(self, *a, **k):
                pass

        class QRunnable:
            def __init__(self, *a, **k):
                pass

        class QThreadPool:
            pass

        qtcore.QObject, qtcore.QRunnable, qtcore.QThreadPool = QObject, QRunnable, QThreadPool
        qtcore.Signal = _Signal
        qtcore.Slot = lambda *a, **k: (lambda f: f)
        pyside.QtCore = qtcore
        sys.modules["PySide6"] = pyside
        sys.modules["PySide6.QtCore"] = qtcore


def load():
    """-> namespace with SpecUNet_2D, NNDetector, ProcessWorker, settings."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from root.code.backend import settings
    from root.code.backend.pytorch_neural_nets import SpecUNet_2D
    from root.code.frontend.NNDetector import NNDetector
    from root.code.backend import worker
    return types.SimpleNamespace(settings=settings, SpecUNet_2D=SpecUNet_2D,
                                 NNDetector=NNDetector, worker=worker)


class _ProjectManager:
    def __init__(self, files):
        self._files = list(files)

    def get_unprocessed_list(self):
        return list(self._files)


def make_detector(ref, state_dict, files=(), threads=None):
    """A real reference `NNDetector` with `state_dict` loaded (strict)."""
    import torch
    det = ref.NNDetector(_ProjectManager(files))
    if threads:
        torch.set_num_threads(threads)
    det.model.load_state_dict(state_dict)
    det.model.eval()
    return det


def extract_ui_classes(names=("DetectionProject", "SilenceWorkerSignals", "SilenceWorker"), extra_globals=None):
    """`DetectionProject` / `SilenceWorker` live in silencer_ui.py, which imports Qt
    widgets at module level; lift just those class definitions out with `ast`
    and execute them (unmodified source text) in a namespace that provides the
    handful of names they use."""
    import ast
    import numpy as np
    import pandas as pd
    install_stubs()
    path = os.path.join(REFERENCE_ROOT, "root", "code", "frontend", "silencer_ui.py")
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    qt = sys.modules["PySide6.QtCore"]
    ns = {"os": os, "np": np, "pd": pd, "QObject": qt.QObject, "QRunnable": qt.QRunnable,
          "Signal": qt.Signal, "Slot": qt.Slot, "librosa": MagicMock(), "sf": MagicMock()}
    if extra_globals:
        ns.update(extra_globals)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns
