"""Compile-time constants of the voice-detector batch path.

Every value restates a constant of the reference (cited file:line, relative to
the reference tree).  The CUDA kernels bake the same numbers in
(`csrc/ss_common.cuh`); `tests/test_abi.py` asserts that both sides agree
through `ss_get_constant`.
"""
import math

# root/code/backend/settings.py:4-6,9,12,13,16,26
N_FFT_SETTING = 512
WIN_LENGTH = 512
HOP_LENGTH = 256
STEP_SIZE_S = 0.6
PREDICTION_BATCH_SIZE = 32
THRESHOLD = 0.1
SAMPLE_RATE = 22050
MINIMUM_DETECTION_LEN = 0.1

# root/code/backend/pytorch_neural_nets.py:87,92-99,150
N_FFT = N_FFT_SETTING * 4          # 2048-point transform of a 512-tap frame
N_FREQS = N_FFT // 2 + 1           # 1025 one-sided bins
N_MELS = 128
F_MAX = 8000.0
N_FRAMES = 256                     # frames kept per window (:150)

# root/code/frontend/NNDetector.py:67-75
WINDOW_S = 3
WINDOW_SAMPLES = SAMPLE_RATE * WINDOW_S                 # 66150
STEP_SAMPLES = math.floor(SAMPLE_RATE * STEP_SIZE_S)    # 13230
# root/code/backend/worker.py:58-62
PAD_SAMPLES = SAMPLE_RATE * 3                           # 66150 zeros each side
# root/code/backend/worker.py:97
BREAK_DURATION_S = 0.5
# merge rule in bins: float(next_start) - float(cur_end) <= 0.5  <=>  gap <= 42 bins
# (42/(256/3) = 0.4922 <= 0.5 < 43/(256/3) = 0.5039)
GAP_BINS = 42

# samples of a window that can influence its 256 kept frames: frame t covers
# x[256 t - 256 .. 256 t + 255]; t = 255 ends at sample 65535.
WINDOW_SAMPLES_USED = 65536


def window_position(i: int) -> int:
    """Start bin of window i on the 256/3 Hz timeline.

    Reference: NNDetector.py:175 `int(round(i * 0.6 / (3 / 256)))`.  The
    integer form `(256 i + 2) // 5` is identical for every i < 2**31 / 256
    (checked exhaustively in tests/test_oracle_postproc.py for i < 200000).
    """
    return (256 * i + 2) // 5
