"""Mirror of the reference's configuration module (root/code/backend/settings.py:1-33).

Same attribute names and values, so code written against the reference's
`settings` keeps working; the kernels bake the same constants in and
`Engine.__init__` asserts they agree.  `cpu_threads` is kept for API
compatibility only — nothing on this path runs on CPU threads.
"""
import os

# STFT settings (settings.py:4-6)
n_fft = 512
win_length = n_fft
hop_length = win_length // 2

# controlling the window step size (settings.py:9)
step_size = 0.6

# batches sent for predictions (settings.py:12-13)
prediction_batch_size = 32
threshold = 0.1

# the application operates at 22050 internally (settings.py:16)
vad_resample = 22050

# model settings (settings.py:19-20) — Windows-style relative path kept verbatim;
# checkpoint.normalise_model_path() makes it usable on POSIX.
model_dir = '.\\root\\models\\spec_unet_2d_pytorch'
model_name = 'model_checkpoint.pth'

# project settings and results (settings.py:23)
project_dir = '.\\projects'

# detection duration must be longer than this to be seen for review (settings.py:26)
minimum_detection_len = 0.1

user_guide_url = 'https://github.com/AVianEco/Softspoken'

cpu_threads = (os.cpu_count() or 2) // 2
