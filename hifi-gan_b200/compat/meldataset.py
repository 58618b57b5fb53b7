"""Top-level `meldataset` shim (reference import style: `from meldataset import mel_spectrogram, MAX_WAV_VALUE,
load_wav`, inference.py:10)."""
from hifigan_b200.meldataset import *  # noqa: F401,F403
from hifigan_b200.meldataset import (MAX_WAV_VALUE, MelDataset, SegmentSampler, get_dataset_filelist,  # noqa: F401
                                     hann_window, load_wav, mel_basis, mel_spectrogram, save_wav, torch_mels)
