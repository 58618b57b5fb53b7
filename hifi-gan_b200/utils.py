"""Helpers with the reference's names and behaviour (src/utils.py:66-101).
The matplotlib plotters of the reference (src/utils.py:16-63) are TensorBoard cosmetics and are out of
scope."""
import glob
import os

import torch
from torch.nn.utils import weight_norm


def init_weights(m, mean=0.0, std=0.01):
    # src/utils.py:66-69.  On a weight_norm-wrapped conv `m.weight` is a derived tensor, so this does not
    # change weight_g / weight_v — but it does advance the RNG, which same-seed construction relies on.
    if "Conv" in type(m).__name__:
        m.weight.data.normal_(mean, std)


def apply_weight_norm(m):
    # src/utils.py:72-75
    if "Conv" in type(m).__name__:
        weight_norm(m)


def get_padding(kernel_size, dilation=1):
    # src/utils.py:78-79 — "same" padding of a dilated odd kernel
    return int((kernel_size * dilation - dilation) / 2)


def load_checkpoint(filepath, device):
    # src/utils.py:82-87 (assert on a missing file is the reference's error convention)
    assert os.path.isfile(filepath)
    print("Loading '{}'".format(filepath))
    checkpoint_dict = torch.load(filepath, map_location=device)
    print("Complete.")
    return checkpoint_dict


def save_checkpoint(filepath, obj):
    # src/utils.py:90-93
    print("Saving checkpoint to {}".format(filepath))
    torch.save(obj, filepath)
    print("Complete.")


def scan_checkpoint(cp_dir, prefix):
    # src/utils.py:96-101 — newest `prefix????????` file or None
    found = sorted(glob.glob(os.path.join(cp_dir, prefix + "????????")))
    return found[-1] if found else None
