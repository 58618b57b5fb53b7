"""Top-level `utils` shim (`from utils import init_weights, get_padding`, models.py:6)."""
from hifigan_b200.utils import (apply_weight_norm, get_padding, init_weights, load_checkpoint,  # noqa: F401
                                save_checkpoint, scan_checkpoint)
