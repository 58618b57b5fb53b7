"""Top-level `env` shim (`from env import AttrDict`, inference.py:9)."""
from hifigan_b200.env import AttrDict, build_env  # noqa: F401
