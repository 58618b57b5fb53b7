"""The UPSTREAM HiFi-GAN hyper-parameter files (config_v1/v2/v3.json).  The fork deleted them although its README
still tells the user to pass them (README.md:33-39) and `Generator` still consumes exactly these keys
(src/models.py:79-96); values restated from jik876/hifi-gan."""
import json
import os

from ..env import AttrDict

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_config(version: str = "v1") -> AttrDict:
    """AttrDict `h` for 'v1' | 'v2' | 'v3' — what `inference.py:74-80` builds from config.json."""
    with open(os.path.join(_HERE, f"config_{version}.json")) as f:
        return AttrDict(json.load(f))
