"""Config carrier — mirrors the reference's src/env.py (AttrDict :5-8, build_env :11-15)."""
import os
import shutil


class AttrDict(dict):
    """dict whose keys are also attributes: the `h` object every model constructor receives."""

    def __init__(self, *args, **kwargs):
        dict.__init__(self, *args, **kwargs)
        self.__dict__ = self


def build_env(config, config_name, path):
    """Copy the config file next to the checkpoints unless it already lives there."""
    target = os.path.join(path, config_name)
    if config == target:
        return
    os.makedirs(path, exist_ok=True)
    shutil.copyfile(config, target)
