"""Importable alias of the package directory `hifi-gan_b200/` (a hyphen is not a valid module name).

`import hifigan_b200` executes hifi-gan_b200/__init__.py under this name, so
`hifigan_b200.models`, `hifigan_b200.meldataset`, ... resolve to the files in that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hifi-gan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
