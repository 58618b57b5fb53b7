"""Recipe for `oracle/_ref/`: the REFERENCE's own hot-path modules, installed beside the oracle so that
`bench.py --impl reference` (and the cpu_baseline leg) can time the reference's code — not this repo's restatement
of it — on the GPU box's host cores, where /root/reference does not exist.

    python oracle/build_ref.py            (run by __graft_entry__.build() in the build container)

What it does: takes src/{models,meldataset,utils,env}.py from the reference tree WHERE THEY LIE (default
/root/reference/src), unmodified, into oracle/_ref/ (git-ignored: reference sources never enter the history; the
directory travels to the GPU box with the snapshot like a built .so), and writes beside them the two import stubs the
reference needs at import time for modules that contribute no arithmetic to this path (matplotlib: utils.py:3-9;
librosa: meldataset.py:7,9 — the same stubs tests/golden/make_golden.py installs).  `load()` imports the result.

Test infrastructure / measurement baseline only: nothing under hifi-gan_b200/ imports it.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("models.py", "meldataset.py", "utils.py", "env.py")


def build(src: str = "/root/reference/src") -> bool:
    """returns False (and leaves any existing oracle/_ref alone) when the reference tree is not present"""
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    with open(os.path.join(REF_DIR, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"copied unmodified from {src} by oracle/build_ref.py: {', '.join(FILES)}\n")
    return True


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in FILES)


def load():
    """import the installed reference modules -> (models, meldataset, env).  Stubs: see the module docstring."""
    if not available():
        raise ImportError("oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)")
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.colors", "librosa", "librosa.util", "librosa.filters"):
        sys.modules.setdefault(name, types.ModuleType(name))
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    mpl.pylab, mpl.colors = sys.modules["matplotlib.pylab"], sys.modules["matplotlib.colors"]
    col = sys.modules["matplotlib.colors"]
    for attr in ("BASE_COLORS", "TABLEAU_COLORS", "CSS4_COLORS"):
        if not hasattr(col, attr):
            setattr(col, attr, {})
    if not hasattr(sys.modules["librosa.util"], "normalize"):
        sys.modules["librosa.util"].normalize = lambda x, **k: x
    if not hasattr(sys.modules["librosa.filters"], "mel"):
        sys.modules["librosa.filters"].mel = lambda *a, **k: None
    # the reference imports its siblings as top-level modules (`from utils import ...`, models.py:6): its directory
    # goes first on the path for the duration of the import, and the modules are kept under private names so they
    # never shadow this repo's compat shims
    saved = {n: sys.modules.pop(n, None) for n in ("models", "meldataset", "utils", "env")}
    sys.path.insert(0, REF_DIR)
    try:
        import env as r_env
        import meldataset as r_mel
        import models as r_models
    finally:
        sys.path.remove(REF_DIR)
        for n in ("models", "meldataset", "utils", "env"):
            mod = sys.modules.pop(n, None)
            if mod is not None:
                sys.modules["_hg_ref_" + n] = mod
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    return r_models, r_mel, r_env


if __name__ == "__main__":
    ok = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src")
    print("oracle/_ref built" if ok else "reference tree not found: oracle/_ref left as it is")
