"""Import the UNMODIFIED reference staged in git-ignored `baseline/_ref/` (tools/make_baseline_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY (like everything under oracle/): used by `bench.py --impl reference`, bench's
`cpu_baseline` leg and the drop-in tests.  The reference's import chain touches third-party packages the model path never
uses; they are stubbed in `sys.modules` before the import."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

STUB_SNIPPET = '''
import sys, types
for _n in ("librosa", "librosa.util", "librosa.filters", "nltk", "g2p_en", "unidecode", "inflect", "configargparse",
           "matplotlib", "matplotlib.pyplot", "tensorboardX", "pyworld", "soundfile"):
    if _n not in sys.modules:
        try:
            __import__(_n)
        except Exception:
            sys.modules[_n] = types.ModuleType(_n)
_m = sys.modules
if not hasattr(_m["g2p_en"], "G2p"): _m["g2p_en"].G2p = object
if not hasattr(_m["unidecode"], "unidecode"): _m["unidecode"].unidecode = lambda s: s
if not hasattr(_m["inflect"], "engine"): _m["inflect"].engine = lambda: None
if not hasattr(_m["matplotlib"], "use"): _m["matplotlib"].use = lambda *a, **k: None
if not hasattr(_m["matplotlib"], "pyplot"): _m["matplotlib"].pyplot = _m["matplotlib.pyplot"]
if not hasattr(_m["tensorboardX"], "SummaryWriter"): _m["tensorboardX"].SummaryWriter = object
if not hasattr(_m["librosa"], "util"): _m["librosa"].util = _m["librosa.util"]
if not hasattr(_m["librosa"], "filters"): _m["librosa"].filters = _m["librosa.filters"]
for _k in ("pad_center", "normalize"):
    if not hasattr(_m["librosa.util"], _k): setattr(_m["librosa.util"], _k, lambda *a, **k: None)
if not hasattr(_m["librosa.util"], "tiny"): _m["librosa.util"].tiny = lambda *a, **k: 0.0
if not hasattr(_m["librosa.filters"], "mel"): _m["librosa.filters"].mel = lambda *a, **k: None
'''


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "fastspeech.py"))


def load_reference():
    """Returns (FeedForwardTransformer class of the reference, hp = HParam("configs/default.yaml")).  Puts baseline/_ref
    first on sys.path and chdirs there for the duration of the import (the config path is relative)."""
    if not available():
        raise RuntimeError(f"{REF_DIR} is missing: run tools/make_baseline_ref.py in the build container")
    exec(STUB_SNIPPET, {})
    cwd = os.getcwd()
    sys.path.insert(0, REF_DIR)
    saved = sys.modules.pop("fastspeech", None)
    try:
        os.chdir(REF_DIR)
        import fastspeech as ref_fastspeech            # noqa: the reference's module
        from utils.hparams import HParam                # noqa
        hp = HParam("configs/default.yaml")
        return ref_fastspeech.FeedForwardTransformer, hp
    finally:
        os.chdir(cwd)
        sys.modules.pop("fastspeech", None)
        if saved is not None:
            sys.modules["fastspeech"] = saved
