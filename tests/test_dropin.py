"""The drop-in module shadows the reference's `fastspeech` module name and exports the same class name with the
reference's constructor and method signatures.  CPU only."""
import importlib.util
import inspect
import os

from conftest import REPO


def test_dropin_module_exports_the_class():
    spec = importlib.util.spec_from_file_location("fastspeech", os.path.join(REPO, "dropin", "fastspeech.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cls = mod.FeedForwardTransformer
    assert cls.__name__ == "FeedForwardTransformer"
    init = list(inspect.signature(cls.__init__).parameters)
    assert init[:4] == ["self", "idim", "odim", "hp"]                                   # fastspeech.py:37
    fwd = list(inspect.signature(cls.forward).parameters)
    assert fwd == ["self", "xs", "ilens", "ys", "olens", "ds", "es", "ps"]               # fastspeech.py:245-254
    low = inspect.signature(cls._forward).parameters
    assert list(low)[:8] == ["self", "xs", "ilens", "olens", "ds", "es", "ps", "is_inference"]   # fastspeech.py:169-178
    assert low["olens"].default is None and low["is_inference"].default is False
    assert list(inspect.signature(cls.inference).parameters) == ["self", "x"]            # fastspeech.py:339
