"""Drop-in module named like the reference's `fastspeech.py`: when this directory precedes the reference checkout on
sys.path, the reference's `from fastspeech import FeedForwardTransformer` / `import fastspeech` (inference.py:9,
evaluation.py:3, train_fastspeech.py:1) resolve to the B200 path.  A script's own directory always comes first on
sys.path, so either run the reference scripts through `python -m fastspeech2_b200.dropin_run <script> ...` or copy this
file over the checkout's `fastspeech.py`."""
from fastspeech2_b200.fastspeech import FeedForwardTransformer  # noqa: F401
