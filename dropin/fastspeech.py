"""Drop-in module: put this directory in front of the reference checkout on sys.path
(`PYTHONPATH=/path/to/fs2-b200/dropin:/path/to/fs2-b200:$PYTHONPATH`) and the reference's
`from fastspeech import FeedForwardTransformer` / `import fastspeech` (inference.py:9,
evaluation.py:3, train_fastspeech.py:1) resolve to the B200 path."""
from fastspeech2_b200.fastspeech import FeedForwardTransformer  # noqa: F401
