"""Run an unmodified reference script with this package's `FeedForwardTransformer` in place of the reference's.

    cd /path/to/FastSpeech2
    python -m fastspeech2_b200.dropin_run inference.py -c configs/default.yaml -p ckpt.pyt --text "..."

Why a launcher: `python inference.py` puts the script's own directory first on `sys.path`, so the reference's
`fastspeech.py` (same directory) always wins over a `PYTHONPATH` entry.  This module puts `<repo>/dropin` (a module named
`fastspeech` that re-exports our class) in front of the script directory and then executes the script as `__main__`.
The other way to switch over is to copy `dropin/fastspeech.py` over the checkout's `fastspeech.py`.
"""
from __future__ import annotations

import os
import runpy
import sys

DROPIN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m fastspeech2_b200.dropin_run <reference script.py> [script args...]")
    script = os.path.abspath(argv[0])
    front = [DROPIN_DIR, os.path.dirname(script)]
    cwd = os.getcwd()
    if cwd not in front:
        front.append(cwd)
    sys.path[:] = front + [p for p in sys.path if p not in front and p != ""]
    sys.modules.pop("fastspeech", None)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
