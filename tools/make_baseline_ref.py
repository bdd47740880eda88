#!/usr/bin/env python
"""Stage the UNMODIFIED reference (Python sources + config; no audio samples, notebooks or filelists beyond the first 64
rows) into git-ignored `baseline/_ref/` so that it travels to the GPU box with the snapshot:

  * `bench.py --impl reference` then times the reference's own `FeedForwardTransformer._forward` on the host cores
    (`cpu_baseline.kind = "reference"`), not the oracle port;
  * the `-m gpu` drop-in tests run the unmodified `inference.synth` / `evaluation.evaluate` against this repo's class.

The reference has no setup.py / pyproject (plain scripts), so there is nothing `pip install --target` could install; this
is a verbatim file copy from where the sources lie.  Nothing is copied into tracked paths.  Run in the build container
(where /root/reference exists): `python tools/make_baseline_ref.py`; `__graft_entry__.build()` calls it too."""
import os
import shutil
import sys

SRC = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["fastspeech.py", "inference.py", "evaluation.py", "train_fastspeech.py", "export_torchscript.py", "LICENSE"]
DIRS = ["core", "utils", "dataset", "configs", "tests"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: baseline/_ref left as is", file=sys.stderr)
        return 0 if os.path.isdir(DST) else 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for f in FILES:
        shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    for d in DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DST, "filelists"))
    for name in ("train_filelist.txt", "valid_filelist.txt"):
        src = os.path.join(SRC, "filelists", name)
        if os.path.exists(src):
            with open(src) as fi, open(os.path.join(DST, "filelists", name), "w") as fo:
                for i, line in enumerate(fi):
                    if i >= 64:
                        break
                    fo.write(line)
    print("staged the reference into", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
