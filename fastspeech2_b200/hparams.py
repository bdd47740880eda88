"""Minimal attribute-access config so the package is usable without the reference's
utils/hparams.py.  The reference's own `HParam` object is accepted unchanged by
`FeedForwardTransformer` (it is a dict subclass with attribute access too)."""
from __future__ import annotations

import os

import yaml

DEFAULT_YAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "default.yaml")


class AttrDict(dict):
    """dict whose keys are also attributes; nested mappings are converted recursively."""

    def __init__(self, mapping=None):
        super().__init__()
        for k, v in (mapping or {}).items():
            self[k] = AttrDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def load_hp(path: str = DEFAULT_YAML) -> AttrDict:
    with open(path) as f:
        return AttrDict(yaml.safe_load(f))


def load_hp_str(hp_str: str) -> AttrDict:
    """The `hp_str` a reference checkpoint carries (train_fastspeech.py:235-244): the YAML text of the config, possibly
    several documents, merged like utils/hparams.py:14-24 does."""
    merged = {}
    for doc in yaml.safe_load_all(hp_str):
        if isinstance(doc, dict):
            merged.update(doc)
    return AttrDict(merged)
