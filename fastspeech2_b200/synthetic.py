"""Seeded synthetic LJSpeech-shaped batches (there is no dataset offline).

Shapes and value ranges follow the reference's data contract (dataset/dataloader.py:96-118,
configs/default.yaml:5-13): phoneme ids in 1..67 (0 = pad), energy in [e_min, e_max], pitch 0
(unvoiced, ~30 %) or in [p_min, p_max], integer durations >= 1 that sum to the utterance's
mel length (LJSpeech: ~8 frames per phoneme on average).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch


def make_batch(B: int, T: int, L: int, seed: int = 1234, ilens: Optional[Sequence[int]] = None,
               olens: Optional[Sequence[int]] = None, n_sym: int = 68, odim: int = 80) -> Dict[str, torch.Tensor]:
    """CPU tensors: xs[B,T] i64, ilens[B], olens[B], ds[B,T] i64, es/ps[B,L] f32, ys[B,L,odim] f32.
    Utterance b has ilens[b] phonemes whose durations sum to olens[b] (defaults: all T / all L)."""
    g = torch.Generator().manual_seed(int(seed))
    il = torch.tensor(list(ilens) if ilens is not None else [T] * B, dtype=torch.int64)
    ol = torch.tensor(list(olens) if olens is not None else [L] * B, dtype=torch.int64)
    assert int(il.max()) == T and int(ol.max()) == L and bool((ol >= il).all())
    xs = torch.zeros(B, T, dtype=torch.int64)
    ds = torch.zeros(B, T, dtype=torch.int64)
    es = torch.zeros(B, L)
    ps = torch.zeros(B, L)
    for b in range(B):
        n, m = int(il[b]), int(ol[b])
        xs[b, :n] = torch.randint(1, n_sym, (n,), generator=g)
        w = torch.rand(n, generator=g) ** 2 + 0.05                      # skewed like real phoneme durations
        extra = torch.multinomial(w / w.sum(), m - n, replacement=True, generator=g) if m > n else torch.empty(0, dtype=torch.int64)
        ds[b, :n] = 1 + torch.bincount(extra, minlength=n)
        es[b, :m] = 0.0179 + torch.rand(m, generator=g) * (130.5 - 0.0179)
        p = 71.0 + torch.rand(m, generator=g) * (676.0 - 71.0)
        p[torch.rand(m, generator=g) < 0.3] = 0.0
        ps[b, :m] = p
    ys = torch.randn(B, L, odim, generator=g)
    return dict(xs=xs, ilens=il, olens=ol, ds=ds, es=es, ps=ps, ys=ys)
