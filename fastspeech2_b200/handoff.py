"""Vocoder hand-off: the layout the reference's callers build before MelGAN / Griffin-Lim.

`inference.py:170-180` synthesises a paragraph sentence by sentence, transposes each `[L,80]` mel to `[80,L]`,
concatenates along time and adds a batch axis: `[1, 80, sum L]`.  With the batched path the sentences of a paragraph come
out of ONE `_forward(is_inference=True)` call as a padded `[B,Lmax,80]` batch plus `olens`; these helpers produce the
same vocoder input from that batch without leaving the device (plain torch indexing: plumbing, a few MB per call).
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def paragraph_mel(mels: Sequence[torch.Tensor]) -> torch.Tensor:
    """[L_i, n_mels] per sentence -> [1, n_mels, sum L_i]; the tensor ops of inference.py:173-180 verbatim."""
    return torch.cat([m.T for m in mels], dim=1).unsqueeze(0)


def batch_to_vocoder(mels: torch.Tensor, olens: torch.Tensor) -> torch.Tensor:
    """Padded batch [B, Lmax, n_mels] + valid lengths [B] -> [1, n_mels, sum olens] (sentences in batch order, padding
    dropped), one boolean-mask gather on the device."""
    B, L, _ = mels.shape
    valid = torch.arange(L, device=mels.device)[None, :] < olens.to(mels.device)[:, None]
    return mels[valid].T.contiguous().unsqueeze(0)


def split_utterances(mels: torch.Tensor, olens: torch.Tensor) -> List[torch.Tensor]:
    """Padded batch -> list of [1, n_mels, L_b] views-then-copies, one vocoder call per utterance (inference.py:187-193)."""
    lens = [int(v) for v in olens.tolist()]
    return [mels[b, :n].T.contiguous().unsqueeze(0) for b, n in enumerate(lens)]
