"""Seeded synthetic side inputs shared by tests/golden/make_golden.py and the tests (so fixtures need not store them)."""
import torch


def seeded_energy_pitch(seed, olens, L):
    """Per-frame energy U(0,130) and pitch (0 w.p. 0.3 else U(71,671)), zero past each utterance's length."""
    g = torch.Generator().manual_seed(seed)
    B = olens.numel()
    es = torch.rand(B, L, generator=g) * 130.0
    ps = torch.rand(B, L, generator=g) * 600.0 + 71.0
    ps[torch.rand(B, L, generator=g) < 0.3] = 0.0
    for b in range(B):
        es[b, olens[b]:] = 0.0
        ps[b, olens[b]:] = 0.0
    return es, ps
