"""Data-parallel synthesis over the GPUs of one box.

Utterances are independent, so the batch is cut into contiguous shards, one process per GPU, weights
replicated.  The path has exactly one exchange step: the all-gather of the final mel shards (plus
their lengths).  When shards have different padded lengths a tiny all-reduce(max) of Lmax comes
first so every rank contributes a `[B_r, Lmax, odim]` block.

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo on CPU for the
host-logic tests (tests/test_sharded_gloo.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of `n_items` owned by `rank`: sizes differ by at most one, earlier ranks larger."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_mels(mel: torch.Tensor, olens: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                equal_shapes: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather `[B_r, L_r, odim]` mel shards (zero-extended to the global Lmax) and `[B_r]` lengths.

    Returns (`[sum B_r, Lmax, odim]`, `[sum B_r]`) on every rank, in rank order.  `equal_shapes=True`
    skips the shape exchange (the benchmark's equal shards): a single collective on the mels.
    """
    world = dist.get_world_size(group)
    if world == 1:
        return mel, olens
    B, L, D = mel.shape
    if equal_shapes:
        out = torch.empty((world * B, L, D), dtype=mel.dtype, device=mel.device)
        dist.all_gather_into_tensor(out, mel.contiguous(), group=group)
        lens = torch.empty((world * B,), dtype=olens.dtype, device=olens.device)
        dist.all_gather_into_tensor(lens, olens.contiguous(), group=group)
        return out, lens
    shape = torch.tensor([B, L], dtype=torch.int64, device=mel.device)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    shapes = [tuple(int(v) for v in s.tolist()) for s in shapes]
    Lmax = max(s[1] for s in shapes)
    Bmax = max(s[0] for s in shapes)
    block = torch.zeros((Bmax, Lmax, D), dtype=mel.dtype, device=mel.device)
    block[:B, :L] = mel
    lens_block = torch.zeros((Bmax,), dtype=olens.dtype, device=olens.device)
    lens_block[:B] = olens
    blocks = torch.empty((world, Bmax, Lmax, D), dtype=mel.dtype, device=mel.device)
    dist.all_gather_into_tensor(blocks.view(world * Bmax, Lmax, D), block, group=group)
    lens_all = torch.empty((world * Bmax,), dtype=olens.dtype, device=olens.device)
    dist.all_gather_into_tensor(lens_all, lens_block, group=group)
    mels = torch.cat([blocks[r, : shapes[r][0]] for r in range(world)], dim=0)
    lens = torch.cat([lens_all[r * Bmax: r * Bmax + shapes[r][0]] for r in range(world)], dim=0)
    return mels, lens


def gather_mels_to_root(mel: torch.Tensor, dst: int = 0, group: Optional[dist.ProcessGroup] = None,
                        out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Gather equal-shape `[B, L, odim]` mel shards on rank `dst` only (the other ranks just send their 1/N share):
    the "gather of the final mel batch" of SURVEY section 8e without the N-fold fan-out of an all-gather.  Returns the
    `[world*B, L, odim]` batch on `dst` (written into `out` when given) and None elsewhere."""
    world = dist.get_world_size(group)
    if world == 1:
        return mel
    B, L, D = mel.shape
    rank = dist.get_rank(group)
    if rank == dst:
        if out is None:
            out = torch.empty((world * B, L, D), dtype=mel.dtype, device=mel.device)
        dist.gather(mel.contiguous(), list(out.view(world, B, L, D).unbind(0)), dst=dst, group=group)
        return out
    dist.gather(mel.contiguous(), None, dst=dst, group=group)
    return None


def synthesize_sharded(model, xs: torch.Tensor, ilens: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Batched `is_inference=True` synthesis of a global batch: every rank passes the same global
    `xs [B,T]` / `ilens [B]`, runs its shard and receives all mels.  Returns (mels, olens)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(xs.shape[0], rank, world)
    il = ilens[lo:hi]
    t = int(il.max()) if hi > lo else 0
    with torch.no_grad():
        _, after, d_outs, _, _ = model._forward(xs[lo:hi, :t].contiguous(), il, is_inference=True, _one_hot=False)
    olens = d_outs.sum(dim=1)
    if world == 1:
        return after, olens
    return gather_mels(after, olens, group)
