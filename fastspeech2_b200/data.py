"""Input side of the path: batch assembly with the reference's tensor contract.

* `collate_tts(batch)` -- same 9-tuple as the reference's dataset/dataloader.py:96-118
  (inputs, ilens, mels, labels, olens, ids, durations, energys, pitches), zero padded to the batch maxima.
* `PinnedCollator` -- the same tensors assembled straight into reusable page-locked buffers, so the seven H2D
  copies the reference issues per step (train_fastspeech.py:101-107) become asynchronous `non_blocking` copies.
* `BucketBatchSampler` -- length-bucketed batches.  The forward path computes the whole padded [B, Lmax] rectangle
  (the reference's convolutions read padded frames), so padding is real work: bucketing by
  mel length cuts it.  (The reference ships a BinnedLengthSampler, dataloader.py:121-150, but never instantiates it.)

Items are the reference's `TTSDataset.__getitem__` tuples (dataloader.py:47-74):
  (phoneme ids [T] int, mel [L, n_mels] float, id str, mel_len int, durations [T] int, energy [L] float, pitch [L] float).
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


def _lens(batch) -> Tuple[List[int], List[int]]:
    return [int(np.shape(it[0])[0]) for it in batch], [int(np.shape(it[1])[0]) for it in batch]


def _fill(batch, inputs, mels, durations, energys, pitches, labels):
    for i, it in enumerate(batch):
        x, mel, dur, e, p = (torch.as_tensor(np.asarray(it[k])) for k in (0, 1, 4, 5, 6))
        T, L = x.shape[0], mel.shape[0]
        inputs[i, :T] = x
        mels[i, :L] = mel
        durations[i, : dur.shape[0]] = dur
        energys[i, : e.shape[0]] = e
        pitches[i, : p.shape[0]] = p
        labels[i, L - 1:] = 1.0          # stop-token labels, dataloader.py:110-112


def collate_tts(batch: Sequence[tuple]):
    """Drop-in for the reference's `collate_tts`."""
    il, ol = _lens(batch)
    B, T, L = len(batch), max(il), max(ol)
    n_mels = int(np.shape(batch[0][1])[1])
    Td = max(int(np.shape(it[4])[0]) for it in batch)
    Le, Lp = max(int(np.shape(it[5])[0]) for it in batch), max(int(np.shape(it[6])[0]) for it in batch)
    inputs = torch.zeros(B, T, dtype=torch.int64)
    mels = torch.zeros(B, L, n_mels, dtype=torch.float32)
    durations = torch.zeros(B, Td, dtype=torch.int64)
    energys = torch.zeros(B, Le, dtype=torch.float32)
    pitches = torch.zeros(B, Lp, dtype=torch.float32)
    labels = torch.zeros(B, L, dtype=torch.float32)
    _fill(batch, inputs, mels, durations, energys, pitches, labels)
    ids = [it[2] for it in batch]
    return inputs, torch.tensor(il, dtype=torch.int64), mels, labels, torch.tensor(ol, dtype=torch.int64), ids, durations, energys, pitches


class _Collated(tuple):
    """The 9-tuple of `collate_tts` plus the pinned slot it lives in."""
    slot = -1


class PinnedCollator:
    """`collate_tts` into reusable pinned host buffers + asynchronous upload.

    `collator(batch)` returns the 9-tuple with pinned CPU tensors (views of one of `slots` buffer sets);
    `collator.to_device(tuple, device)` issues the H2D copies with `non_blocking=True` on the current stream and records a
    CUDA event for that slot.  Before a slot is refilled its event is waited for, so a loader that runs ahead of the GPU
    can never overwrite host memory an earlier asynchronous copy is still reading (with `slots` batches in flight).
    """

    def __init__(self, max_batch: int, max_T: int, max_L: int, n_mels: int = 80, pin: Optional[bool] = None, slots: int = 2):
        pin = torch.cuda.is_available() if pin is None else pin
        mk = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory() if pin else torch.zeros(shape, dtype=dt)
        self.cap = (max_batch, max_T, max_L)
        self.n_mels = n_mels
        self.slots = [dict(inputs=mk((max_batch, max_T), torch.int64), durations=mk((max_batch, max_T), torch.int64),
                           mels=mk((max_batch, max_L, n_mels), torch.float32), labels=mk((max_batch, max_L), torch.float32),
                           energys=mk((max_batch, max_L), torch.float32), pitches=mk((max_batch, max_L), torch.float32),
                           ilens=mk((max_batch,), torch.int64), olens=mk((max_batch,), torch.int64), event=None)
                      for _ in range(max(1, int(slots)))]
        self._next = 0

    def __call__(self, batch: Sequence[tuple]):
        il, ol = _lens(batch)
        B, T, L = len(batch), max(il), max(ol)
        if B > self.cap[0] or T > self.cap[1] or L > self.cap[2]:
            raise ValueError(f"batch {B}x{T}x{L} exceeds the pinned capacity {self.cap}")
        idx = self._next
        self._next = (idx + 1) % len(self.slots)
        s = self.slots[idx]
        if s["event"] is not None:          # the upload that last read this slot must have finished
            s["event"].synchronize()
            s["event"] = None
        # dense [B, T] / [B, L] windows of the big buffers are not contiguous; stage into contiguous prefixes instead
        n_mels = self.n_mels
        inputs = s["inputs"].view(-1)[: B * T].view(B, T).zero_()
        durations = s["durations"].view(-1)[: B * T].view(B, T).zero_()
        mels = s["mels"].view(-1)[: B * L * n_mels].view(B, L, n_mels).zero_()
        labels = s["labels"].view(-1)[: B * L].view(B, L).zero_()
        energys = s["energys"].view(-1)[: B * L].view(B, L).zero_()
        pitches = s["pitches"].view(-1)[: B * L].view(B, L).zero_()
        _fill(batch, inputs, mels, durations, energys, pitches, labels)
        ilens, olens = s["ilens"][:B], s["olens"][:B]
        ilens.copy_(torch.tensor(il)); olens.copy_(torch.tensor(ol))
        out = _Collated((inputs, ilens, mels, labels, olens, [it[2] for it in batch], durations, energys, pitches))
        out.slot = idx
        return out

    def to_device(self, collated, device):
        """Asynchronous H2D of every tensor of `collated` on the current stream of `device`; remembers an event so the
        slot is not refilled before the copies have run."""
        out = tuple(t.to(device, non_blocking=True) if torch.is_tensor(t) else t for t in collated)
        slot = getattr(collated, "slot", -1)
        if slot >= 0 and torch.device(device).type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            self.slots[slot]["event"] = ev
        return out

    def upload_into(self, collated, dst: Sequence[Optional[torch.Tensor]]):
        """Like `to_device`, but copies into existing device tensors (e.g. the static inputs of a captured CUDA graph):
        `dst[i]` receives `collated[i]` (None / non-tensor entries are skipped).  Asynchronous, same slot bookkeeping."""
        dev = None
        for src, d in zip(collated, dst):
            if d is not None and torch.is_tensor(src):
                d.copy_(src, non_blocking=True)
                dev = d.device
        slot = getattr(collated, "slot", -1)
        if slot >= 0 and dev is not None and dev.type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self.slots[slot]["event"] = ev


class BucketBatchSampler:
    """Batches of indices with similar length: sort by length, cut into bins of `bin_size`, shuffle inside and
    across bins with a seeded generator, then chunk into batches.  `padding_waste(lengths)` reports the fraction
    of padded positions the resulting batches contain (what the rectangle computation pays for)."""

    def __init__(self, lengths: Sequence[int], batch_size: int, bin_size: Optional[int] = None, seed: int = 0, drop_last: bool = False):
        self.lengths = torch.as_tensor(list(lengths), dtype=torch.int64)
        self.batch_size = int(batch_size)
        self.bin_size = int(bin_size or 8 * batch_size)
        if self.bin_size % self.batch_size:
            raise ValueError("bin_size must be a multiple of batch_size")
        self.seed, self.epoch, self.drop_last = seed, 0, drop_last

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __iter__(self) -> Iterator[List[int]]:
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        order = torch.argsort(self.lengths, stable=True)
        bins = [order[i: i + self.bin_size] for i in range(0, len(order), self.bin_size)]
        bins = [b[torch.randperm(len(b), generator=g)] for b in bins]
        batches: List[List[int]] = []
        for b in bins:
            for i in range(0, len(b), self.batch_size):
                chunk = b[i: i + self.batch_size].tolist()
                if len(chunk) == self.batch_size or not self.drop_last:
                    batches.append(chunk)
        for i in torch.randperm(len(batches), generator=g).tolist():
            yield batches[i]

    def __len__(self) -> int:
        n = len(self.lengths)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def padding_waste(self) -> float:
        pad = tot = 0
        for batch in self:
            ls = self.lengths[batch]
            tot += int(ls.max()) * len(batch)
            pad += int(ls.max()) * len(batch) - int(ls.sum())
        return pad / max(tot, 1)
