"""N>1 host logic on CPU: world_size-2 gloo run of the shard partition + the single mel gather
(fastspeech2_b200/sharded.py).  The model itself is not involved (it needs a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastspeech2_b200.sharded import gather_mels, gather_mels_to_root, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, equal):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        n = 10
        full_lens = torch.randint(3, 9, (n,), generator=g) if not equal else torch.full((n,), 6)
        full = torch.randn(n, int(full_lens.max()), 5, generator=g)
        for i in range(n):
            full[i, full_lens[i]:] = 0.0
        lo, hi = shard_bounds(n, rank, world)
        lens = full_lens[lo:hi]
        local = full[lo:hi, : int(lens.max())].contiguous()   # a shard is padded to ITS OWN Lmax
        mels, out_lens = gather_mels(local, lens, equal_shapes=equal)
        assert torch.equal(out_lens, full_lens)
        assert mels.shape == full.shape and torch.equal(mels, full)
    finally:
        dist.destroy_process_group()


def test_gather_mels_world2_ragged():
    mp.spawn(_worker, args=(2, _free_port(), False), nprocs=2, join=True)


def test_gather_mels_world2_equal_single_collective():
    mp.spawn(_worker, args=(2, _free_port(), True), nprocs=2, join=True)


def _root_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.randn(8, 6, 5, generator=torch.Generator().manual_seed(1))
        lo, hi = shard_bounds(8, rank, world)
        got = gather_mels_to_root(full[lo:hi].contiguous(), dst=0)
        if rank == 0:
            assert torch.equal(got, full)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def test_gather_mels_to_root_world2():
    mp.spawn(_root_worker, args=(2, _free_port()), nprocs=2, join=True)


def _peer_worker(rank, world, port):
    """PeerGather's protocol (push per step, root waits, `gathered` view) with CPU tensors: the class degrades to the gloo
    gather, so the host-side sequencing of bench.py / serving loops is exercised without a GPU."""
    from fastspeech2_b200.sharded import PeerGather
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pg = PeerGather((4, 6, 5), "cpu")
        for step in (1, 2, 3):
            full = torch.randn(8, 6, 5, generator=torch.Generator().manual_seed(step))
            lo, hi = shard_bounds(8, rank, world)
            pg.push(full[lo:hi].contiguous(), step)
            pg.wait(step)
            if rank == 0:
                assert torch.equal(pg.gathered, full)
            else:
                assert pg.gathered is None
        pg.close()
    finally:
        dist.destroy_process_group()


def test_peer_gather_protocol_world2_cpu_fallback():
    mp.spawn(_peer_worker, args=(2, _free_port()), nprocs=2, join=True)
