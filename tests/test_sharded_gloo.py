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


def _grad_worker(rank, world, port):
    """GradientSync (data-parallel training) on a stand-in module: after the one bucket all-reduce every rank
    holds the mean of the per-rank gradients, equal to the gradient of the whole batch when shards are equally sized; the
    optimizer then keeps the replicas bit-identical."""
    from fastspeech2_b200.sharded import GradientSync
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                       # replicas start DIFFERENT: broadcast must fix that
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        sync = GradientSync(net)
        sync.broadcast_parameters(src=0)
        ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        torch.manual_seed(100)
        ref0 = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        for a, b in zip(net.parameters(), ref0.parameters()):
            assert torch.equal(a, b)
        net.eval(); ref0.eval()                             # BatchNorm on running statistics: shard losses add up exactly
        g = torch.Generator().manual_seed(5)
        x, y = torch.randn(8, 7, generator=g), torch.randn(8, 3, generator=g)
        lo, hi = shard_bounds(8, rank, world)
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        for _ in range(2):
            sync.zero_grad()
            torch.nn.functional.mse_loss(net(x[lo:hi]), y[lo:hi]).backward()
            sync.all_reduce()
            for p in ref0.parameters():
                p.grad = None
            torch.nn.functional.mse_loss(ref0(x), y).backward()
            for a, b in zip(net.parameters(), ref0.parameters()):
                assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7), (a.grad, b.grad)
            opt.step()
            with torch.no_grad():
                for p in ref0.parameters():
                    p -= 0.1 * p.grad
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                # replicas stayed in lock step
        # a .grad that was re-created (set_to_none) no longer aliases the bucket: loud error, not a silent no-op
        opt.zero_grad(set_to_none=True)
        try:
            sync.all_reduce()
            raise AssertionError("expected GradientSync to notice the detached .grad")
        except RuntimeError:
            pass
        del ref
    finally:
        dist.destroy_process_group()


def test_gradient_sync_world2():
    mp.spawn(_grad_worker, args=(2, _free_port()), nprocs=2, join=True)
