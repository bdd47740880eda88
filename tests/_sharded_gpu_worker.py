"""Worker of tests/test_gpu_robustness.py::test_sharded_two_ranks_vs_per_shard_oracle (launched under torchrun, one rank
per GPU, NCCL): every rank runs `synthesize_sharded` on the same global batch; rank 0 compares the gathered mels with the
CPU oracle evaluated PER SHARD with the same partition (an utterance's result depends on the padded
lengths of its shard)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from fastspeech2_b200 import FeedForwardTransformer, synthetic_state_dict
    from fastspeech2_b200.hparams import load_hp
    from fastspeech2_b200.sharded import gather_mels_to_root, shard_bounds, synthesize_sharded
    from oracle import fs2_oracle as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = synthetic_state_dict(7)
    g = torch.Generator().manual_seed(123)
    ilens = [41, 38, 30, 22, 17, 9, 5]                      # 7 utterances over 2 ranks: shards of 4 and 3, different Tmax
    xs = torch.zeros(len(ilens), max(ilens), dtype=torch.int64)
    for b, n in enumerate(ilens):
        xs[b, :n] = torch.randint(1, 68, (n,), generator=g)
    il = torch.tensor(ilens)
    ok = True
    for prec, tol in (("fp32", 1e-4), ("3xf16", 1e-4)):
        m = FeedForwardTransformer(68, 80, load_hp(), precision=prec)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        mels, olens = synthesize_sharded(m, xs.to(dev), il.to(dev))
        root = gather_mels_to_root(mels[shard_bounds(len(ilens), rank, world)[0]:shard_bounds(len(ilens), rank, world)[0] + 1].contiguous(), dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            assert root is not None and root.shape[0] == world
            want_parts, want_lens = [], []
            for r in range(world):
                lo, hi = shard_bounds(len(ilens), r, world)
                t = int(il[lo:hi].max())
                with torch.no_grad():
                    w = O.forward_path(sd, xs[lo:hi, :t], il[lo:hi], is_inference=True)
                want_parts.append(w[1]); want_lens.append(w[2].sum(1))
            Lmax = max(p.shape[1] for p in want_parts)
            want = torch.cat([torch.nn.functional.pad(p, (0, 0, 0, Lmax - p.shape[1])) for p in want_parts], 0)
            want_lens = torch.cat(want_lens)
            assert torch.equal(olens.cpu(), want_lens), (prec, olens.cpu(), want_lens)
            got = mels.cpu()
            assert got.shape == want.shape, (got.shape, want.shape)
            valid = torch.arange(Lmax)[None] < want_lens[:, None]
            err = float((got - want).abs()[valid].max())
            print(f"sharded[{prec}]: max-abs err vs per-shard oracle {err:.3e}", flush=True)
            ok = ok and err <= tol
    # fused exchange step: every rank's teacher-forced forward writes its mels straight into the root's receive buffer
    # (peer-mapped output pointer of the last Postnet kernel), then publishes the step with a flag store; the root's view
    # must equal what an NCCL all-gather of locally computed mels delivers, bit for bit
    from fastspeech2_b200.sharded import PeerGather
    from fastspeech2_b200.synthetic import make_batch
    m = FeedForwardTransformer(68, 80, load_hp())
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    B, T, L = 3, 25, 210
    bt = make_batch(B, T, L, seed=900 + rank)
    args = [bt[k].to(dev) for k in ("xs", "ilens", "olens", "ds", "es", "ps")]
    pg = PeerGather((B, L, 80), dev, buffers=2)
    with torch.no_grad():
        local = m._forward(*args, is_inference=False)[1].clone()
        for step, buf in ((1, 0), (2, 1), (3, 0)):
            out = m._forward(*args, is_inference=False, _after_out=pg.slot(buf))[1]
            assert out.data_ptr() == pg.slot(buf).data_ptr()
            pg.signal(step)
            pg.wait(step)
            both = torch.empty((world * B, L, 80), device=dev)
            dist.all_gather_into_tensor(both, local)
            torch.cuda.synchronize()
            if rank == 0:
                assert torch.equal(pg.gathered_buffer(buf), both), f"fused gather differs at step {step}"
        g = m.graphed_forward(*args, after_out=pg.slot(0))          # and through a captured graph
        g(*args)
        pg.signal(4); pg.wait(4)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            assert torch.equal(pg.gathered_buffer(0), both)
            print("fused peer_store gather: bit-identical to all_gather of local results", flush=True)
    pg.close()
    # data-parallel training step: every rank back-propagates its own batch, ONE all-reduce of the flat gradient
    # bucket over NCCL; afterwards every rank holds the mean of the per-rank gradients and the replicas stay identical
    from fastspeech2_b200.sharded import GradientSync
    torch.manual_seed(1000 + rank)
    mt = FeedForwardTransformer(68, 80, load_hp(), precision="fp32")
    mt.load_state_dict(synthetic_state_dict(7 + rank), strict=True)      # replicas start different on purpose
    mt = mt.to(dev).train()
    sync = GradientSync(mt)
    sync.broadcast_parameters(src=0)
    keys = ("xs", "ilens", "ys", "olens", "ds", "es", "ps")
    bt = make_batch(2, 20, 150, seed=40 + rank)
    targs = [bt[k].to(dev) for k in keys]
    opt = torch.optim.Adam(mt.parameters(), lr=1e-3)
    for it in range(2):
        sync.zero_grad()
        loss, _ = mt(*targs)
        loss.backward()
        mine = sync.flat.clone()
        both = torch.empty((world,) + mine.shape, device=dev)
        dist.all_gather_into_tensor(both.view(-1), mine)
        sync.all_reduce()
        torch.cuda.synchronize()
        want = both.mean(0)
        err = float((sync.flat - want).abs().max()) / (float(want.abs().max()) + 1e-20)
        assert err <= 1e-6, f"gradient all-reduce: relative error {err:.3e}"
        assert float(mine.abs().max()) > 0 and not torch.equal(both[0], both[1])
        torch.nn.utils.clip_grad_norm_(mt.parameters(), 1.0)
        opt.step()
    flat_w = torch.cat([p_.detach().reshape(-1) for p_ in mt.parameters()])
    ws = torch.empty((world,) + flat_w.shape, device=dev)
    dist.all_gather_into_tensor(ws.view(-1), flat_w)
    assert torch.equal(ws[0], ws[1]), "replicas diverged after synchronised steps"
    if rank == 0:
        print("data-parallel train step: gradient bucket all-reduce == mean of per-rank gradients, replicas identical", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        assert ok
        print("SHARDED_OK", flush=True)


if __name__ == "__main__":
    main()
