"""Robustness of the CUDA path beyond the synthetic-default checkpoint ("what's weak" 2-4):

  * trained-checkpoint statistics -- per-layer weights of magnitude 1e-3 .. 1e-2, LayerNorm gamma << 1, large BatchNorm
    running variances: durations and bucket ids bit-exact against the CPU oracle in every tensor-core mode;
  * BASELINE config 4 at its full batch (B=32, L=2000), config 2 at full size with the batch-independence property;
  * a CUDA-graph replay after an eager call that re-allocated the model's workspace;
  * two devices driven from one process (per-device kernel attributes);
  * the sharded path on two GPUs against the per-shard oracle (SURVEY 8e caveat).
Needs a B200: run with `-m gpu`.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from fastspeech2_b200 import FeedForwardTransformer
from fastspeech2_b200.hparams import load_hp
from fastspeech2_b200.synthetic import make_batch
from fastspeech2_b200.weights import ModelDims, synthetic_state_dict
from oracle import fs2_oracle as O
from test_gpu_parity import TOL, close

pytestmark = pytest.mark.gpu
TC_MODES = ["3xtf32", "f16", "tf32"]
KEYS = ("xs", "ilens", "olens", "ds", "es", "ps")


def trained_like_state_dict(seed: int):
    """Checkpoint with the statistics of a trained model rather than a fresh init: every conv / linear weight tensor
    scaled so that |w| ~ 1e-3 .. 1e-2 (a different factor per layer), LayerNorm gains ~0.1, BatchNorm running
    variances of 10 .. 100, and predictor heads strong enough that durations / bucket ids vary along the sequence."""
    sd = synthetic_state_dict(seed, ModelDims())
    g = torch.Generator().manual_seed(seed + 1000)
    for k, v in sd.items():
        if k.endswith("linear.weight") and "predictor" in k:
            sd[k] = v * 6.0                                            # scalar heads: keep the integers lively
        elif k.endswith(".weight") and v.dim() >= 2 and "embed.0" not in k and "_embed" not in k:
            target = 10 ** float(torch.empty(1).uniform_(-3.0, -2.0, generator=g))   # max |w| of this layer
            sd[k] = v * (target / float(v.abs().max()))
        elif ("norm" in k and k.endswith(".weight")) or k == "decoder.embed.1.weight":
            sd[k] = 0.1 * (1.0 + 0.1 * torch.randn(v.shape, generator=g))
        elif k.endswith("running_var"):
            sd[k] = 10.0 + 90.0 * torch.rand(v.shape, generator=g)
    return sd


@pytest.fixture(scope="module")
def trained_models():
    sd = trained_like_state_dict(11)
    out = {}
    for prec in TC_MODES + ["fp32"]:
        m = FeedForwardTransformer(68, 80, load_hp(), precision=prec)
        m.load_state_dict(sd, strict=True)
        out[prec] = m.cuda().eval()
    return sd, out


def test_trained_like_weights_integers_bit_exact(trained_models):
    sd, models = trained_models
    g = torch.Generator().manual_seed(5)
    ilens = [70, 64, 51, 33, 20, 9]
    xs = torch.zeros(6, 70, dtype=torch.int64)
    for b, n in enumerate(ilens):
        xs[b, :n] = torch.randint(1, 68, (n,), generator=g)
    il = torch.tensor(ilens)
    with torch.no_grad():
        want = O.forward_path(sd, xs, il, is_inference=True)
    assert int(want[2].max()) > int(want[2][want[2] > 0].min()), "test weights give constant durations: not a useful case"
    scale = float(want[1].abs().max())
    for prec in TC_MODES + ["fp32"]:
        with torch.no_grad():
            got = models[prec]._forward(xs.cuda(), il.cuda(), is_inference=True)
        assert torch.equal(got[2].cpu(), want[2]), f"{prec}: durations differ"
        assert torch.equal(got[3].argmax(-1).cpu(), want[3].argmax(-1)), f"{prec}: energy bucket ids differ"
        assert torch.equal(got[4].argmax(-1).cpu(), want[4].argmax(-1)), f"{prec}: pitch bucket ids differ"
        tol = {k: v * max(1.0, scale) for k, v in TOL[prec].items()}
        close(got[1], want[1], tol, f"after ({prec}, trained-like weights)")


def test_trained_like_weights_teacher_forced(trained_models):
    sd, models = trained_models
    bt = make_batch(4, 40, 330, seed=12, ilens=[40, 31, 17, 40], olens=[330, 250, 140, 329])
    with torch.no_grad():
        want = O.forward_path(sd, bt["xs"], bt["ilens"], bt["olens"], bt["ds"].clone(), bt["es"], bt["ps"], False)
    scale = max(1.0, float(want[1].abs().max()))
    for prec in TC_MODES + ["fp32"]:
        with torch.no_grad():
            got = models[prec]._forward(*[bt[k].cuda() for k in KEYS], is_inference=False)
        tol = {k: v * scale for k, v in TOL[prec].items()}
        close(got[0], want[0], tol, f"before ({prec})"); close(got[1], want[1], tol, f"after ({prec})")
        close(got[2], want[2], TOL["fp32"], f"d_outs ({prec})")
        close(got[3], want[3], dict(max=2e-4, mean=2e-5), f"e_outs ({prec})"); close(got[4], want[4], dict(max=2e-4, mean=2e-5), f"p_outs ({prec})")


def test_config4_full_batch(weights):
    """BASELINE config 4 at full size: B=32, T=250, L=2000 (ragged 1500..2000) against the CPU oracle."""
    g = torch.Generator().manual_seed(44)
    olens = [2000] + [int(v) for v in torch.randint(1500, 2001, (31,), generator=g)]
    ilens = [250] + [int(v) for v in torch.randint(180, 251, (31,), generator=g)]
    bt = make_batch(32, 250, 2000, seed=45, ilens=ilens, olens=olens)
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    with torch.no_grad():
        want = O.forward_path(weights, bt["xs"], bt["ilens"], bt["olens"], bt["ds"].clone(), bt["es"], bt["ps"], False)
    valid = torch.arange(2000)[None] < bt["olens"][:, None]
    for prec in ("3xtf32", "f16"):
        m = FeedForwardTransformer(68, 80, load_hp(), precision=prec)
        m.load_state_dict(weights, strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            got = m._forward(*[bt[k].cuda() for k in KEYS], is_inference=False)
        close(got[1].cpu()[valid], want[1][valid], TOL[prec], f"c4 after ({prec})")
        close(got[2], want[2], TOL["fp32"], f"c4 d_outs ({prec})")
        del m
        torch.cuda.empty_cache()


def test_config2_full_batch_oracle_and_batch_independence(weights):
    """BASELINE config 2 -- the benchmarked shape, B=64, T=100, every utterance 800 frames -- at full size: (i) against the CPU
    oracle, (ii) a size-independent property: with no padding in the batch an utterance's result does not depend on its batch
    mates (every GEMM row, attention head, convolution window and LayerNorm row reads only its own utterance, and the K-loop
    order of a row does not depend on which tile it sits in), so utterance b of the 64-batch must equal, BIT FOR BIT, the same
    utterance synthesised alone -- in the fp32-class mode and in the f16 mode, and also through the captured CUDA graph."""
    bt = make_batch(64, 100, 800, seed=1234)
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    with torch.no_grad():
        want = O.forward_path(weights, bt["xs"], bt["ilens"], bt["olens"], bt["ds"].clone(), bt["es"], bt["ps"], False)
    for prec in ("3xtf32", "f16"):
        m = FeedForwardTransformer(68, 80, load_hp(), precision=prec)
        m.load_state_dict(weights, strict=True)
        m = m.cuda().eval()
        args = [bt[k].cuda() for k in KEYS]
        with torch.no_grad():
            got = [t.clone() for t in m._forward(*args, is_inference=False)]
            close(got[1].cpu(), want[1], TOL[prec], f"c2 after ({prec})")
            close(got[0].cpu(), want[0], TOL[prec], f"c2 before ({prec})")
            close(got[2], want[2], TOL["fp32"], f"c2 d_outs ({prec})")
            for b in (0, 17, 63):
                one = m._forward(*[a[b:b + 1].contiguous() for a in args], is_inference=False)
                for name, x, y in zip(("before", "after", "d_outs", "e_outs", "p_outs"), one, got):
                    assert torch.equal(x[0], y[b]), f"{prec}: {name} of utterance {b} depends on its batch mates"
            g = m.graphed_forward(*args)
            rep = g(*args)
            assert all(torch.equal(a, b_) for a, b_ in zip(rep, got)), f"{prec}: graph replay differs from the eager step at full size"
        del m, g
        torch.cuda.empty_cache()


def test_graph_replay_after_workspace_growth(weights):
    """the captured graph must not be corrupted by a later, larger eager call that re-allocates the
    model's scratch (and anything else the caching allocator hands out in between)."""
    m = FeedForwardTransformer(68, 80, load_hp())
    m.load_state_dict(weights, strict=True)
    m = m.cuda().eval()
    small = make_batch(2, 20, 170, seed=61, ilens=[20, 13], olens=[170, 101])
    big = make_batch(6, 60, 700, seed=62)
    S = [small[k].cuda() for k in KEYS]
    with torch.no_grad():
        ref = [t.clone() for t in m._forward(*S, is_inference=False)]
        gr = m.graphed_forward(*S)
        m._forward(*[big[k].cuda() for k in KEYS], is_inference=False)        # grows + re-allocates model._workspace
        junk = [torch.randn(1 << 22, device="cuda") for _ in range(8)]         # allocator traffic over any freed block
        out = gr(*S)
        torch.cuda.synchronize()
    for r, o in zip(ref, out):
        assert torch.equal(r, o)
    del junk
    # weight updates the version counters cannot see are caught through invalidate()
    with torch.no_grad():
        m.feat_out.bias.data.add_(1.0)
    m.invalidate()
    with pytest.raises(RuntimeError, match="parameters changed"):
        gr(*S)


def test_deferred_validation_reports_on_next_call(weights):
    m = FeedForwardTransformer(68, 80, load_hp())
    m.load_state_dict(weights, strict=True)
    m = m.cuda().eval()
    a = make_batch(2, 16, 120, seed=71)
    A = [a[k].cuda() for k in KEYS]
    gr = m.graphed_forward(*A)
    gr(*A, validate="deferred")
    gr.flush()
    bad = [t.clone() for t in A]
    bad[3][0, 0] += 3
    gr(*bad, validate="deferred")                    # queued without a host sync ...
    with pytest.raises(RuntimeError, match="length mismatch"):
        gr(*A, validate="deferred")                  # ... and reported at the start of the next call


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_one_process(weights):
    """Kernel attributes (dynamic shared memory opt-in) are per device: a second model on cuda:1 in the same process
    must launch, match cuda:0 bit for bit, and leave torch's current device alone."""
    bt = make_batch(2, 24, 200, seed=81, ilens=[24, 11], olens=[200, 93])
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = FeedForwardTransformer(68, 80, load_hp())
        m.load_state_dict(weights, strict=True)
        m = m.to(dev).eval()
        with torch.no_grad():
            outs.append(m._forward(*[bt[k].to(dev) for k in KEYS], is_inference=False)[1].cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_two_ranks_vs_per_shard_oracle():
    """the oracle of a sharded batch is the reference run per shard (padding leaks into valid frames)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(root, "tests", "_sharded_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_OK" in r.stdout
