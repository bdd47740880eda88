"""Train-mode forward + backward against the reference's own autograd.

* unit level: each autograd stage of fastspeech2_b200/train.py against the same torch op's autograd in float64;
* model level: `model.train(); loss, report = model(...); loss.backward()` against the UNMODIFIED reference class (from
  baseline/_ref, CPU, fp32) with identical weights, inputs and dropout masks -- `torch.nn.functional.dropout` is patched in
  the reference run to draw masks from a seeded generator and record them; the same masks are injected into our path
  (`model.dropout_masks`).  Compared: the loss, the seven report values, the gradient of every parameter, BatchNorm's
  updated running statistics.
Stated tolerance: fp32 arithmetic with different summation orders (weight gradients are sums over thousands of frames,
split-K with atomics here, one long chain in the reference) -- loss rel 1e-4; gradients max-abs <= 1e-2 * max|g_ref| + 1e-6
(observed: worst 5.3e-3 on a decoder conv-FFN weight, typically 1e-4).
Needs a B200: run with `-m gpu`."""
import math
import os

import pytest
import torch

from fastspeech2_b200 import FeedForwardTransformer
from fastspeech2_b200 import train as T
from fastspeech2_b200.hparams import load_hp
from fastspeech2_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu
KEYS = ("xs", "ilens", "ys", "olens", "ds", "es", "ps")


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / (want.abs().max() + 1e-12))


# ---- unit level ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 70, 256, 1024, 9, 1), (2, 133, 384, 384, 1, 0), (4, 41, 80, 256, 5, 0), (2, 50, 256, 256, 3, 1)])
def test_conv_fn_gradients(shape):
    B, L, K, N, taps, act = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, L, K, generator=g); w = torch.randn(N, K, taps, generator=g) / math.sqrt(K * taps); b = torch.randn(N, generator=g)
    gy = torch.randn(B, L, N, generator=g)
    xr, wr, br = (t.double().requires_grad_() for t in (x, w, b))
    y = torch.nn.functional.conv1d(xr.transpose(1, 2), wr, br, padding=(taps - 1) // 2).transpose(1, 2)
    y = torch.relu(y) if act else y
    y.backward(gy.double())
    xc, wc, bc = (t.cuda().requires_grad_() for t in (x, w, b))
    out = T.ConvFn.apply(xc, wc, bc, act, None)
    out.backward(gy.cuda())
    assert rel_err(out, y) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 1e-5 and rel_err(wc.grad, wr.grad) < 1e-5 and rel_err(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("C", [256, 384])
def test_layernorm_fn_gradients(C):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(5, 77, C, generator=g) * 2 + 0.3; w = 1 + 0.2 * torch.randn(C, generator=g); b = torch.randn(C, generator=g)
    gy = torch.randn(5, 77, C, generator=g)
    xr, wr, br = (t.double().requires_grad_() for t in (x, w, b))
    torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5).backward(gy.double())
    xc, wc, bc = (t.cuda().requires_grad_() for t in (x, w, b))
    T.LayerNormFn.apply(xc, wc, bc, 1e-5).backward(gy.cuda())
    assert rel_err(xc.grad, xr.grad) < 1e-5 and rel_err(wc.grad, wr.grad) < 1e-5 and rel_err(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("C,L,heads", [(256, 100, 2), (384, 333, 2)])
def test_attention_fn_gradients(C, L, heads):
    B, dk = 3, C // heads
    g = torch.Generator().manual_seed(C + L)
    q, k, v = (torch.randn(B, L, C, generator=g) for _ in range(3))
    lens = torch.tensor([L, L // 2, 7])
    drop = torch.rand(B, heads, L, L, generator=g) >= 0.2
    gy = torch.randn(B, L, C, generator=g)
    qr, kr, vr = (t.double().requires_grad_() for t in (q, k, v))
    split = lambda t: t.view(B, L, heads, dk).transpose(1, 2)
    s = split(qr) @ split(kr).transpose(-1, -2) / math.sqrt(dk)
    valid = torch.arange(L)[None] < lens[:, None]
    m = ~(valid[:, None, :] & valid[:, :, None])[:, None]
    p = torch.softmax(s.masked_fill(m, -float("inf")), -1).masked_fill(m, 0.0)
    pd = p * drop.double() / 0.8
    want = (pd @ split(vr)).transpose(1, 2).reshape(B, L, C)
    want.backward(gy.double())
    qc, kc, vc = (t.cuda().requires_grad_() for t in (q, k, v))
    got = T.AttentionFn.apply(qc, kc, vc, lens.cuda(), heads, 0.2, drop.to(torch.uint8).cuda())
    got.backward(gy.cuda())
    assert rel_err(got, want) < 1e-5
    for a, r in ((qc, qr), (kc, kr), (vc, vr)):
        assert torch.isfinite(a.grad).all() and rel_err(a.grad, torch.nan_to_num(r.grad)) < 2e-5


def test_batchnorm_fn_matches_torch_train_mode():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 90, 256, generator=g) * 1.5 + 0.2
    bn = torch.nn.BatchNorm1d(256).double()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(256, generator=g)); bn.bias.copy_(torch.randn(256, generator=g))
    gy = torch.randn(4, 90, 256, generator=g)
    xr = x.double().requires_grad_()
    torch.tanh(bn(xr.transpose(1, 2))).transpose(1, 2).backward(gy.double())
    xc = x.cuda().requires_grad_()
    w, b = bn.weight.detach().float().cuda().requires_grad_(), bn.bias.detach().float().cuda().requires_grad_()
    rm, rv = torch.zeros(256).cuda(), torch.ones(256).cuda()
    T.BatchNormFn.apply(xc, w, b, rm, rv, 1e-5, 0.1, T.ACT_TANH).backward(gy.cuda())
    assert rel_err(xc.grad, xr.grad) < 2e-5 and rel_err(w.grad, bn.weight.grad) < 2e-5 and rel_err(b.grad, bn.bias.grad) < 2e-5
    assert rel_err(rm, bn.running_mean) < 1e-5 and rel_err(rv, bn.running_var) < 1e-5


def test_philox_mask_rate_and_determinism():
    src = T.MaskSource(seed=1234)
    m1 = src.next((7, 333, 256), 0.2, torch.device("cuda"))
    m2 = T.MaskSource(seed=1234).next((7, 333, 256), 0.2, torch.device("cuda"))
    assert torch.equal(m1, m2) and abs(float(m1.float().mean()) - 0.8) < 5e-3
    assert not torch.equal(m1, T.MaskSource(seed=99).next((7, 333, 256), 0.2, torch.device("cuda")))


# ---- model level: the unmodified reference in train mode with shared dropout masks ------------------------------------
class _Recorded(T.MaskSource):
    """Masks recorded from the reference run, in call order; sites where the reference drops a channel-first [B, C, time]
    tensor (conv predictors, Postnet) are permuted to this path's [B, time, C] layout."""

    def __init__(self, masks):
        super().__init__(seed=0, injected=None)
        self.recorded = list(masks)

    def next(self, shape, p, device):
        self.calls += 1
        m = self.recorded.pop(0)
        if tuple(m.shape) != tuple(shape):
            assert m.dim() == 3 and tuple(m.permute(0, 2, 1).shape) == tuple(shape), (self.calls, tuple(m.shape), tuple(shape))
            m = m.permute(0, 2, 1)
        return m.to(torch.uint8).contiguous().to(device)


@pytest.mark.parametrize("ragged", [False, True])
def test_train_step_matches_reference_autograd(weights, ragged):
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("baseline/_ref not staged")
    cls, hp = ref_import.load_reference()
    if ragged:
        bt = make_batch(3, 23, 181, seed=17, ilens=[23, 17, 9], olens=[181, 140, 66])
    else:
        bt = make_batch(2, 20, 150, seed=16)
    ref = cls(68, 80, hp)
    ref.load_state_dict(weights, strict=True)
    ref.train()
    recorded = []
    gen = torch.Generator().manual_seed(5)
    real = torch.nn.functional.dropout

    def shared_dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        m = torch.rand(input.shape, generator=gen) >= p
        recorded.append(m)
        return input * m.to(input.dtype) / (1.0 - p)

    torch.set_num_threads(min(16, os.cpu_count() or 1))
    torch.nn.functional.dropout = shared_dropout
    try:
        loss_ref, rep_ref = ref(*[bt[k] for k in KEYS])
        loss_ref.backward()
    finally:
        torch.nn.functional.dropout = real

    ours = FeedForwardTransformer(68, 80, load_hp(), precision="fp32")
    ours.load_state_dict(weights, strict=True)
    ours = ours.cuda().train()
    ours.dropout_masks = _Recorded(recorded)
    loss, rep = ours(*[bt[k].cuda() for k in KEYS])
    loss.backward()
    torch.cuda.synchronize()
    assert not ours.dropout_masks.recorded, "the reference made more dropout calls than this path"

    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    assert [list(r)[0] for r in rep] == [list(r)[0] for r in rep_ref]
    for a, b in zip(rep, rep_ref):
        va, vb = list(a.values())[0], list(b.values())[0]
        assert abs(va - vb) <= 1e-4 * max(1.0, abs(vb)), (a, b)
    ref_params = dict(ref.named_parameters())
    worst = ("", 0.0)
    for name, p in ours.named_parameters():
        gr = ref_params[name].grad
        if gr is None:
            assert p.grad is None, f"{name}: the reference leaves no gradient here"
            continue
        assert p.grad is not None, f"{name}: missing gradient"
        assert torch.isfinite(p.grad).all(), name
        err = float((p.grad.cpu() - gr).abs().max())
        scale = float(gr.abs().max())
        if err / (scale + 1e-12) > worst[1]:
            worst = (name, err / (scale + 1e-12))
        assert err <= 1e-2 * scale + 1e-6, f"{name}: grad max-abs err {err:.3e} vs scale {scale:.3e}"
    print("worst relative gradient error:", worst)
    for (n1, b1), (n2, b2) in zip(ours.named_buffers(), ref.named_buffers()):
        if "running" in n1 or "num_batches" in n1:
            assert n1 == n2 and torch.allclose(b1.cpu().double(), b2.double(), rtol=1e-4, atol=1e-6), n1


def test_optimizer_step_through_the_reference_training_recipe(weights):
    """train_fastspeech.py:100-131 in miniature: forward, backward, clip_grad_norm_, Adam step, zero_grad, then an eval
    forward that sees the updated weights (the repack fingerprint follows the optimizer's in-place updates)."""
    m = FeedForwardTransformer(68, 80, load_hp(), precision="fp32")
    m.load_state_dict(weights, strict=True)
    m = m.cuda()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    bt = make_batch(2, 20, 150, seed=21)
    args = [bt[k].cuda() for k in KEYS]
    m.eval()
    with torch.no_grad():
        l0, _ = m(*args)
    m.train()
    losses = []
    for _ in range(3):
        loss, report = m(*args)
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        assert math.isfinite(float(gn))
        opt.step(); opt.zero_grad()
        losses.append(float(loss))
    m.eval()
    with torch.no_grad():
        l1, _ = m(*args)
    assert float(l1) < float(l0), (float(l0), float(l1), losses)
