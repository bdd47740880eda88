"""CPU oracle for the FastSpeech2 mel-synthesis forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fastspeech2_b200/`` imports this
file.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may call it, and only as the
checker / the timed CPU baseline -- never as the product path.

What it is: a functional restatement of the reference's eval-mode
``FeedForwardTransformer._forward`` / ``forward`` / ``inference``
(``/root/reference/fastspeech.py:169-357``) driven purely by a ``state_dict``
(name -> CPU fp32 tensor, the reference's own 225 keys).  The reference is
~3.5 k lines of pure Python whose every FLOP executes inside PyTorch's CPU
kernels (MKL / oneDNN), so the faithful CPU restatement
issues the *same ATen calls in the same order* -- that is what makes it
bit-identical to the reference on the same inputs, and what makes its timing
the reference's CPU timing.

Pinning: ``tests/golden/make_golden.py`` imports the *unmodified* reference
from ``/root/reference`` in the build container, runs it on seeded inputs and
weights, and commits inputs+outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this file against those vectors
(bit-exact for integer outputs, <= 1e-6 abs for floats).  The reference's own
test (``tests/test_fastspeech2.py``) pins no numbers, so the golden vectors
generated from the live reference are the pin.

Every function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# masks  (utils/util.py:178-286 make_pad_mask, :294-376 make_non_pad_mask)
# --------------------------------------------------------------------------
def pad_mask(lengths: torch.Tensor) -> torch.Tensor:
    """True at padded positions; width = max(lengths) (util.py:262-272)."""
    lengths = lengths.to(torch.int64).cpu()
    maxlen = int(lengths.max())
    return torch.arange(maxlen).unsqueeze(0) >= lengths.unsqueeze(-1)


def source_mask(lengths: torch.Tensor) -> torch.Tensor:
    """[B,T,T] outer-AND of non-pad flags (fastspeech.py:359-376)."""
    m = ~pad_mask(lengths)
    return m.unsqueeze(-2) & m.unsqueeze(-1)


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def positional_table(n_pos: int, d_model: int) -> torch.Tensor:
    """Sinusoid table [1,n_pos,d] (core/embedding.py:57-65)."""
    pe = torch.zeros(n_pos, d_model)
    position = torch.arange(0, n_pos, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(
        torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model)
    )
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def scaled_pos_enc(x: torch.Tensor, alpha: torch.Tensor, pe: torch.Tensor) -> torch.Tensor:
    """x + alpha * pe[:, :T]  (core/embedding.py:105-120; dropout = identity in eval)."""
    if pe.size(1) < x.size(1):  # extend_pe, embedding.py:48-66
        pe = positional_table(x.size(1), x.size(2))
    return x + alpha * pe[:, : x.size(1)]


def mha(sd: SD, p: str, x: torch.Tensor, mask: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """MultiHeadedAttention.forward (core/attention.py:30-74)."""
    B, T, C = x.shape
    dk = C // heads
    q = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    k = F.linear(x, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    v = F.linear(x, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)  # attention.py:55-57
    if mask is not None:
        m = mask.unsqueeze(1).eq(0)
        scores = scores.masked_fill_(m, -float("inf"))  # :58-62
        attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)  # :63-65 (NaN rows -> 0)
    else:
        attn = torch.softmax(scores, dim=-1)  # :67
    ctx = torch.matmul(attn, v)  # :70
    ctx = ctx.transpose(1, 2).contiguous().view(B, -1, heads * dk)  # :71-73
    return F.linear(ctx, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])  # :74


def conv_ffn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """MultiLayeredConv1d.forward (core/modules.py:237-248), k odd, 'same' zero pad."""
    w1 = sd[p + "w_1.weight"]
    h = torch.relu(F.conv1d(x.transpose(-1, 1), w1, sd[p + "w_1.bias"], padding=(w1.size(2) - 1) // 2)).transpose(-1, 1)
    return F.conv1d(h.transpose(-1, 1), sd[p + "w_2.weight"], sd[p + "w_2.bias"]).transpose(-1, 1)


def fft_block(sd: SD, p: str, x: torch.Tensor, mask: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """EncoderLayer.forward, post-LN branch (core/encoder.py:46-71 with
    normalize_before=False, concat_after=False; LN eps 1e-5 :37-38)."""
    C = x.size(-1)
    x = F.layer_norm(x + mha(sd, p + "self_attn.", x, mask, heads), (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    x = F.layer_norm(x + conv_ffn(sd, p + "feed_forward.", x), (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    return x


def _n_layers(sd: SD, prefix: str) -> int:
    n = 0
    while (prefix + "%d.norm1.weight" % n) in sd:
        n += 1
    return n


def encoder(sd: SD, xs: torch.Tensor, ilens: torch.Tensor, heads: int = 2) -> torch.Tensor:
    """Phoneme encoder: Embedding(pad 0) + scaled PE + FFT blocks
    (fastspeech.py:65-84,184; core/encoder.py:136-143,185-204; after_norm skipped :201)."""
    x = F.embedding(xs, sd["encoder.embed.0.weight"], padding_idx=0)
    x = scaled_pos_enc(x, sd["encoder.embed.1.alpha"], sd["encoder.embed.1.pe"])
    mask = source_mask(ilens)
    for i in range(_n_layers(sd, "encoder.encoders_.")):
        x = fft_block(sd, "encoder.encoders_.%d." % i, x, mask, heads)
    return x


def decoder(sd: SD, hs: torch.Tensor, olens: Optional[torch.Tensor], heads: int = 2) -> torch.Tensor:
    """Mel decoder: Linear->LN->ReLU->scaled PE (core/encoder.py:118-125) + FFT blocks;
    olens None => no mask at all (fastspeech.py:221-226)."""
    C = sd["decoder.embed.0.weight"].size(0)
    x = F.linear(hs, sd["decoder.embed.0.weight"], sd["decoder.embed.0.bias"])
    x = F.layer_norm(x, (C,), sd["decoder.embed.1.weight"], sd["decoder.embed.1.bias"], 1e-5)
    x = torch.relu(x)
    x = scaled_pos_enc(x, sd["decoder.embed.4.alpha"], sd["decoder.embed.4.pe"])
    mask = source_mask(olens) if olens is not None else None
    for i in range(_n_layers(sd, "decoder.encoders_.")):
        x = fft_block(sd, "decoder.encoders_.%d." % i, x, mask, heads)
    return x


def predictor(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """Conv stack shared by Duration/VariancePredictor: n x [Conv1d k3 -> ReLU ->
    channel LayerNorm eps 1e-12] -> Linear(->1)  (duration_predictor.py:45-75,
    variance_predictor.py:20-51, core/modules.py:112-120).  Returns [B,time]."""
    y = x.transpose(1, -1)
    i = 0
    while (p + "conv.%d.0.weight" % i) in sd:
        w = sd[p + "conv.%d.0.weight" % i]
        y = torch.relu(F.conv1d(y, w, sd[p + "conv.%d.0.bias" % i], padding=(w.size(2) - 1) // 2))
        y = F.layer_norm(y.transpose(1, -1), (w.size(0),), sd[p + "conv.%d.2.layer_norm.weight" % i],
                         sd[p + "conv.%d.2.layer_norm.bias" % i], 1e-12).transpose(1, -1)
        i += 1
    return F.linear(y.transpose(1, -1), sd[p + "linear.weight"], sd[p + "linear.bias"]).squeeze(-1)


def durations_from_log(d_log: torch.Tensor, offset: float = 1.0) -> torch.Tensor:
    """clamp(round(exp(x) - offset), min=0).long()  (duration_predictor.py:77-81)."""
    return torch.clamp(torch.round(d_log.exp() - offset), min=0).long()


def length_regulator(xs: torch.Tensor, ds: torch.Tensor, ilens: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """LengthRegulator.forward (core/duration_modeling/length_regulator.py:38-95)
    + pad_2d_tensor (utils/util.py:91-104).  Mutates ``ds`` in place in the all-zero
    case exactly like the reference (``d.fill_(1)`` on a view, :86-88) when alpha == 1.
    Uses repeat_interleave per utterance instead of the reference's per-phoneme
    Python loop (:90-95); int(d_) truncation and the d_ != 0 skip are preserved."""
    assert alpha > 0
    if alpha != 1.0:
        ds = torch.round(ds.float() * alpha).long()  # :58-59
    outs: List[torch.Tensor] = []
    for x, d, ilen in zip(xs, ds, ilens):
        n = int(ilen)
        x, d = x[:n], d[:n]  # :60-61
        if d.sum() == 0:  # :86-88
            d = d.fill_(1)
        reps = d.to(torch.int64) if not d.is_floating_point() else d.trunc().to(torch.int64)  # int(d_), :93
        if bool((reps < 0).any()):
            raise RuntimeError("negative duration")  # x_.repeat(negative) raises in the reference
        outs.append(torch.repeat_interleave(x, reps, dim=0))
    max_len = max(o.size(0) for o in outs)
    return torch.stack([F.pad(o, (0, 0, 0, max_len - o.size(0)), "constant", 0.0) for o in outs])


def bucket_ids(values: torch.Tensor, bins: torch.Tensor) -> torch.Tensor:
    """torch.bucketize(x, bins) (variance_predictor.py:158,231): #{i: bins[i] < x}."""
    return torch.bucketize(values, bins)


def postnet(sd: SD, before: torch.Tensor) -> torch.Tensor:
    """Postnet.forward (core/modules.py:350-359; layers built :283-348): 5 x
    [Conv1d k5 no bias -> BatchNorm1d(eval, eps 1e-5) -> Tanh (not on the last)]."""
    y = before.transpose(1, 2)
    n = 0
    while ("postnet.postnet.%d.0.weight" % n) in sd:
        n += 1
    for i in range(n):
        p = "postnet.postnet.%d." % i
        w = sd[p + "0.weight"]
        y = F.conv1d(y, w, None, padding=(w.size(2) - 1) // 2)
        if (p + "1.running_mean") in sd:
            y = F.batch_norm(y, sd[p + "1.running_mean"], sd[p + "1.running_var"], sd[p + "1.weight"], sd[p + "1.bias"],
                             False, 0.1, 1e-5)
        if i < n - 1:
            y = torch.tanh(y)
    return y.transpose(1, 2)


# --------------------------------------------------------------------------
# the path
# --------------------------------------------------------------------------
def forward_path(
    sd: SD,
    xs: torch.Tensor,
    ilens: torch.Tensor,
    olens: Optional[torch.Tensor] = None,
    ds: Optional[torch.Tensor] = None,
    es: Optional[torch.Tensor] = None,
    ps: Optional[torch.Tensor] = None,
    is_inference: bool = False,
    heads: int = 2,
) -> Tuple[torch.Tensor, ...]:
    """FeedForwardTransformer._forward, eval mode (fastspeech.py:169-243)."""
    hs = encoder(sd, xs, ilens, heads)  # :180-186
    d_masks = pad_mask(ilens)  # :190
    e_bins, p_bins = sd["energy_predictor.energy_bins"], sd["pitch_predictor.pitch_bins"]
    if is_inference:
        d_outs = durations_from_log(predictor(sd, "duration_predictor.", hs)).masked_fill(d_masks, 0)  # :193
        hs = length_regulator(hs, d_outs, ilens)  # :194
        e_val = predictor(sd, "energy_predictor.predictor.", hs)  # :195 (no mask, variance_predictor.py:80-95)
        p_val = predictor(sd, "pitch_predictor.predictor.", hs)  # :196
        e_ids, p_ids = bucket_ids(e_val, e_bins), bucket_ids(p_val, p_bins)
        e_ret = F.one_hot(e_ids.long(), 256).float()
        p_ret = F.one_hot(p_ids.long(), 256).float()
    else:
        e_ids, p_ids = bucket_ids(es, e_bins), bucket_ids(ps, p_bins)  # :200-206
        mel_masks = pad_mask(olens)  # :208
        d_outs = predictor(sd, "duration_predictor.", hs).masked_fill(d_masks, 0.0)  # :210
        hs = length_regulator(hs, ds, ilens)  # :212
        e_ret = predictor(sd, "energy_predictor.predictor.", hs).masked_fill(mel_masks, 0.0)  # :214
        p_ret = predictor(sd, "pitch_predictor.predictor.", hs).masked_fill(mel_masks, 0.0)  # :216
    # one-hot x Linear == column gather + bias (:218-219); keep the reference's GEMM form
    hs = hs + F.linear(F.one_hot(p_ids.long(), 256).float(), sd["pitch_embed.weight"], sd["pitch_embed.bias"])
    hs = hs + F.linear(F.one_hot(e_ids.long(), 256).float(), sd["energy_embed.weight"], sd["energy_embed.bias"])
    zs = decoder(sd, hs, olens, heads)  # :221-226
    odim = sd["feat_out.weight"].size(0)
    before = F.linear(zs, sd["feat_out.weight"], sd["feat_out.bias"]).view(zs.size(0), -1, odim)  # :228-230
    after = before + postnet(sd, before)  # :236-238
    return before, after, d_outs, e_ret, p_ret


def inference(sd: SD, x: torch.Tensor, heads: int = 2) -> torch.Tensor:
    """FeedForwardTransformer.inference (fastspeech.py:339-357)."""
    ilens = torch.tensor([x.shape[0]], dtype=torch.long)
    return forward_path(sd, x.unsqueeze(0), ilens, is_inference=True, heads=heads)[1][0]


def forward_loss(sd: SD, xs, ilens, ys, olens, ds, es, ps, heads: int = 2):
    """FeedForwardTransformer.forward with use_masking=True, use_weighted_masking=False
    (fastspeech.py:245-337).  Returns (loss, report_keys)."""
    xs = xs[:, : int(max(ilens))]  # :266
    ys = ys[:, : int(max(olens))]  # :267
    before, after, d_outs, e_outs, p_outs = forward_path(sd, xs, ilens, olens, ds, es, ps, False, heads)
    in_m = ~pad_mask(ilens)
    mel_m = ~pad_mask(olens)
    out_m = mel_m.unsqueeze(-1)
    d_sel, ds_sel = d_outs.masked_select(in_m), ds.masked_select(in_m)  # :281-283
    b_sel, a_sel, y_sel = before.masked_select(out_m), after.masked_select(out_m), ys.masked_select(out_m)
    es_sel, ps_sel = es.masked_select(mel_m), ps.masked_select(mel_m)
    e_sel, p_sel = e_outs.masked_select(mel_m), p_outs.masked_select(mel_m)
    before_loss = F.l1_loss(b_sel, y_sel)  # :298
    after_loss = F.l1_loss(a_sel, y_sel)  # :301
    l1_loss = before_loss + after_loss
    duration_loss = F.mse_loss(d_sel, torch.log(ds_sel.float() + 1.0))  # duration_predictor.py:148-149
    energy_loss = F.mse_loss(e_sel, es_sel)  # variance_predictor.py:235-275
    pitch_loss = F.mse_loss(p_sel, ps_sel)
    loss = l1_loss + duration_loss + energy_loss + pitch_loss  # :324
    report = [
        {"l1_loss": l1_loss.item()}, {"before_loss": before_loss.item()}, {"after_loss": after_loss.item()},
        {"duration_loss": duration_loss.item()}, {"energy_loss": energy_loss.item()},
        {"pitch_loss": pitch_loss.item()}, {"loss": loss.item()},
    ]
    return loss, report
