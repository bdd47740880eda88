"""Shapes of the checkpoint and a seeded synthetic checkpoint.

The reference publishes no checkpoint that is reachable offline (README.md:64 is
a Drive link), so benchmarks and parity tests run on *synthetic* weights.  The
factory below does not construct any module: it walks the checkpoint layout
(`state_dict_spec`) in a fixed order with one seeded generator, so the same
seed yields the same 225 tensors on every machine, for the reference
(`tests/golden/make_golden.py`), for the CPU oracle and for the B200 path.

Values are deliberately *not* a fresh init: LayerNorm/BatchNorm affines and
running statistics are perturbed, both positional-encoding alphas differ from
1 and the duration head bias sits near log(9), so that folding / ordering bugs
show up and `is_inference=True` produces realistic lengths (a fresh init
predicts durations ~0).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class ModelDims:
    """The subset of `hp` that fixes tensor shapes (configs/default.yaml:38-106)."""
    idim: int = 68
    odim: int = 80
    adim: int = 256
    aheads: int = 2
    elayers: int = 4
    eunits: int = 1024
    ddim: int = 384
    dlayers: int = 4
    dunits: int = 1024
    ffn_kernel: int = 9
    pred_layers: int = 2      # Energy/PitchPredictor ignore hp and use 2x256xk3 (variance_predictor.py:125,198)
    pred_chans: int = 256
    pred_kernel: int = 3
    postnet_layers: int = 5
    postnet_chans: int = 256
    postnet_filts: int = 5
    n_bins: int = 256
    pe_len: int = 5000
    e_min: float = 0.01786651276051998
    e_max: float = 130.5338592529297
    p_min: float = 71.0
    p_max: float = 676.2260946528305

    def as_dict(self):
        return asdict(self)


def positional_table(n_pos: int, d_model: int) -> torch.Tensor:
    """fp32 sinusoid table [1, n_pos, d_model] (formula of core/embedding.py:57-64)."""
    pos = torch.arange(n_pos, dtype=torch.float32)[:, None]
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    tab = torch.empty(n_pos, d_model)
    tab[:, 0::2] = torch.sin(pos * freq)
    tab[:, 1::2] = torch.cos(pos * freq)
    return tab[None]


def variance_bins(dims: ModelDims) -> Tuple[torch.Tensor, torch.Tensor]:
    """energy: linspace(e_min,e_max,255) (variance_predictor.py:124);
    pitch: exp(linspace(log p_min, log p_max, 255)) (:188-197)."""
    e = torch.linspace(dims.e_min, dims.e_max, dims.n_bins - 1)
    p = torch.exp(torch.linspace(torch.log(torch.tensor(dims.p_min)), torch.log(torch.tensor(dims.p_max)), dims.n_bins - 1))
    return e, p


def state_dict_spec(dims: ModelDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (key, shape, kind) list of the reference checkpoint (225 entries with
    default dims).  kind in {w, b, ln_w, ln_b, bn_w, bn_b, bn_mean, bn_var, bn_count,
    alpha, pe, e_bins, p_bins, emb, dur_b}."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def stack(prefix: str, C: int, H: int, n: int):
        for i in range(n):
            p = f"{prefix}.encoders_.{i}."
            for nm in ("q", "k", "v", "out"):
                out.append((p + f"self_attn.linear_{nm}.weight", (C, C), "w"))
                out.append((p + f"self_attn.linear_{nm}.bias", (C,), "b"))
            out.append((p + "feed_forward.w_1.weight", (H, C, dims.ffn_kernel), "w"))
            out.append((p + "feed_forward.w_1.bias", (H,), "b"))
            out.append((p + "feed_forward.w_2.weight", (C, H, 1), "w"))
            out.append((p + "feed_forward.w_2.bias", (C,), "b"))
            for nm in ("norm1", "norm2"):
                out.append((p + nm + ".weight", (C,), "ln_w"))
                out.append((p + nm + ".bias", (C,), "ln_b"))
            out.append((p + "concat_linear.weight", (C, 2 * C), "w"))
            out.append((p + "concat_linear.bias", (C,), "b"))

    def pred(prefix: str, head_kind: str):
        for i in range(dims.pred_layers):
            cin = dims.adim if i == 0 else dims.pred_chans
            out.append((f"{prefix}conv.{i}.0.weight", (dims.pred_chans, cin, dims.pred_kernel), "w"))
            out.append((f"{prefix}conv.{i}.0.bias", (dims.pred_chans,), "b"))
            out.append((f"{prefix}conv.{i}.2.layer_norm.weight", (dims.pred_chans,), "ln_w"))
            out.append((f"{prefix}conv.{i}.2.layer_norm.bias", (dims.pred_chans,), "ln_b"))
        out.append((prefix + "linear.weight", (1, dims.pred_chans), "w"))
        out.append((prefix + "linear.bias", (1,), head_kind))

    A, D = dims.adim, dims.ddim
    out.append(("encoder.after_norm.weight", (A,), "ln_w"))
    out.append(("encoder.after_norm.bias", (A,), "ln_b"))
    out.append(("encoder.embed.0.weight", (dims.idim, A), "emb"))
    out.append(("encoder.embed.1.alpha", (), "alpha"))
    out.append(("encoder.embed.1.pe", (1, dims.pe_len, A), "pe"))
    stack("encoder", A, dims.eunits, dims.elayers)
    pred("duration_predictor.", "dur_b")
    out.append(("energy_predictor.energy_bins", (dims.n_bins - 1,), "e_bins"))
    pred("energy_predictor.predictor.", "b")
    out.append(("energy_embed.weight", (A, dims.n_bins), "w"))
    out.append(("energy_embed.bias", (A,), "b"))
    out.append(("pitch_predictor.pitch_bins", (dims.n_bins - 1,), "p_bins"))
    pred("pitch_predictor.predictor.", "b")
    out.append(("pitch_embed.weight", (A, dims.n_bins), "w"))
    out.append(("pitch_embed.bias", (A,), "b"))
    out.append(("decoder.after_norm.weight", (D,), "ln_w"))
    out.append(("decoder.after_norm.bias", (D,), "ln_b"))
    out.append(("decoder.embed.0.weight", (D, A), "w"))
    out.append(("decoder.embed.0.bias", (D,), "b"))
    out.append(("decoder.embed.1.weight", (D,), "ln_w"))
    out.append(("decoder.embed.1.bias", (D,), "ln_b"))
    out.append(("decoder.embed.4.alpha", (), "alpha"))
    out.append(("decoder.embed.4.pe", (1, dims.pe_len, D), "pe"))
    stack("decoder", D, dims.dunits, dims.dlayers)
    for i in range(dims.postnet_layers):
        cin = dims.odim if i == 0 else dims.postnet_chans
        cout = dims.odim if i == dims.postnet_layers - 1 else dims.postnet_chans
        p = f"postnet.postnet.{i}."
        out.append((p + "0.weight", (cout, cin, dims.postnet_filts), "w"))
        out.append((p + "1.weight", (cout,), "bn_w"))
        out.append((p + "1.bias", (cout,), "bn_b"))
        out.append((p + "1.running_mean", (cout,), "bn_mean"))
        out.append((p + "1.running_var", (cout,), "bn_var"))
        out.append((p + "1.num_batches_tracked", (), "bn_count"))
    out.append(("feat_out.weight", (dims.odim, D), "w"))
    out.append(("feat_out.bias", (dims.odim,), "b"))
    return out


def synthetic_state_dict(seed: int = 0, dims: ModelDims = ModelDims()) -> Dict[str, torch.Tensor]:
    """Seeded CPU fp32 checkpoint with the reference's keys (see module docstring)."""
    g = torch.Generator().manual_seed(int(seed))
    e_bins, p_bins = variance_bins(dims)
    sd: Dict[str, torch.Tensor] = {}

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound

    for key, shape, kind in state_dict_spec(dims):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = uniform(shape, 1.0 / math.sqrt(fan_in))
        elif kind == "b":
            t = uniform(shape, 0.05)
        elif kind == "dur_b":
            t = torch.full(shape, math.log(9.0)) + uniform(shape, 0.05)
        elif kind == "emb":
            t = torch.randn(shape, generator=g)
            t[0].zero_()  # padding_idx = 0 (fastspeech.py:57-67)
        elif kind in ("ln_w", "bn_w"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind in ("ln_b", "bn_b"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "bn_count":
            t = torch.tensor(100, dtype=torch.int64)
        elif kind == "alpha":
            t = torch.tensor(1.0) + uniform((), 0.25)
        elif kind == "pe":
            t = positional_table(shape[1], shape[2])
        elif kind == "e_bins":
            t = e_bins.clone()
        elif kind == "p_bins":
            t = p_bins.clone()
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[key] = t.contiguous()
    return sd
