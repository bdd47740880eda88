"""`FeedForwardTransformer` -- drop-in for the class of the same name in the reference's
fastspeech.py (:28-387), with the mel-synthesis forward path running on libfs2b200.so.

Surface kept from the reference:
  __init__(idim, odim, hp)          hp = the reference's HParam Dotdict (or any mapping)
  forward(xs, ilens, ys, olens, ds, es, ps) -> (loss, report_keys)     fastspeech.py:245-337
  inference(x) -> [L, odim]                                            fastspeech.py:339-357
  _forward(xs, ilens, olens=None, ds=None, es=None, ps=None, is_inference=False) -> 5-tuple
  nn.Module behaviour: parameters()/state_dict()/load_state_dict() with the reference's 225
  keys, .to(), .eval()/.train().

The sub-modules below are *parameter holders* laid out so that `state_dict()` has the
reference's keys in the reference's order; they carry no PyTorch compute graph.  All
arithmetic happens in hand-written sm_100a kernels behind the C ABI (include/fs2_b200.h).
There is no CPU path and no PyTorch fallback: CPU tensors or a missing library raise.
`forward()` in train mode (`model.train()`, train_fastspeech.py:100-123) runs the train path of
fastspeech2_b200/train.py: dropout, BatchNorm batch statistics and a backward through every stage,
all on the library's kernels with torch.autograd as the graph only; `_forward` / `inference` in train
mode raise (the reference's scripts call those under `model.eval()`).

Precision (`precision=` / FS2_PRECISION): "3xf16" (default; alias "3xtf32") is the reference-precision
mode -- every contraction, attention included, error-compensated on the tensor cores, fp32-class
results; "f16" and "tf32" are the 10-bit-mantissa fast modes for the decoder side; "fp32" is exact
fp32 FMA on CUDA cores (include/fs2_b200.h, FS2_MATH_*).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from . import length_regulator as _lr
from .weights import ModelDims, positional_table, variance_bins

DEFAULT_PRECISION = os.environ.get("FS2_PRECISION", "3xf16")


def _get(node: Any, key: str, default: Any = None) -> Any:
    if isinstance(node, dict):
        return node.get(key, default)
    return getattr(node, key, default)


def dims_from_hp(idim: int, odim: int, hp: Any) -> ModelDims:
    """Read the shape-defining subset of `hp` (fastspeech.py:53-160) and reject what the kernels
    do not implement, loudly."""
    m, d = _get(hp, "model"), _get(hp, "data")
    if m is None or d is None:
        raise ValueError("hp must provide .model and .data (utils/hparams.py HParam)")

    def need(cond: bool, what: str):
        if not cond:
            raise NotImplementedError(f"fastspeech2_b200 supports the configs/default.yaml architecture only: {what}")

    need(_get(m, "positionwise_layer_type", "conv1d") == "conv1d", "positionwise_layer_type must be 'conv1d'")
    need(not _get(m, "encoder_normalize_before", False) and not _get(m, "decoder_normalize_before", False), "post-LN blocks only")
    need(not _get(m, "encoder_concat_after", False) and not _get(m, "decoder_concat_after", False), "concat_after=False only")
    need(bool(_get(m, "use_scaled_pos_enc", True)), "use_scaled_pos_enc=True only")
    need(bool(_get(m, "use_batch_norm", True)), "use_batch_norm=True only")
    need(int(_get(m, "reduction_factor", 1)) == 1, "reduction_factor=1 only")
    need(int(_get(m, "postnet_layers", 5)) >= 1, "postnet_layers >= 1")
    # Energy/PitchPredictor ignore hp and are always VariancePredictor(idim) = 2 x 256 x k3
    # (variance_predictor.py:125,198); the kernels share one predictor shape, so the duration predictor must match it
    need((int(_get(m, "duration_predictor_layers")), int(_get(m, "duration_predictor_chans")), int(_get(m, "duration_predictor_kernel_size"))) == (2, 256, 3),
         "duration_predictor_{layers,chans,kernel_size} must be (2, 256, 3), the fixed shape of the energy / pitch predictors")
    return ModelDims(
        idim=int(idim), odim=int(odim), adim=int(_get(m, "adim")),
        aheads=int(_get(m, "aheads")), elayers=int(_get(m, "elayers")), eunits=int(_get(m, "eunits")),
        ddim=int(_get(m, "ddim")), dlayers=int(_get(m, "dlayers")), dunits=int(_get(m, "dunits")),
        ffn_kernel=int(_get(m, "positionwise_conv_kernel_size")),
        # DurationPredictor honours hp; Energy/PitchPredictor hard-code VariancePredictor(idim)
        # defaults 2 x 256 x k3 (variance_predictor.py:125,198).  One shape for all three is
        # what default.yaml yields; anything else was rejected above.
        pred_layers=int(_get(m, "duration_predictor_layers")), pred_chans=int(_get(m, "duration_predictor_chans")),
        pred_kernel=int(_get(m, "duration_predictor_kernel_size")),
        postnet_layers=int(_get(m, "postnet_layers")), postnet_chans=int(_get(m, "postnet_chans")),
        postnet_filts=int(_get(m, "postnet_filts")),
        e_min=float(_get(d, "e_min")), e_max=float(_get(d, "e_max")), p_min=float(_get(d, "p_min")), p_max=float(_get(d, "p_max")),
    )


# ------------------------------------------------------------------------------------------------
# parameter holders (checkpoint layout only)
# ------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the forward path runs in libfs2b200.so, not in torch modules")


class _ScaledPosEnc(_Holder):
    def __init__(self, d_model: int, max_len: int):
        super().__init__()
        self.alpha = nn.Parameter(torch.tensor(1.0))
        self.register_buffer("pe", positional_table(max_len, d_model))


class _SelfAttn(_Holder):
    def __init__(self, C_: int):
        super().__init__()
        self.linear_q, self.linear_k = nn.Linear(C_, C_), nn.Linear(C_, C_)
        self.linear_v, self.linear_out = nn.Linear(C_, C_), nn.Linear(C_, C_)


class _ConvFFN(_Holder):
    def __init__(self, C_: int, H: int, k: int):
        super().__init__()
        self.w_1 = nn.Conv1d(C_, H, k, padding=(k - 1) // 2)
        self.w_2 = nn.Conv1d(H, C_, 1)


class _FFTBlock(_Holder):
    def __init__(self, C_: int, H: int, k: int):
        super().__init__()
        self.self_attn = _SelfAttn(C_)
        self.feed_forward = _ConvFFN(C_, H, k)
        self.norm1, self.norm2 = nn.LayerNorm(C_), nn.LayerNorm(C_)
        self.concat_linear = nn.Linear(2 * C_, C_)  # present in the checkpoint, unused (concat_after=False)


class _FFTStack(_Holder):
    def __init__(self, embed: nn.Sequential, C_: int, H: int, k: int, n: int):
        super().__init__()
        self.after_norm = nn.LayerNorm(C_)  # present in the checkpoint, unused (normalize_before=False)
        self.embed = embed
        self.encoders_ = nn.ModuleList([_FFTBlock(C_, H, k) for _ in range(n)])


class _ChannelNorm(_Holder):
    def __init__(self, n: int):
        super().__init__()
        self.layer_norm = nn.LayerNorm(n, eps=1e-12)


class _ConvPredictor(_Holder):
    def __init__(self, cin: int, layers: int, chans: int, k: int):
        super().__init__()
        self.conv = nn.ModuleList([
            nn.Sequential(nn.Conv1d(cin if i == 0 else chans, chans, k, padding=(k - 1) // 2), nn.ReLU(), _ChannelNorm(chans),
                          nn.Dropout(0.5)) for i in range(layers)])
        self.linear = nn.Linear(chans, 1)


class _BinnedPredictor(_Holder):
    """Energy / Pitch predictor: bins buffer + conv predictor (variance_predictor.py:98-232)."""

    def __init__(self, bins_name: str, bins: torch.Tensor, dims: ModelDims):
        super().__init__()
        self.register_buffer(bins_name, bins)
        self.predictor = _ConvPredictor(dims.adim, dims.pred_layers, dims.pred_chans, dims.pred_kernel)
        self._bins_name = bins_name
        self._n_bins = dims.n_bins

    def to_one_hot(self, x: torch.Tensor) -> torch.Tensor:
        """bucketize + one_hot(256).float() (variance_predictor.py:154-159 / :227-232) on the GPU kernels."""
        lib = _lib.load()
        bins = getattr(self, self._bins_name)
        xs = x.contiguous().float()
        ids = torch.empty(xs.shape, dtype=torch.int64, device=xs.device)
        st = _lib.stream_ptr(xs.device)
        _lib.check(lib.fs2_bucketize(_lib.ptr(xs), _lib.ptr(bins), bins.numel(), xs.numel(), _lib.ptr(ids), st), "fs2_bucketize")
        out = torch.empty(tuple(xs.shape) + (self._n_bins,), dtype=torch.float32, device=xs.device)
        _lib.check(lib.fs2_one_hot(_lib.ptr(ids), ids.numel(), self._n_bins, _lib.ptr(out), st), "fs2_one_hot")
        return out


class _Postnet(_Holder):
    def __init__(self, dims: ModelDims):
        super().__init__()
        layers = []
        for i in range(dims.postnet_layers):
            cin = dims.odim if i == 0 else dims.postnet_chans
            cout = dims.odim if i == dims.postnet_layers - 1 else dims.postnet_chans
            mods: List[nn.Module] = [nn.Conv1d(cin, cout, dims.postnet_filts, padding=(dims.postnet_filts - 1) // 2, bias=False),
                                     nn.BatchNorm1d(cout)]
            if i < dims.postnet_layers - 1:
                mods.append(nn.Tanh())
            mods.append(nn.Dropout(0.5))
            layers.append(nn.Sequential(*mods))
        self.postnet = nn.ModuleList(layers)


def _init_like_reference(model: nn.Module, init_type: str) -> None:
    """core/modules.py:51-81 `initialize`."""
    if init_type == "pytorch":
        return
    fns = {"xavier_uniform": nn.init.xavier_uniform_, "xavier_normal": nn.init.xavier_normal_,
           "kaiming_uniform": lambda p: nn.init.kaiming_uniform_(p, nonlinearity="relu"),
           "kaiming_normal": lambda p: nn.init.kaiming_normal_(p, nonlinearity="relu")}
    if init_type not in fns:
        raise ValueError("Unknown initialization: " + init_type)
    for p in model.parameters():
        if p.dim() > 1:
            fns[init_type](p.data)
    for p in model.parameters():
        if p.dim() == 1:
            p.data.zero_()
    for m in model.modules():
        if isinstance(m, (nn.Embedding, nn.LayerNorm)):
            m.reset_parameters()


# ------------------------------------------------------------------------------------------------
class FeedForwardTransformer(nn.Module):
    """Feed-forward Transformer TTS (FastSpeech2) on B200.  See module docstring."""

    @classmethod
    def from_checkpoint(cls, checkpoint, hp=None, precision: Optional[str] = None, device=None):
        """Build the model from a reference checkpoint: a path or the loaded dict `{"model": state_dict, "optim": ..,
        "step": .., "hp_str": .., "githash": ..}` that train_fastspeech.py:235-244 writes (or a bare state_dict, the
        `--old_model` case of inference.py:161-163).  `hp` defaults to the checkpoint's own `hp_str` like
        inference.py:148-152; idim is read off the embedding table instead of the text front-end's symbol list."""
        from .hparams import load_hp_str
        if isinstance(checkpoint, (str, bytes, os.PathLike)):
            checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=True)
        sd = checkpoint["model"] if "model" in checkpoint else checkpoint
        if hp is None:
            if "hp_str" not in checkpoint:
                raise ValueError("checkpoint carries no hp_str: pass hp=")
            hp = load_hp_str(checkpoint["hp_str"])
        idim = int(sd["encoder.embed.0.weight"].shape[0])
        odim = int(_get(_get(hp, "audio"), "num_mels"))
        model = cls(idim, odim, hp, precision=precision)
        model.load_state_dict(sd, strict="model" in checkpoint)
        model.eval()
        return model.to(device) if device is not None else model

    def __init__(self, idim: int, odim: int, hp: Dict, precision: Optional[str] = None):
        super().__init__()
        dims = dims_from_hp(idim, odim, hp)
        self.dims = dims
        self.idim, self.odim = idim, odim
        m = _get(hp, "model")
        self.use_scaled_pos_enc = bool(_get(m, "use_scaled_pos_enc", True))
        self.use_masking = bool(_get(m, "use_masking", True))
        self.use_weighted_masking = bool(_get(m, "use_weighted_masking", False))
        if not self.use_masking or self.use_weighted_masking:
            raise NotImplementedError("loss kernels implement use_masking=True, use_weighted_masking=False (configs/default.yaml:57-58)")
        self.precision = precision or DEFAULT_PRECISION
        if self.precision not in _lib.MATH_MODES:
            raise ValueError(f"precision must be one of {sorted(_lib.MATH_MODES)}")

        A, D = dims.adim, dims.ddim
        self.encoder = _FFTStack(nn.Sequential(nn.Embedding(idim, A, padding_idx=0), _ScaledPosEnc(A, dims.pe_len)),
                                 A, dims.eunits, dims.ffn_kernel, dims.elayers)
        self.duration_predictor = _ConvPredictor(A, dims.pred_layers, dims.pred_chans, dims.pred_kernel)
        e_bins, p_bins = variance_bins(dims)
        self.energy_predictor = _BinnedPredictor("energy_bins", e_bins, dims)
        self.energy_embed = nn.Linear(dims.n_bins, A)
        self.pitch_predictor = _BinnedPredictor("pitch_bins", p_bins, dims)
        self.pitch_embed = nn.Linear(dims.n_bins, A)
        self.length_regulator = _lr.LengthRegulator()
        self.decoder = _FFTStack(nn.Sequential(nn.Linear(A, D), nn.LayerNorm(D), nn.Dropout(0.2), nn.ReLU(), _ScaledPosEnc(D, dims.pe_len)),
                                 D, dims.dunits, dims.ffn_kernel, dims.dlayers)
        self.postnet = _Postnet(dims)
        self.feat_out = nn.Linear(D, odim)

        _init_like_reference(self, str(_get(m, "transformer_init", "pytorch")))
        self.encoder.embed[-1].alpha.data = torch.tensor(float(_get(m, "initial_encoder_alpha", 1.0)))
        self.decoder.embed[-1].alpha.data = torch.tensor(float(_get(m, "initial_decoder_alpha", 1.0)))

        self.postnet_dropout_rate = float(_get(m, "postnet_dropout_rate", 0.5))
        self.duration_dropout_rate = float(_get(m, "duration_predictor_dropout_rate", 0.1))
        self.dropout_masks = None           # train.MaskSource override (tests inject the reference's masks)
        self._epoch = 0
        self._sd_cache: Optional[list] = None
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[torch.device] = None
        self._fingerprint: Optional[Tuple] = None
        self._workspace: Optional[torch.Tensor] = None
        self._keepalive: List[Any] = []

    # -- library plumbing --------------------------------------------------------------------
    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().fs2_destroy(self._handle)
        except Exception:
            pass

    def _device(self) -> torch.device:
        return self.feat_out.weight.device

    def _current_fingerprint(self) -> Tuple:
        """(data_ptr, _version) of every checkpoint tensor plus the explicit epoch.  In-place writes through `.data`
        (`p.data.copy_()`, EMA, the reference's own `initialize()`) do not bump `_version`: call `invalidate()` after
        such an update (`load_state_dict` and `.to()` / `_apply` do it themselves)."""
        if self._sd_cache is None:
            self._sd_cache = list(self.state_dict(keep_vars=True).values())
        return (self._epoch,) + tuple((t.data_ptr(), t._version) for t in self._sd_cache)

    def invalidate(self) -> None:
        """Force a repack of the checkpoint into the kernel layout on the next forward (after weight updates the
        version counters cannot see).  Captured `GraphedForward` objects refuse to replay afterwards."""
        self._epoch += 1
        self._sd_cache = None

    repack = invalidate

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.invalidate()
        return out

    def _extend_pe(self, stack: "_FFTStack", n: int) -> None:
        """core/embedding.py:48-66 `extend_pe`: regenerate a longer sinusoid table when the input outgrows it."""
        pos = stack.embed[-1]
        if pos.pe.shape[1] >= n:
            return
        pos.pe = positional_table(n, pos.pe.shape[2]).to(device=pos.pe.device, dtype=pos.pe.dtype)
        self.invalidate()

    def _ready(self, like: torch.Tensor) -> C.c_void_p:
        """Handle with weights packed for the current parameters (repacks after load_state_dict,
        .to(), optimizer steps ...)."""
        if self.training:
            raise NotImplementedError(
                "_forward / inference run the eval-mode path; call model.eval() (the reference's scripts do: inference.py:115, "
                "evaluation.py:19, train_fastspeech.py:152).  In train mode use model(xs, ilens, ys, olens, ds, es, ps), which "
                "runs the train path (dropout, BatchNorm batch statistics, backward).")
        dev = self._device()
        if dev.type != "cuda":
            raise _lib.Fs2Error("model parameters are on %s: the B200 path has no CPU fallback, call model.to('cuda')" % dev)
        if like.device != dev:
            raise _lib.Fs2Error(f"input on {like.device} but model on {dev}")
        lib = _lib.load()
        if self._handle is None or self._handle_device != dev:
            if self._handle is not None:
                lib.fs2_destroy(self._handle)
            d = self.dims
            cfg = _lib.Config(d.idim, d.odim, d.adim, d.aheads, d.elayers, d.eunits, d.ddim, d.dlayers, d.dunits, d.ffn_kernel,
                              d.pred_layers, d.pred_chans, d.pred_kernel, d.postnet_layers, d.postnet_chans, d.postnet_filts,
                              d.n_bins, d.pe_len, _lib.MATH_MODES[self.precision])
            h = C.c_void_p()
            _lib.check(lib.fs2_create(C.byref(h), C.byref(cfg), dev.index if dev.index is not None else torch.cuda.current_device()), "fs2_create")
            self._handle, self._handle_device, self._fingerprint = h, dev, None
        fp = self._current_fingerprint()
        if fp != self._fingerprint:
            sd = self.state_dict(keep_vars=True)
            descs = (_lib.WeightDesc * len(sd))()
            keep = []
            for i, (k, t) in enumerate(sd.items()):
                t = t.detach()
                if t.dtype not in (torch.float32, torch.int64):
                    raise _lib.Fs2Error(f"parameter {k} is {t.dtype}: the path is fp32-only like the reference (variance_predictor.py:159)")
                t = t.contiguous()
                keep.append(t)
                name = k.encode()
                keep.append(name)
                shape = (C.c_int64 * 4)(*([int(s) for s in t.shape] + [0] * (4 - t.dim())))
                descs[i] = _lib.WeightDesc(name, t.data_ptr(), t.dim(), shape, 0 if t.dtype == torch.float32 else 1)
            _lib.check(lib.fs2_load_weights(self._handle, descs, len(sd), _lib.stream_ptr(dev)), "fs2_load_weights")
            self._fingerprint = fp
        _lib.check(lib.fs2_set_math_mode(self._handle, _lib.MATH_MODES[self.precision]), "fs2_set_math_mode")
        return self._handle

    def _ws(self, B: int, T: int, L: int) -> torch.Tensor:
        lib = _lib.load()
        need = C.c_size_t()
        _lib.check(lib.fs2_workspace_bytes(self._handle, B, T, L, C.byref(need)), "fs2_workspace_bytes")
        dev = self._device()
        if self._workspace is None or self._workspace.device != dev or self._workspace.numel() < need.value:
            self._workspace = None
            self._workspace = torch.empty(int(need.value), dtype=torch.uint8, device=dev)
        return self._workspace

    # -- the path ----------------------------------------------------------------------------
    def _forward(self, xs: torch.Tensor, ilens: torch.Tensor, olens: torch.Tensor = None, ds: torch.Tensor = None,
                 es: torch.Tensor = None, ps: torch.Tensor = None, is_inference: bool = False,
                 _one_hot: bool = True, _defer_check: Optional[list] = None, _after_out: Optional[torch.Tensor] = None) -> Sequence[torch.Tensor]:
        # the handle-less ABI stages (LengthRegulator, losses, ...) run on the CURRENT device like any CUDA library call:
        # select the device the data lives on for the whole call, leave the caller's current device untouched
        if xs.is_cuda and torch.cuda.current_device() != (xs.device.index or 0):
            with torch.cuda.device(xs.device):
                return self._forward(xs, ilens, olens, ds, es, ps, is_inference, _one_hot, _defer_check, _after_out)
        h = self._ready(xs)
        lib = _lib.load()
        dev, d = xs.device, self.dims
        st = _lib.stream_ptr(dev)
        if xs.dim() != 2:
            raise ValueError("xs must be [B, Tmax]")
        B, T = xs.shape
        if T > self.encoder.embed[-1].pe.shape[1] or (es is not None and es.dim() == 2 and es.shape[1] > self.decoder.embed[-1].pe.shape[1]):
            self._extend_pe(self.encoder, T)
            if es is not None and es.dim() == 2:
                self._extend_pe(self.decoder, int(es.shape[1]))
            h = self._ready(xs)
        xs = xs.to(torch.int64).contiguous()
        ilens = ilens.to(device=dev, dtype=torch.int64).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)

        if is_inference:
            L_known = None
        else:
            if olens is None or ds is None or es is None or ps is None:
                raise ValueError("teacher-forced _forward needs olens, ds, es and ps (fastspeech.py:197-216)")
            L_known = int(es.shape[1])
            if ps.shape != es.shape or es.shape[0] != B:
                raise ValueError(f"es {tuple(es.shape)} / ps {tuple(ps.shape)} must both be [B, Lmax]")

        # stage 1: encoder + duration predictor
        ws = self._ws(B, T, L_known or 0)
        hs = torch.empty((B, T, d.adim), **f32)
        d_log = None if is_inference else torch.empty((B, T), **f32)
        d_int = torch.empty((B, T), dtype=torch.int64, device=dev) if is_inference else None
        _lib.check(lib.fs2_encode(h, _lib.ptr(xs), _lib.ptr(ilens), B, T, _lib.ptr(hs), _lib.ptr(d_log), _lib.ptr(d_int),
                                  _lib.ptr(ws), ws.numel(), st), "fs2_encode")

        # stage 2: length regulator
        cum, olens_lr, stats, _ = _lr.plan(hs, d_int if is_inference else ds, ilens, 1.0)
        if is_inference:
            lmax, n_neg = stats.tolist()  # the single host sync of inference: sizes the mel buffers
            L = int(lmax)
            if L <= 0:
                raise RuntimeError("inference produced zero frames")
            if L > self.decoder.embed[-1].pe.shape[1]:
                self._extend_pe(self.decoder, L)
                h = self._ready(xs)
            olens_dec = None  # decoder unmasked, fastspeech.py:221-224
        else:
            L = L_known
            olens_dec = olens.to(device=dev, dtype=torch.int64).contiguous()
        hm = _lr.gather(hs, cum, ilens, L)

        # stage 3: variance adaptor + decoder + postnet
        ws = self._ws(B, T, L)
        before = torch.empty((B, L, d.odim), **f32)
        if _after_out is not None:      # caller-provided destination of the final mels, e.g. this rank's slot of the root rank's
            after = _after_out          # receive buffer mapped over NVLink (sharded.PeerGather): the last Postnet epilogue stores there
            if tuple(after.shape) != (B, L, d.odim) or after.dtype != torch.float32 or not after.is_contiguous():
                raise ValueError(f"_after_out must be a contiguous float32 [{B}, {L}, {d.odim}] tensor")
        else:
            after = torch.empty((B, L, d.odim), **f32)
        e_out, p_out = torch.empty((B, L), **f32), torch.empty((B, L), **f32)
        want_ids = is_inference and _one_hot
        e_ids = torch.empty((B, L), dtype=torch.int64, device=dev) if want_ids else None
        p_ids = torch.empty((B, L), dtype=torch.int64, device=dev) if want_ids else None
        es_c = None if is_inference else es.to(**f32).contiguous()
        ps_c = None if is_inference else ps.to(**f32).contiguous()
        _lib.check(lib.fs2_decode(h, _lib.ptr(hm), _lib.ptr(olens_dec), _lib.ptr(es_c), _lib.ptr(ps_c), B, L, _lib.ptr(before),
                                  _lib.ptr(after), _lib.ptr(e_out), _lib.ptr(p_out), _lib.ptr(e_ids), _lib.ptr(p_ids),
                                  _lib.ptr(ws), ws.numel(), st), "fs2_decode")

        if is_inference:
            if not _one_hot:
                return before, after, d_int, None, None
            oh_e = torch.empty((B, L, d.n_bins), **f32)
            oh_p = torch.empty((B, L, d.n_bins), **f32)
            _lib.check(lib.fs2_one_hot(_lib.ptr(e_ids), B * L, d.n_bins, _lib.ptr(oh_e), st), "fs2_one_hot")
            _lib.check(lib.fs2_one_hot(_lib.ptr(p_ids), B * L, d.n_bins, _lib.ptr(oh_p), st), "fs2_one_hot")
            return before, after, d_int, oh_e, oh_p

        # teacher-forced: validate what the reference would have tripped over with shape errors
        # (mask widths are max(lengths): utils/util.py:262-272), one host read at the end.
        chk_dev = torch.stack([stats[0], stats[1], ilens.max(), olens_dec.max()])
        if _defer_check is not None:          # CUDA-graph capture: no host read here, the caller validates after replay
            _defer_check.append((chk_dev, T, L))
            return before, after, d_log, e_out, p_out
        self._validate_lengths(chk_dev.tolist(), T, L)
        return before, after, d_log, e_out, p_out

    @staticmethod
    def _validate_lengths(chk, T: int, L: int) -> None:
        if chk[1]:
            raise RuntimeError(f"LengthRegulator: {chk[1]} negative duration(s)")
        if chk[2] != T:
            raise RuntimeError(f"xs has Tmax={T} but max(ilens)={chk[2]} (the reference's masks are max(ilens) wide)")
        if chk[0] != L or chk[3] != L:
            raise RuntimeError(f"length mismatch: es/ps have Lmax={L}, max(sum(ds))={chk[0]}, max(olens)={chk[3]}")

    def graphed_forward(self, xs, ilens, olens, ds, es, ps, after_out: Optional[torch.Tensor] = None) -> "GraphedForward":
        """Capture the teacher-forced `_forward` for these shapes into one CUDA graph (see GraphedForward).  `after_out`:
        where the final mels are written (default: a fresh tensor), e.g. a peer-mapped slot of `sharded.PeerGather`."""
        return GraphedForward(self, xs, ilens, olens, ds, es, ps, after_out=after_out)

    def forward(self, xs: torch.Tensor, ilens: torch.Tensor, ys: torch.Tensor, olens: torch.Tensor, ds: torch.Tensor,
                es: torch.Tensor, ps: torch.Tensor) -> Tuple[torch.Tensor, List[Dict[str, float]]]:
        """Loss computation (fastspeech.py:245-337). Returns (loss, report_keys).  In train mode the loss is attached to
        the autograd graph of the train path (fastspeech2_b200/train.py) so `loss.backward()` fills `.grad` of every
        parameter the reference trains; `self.dropout_masks` (a train.MaskSource) may be set to inject masks."""
        if xs.is_cuda and torch.cuda.current_device() != (xs.device.index or 0):
            with torch.cuda.device(xs.device):
                return self.forward(xs, ilens, ys, olens, ds, es, ps)
        if self.training:
            from .train import train_forward
            return train_forward(self, xs, ilens, ys, olens, ds, es, ps, masks=self.dropout_masks)
        self._ready(xs)
        lib = _lib.load()
        dev = xs.device
        ilens = ilens.to(device=dev, dtype=torch.int64)
        olens = olens.to(device=dev, dtype=torch.int64)
        tmax, lmax = torch.stack([ilens.max(), olens.max()]).tolist()
        xs = xs[:, :tmax]  # fastspeech.py:266-267
        ds_t = ds[:, :tmax].contiguous() if ds.shape[1] != tmax else ds
        es_t, ps_t = es[:, :lmax], ps[:, :lmax]
        before, after, d_outs, e_outs, p_outs = self._forward(xs, ilens, olens, ds_t, es_t, ps_t, is_inference=False)
        if ds_t is not ds and ds_t.shape == ds[:, :tmax].shape:
            ds[:, :tmax].copy_(ds_t)
        B, T = xs.shape
        L = before.shape[1]
        ys_c = ys.to(dtype=torch.float32, device=dev).contiguous()
        es_c, ps_c = es_t.to(torch.float32).contiguous(), ps_t.to(torch.float32).contiguous()
        ds_c = ds_t.contiguous()
        out7 = torch.empty((7,), dtype=torch.float32, device=dev)
        scratch = torch.empty((16,), dtype=torch.float64, device=dev)
        _lib.check(lib.fs2_masked_losses(_lib.ptr(before), _lib.ptr(after), _lib.ptr(ys_c), int(ys_c.shape[1]), _lib.ptr(d_outs),
                                         _lib.ptr(ds_c), _lib.dur_dtype(ds_c), _lib.ptr(e_outs), _lib.ptr(p_outs), _lib.ptr(es_c),
                                         _lib.ptr(ps_c), _lib.ptr(ilens.contiguous()), _lib.ptr(olens.contiguous()), B, T, L,
                                         self.odim, _lib.ptr(out7), _lib.ptr(scratch), _lib.stream_ptr(dev)), "fs2_masked_losses")
        vals = out7.tolist()
        names = ["l1_loss", "before_loss", "after_loss", "duration_loss", "energy_loss", "pitch_loss", "loss"]
        return out7[6], [{k: v} for k, v in zip(names, vals)]

    def inference(self, x: torch.Tensor) -> torch.Tensor:
        """x [T] int64 -> mel [L, odim] (fastspeech.py:339-357)."""
        ilens = torch.tensor([x.shape[0]], dtype=torch.long, device=x.device)
        _, outs, _, _, _ = self._forward(x.unsqueeze(0), ilens, is_inference=True, _one_hot=False)
        return outs[0]


class GraphedForward:
    """Teacher-forced `_forward` of one fixed shape replayed as a single CUDA graph.

    The ~90 kernel launches of a step (each with its TMA descriptors baked into the launch parameters) are captured
    once; a call copies the inputs into the graph's static buffers (`self.inputs`; callers may also fill those directly
    and call `replay()`), replays, and returns the static output tensors `(before, after, d_outs, e_outs, p_outs)`
    (valid until the next replay).

    Validation of the length words (what the eager path reads back after every forward):
      validate=True        read them now (one host sync per call) and raise like the eager path;
      validate="deferred"  copy them to pinned host memory asynchronously; the check of replay i runs at the start of
                           call i+1 (or in `flush()`), so replays queue back to back with no host bubble;
      validate=False       no check.
    The graph owns references to the workspace and static buffers it was captured with, so later eager calls that grow
    the model's workspace cannot hand that memory to anyone else.  Weight updates require a new capture (the packed
    weight arena is part of the graph): `__call__` raises if the model was repacked or its parameters changed."""

    def __init__(self, model: FeedForwardTransformer, xs, ilens, olens, ds, es, ps, after_out: Optional[torch.Tensor] = None):
        self.model = model
        self._after_out = after_out
        dev = xs.device
        self.inputs = [t.detach().clone().contiguous() for t in (xs, ilens.to(dev), olens.to(dev), ds, es, ps)]
        with torch.no_grad():
            for _ in range(2):                      # warm-up: packs weights, sizes the workspace, sets kernel attributes
                model._forward(*self.inputs, is_inference=False, _after_out=after_out)
            torch.cuda.synchronize(dev)
            self._fingerprint = model._current_fingerprint()
            self._params = list(model._sd_cache)
            self._versions = sum(t._version for t in self._params)
            self._workspace = model._workspace      # keep the captured scratch alive for the life of the graph
            self.graph = torch.cuda.CUDAGraph()
            deferred: list = []
            with torch.cuda.graph(self.graph):
                self.outputs = model._forward(*self.inputs, is_inference=False, _defer_check=deferred, _after_out=after_out)
            self._chk, self._T, self._L = deferred[0]
        self._chk_host = torch.empty(self._chk.shape, dtype=self._chk.dtype).pin_memory()
        self._chk_event = torch.cuda.Event()
        self._pending = False

    def _check_model(self) -> None:
        m = self.model
        if m._epoch != self._fingerprint[0] or sum(t._version for t in self._params) != self._versions:
            raise RuntimeError("model parameters changed since capture: build a new GraphedForward")

    def flush(self) -> None:
        """Run the outstanding deferred validation (waits for the replay it belongs to)."""
        if self._pending:
            self._pending = False
            self._chk_event.synchronize()
            FeedForwardTransformer._validate_lengths(self._chk_host.tolist(), self._T, self._L)

    def replay(self, validate=True):
        """Replay on the current contents of `self.inputs`."""
        self._check_model()
        self.flush()
        self.graph.replay()
        if validate == "deferred":
            self._chk_host.copy_(self._chk, non_blocking=True)
            self._chk_event.record()
            self._pending = True
        elif validate:
            FeedForwardTransformer._validate_lengths(self._chk.tolist(), self._T, self._L)
        return self.outputs

    def __call__(self, xs, ilens, olens, ds, es, ps, validate=True):
        for dst, src in zip(self.inputs, (xs, ilens, olens, ds, es, ps)):
            if dst.shape != src.shape:
                raise ValueError(f"captured for shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        return self.replay(validate)
