"""Vocoder hand-off on the device: the STFT / inverse-STFT pair and the Griffin-Lim loop the
reference falls back to when no neural vocoder is configured (inference.py:188-193 -> utils/stft.py:41-156,
dataset/audio_processing.py:224-240).

Same class surface as the reference's `STFT` (`transform`, `inverse`, `forward`, buffers `forward_basis` / `inverse_basis`
built the same way: windowed real/imag Fourier rows and their scaled pseudo-inverse), but the arithmetic runs on
libfs2b200.so: the reference's strided `F.conv1d` / `F.conv_transpose1d` are GEMMs over a [frames, n_fft] matrix, issued
through the library's tap-GEMM (fp32-class 3xF16 on tcgen05 by default) with hand-written kernels for reflect-padding +
framing, magnitude / phase, recombination and overlap-add + window-sum normalisation (csrc/stft.cu).  No cuFFT, no torch
compute.  Everything stays on the GPU between the mel batch and the waveform (the reference round-trips through `.cpu()`
every iteration, utils/stft.py:97-103).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _hann(win_length: int, n_fft: int) -> np.ndarray:
    """scipy.signal.get_window("hann", win_length, fftbins=True) zero-padded to n_fft about its centre (librosa pad_center)."""
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad))


def window_sumsquare(n_frames: int, hop_length: int, win_length: int, n_fft: int) -> np.ndarray:
    """dataset/audio_processing.py:169-221 (librosa 0.6): sum of squared windows at every sample, float32."""
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=np.float32)
    win_sq = _hann(win_length, n_fft) ** 2
    for i in range(n_frames):
        s = i * hop_length
        x[s: min(n, s + n_fft)] += win_sq[: max(0, min(n_fft, n - s))]
    return x


class STFT(torch.nn.Module):
    """Drop-in for utils/stft.py:41-156 (hann window only), on the B200 kernels.  `math_mode`: "3xf16" (default,
    fp32-class), "fp32" (CUDA cores) or "f16" / "tf32" for the two GEMMs."""

    def __init__(self, filter_length: int = 800, hop_length: int = 200, win_length: int = 800, window: str = "hann", math_mode: str = "3xf16"):
        super().__init__()
        if window != "hann":
            raise NotImplementedError("only the hann window the reference uses is implemented")
        assert filter_length >= win_length and filter_length % 80 == 0 or filter_length % 128 == 0, "n_fft must be a multiple of 80 or 128 (GEMM tile widths)"
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window
        self.math_mode = _lib.MATH_MODES[math_mode]
        scale = filter_length / hop_length
        fourier = np.fft.fft(np.eye(filter_length))
        self.cutoff = cutoff = filter_length // 2 + 1
        fourier = np.vstack([np.real(fourier[:cutoff, :]), np.imag(fourier[:cutoff, :])])          # [2*cutoff, n_fft]
        win = _hann(win_length, filter_length)
        fwd = torch.FloatTensor(fourier[:, None, :]) * torch.from_numpy(win).float()
        inv = torch.FloatTensor(np.linalg.pinv(scale * fourier).T[:, None, :]) * torch.from_numpy(win).float()
        self.register_buffer("forward_basis", fwd.float())                                         # [2*cutoff, 1, n_fft]
        self.register_buffer("inverse_basis", inv.float())
        # GEMM operands in the library's [taps=1][N][K] layout, N / K padded with zero rows / columns to the next tile multiple
        self.cpad = (2 * cutoff + 63) // 64 * 64
        w_f = torch.zeros(self.cpad, filter_length); w_f[: 2 * cutoff] = fwd[:, 0, :]
        w_i = torch.zeros(filter_length, self.cpad); w_i[:, : 2 * cutoff] = inv[:, 0, :].T
        self.register_buffer("_w_forward", w_f.contiguous(), persistent=False)
        self.register_buffer("_w_inverse", w_i.contiguous(), persistent=False)
        self.register_buffer("_zero_bias_f", torch.zeros(self.cpad), persistent=False)
        self.register_buffer("_zero_bias_i", torch.zeros(filter_length), persistent=False)
        self._wsum = {}

    def _gemm(self, x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
        B, L, K = x.shape
        N = w.shape[0]
        out = torch.empty((B, L, N), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().fs2_op_tap_gemm(self.math_mode, _lib.ptr(x), B, L, K, _lib.ptr(w), _lib.ptr(bias), N, 1, 0, None, _lib.ptr(out),
                                               _lib.stream_ptr(x.device)), "fs2_op_tap_gemm")
        return out

    def transform(self, input_data: torch.Tensor):
        """[B, n] samples -> (magnitude, phase), each [B, n_fft/2+1, frames] (utils/stft.py:82-112)."""
        lib = _lib.load()
        x = input_data.to(dtype=torch.float32).contiguous()
        if not x.is_cuda:
            raise _lib.Fs2Error("STFT.transform: CUDA tensor required (no CPU fallback)")
        B, n = x.shape
        self.num_samples = n
        frames = n // self.hop_length + 1
        st = _lib.stream_ptr(x.device)
        fr = torch.empty((B, frames, self.filter_length), dtype=torch.float32, device=x.device)
        _lib.check(lib.fs2_stft_frames(_lib.ptr(x), B, n, self.filter_length, self.hop_length, frames, _lib.ptr(fr), st), "fs2_stft_frames")
        spec = self._gemm(fr, self._w_forward, self._zero_bias_f)
        mag = torch.empty((B, self.cutoff, frames), dtype=torch.float32, device=x.device)
        phase = torch.empty_like(mag)
        _lib.check(lib.fs2_stft_magphase(_lib.ptr(spec), self.cpad, B, self.cutoff, frames, _lib.ptr(mag), _lib.ptr(phase), st), "fs2_stft_magphase")
        return mag, phase

    def inverse(self, magnitude: torch.Tensor, phase: torch.Tensor) -> torch.Tensor:
        """(magnitude, phase) [B, cutoff, frames] -> [B, 1, (frames-1)*hop] samples (utils/stft.py:114-151)."""
        lib = _lib.load()
        mag, ph = magnitude.to(torch.float32).contiguous(), phase.to(torch.float32).contiguous()
        if not mag.is_cuda:
            raise _lib.Fs2Error("STFT.inverse: CUDA tensors required (no CPU fallback)")
        B, cutoff, frames = mag.shape
        assert cutoff == self.cutoff, f"expected {self.cutoff} frequency rows, got {cutoff}"
        st = _lib.stream_ptr(mag.device)
        rec = torch.empty((B, frames, self.cpad), dtype=torch.float32, device=mag.device)
        _lib.check(lib.fs2_istft_recombine(_lib.ptr(mag), _lib.ptr(ph), B, cutoff, frames, self.cpad, _lib.ptr(rec), st), "fs2_istft_recombine")
        fr = self._gemm(rec, self._w_inverse, self._zero_bias_i)
        key = (frames, str(mag.device))
        if key not in self._wsum:
            self._wsum[key] = torch.from_numpy(window_sumsquare(frames, self.hop_length, self.win_length, self.filter_length)).to(mag.device)
        y = torch.empty((B, 1, (frames - 1) * self.hop_length), dtype=torch.float32, device=mag.device)
        _lib.check(lib.fs2_istft_overlap_add(_lib.ptr(fr), B, self.filter_length, self.hop_length, frames, _lib.ptr(self._wsum[key]),
                                             float(np.finfo(np.float32).tiny), _lib.ptr(y), st), "fs2_istft_overlap_add")
        return y

    def forward(self, input_data: torch.Tensor) -> torch.Tensor:
        self.magnitude, self.phase = self.transform(input_data)
        return self.inverse(self.magnitude, self.phase)


def griffin_lim(magnitudes: torch.Tensor, stft_fn: STFT, n_iters: int = 30, angles: torch.Tensor = None) -> torch.Tensor:
    """dataset/audio_processing.py:224-240: random initial phases, then n_iters x (transform -> keep phase -> inverse).
    `angles` may be given for reproducibility (the reference draws them with np.random)."""
    if angles is None:
        angles = np.angle(np.exp(2j * np.pi * np.random.rand(*magnitudes.size()))).astype(np.float32)
        angles = torch.from_numpy(angles)
    angles = angles.to(magnitudes.device)
    signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
    for _ in range(n_iters):
        _, angles = stft_fn.transform(signal)
        signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
    return signal
