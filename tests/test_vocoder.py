"""STFT / inverse STFT / Griffin-Lim on the B200 kernels against the CPU restatement of
utils/stft.py + dataset/audio_processing.py (oracle/stft_oracle.py).  fp32-class GEMMs (3xF16): stated tolerance 2e-4
relative to the signal scale for one transform / inverse; Griffin-Lim iterates 5 times (phase wrap makes single samples
chaotic at atan2's branch cut, so the loop is compared through the reconstruction error it achieves)."""
import numpy as np
import pytest
import torch

from fastspeech2_b200.vocoder import STFT, griffin_lim, window_sumsquare
from oracle import stft_oracle as O


def test_bases_and_window_sum_match_the_restatement():
    ours, ref = STFT(1024, 256, 1024), O.STFT(1024, 256, 1024)
    assert torch.equal(ours.forward_basis, ref.forward_basis) and torch.equal(ours.inverse_basis, ref.inverse_basis)
    assert ours.forward_basis.shape == (1026, 1, 1024)
    w = window_sumsquare(9, 256, 1024, 1024)
    assert w.shape == (1024 + 256 * 8,) and w.dtype == np.float32 and abs(float(w[1024]) - 1.5) < 1e-5


def test_cpu_tensors_fail_loudly():
    from fastspeech2_b200 import _lib
    with pytest.raises(_lib.Fs2Error, match="no CPU fallback"):
        STFT(1024, 256, 1024).transform(torch.zeros(1, 4096))


@pytest.mark.gpu
@pytest.mark.parametrize("n_fft,hop,n", [(1024, 256, 22050), (800, 200, 7777)])
def test_transform_and_inverse_vs_oracle(n_fft, hop, n):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(3, n, generator=g) * 0.3
    ours, ref = STFT(n_fft, hop, n_fft).cuda(), O.STFT(n_fft, hop, n_fft)
    mag, ph = ours.transform(x.cuda())
    mag_r, ph_r = ref.transform(x)
    assert mag.shape == mag_r.shape
    assert float((mag.cpu() - mag_r).abs().max()) <= 2e-4 * float(mag_r.abs().max())
    strong = mag_r > 1e-2 * mag_r.max()                                  # phase is ill-conditioned where the bin is ~0
    dphi = (ph.cpu() - ph_r)[strong]
    dphi = torch.atan2(torch.sin(dphi), torch.cos(dphi))
    assert float(dphi.abs().max()) <= 2e-3
    y = ours.inverse(mag_r.cuda(), ph_r.cuda())                           # same inputs to both inverses
    y_r = ref.inverse(mag_r, ph_r)
    assert y.shape == y_r.shape
    assert float((y.cpu() - y_r).abs().max()) <= 2e-4 * float(y_r.abs().max())
    rec = ours(x.cuda()).squeeze(1).cpu()                                 # analysis -> synthesis reproduces the signal
    m = min(rec.shape[1], n)
    assert float((rec[:, n_fft:m - n_fft] - x[:, n_fft:m - n_fft]).abs().max()) <= 1e-3


@pytest.mark.gpu
def test_griffin_lim_converges_like_the_oracle():
    g = torch.Generator().manual_seed(1)
    t = torch.arange(16384) / 22050.0
    x = (0.4 * torch.sin(2 * np.pi * 220 * t) + 0.2 * torch.sin(2 * np.pi * 1330 * t + 0.3))[None] + 0.01 * torch.randn(1, 16384, generator=g)
    ref = O.STFT(1024, 256, 1024)
    mag, _ = ref.transform(x)
    angles = (torch.rand(mag.shape, generator=g) * 2 - 1) * np.pi
    ours = STFT(1024, 256, 1024).cuda()
    sig = griffin_lim(mag.cuda(), ours, 5, angles=angles)
    sig_r = O.griffin_lim(mag, ref, 5, angles)
    assert sig.shape == sig_r.shape
    err = lambda s, st: float(((st.transform(s)[0].cpu() if s.is_cuda else st.transform(s)[0]) - mag).pow(2).mean().sqrt())
    e_ours, e_ref = err(sig, ours), err(sig_r, ref)
    assert abs(e_ours - e_ref) <= 0.05 * e_ref + 1e-4, (e_ours, e_ref)
