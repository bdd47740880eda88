"""CPU restatement of the reference's STFT / inverse STFT / Griffin-Lim (utils/stft.py:41-156,
dataset/audio_processing.py:169-240) with the same ATen calls (reflect F.pad, F.conv1d at stride hop, F.conv_transpose1d),
minus the `.cuda()` / `.cpu()` hops.  TEST INFRASTRUCTURE ONLY.  The reference module itself cannot be imported offline
(it needs librosa's pad_center / tiny and scipy's get_window at import time); the window construction below is those two
functions' published definitions (periodic hann; centre zero-padding)."""
import numpy as np
import torch
import torch.nn.functional as F


def hann_padded(win_length, n_fft):
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)      # scipy get_window("hann", fftbins=True)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad))                           # librosa.util.pad_center


class STFT:
    def __init__(self, filter_length=800, hop_length=200, win_length=800):
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        scale = filter_length / hop_length                                         # utils/stft.py:54
        fourier_basis = np.fft.fft(np.eye(filter_length))
        cutoff = int(filter_length / 2 + 1)
        fourier_basis = np.vstack([np.real(fourier_basis[:cutoff, :]), np.imag(fourier_basis[:cutoff, :])])
        forward_basis = torch.FloatTensor(fourier_basis[:, None, :])
        inverse_basis = torch.FloatTensor(np.linalg.pinv(scale * fourier_basis).T[:, None, :])
        win = torch.from_numpy(hann_padded(win_length, filter_length)).float()
        self.forward_basis = (forward_basis * win).float()                         # :74-77
        self.inverse_basis = (inverse_basis * win).float()

    def transform(self, x):                                                         # :82-112
        B, n = x.shape
        x = F.pad(x.view(B, 1, n).unsqueeze(1), (self.filter_length // 2, self.filter_length // 2, 0, 0), mode="reflect").squeeze(1)
        ft = F.conv1d(x, self.forward_basis, stride=self.hop_length, padding=0)
        cutoff = self.filter_length // 2 + 1
        re, im = ft[:, :cutoff, :], ft[:, cutoff:, :]
        return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)

    def inverse(self, magnitude, phase):                                            # :114-151
        rec = torch.cat([magnitude * torch.cos(phase), magnitude * torch.sin(phase)], dim=1)
        out = F.conv_transpose1d(rec, self.inverse_basis, stride=self.hop_length, padding=0)
        n_frames = magnitude.size(-1)
        n = self.filter_length + self.hop_length * (n_frames - 1)
        wsum = np.zeros(n, dtype=np.float32)                                        # window_sumsquare, audio_processing.py:169-221
        win_sq = hann_padded(self.win_length, self.filter_length) ** 2
        for i in range(n_frames):
            s = i * self.hop_length
            wsum[s: min(n, s + self.filter_length)] += win_sq[: max(0, min(self.filter_length, n - s))]
        nz = torch.from_numpy(np.where(wsum > np.finfo(np.float32).tiny)[0])
        out[:, :, nz] /= torch.from_numpy(wsum)[nz]
        out *= float(self.filter_length) / self.hop_length
        out = out[:, :, self.filter_length // 2:]
        out = out[:, :, : -(self.filter_length // 2)]
        return out


def griffin_lim(magnitudes, stft_fn, n_iters, angles):                               # audio_processing.py:224-240 (angles injected)
    signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
    for _ in range(n_iters):
        _, angles = stft_fn.transform(signal)
        signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
    return signal
