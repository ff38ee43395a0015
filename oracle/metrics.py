"""Evaluation metrics for the parity bars of BASELINE.json (test infrastructure).

`mel_spectrogram` restates reference mel_processing.py:85-112 (`mel_spectrogram_torch`, the metric train.py:303-313 logs):
reflect pad (n_fft-hop)/2, STFT n_fft=win=2048, hop 512, hann, center=False, magnitude sqrt(re^2+im^2+1e-6), 80-band
Slaney mel 0..sr/2, log(clamp(., 1e-5)).  The reference takes the filterbank from librosa (not installed here);
torchaudio's Slaney-normalised Slaney-scale filterbank is the same definition.
"""
from __future__ import annotations

import torch


def mel_spectrogram(y: torch.Tensor, n_fft=2048, n_mels=80, sr=44100, hop=512, win=2048, fmin=0.0, fmax=None) -> torch.Tensor:
    import torchaudio
    y = y.reshape(1, -1).float()
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, float(fmax or sr / 2), n_mels, sr, norm="slaney",
                                               mel_scale="slaney").t()
    pad = (n_fft - hop) // 2
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), center=False,
                      normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-6)
    return torch.log(torch.clamp(torch.matmul(fb, mag[0]), min=1e-5))


def snr_db(ref: torch.Tensor, x: torch.Tensor) -> float:
    ref, x = ref.double().reshape(-1), x.double().reshape(-1)
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum().clamp_min(1e-300)))
