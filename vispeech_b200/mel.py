"""Linear and log-mel spectrograms on the GPU with the reference's call surface:
`spectrogram_torch` (mel_processing.py:50-70, the input of `voice_conversion`), `spec_to_mel_torch` (:73-82) and
`mel_spectrogram_torch` (:85-112, the metric train.py:303-313 logs and the "mel within 1e-2" bar is stated on).

The STFT the reference takes from `torch.stft` (n_fft = win = 2048, hop 512, hann, center=False after a reflect pad of
(n_fft - hop) / 2) is a 4-tap convolution over rows of one hop each whose weights are the windowed DFT basis, so it runs
as one 3xTF32 tcgen05 GEMM of csrc/umma_tf32.cu (fp32-level accuracy); the 80-band Slaney mel projection is a second one.
The filterbank is librosa's `filters.mel` definition (Slaney scale, Slaney area normalisation), restated here because
librosa is not a dependency; tests/test_abi_and_host.py checks it against torchaudio's implementation of the same.
CUDA only, like the rest of the package.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .layout import make_rows
from .packing import pack_tf32


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sampling_rate: int, n_fft: int, num_mels: int, fmin: float = 0.0, fmax: Optional[float] = None) -> np.ndarray:
    """[num_mels, n_fft // 2 + 1] float32: triangular filters on the Slaney mel scale, each normalised to unit area
    (what `librosa_mel_fn(sampling_rate, n_fft, num_mels, fmin, fmax)` returns, mel_processing.py:78)."""
    fmax = float(fmax) if fmax else sampling_rate / 2.0
    freqs = np.linspace(0.0, sampling_rate / 2.0, n_fft // 2 + 1)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), num_mels + 2))
    fdiff = np.diff(pts)
    ramps = pts[:, None] - freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (pts[2:num_mels + 2] - pts[:num_mels]))[:, None]
    return w.astype(np.float32)


def dft_basis(n_fft: int, hop: int) -> Tuple[torch.Tensor, int]:
    """Windowed DFT basis as conv weights [taps = n_fft / hop][K][2 * half]: column f = hann[n] cos(2 pi f n / n_fft),
    column half + f = -hann[n] sin(...), n = tap * hop + row; K = hop rounded up to 48, half = n_bins rounded up to 192."""
    taps, n_bins = n_fft // hop, n_fft // 2 + 1
    k = -(-hop // 48) * 48
    half = -(-n_bins // 192) * 192
    n = np.arange(n_fft, dtype=np.float64)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)             # torch.hann_window(periodic=True), mel_processing.py:61
    ang = 2 * np.pi * np.outer(n, np.arange(n_bins)) / n_fft
    w = np.zeros((taps, k, 2 * half), np.float32)
    w[:, :hop, :n_bins] = (win[:, None] * np.cos(ang)).reshape(taps, hop, n_bins)
    w[:, :hop, half:half + n_bins] = (-win[:, None] * np.sin(ang)).reshape(taps, hop, n_bins)
    return torch.from_numpy(w), half


class MelSpectrogram:
    """Device-resident bases for one (n_fft, hop, sampling_rate, num_mels, fmin, fmax); call it on waveforms."""

    def __init__(self, n_fft: int = 2048, num_mels: int = 80, sampling_rate: int = 44100, hop_size: int = 512,
                 win_size: int = 2048, fmin: float = 0.0, fmax: Optional[float] = None, device="cuda:0"):
        if not torch.cuda.is_available():
            raise _lib.VsError("vispeech_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if n_fft != win_size or n_fft != 4 * hop_size or hop_size % 4 or num_mels > 96:
            raise _lib.VsError("only n_fft = win_size = 4 * hop_size (configs/config.json: 2048 / 512) and <= 96 mels are built")
        self.n_fft, self.hop, self.num_mels, self.n_bins = n_fft, hop_size, num_mels, n_fft // 2 + 1
        self.device = torch.device(device)
        self._lib = _lib.load()
        w, self.half = dft_basis(n_fft, hop_size)
        self.dft = pack_tf32(w, split3=True).to(self.device)
        ld_mag = -(-self.n_bins // 48) * 48
        fb = np.zeros((1, ld_mag, 96), np.float32)
        fb[0, :self.n_bins, :num_mels] = mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax).T
        self.mel = pack_tf32(torch.from_numpy(fb), split3=True).to(self.device)
        self._ws: Optional[torch.Tensor] = None

    @torch.no_grad()
    def __call__(self, y: torch.Tensor, lengths: Optional[Sequence[int]] = None, want: str = "mel"):
        """y: [B, T] (or [T]) float waveform on the device; lengths: valid samples per row (default T).
        want = "mel" -> log-mel [B, num_mels, n_frames]; "spec" -> magnitude [B, n_fft/2+1, n_frames]; "both" -> (spec, mel).
        n_frames = lengths // hop, as torch.stft(center=False) gives after the reference's padding."""
        if y.dim() == 1:
            y = y[None]
        y = y.to(self.device, torch.float32).contiguous()
        B, T = y.shape
        lens = np.full(B, T, np.int64) if lengths is None else np.asarray(lengths, np.int64).reshape(-1)
        pad = (self.n_fft - self.hop) // 2
        if lens.shape[0] != B or lens.max() > T or lens.min() <= pad:
            raise ValueError("lengths must be in (%d, T] (reflect padding needs more than %d samples)" % (pad, pad))
        frames = (lens // self.hop).astype(np.int32)
        dev = self.device
        with torch.cuda.device(dev):
            rows = make_rows(frames + 3, np.zeros(B, np.int32), 4, dev)
            R = rows.n_rows
            ld_x, ld_mag = -(-self.hop // 48) * 48, -(-self.n_bins // 48) * 48
            need = 4 * R * (ld_x + 2 * self.half + ld_mag + 96) + 4 * 256
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            ns = torch.from_numpy(lens.astype(np.int32)).to(dev)
            fmax_ = int(frames.max())
            spec = torch.empty(B, self.n_bins, fmax_, dtype=torch.float32, device=dev) if want in ("spec", "both") else None
            mel = torch.empty(B, self.num_mels, fmax_, dtype=torch.float32, device=dev) if want in ("mel", "both") else None
            check(self._lib.vs_mel_spectrogram(ctypes.byref(rows.struct), ptr(y), T, ptr(ns), self.hop, self.n_bins,
                                               self.num_mels, ptr(self.dft), ptr(self.mel), fmax_, ptr(spec), ptr(mel),
                                               ptr(self._ws), self._ws.numel(), torch.cuda.current_stream(dev).cuda_stream),
                  "vs_mel_spectrogram")
        return {"mel": mel, "spec": spec, "both": (spec, mel)}[want]


_CACHE: Dict[tuple, MelSpectrogram] = {}


def _get(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax) -> MelSpectrogram:
    key = (n_fft, num_mels, sampling_rate, hop_size, win_size, float(fmin), fmax, str(y.device))
    if key not in _CACHE:
        _CACHE[key] = MelSpectrogram(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, device=y.device)
    return _CACHE[key]


def spectrogram_torch(y, n_fft, sampling_rate, hop_size, win_size, center=False):
    """Same arguments and result as reference mel_processing.py:50-70 (y: [B, T] on the GPU) -> [B, n_fft/2+1, frames]."""
    if center:
        raise ValueError("center=True is not what the reference path uses")
    return _get(y, n_fft, 80, sampling_rate, hop_size, win_size, 0.0, None)(y, want="spec")


def mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """Same arguments and result as reference mel_processing.py:85-112 (y: [B, T] on the GPU) -> [B, num_mels, frames]."""
    if center:
        raise ValueError("center=True is not what the reference path uses")
    return _get(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)(y, want="mel")
