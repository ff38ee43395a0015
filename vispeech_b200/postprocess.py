"""Waveform post-processing on the GPU: what reference inference_api.py:50-51 does with a disk
round trip (scipy `wavfile.write` of float32 at 44.1 kHz, then `ffmpeg -i c.wav -ar 22050`), as one kernel:
float32 -> [2:1 low-pass FIR decimation] -> signed 16-bit PCM, plus a RIFF/WAVE header helper.

Resampler definition.  `ffmpeg -ar 22050` on a float WAV runs libswresample (a dependency of the reference's shell command,
not vendored; ffmpeg is absent from this image, so no fixture can be generated) with its defaults: internal format float,
polyphase FIR designed by `build_filter` with filter_size 32, cutoff 0.97, Kaiser window beta 9, then float -> s16 as
clip(lrintf(32768 v)) without dither.  For 44100 -> 22050 one filter phase is used; restated (`swr_kaiser_fir`):
    factor = 0.97 * 22050 / 44100,  taps = ceil(32 / factor) = 66,  center = (taps - 1) // 2 = 32,
    h[i] ~ sinc(pi (i - center) factor) * I0(9 sqrt(1 - w^2)),  w = 2 (i - center) / taps,   sum h = 1,
    out[t] = s16( sum_i h[i] x[2 t + i - center] ).
PIN: the kernel is checked against a float64 statement of exactly this (tests/test_gpu_infer.py) and that statement against
an independent implementation, `scipy.signal.correlate(...)[::2]` with the same taps (tests/test_serving_host.py).  Not
reproduced: libswresample reflects ~33 samples at either end of the stream, here the signal is zero-extended (the first and
last 1.5 ms differ), and its Bessel I0 is a polynomial approximation (relative 1e-7).  `halfband_fir` (63-tap, beta 8.6,
cutoff 0.95) is round 1's own design, kept as an option.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def halfband_fir(n_taps: int = 63, beta: float = 8.6, cutoff: float = 0.95) -> np.ndarray:
    n = np.arange(n_taps) - (n_taps - 1) / 2
    fc = 0.25 * cutoff                       # cycles/sample at the input rate: (new Nyquist = 0.25) * cutoff
    h = 2 * fc * np.sinc(2 * fc * n) * np.kaiser(n_taps, beta)
    return (h / h.sum()).astype(np.float32)


def swr_kaiser_fir(rate_in: int = 44100, rate_out: int = 22050, filter_size: int = 32, cutoff: float = 0.97,
                   beta: float = 9.0) -> np.ndarray:
    """libswresample's default low-pass design (resample.c build_filter, Kaiser type) for an integer decimation: phase 0."""
    factor = min(rate_out * cutoff / rate_in, 1.0)
    taps = max(int(np.ceil(filter_size / factor)), 1)
    center = (taps - 1) // 2
    x = np.pi * (np.arange(taps) - center) * factor
    y = np.where(x == 0, 1.0, np.sin(x) / np.where(x == 0, 1.0, x))
    w = 2.0 * x / (factor * taps * np.pi)
    y = y * np.i0(beta * np.sqrt(np.maximum(1.0 - w * w, 0.0)))
    return (y / y.sum()).astype(np.float32)


def default_fir(rate_in: int = 44100, rate_out: int = 22050) -> np.ndarray:
    return swr_kaiser_fir(rate_in, rate_out)


def decimate_reference(x: np.ndarray, h: np.ndarray, dec: int = 2) -> np.ndarray:
    """float64 statement of the kernel's definition: out[t] = sum_i h[i] x[dec t + i - (len(h) - 1) // 2], zero-extended."""
    x, h = np.asarray(x, np.float64), np.asarray(h, np.float64)
    c = (h.size - 1) // 2
    xz = np.concatenate([np.zeros(c), x, np.zeros(h.size)])
    t_out = (x.size + dec - 1) // dec
    return np.array([np.dot(h, xz[dec * t: dec * t + h.size]) for t in range(t_out)])


@torch.no_grad()
def to_pcm16(o: torch.Tensor, n_samples, rate_in: int = 44100, rate_out: int = 44100, fir: str = "swr") -> torch.Tensor:
    """o: [B,1,T] or [B,T] fp32 CUDA tensor (the first output of `infer`); n_samples: valid samples per utterance.
    Returns int16 [B, T_out] on the same device (T_out = ceil(T / decimate))."""
    if rate_out not in (rate_in, rate_in // 2):
        raise ValueError("only 1:1 and 2:1 are built")
    x = o.reshape(o.shape[0], -1).contiguous()
    B, T = x.shape
    dec = rate_in // rate_out
    t_out = (T + dec - 1) // dec
    dev = x.device
    # pinned + asynchronous uploads: a copy from pageable memory would synchronise the stream first (synthesizer._upload)
    ns = torch.from_numpy(np.asarray(n_samples, dtype=np.int32)).pin_memory().to(dev, non_blocking=True)
    fir = _fir_on(dev, default_fir(rate_in, rate_out) if fir == "swr" else halfband_fir(), fir) if dec == 2 else None
    out = torch.empty(B, t_out, dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().vs_wave_pcm16(ptr(x), B, T, ptr(ns), dec, ptr(fir), 0 if fir is None else fir.numel(), ptr(out),
                                        t_out, torch.cuda.current_stream(dev).cuda_stream), "vs_wave_pcm16")
    return out


_FIR_CACHE: dict = {}


def _fir_on(dev, taps: np.ndarray, name: str) -> torch.Tensor:
    key = (str(dev), name, taps.size)
    if key not in _FIR_CACHE:
        _FIR_CACHE[key] = torch.from_numpy(taps).to(dev)
    return _FIR_CACHE[key]


def wav_bytes(pcm: np.ndarray, rate: int) -> bytes:
    """Mono 16-bit RIFF/WAVE container around int16 samples."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    return (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16)
            + b"data" + struct.pack("<I", len(data)) + data)
