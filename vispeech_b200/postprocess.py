"""Waveform post-processing on the GPU: what reference inference_api.py:50-51 does with a disk
round trip (scipy `wavfile.write` of float32 at 44.1 kHz, then `ffmpeg -i c.wav -ar 22050`), as one kernel:
float32 -> [2:1 low-pass FIR decimation] -> signed 16-bit PCM, plus a RIFF/WAVE header helper.

Resampler definition (ffmpeg's libswresample is not reproduced bit for bit - it is not available here, so this part of
the parity is unpinned): linear-phase Kaiser-windowed sinc, 63 taps, beta 8.6, cutoff 0.475 * 44.1 kHz / 2 ... i.e.
0.95 of the new Nyquist; out[t] = sum_k h[k] x[2t + k - 31].  Quantisation: clip(rint(32768 * v)), ffmpeg's float->s16 rule.
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


@torch.no_grad()
def to_pcm16(o: torch.Tensor, n_samples, rate_in: int = 44100, rate_out: int = 44100) -> torch.Tensor:
    """o: [B,1,T] or [B,T] fp32 CUDA tensor (the first output of `infer`); n_samples: valid samples per utterance.
    Returns int16 [B, T_out] on the same device (T_out = ceil(T / decimate))."""
    if rate_out not in (rate_in, rate_in // 2):
        raise ValueError("only 1:1 and 2:1 are built")
    x = o.reshape(o.shape[0], -1).contiguous()
    B, T = x.shape
    dec = rate_in // rate_out
    t_out = (T + dec - 1) // dec
    dev = x.device
    ns = torch.as_tensor(np.asarray(n_samples, dtype=np.int32)).to(dev)
    fir = torch.from_numpy(halfband_fir()).to(dev) if dec == 2 else None
    out = torch.empty(B, t_out, dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().vs_wave_pcm16(ptr(x), B, T, ptr(ns), dec, ptr(fir), 0 if fir is None else fir.numel(), ptr(out),
                                        t_out, torch.cuda.current_stream(dev).cuda_stream), "vs_wave_pcm16")
    return out


def wav_bytes(pcm: np.ndarray, rate: int) -> bytes:
    """Mono 16-bit RIFF/WAVE container around int16 samples."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    return (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16)
            + b"data" + struct.pack("<I", len(data)) + data)
