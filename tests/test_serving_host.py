"""CPU-only checks of the serving queue (8f rank 1) and the PCM helpers (8f rank 2): batching logic with a fake
synthesis function, WAV container, FIR definition."""
import struct
import threading
import time

import numpy as np
import torch

from vispeech_b200.postprocess import decimate_reference, default_fir, halfband_fir, swr_kaiser_fir, wav_bytes
from vispeech_b200.serving import BatchingSynthesizer


def test_queue_batches_and_never_rejects():
    calls = []

    def fake(reqs):
        calls.append(len(reqs))
        time.sleep(0.02)                                   # a "batch time": requests pile up meanwhile
        return [np.full(int(r.ids.numel()), r.sid, np.int16) for r in reqs]

    s = BatchingSynthesizer(synth_batch=fake, max_batch=8, max_wait_ms=10)
    futs = []
    for i in range(40):                                    # the reference would answer "server busy" to 39 of these
        futs.append(s.submit(list(range(2 + i % 5)), sid=i, duration=[3] * (2 + i % 5)))
    outs = [f.result(timeout=10) for f in futs]
    s.close()
    assert all(o.shape[0] == 2 + i % 5 and int(o[0]) == i for i, o in enumerate(outs))
    assert s.stats["requests"] == 40 and max(calls) > 1 and max(calls) <= 8
    assert sum(calls) == 40


def test_groups_do_not_mix_control_signatures():
    seen = []

    def fake(reqs):
        seen.append({(r.noise_scale, r.duration is None) for r in reqs})
        return [np.zeros(1, np.int16) for _ in reqs]

    s = BatchingSynthesizer(synth_batch=fake, max_batch=16, max_wait_ms=30)
    futs = [s.submit([1, 2, 3], 0, duration=[1, 1, 1]), s.submit([1, 2, 3], 0), s.submit([1, 2, 3], 0, duration=[2, 2, 2], noise_scale=0.5)]
    [f.result(timeout=10) for f in futs]
    s.close()
    assert all(len(g) == 1 for g in seen)


def test_single_phoneme_is_rejected_like_the_reference_crash():
    s = BatchingSynthesizer(synth_batch=lambda r: [np.zeros(1, np.int16)] * len(r))
    f = s.submit([5], 0)
    s.close()
    assert isinstance(f.exception(timeout=5), ValueError)


def test_wav_container_and_fir():
    pcm = (np.arange(1000) % 200 - 100).astype(np.int16)
    b = wav_bytes(pcm, 22050)
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt " and struct.unpack("<I", b[24:28])[0] == 22050
    assert struct.unpack("<I", b[40:44])[0] == 2000 and len(b) == 44 + 2000
    h = halfband_fir()
    assert h.shape == (63,) and abs(h.sum() - 1) < 1e-6 and np.allclose(h, h[::-1])
    H = np.abs(np.fft.rfft(h, 4096))
    f = np.fft.rfftfreq(4096)                               # cycles/sample at 44.1 kHz
    assert H[f < 0.20].min() > 0.99 and H[f > 0.30].max() < 1e-3     # flat to 8.8 kHz, > 60 dB down above 13.2 kHz


def test_swr_default_filter_design_and_decimation_pin():
    """The resampler is pinned to a definition, not to itself: (a) the restated libswresample default design (filter_size 32,
    cutoff 0.97, Kaiser beta 9 -> 66 taps, peak at tap 32, unit DC gain, flat to 8.5 kHz, -6 dB at 0.97 x the new Nyquist, > 80 dB down above 13 kHz);
    (b) the float64 statement the kernel is tested against equals an independent implementation (scipy's correlate + slicing)."""
    from scipy import signal
    h = swr_kaiser_fir()
    assert h.shape == (66,) and int(np.argmax(h)) == 32 and abs(float(h.sum()) - 1) < 1e-6 and np.array_equal(h, default_fir())
    assert np.allclose(h[32 - 30: 32], h[32 + 30: 32: -1], atol=1e-8)          # symmetric around the peak (even length: one extra tap)
    H = np.abs(np.fft.rfft(h.astype(np.float64), 8192))
    f = np.fft.rfftfreq(8192) * 44100
    assert H[f < 8500].min() > 0.9999 and abs(H[np.argmin(np.abs(f - 0.97 * 11025))] - 0.5) < 0.02 and H[f > 13000].max() < 1e-4
    g = np.random.default_rng(0)
    x = g.standard_normal(4001)
    want = decimate_reference(x, h)
    c = (h.size - 1) // 2
    xz = np.concatenate([np.zeros(c), x, np.zeros(h.size)])
    alt = signal.correlate(xz, h.astype(np.float64), mode="valid")[::2][: want.size]
    assert want.shape == (2001,) and np.abs(want - alt).max() < 1e-12
    assert np.abs(decimate_reference(x, halfband_fir())[5:-5] - signal.correlate(np.concatenate([np.zeros(31), x, np.zeros(63)]), halfband_fir().astype(np.float64), mode="valid")[::2][5:1996]).max() < 1e-12
