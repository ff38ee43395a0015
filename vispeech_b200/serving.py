"""Serving surface with dynamic batching.

The reference's Flask app (inference_api.py:12-69) synthesises one request at a time behind a NON-blocking global mutex
and answers "server busy" to everything that arrives meanwhile (inference_api.py:13,37,63-64).  On a B200 a batch of 64
utterances costs about as much wall time as a handful, so here requests queue up and a worker thread drains the queue
into length-bucketed batches (`sharding.bucket_batches`): nothing is rejected, latency under load is one batch time.

`BatchingSynthesizer.submit()` returns a `concurrent.futures.Future` of the int16 PCM (22.05 kHz by default, like the
reference's `ffmpeg -ar 22050` output); `submit_text()` puts `text.TextFrontend` (symbol table + LRU phoneme cache, the
reference's G2P when importable) in front of it.  `create_app()` wraps both in an ASGI app with the reference's route shape:
`GET /tts?text=...` (inference_api.py:59-65; `text` may also be a phoneme string), or `GET /tts?ids=...` for callers that
run their own front end.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .sharding import bucket_batches, frames_from_durations


@dataclass
class Request:
    ids: torch.Tensor                      # [Tp] phoneme ids
    sid: int
    duration: Optional[torch.Tensor] = None  # [Tp] frames; None -> duration predictor
    f0: Optional[torch.Tensor] = None
    energy: Optional[torch.Tensor] = None
    noise_scale: float = 0.667             # inference.py:44 / inference_api.py:46
    future: Future = field(default_factory=Future)
    t_submit: float = field(default_factory=time.perf_counter)


class BatchingSynthesizer:
    """Queue + worker thread.  `synth_batch(requests) -> list of int16 numpy arrays` does the actual work; the default
    one runs `SynthesizerTrn.infer` + `postprocess.to_pcm16` on the GPU.  Requests with different `noise_scale` or with
    predicted vs given durations are never mixed in one `infer` call."""

    def __init__(self, net=None, max_batch: int = 64, max_wait_ms: float = 5.0, max_frames_per_batch: int = 65536,
                 rate_out: int = 22050, synth_batch: Optional[Callable[[List[Request]], List[np.ndarray]]] = None):
        self.net, self.max_batch, self.max_wait, self.max_frames = net, max_batch, max_wait_ms / 1000.0, max_frames_per_batch
        self.rate_out = rate_out
        self._synth = synth_batch or self._synth_gpu
        self._q: "queue.Queue[Optional[Request]]" = queue.Queue()
        self.stats = {"requests": 0, "batches": 0, "max_batch_seen": 0}
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    # ---- client side
    def submit(self, ids, sid: int, duration=None, f0=None, energy=None, noise_scale: float = 0.667) -> Future:
        r = Request(torch.as_tensor(ids).long().reshape(-1), int(sid),
                    None if duration is None else torch.as_tensor(duration).reshape(-1),
                    None if f0 is None else torch.as_tensor(f0).float().reshape(-1),
                    None if energy is None else torch.as_tensor(energy).float().reshape(-1), float(noise_scale))
        err = self._validate(r)
        if err is not None:                      # a malformed request fails alone, never its batch
            r.future.set_exception(ValueError(err))
            return r.future
        self._q.put(r)
        return r.future

    def _validate(self, r: Request) -> Optional[str]:
        n = int(r.ids.numel())
        if n < 2:
            return "need at least 2 phonemes (the reference fails on 1, models.py:420)"
        n_vocab = getattr(self.net, "n_vocab", None)
        n_spk = getattr(self.net, "n_speakers", None)
        if n_vocab is not None and (int(r.ids.min()) < 0 or int(r.ids.max()) >= n_vocab):
            return "phoneme id out of range [0, %d)" % n_vocab
        if n_spk is not None and not (0 <= r.sid < n_spk):
            return "sid out of range [0, %d)" % n_spk
        for name, c in (("duration", r.duration), ("f0", r.f0), ("energy", r.energy)):
            if c is not None and int(c.numel()) != n:
                return "%s must have one entry per phoneme (%d), got %d" % (name, n, int(c.numel()))
        if not (r.noise_scale >= 0.0):
            return "noise_scale must be >= 0"
        return None

    def submit_text(self, text: str, sid: int, frontend=None, **kw) -> Future:
        """What inference_api.py:15-19,40-47 does per request: text -> ids (cached) -> synthesis.  A front-end error
        (unknown symbols, no G2P for raw text) fails this request's future only."""
        fe = frontend if frontend is not None else self._frontend()
        try:
            ids = fe.text_to_sequence(text)
        except (ValueError, KeyError) as e:
            f: Future = Future()
            f.set_exception(ValueError(str(e)))
            return f
        return self.submit(ids, sid, **kw)

    def _frontend(self):
        if getattr(self, "_fe", None) is None:
            from .text import TextFrontend
            self._fe = TextFrontend()
        return self._fe

    def close(self):
        self._q.put(None)
        self._worker.join(timeout=30)

    # ---- worker side
    def _take(self) -> Optional[List[Request]]:
        first = self._q.get()
        if first is None:
            return None
        batch, deadline = [first], time.perf_counter() + self.max_wait
        while len(batch) < self.max_batch:
            left = deadline - time.perf_counter()
            try:
                r = self._q.get(timeout=max(left, 0)) if left > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if r is None:
                self._q.put(None)
                break
            batch.append(r)
        return batch

    def _run(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            self.stats["requests"] += len(batch)
            groups: Dict[tuple, List[Request]] = {}
            for r in batch:   # one infer call needs a common noise_scale and a common control signature
                groups.setdefault((r.noise_scale, r.duration is None, r.f0 is None, r.energy is None), []).append(r)
            for reqs in groups.values():
                for sub in self._split(reqs):
                    self.stats["batches"] += 1
                    self.stats["max_batch_seen"] = max(self.stats["max_batch_seen"], len(sub))
                    try:
                        outs = self._synth(sub)
                        for r, o in zip(sub, outs):
                            r.future.set_result(o)
                    except Exception as e:           # a failed batch must not take the server down ...
                        if len(sub) == 1:
                            sub[0].future.set_exception(e)
                            continue
                        for r in sub:                # ... nor its innocent members: retry them one by one
                            if r.future.done():
                                continue
                            try:
                                r.future.set_result(self._synth([r])[0])
                            except Exception as e1:
                                r.future.set_exception(e1)

    def _split(self, reqs: List[Request]) -> List[List[Request]]:
        if reqs[0].duration is None:                 # frame counts unknown before the duration predictor ran
            return [reqs]
        frames = frames_from_durations([r.duration for r in reqs])
        return [[reqs[i] for i in b] for b in bucket_batches(range(len(reqs)), frames, self.max_frames)]

    def _synth_gpu(self, reqs: List[Request]) -> List[np.ndarray]:
        from .postprocess import to_pcm16
        net = self.net
        B = len(reqs)
        tp = max(int(r.ids.numel()) for r in reqs)

        def pad(get, dtype):
            out = torch.zeros(B, tp, dtype=dtype)
            for b, r in enumerate(reqs):
                v = get(r)
                out[b, : v.numel()] = v.to(dtype)
            return out

        kw = {}
        if reqs[0].duration is not None:
            fl = any(r.duration.is_floating_point() for r in reqs)
            kw["duration_control"] = pad(lambda r: r.duration, torch.float64 if fl else torch.long)
        if reqs[0].f0 is not None:
            kw["pitch_control"] = pad(lambda r: r.f0, torch.float32)
        if reqs[0].energy is not None:
            kw["energy_control"] = pad(lambda r: r.energy, torch.float32)
        o, x_mask, *_ = net.infer(pad(lambda r: r.ids, torch.long), torch.LongTensor([int(r.ids.numel()) for r in reqs]),
                                  sid=torch.LongTensor([r.sid for r in reqs]), noise_scale=reqs[0].noise_scale,
                                  outputs="audio", **kw)
        n = (x_mask.sum(dim=(1, 2)) * net.hop_length).to(torch.int32).cpu().numpy()
        pcm = to_pcm16(o, n, net.sampling_rate, self.rate_out).cpu().numpy()
        dec = net.sampling_rate // self.rate_out
        return [pcm[b, : (int(n[b]) + dec - 1) // dec].copy() for b in range(B)]


def create_app(synth: BatchingSynthesizer, frontend=None):
    """ASGI app with the reference's route shape (inference_api.py:59-65):
        GET /tts?text=<raw text or phoneme string>[&sid=1]        -> audio/wav (s16, 22.05 kHz)
        GET /tts?ids=12,7,33&sid=1[&durations=4,6,5]               -> the same, ids from the caller's own front end
    The reference hard-codes speaker 1 (inference_api.py:44); `sid` defaults to that."""
    from fastapi import FastAPI, HTTPException
    from fastapi.responses import Response

    from .postprocess import wav_bytes
    app = FastAPI(title="vispeech_b200")

    @app.get("/tts")
    def tts(text: Optional[str] = None, ids: Optional[str] = None, sid: int = 1, durations: Optional[str] = None):
        try:
            if (text is None) == (ids is None):
                raise ValueError("pass exactly one of text= and ids=")
            dur = None if durations is None else [float(v) for v in durations.split(",")]
            if text is not None:
                fut = synth.submit_text(text, sid, frontend=frontend, duration=dur)
            else:
                fut = synth.submit([int(v) for v in ids.split(",") if v != ""], sid, duration=dur)
            pcm = fut.result(timeout=120)
        except ValueError as e:
            raise HTTPException(status_code=400, detail=str(e))
        return Response(content=wav_bytes(pcm, synth.rate_out), media_type="audio/wav")

    return app
