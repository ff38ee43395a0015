"""`SynthesizerTrn` with the reference's inference call surface, running on libvispeech_b200 (CUDA only).

Mirrors reference models.py:537-561 (ctor) and models.py:672-722 (`infer`), i.e. the surface used by
inference.py:26-44, inference_api.py:21-47, gui.py:96-100 and train.py:300-301.  Differences, all deliberate:
  * `infer` takes two extra keyword arguments: `noise` (the eps of models.py:718, needed for parity - the
    reference can only be seeded) and `outputs` ("all" | "audio") to skip unpacking the latents;
  * batches are processed with per-utterance (batch-1) semantics: pad positions never influence an
    utterance - every reference call site is batch 1, where both agree;
  * there is no CPU path: construction fails without a CUDA device / the built library.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import VsConfig, check, ptr
from .layout import FRAME_GAP, PHONEME_GAP, RaggedRows, make_rows
from .packing import pack_state_dict

Control = Union[None, float, int, torch.Tensor]
_DTYPES = {torch.float32: _lib.VS_DTYPE_F32, torch.float16: _lib.VS_DTYPE_F16}


class SynthesizerTrn:
    """Inference-only drop-in for reference `models.SynthesizerTrn` (same positional/keyword ctor args)."""

    def __init__(self, n_vocab, spec_channels, hop_length, sampling_rate, segment_size, inter_channels,
                 hidden_channels, filter_channels, n_heads, n_layers, kernel_size, p_dropout, resblock,
                 resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates, upsample_initial_channel,
                 upsample_kernel_sizes, n_speakers=0, gin_channels=0, use_sdp=False, freeze_textencoder=False,
                 freeze_decoder=False, device: Union[str, torch.device, None] = None, **kwargs):
        if not torch.cuda.is_available():
            raise _lib.VsError("vispeech_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _lib.VsError("vispeech_b200 runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if (inter_channels, hidden_channels, filter_channels, n_heads, kernel_size) != (192, 192, 768, 2, 3) or \
                str(resblock) != "1" or list(resblock_kernel_sizes) != [3, 7, 11] or \
                [list(d) for d in resblock_dilation_sizes] != [[1, 3, 5]] * 3 or \
                list(upsample_rates) != [8, 8, 4, 2] or list(upsample_kernel_sizes) != [16, 16, 4, 4] or \
                upsample_initial_channel != 512 or gin_channels != 256 or n_speakers <= 1 or hop_length != 512:
            raise _lib.VsError("only the configs/config.json architecture is built (see include/vispeech_b200.h)")
        self.n_vocab, self.n_speakers, self.hop_length, self.sampling_rate = n_vocab, n_speakers, hop_length, sampling_rate
        self.inter_channels, self.hidden_channels = inter_channels, hidden_channels
        self.n_layers = n_layers
        self.decoder_precision = 0          # 0 = bf16 tcgen05 (product); 1 = fp32 cross-check (tests only)
        self._lib = _lib.load()
        self._cfg = VsConfig(n_vocab=n_vocab, hidden=hidden_channels, filter=filter_channels, n_heads=n_heads,
                             n_layers=n_layers, pitch_layers=6, window=4, gin=gin_channels, n_speakers=n_speakers,
                             flow_layers=4, n_flows=4, upsample_initial=upsample_initial_channel, hop=hop_length)
        self._model = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.vs_model_create(ctypes.byref(self._cfg), ctypes.byref(self._model)), "vs_model_create")
        self._weights: Dict[str, torch.Tensor] = {}
        self._ws: Optional[torch.Tensor] = None
        self._ws_lat: Optional[torch.Tensor] = None
        self._lat_stream: Optional[torch.cuda.Stream] = None
        # Cross-call software pipeline (throughput mode): the latent stages (text encoder ... flow, small / latency-bound
        # kernels) of call i+1 run on a second stream while the decoder of call i owns the SMs.  Same kernels, same
        # results; only the enqueue order across calls changes.  Off by default (lowest single-call latency).
        self.overlap_calls = False
        self._in_flight: List[torch.cuda.Event] = []      # overlap_calls: decoder-done events of the calls not yet retired
        self._loaded = False

    # -- nn.Module-ish surface used by the reference scripts
    def eval(self):
        return self

    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise _lib.VsError("vispeech_b200 runs on CUDA devices only")
        if d.index is None:                      # "cuda" means the current device, like torch
            d = torch.device("cuda", torch.cuda.current_device())
        if d != self.device and self._loaded:
            raise _lib.VsError("move before loading weights")
        self.device = d
        return self

    def set_option(self, name: str, value: int) -> None:
        """Per-model runtime option (include/vispeech_b200.h lists them); overrides the process default for this handle."""
        check(self._lib.vs_model_set_option(self._model, name.encode(), int(value)), "vs_model_set_option(%s)" % name)

    def __del__(self):
        try:
            if getattr(self, "_model", None):
                self._lib.vs_model_destroy(self._model)
        except Exception:
            pass

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = False):
        """Takes the reference's own state dict (un-folded weight norm; extra keys such as enc_q.* ignored)."""
        packed = pack_state_dict(state_dict, n_layers=self.n_layers)
        with torch.cuda.device(self.device):
            for name, t in packed.items():
                d = t.to(self.device)
                self._weights[name] = d
                check(self._lib.vs_model_set_tensor(self._model, name.encode(), d.data_ptr(), d.numel(),
                                                    _DTYPES[d.dtype]), "vs_model_set_tensor(%s)" % name)
            check(self._lib.vs_model_finalize(self._model), "vs_model_finalize")
        self._loaded = True
        return self

    # -- helpers
    def _workspace(self, rp: int, rf: int) -> torch.Tensor:
        need = max(int(self._lib.vs_workspace_bytes_latent(self._model, rp, rf)),
                   int(self._lib.vs_workspace_bytes_decoder(self._model, rf, int(self.decoder_precision))))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _workspace_lat(self, rp: int, rf: int) -> torch.Tensor:
        need = int(self._lib.vs_workspace_bytes_latent(self._model, rp, rf))      # the side stream never runs a decoder
        if self._ws_lat is None or self._ws_lat.numel() < need:
            torch.cuda.synchronize(self.device)          # an older, smaller buffer may still be in use on the side stream
            self._ws_lat = None
            self._ws_lat = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws_lat

    @staticmethod
    def _per_utt(c: torch.Tensor, B: int, lengths: np.ndarray, name: str) -> List[np.ndarray]:
        c = c.detach()
        if c.dim() == 3:                      # durations also arrive as [B,1,Tp] (models.py:420 squeezes)
            c = c.reshape(c.shape[0], -1)
        if c.dim() == 1:
            c = c[None]
        if c.shape[0] != B or c.shape[1] < int(lengths.max()):
            raise ValueError("%s: expected [B=%d, Tp>=%d], got %s" % (name, B, int(lengths.max()), tuple(c.shape)))
        if c.is_floating_point():
            c = c.to(torch.float64)
        else:
            c = c.to(torch.int64)
        c = c.cpu().numpy()
        return [c[b, :lengths[b]] for b in range(B)]

    # ------------------------------------------------------------------------------------------------
    # infer = prepare (host: validate, lay out ragged rows, upload inputs) + run (device work) + outputs
    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prepare(self, phonemes, phonemes_lengths, sid=None, noise_scale=1, max_len=None, energy_control: Control = None,
                pitch_control: Control = None, duration_control: Control = None, noise=None,
                bucket: Optional[tuple] = None, static: Optional["Prepared"] = None) -> "Prepared":
        """Host side of `infer`: everything up to (and including) the host->device copies of the inputs.
        `bucket` = (phoneme rows, frame rows): lay the batch out in row buckets of these sizes (layout.make_rows);
        `static`: a Prepared of the same bucket whose device buffers receive the uploads (CUDA-graph replays read them)."""
        if not self._loaded:
            raise _lib.VsError("load_state_dict()/load_checkpoint() first")
        if sid is None:
            raise ValueError("sid is required (n_speakers > 0, models.py:674)")
        dev = self.device
        phon = phonemes.detach().cpu().numpy()
        if phon.ndim == 1:
            phon = phon[None]
        B, Tp = phon.shape
        lens = phonemes_lengths.detach().cpu().numpy().astype(np.int32).reshape(-1)
        sids = sid.detach().cpu().numpy().astype(np.int32).reshape(-1)
        if lens.shape[0] != B or sids.shape[0] != B:
            raise ValueError("phonemes_lengths / sid must have one entry per utterance")
        if lens.min() < 1 or lens.max() > Tp:
            raise ValueError("phonemes_lengths out of range")
        if sids.min() < 0 or sids.max() >= self.n_speakers:
            raise ValueError("sid out of range")
        for b in range(B):                         # nn.Embedding raises IndexError on these (models.py:169); so do we,
            v = phon[b, :lens[b]]                  # instead of synthesising from a zero embedding
            if v.min() < 0 or v.max() >= self.n_vocab:
                raise IndexError("phoneme id out of range [0, %d) in utterance %d" % (self.n_vocab, b))
        P = Prepared()
        P.B, P.Tp, P.lens, P.sids = B, Tp, lens, sids
        P.noise_scale, P.max_len = float(noise_scale), (-1 if max_len is None else int(max_len))
        P.duration_control = duration_control
        # uploads go to the stream that consumes them first (the side stream under overlap_calls: they must not queue
        # behind the previous call's decoder)
        up = self._side_stream() if self.overlap_calls else torch.cuda.current_stream(dev)
        with torch.cuda.device(dev), torch.cuda.stream(up):
            P.bucket, P.static = bucket, static
            P.rp = make_rows(lens, sids, PHONEME_GAP, dev, n_rows=bucket[0] if bucket else 0,
                             out=static.rp if static is not None else None)
            P.ids_rows = _upload(P.rp.scatter([phon[b] for b in range(B)], np.int32, fill=-1), dev,
                                 static.ids_rows if static is not None else None)

            def control(c, name, dtype, prev):
                if isinstance(c, torch.Tensor):
                    per = self._per_utt(c, B, lens, name)
                    return 2, 1.0, _upload(P.rp.scatter(per, dtype), dev, prev), per
                return 0, 1.0 if c is None else float(c), None, None

            st = static
            P.d_mode, P.d_scale, P.d_ctrl, d_per = control(duration_control, "duration_control", np.float64, st.d_ctrl if st else None)
            P.p_mode, P.p_scale, P.p_ctrl, _ = control(pitch_control, "pitch_control", np.float32, st.p_ctrl if st else None)
            P.e_mode, P.e_scale, P.e_ctrl, _ = control(energy_control, "energy_control", np.float32, st.e_ctrl if st else None)
            P.noise = noise
            P.frames = None
            P.rf = None
            P.eps = None
            if d_per is not None:
                # durations given: frame counts follow on the host by the same rule the kernel applies
                # (n_i = max(int(d_i), 0), models.py:421-423), so the frame layout needs no device round trip
                P.frames = np.asarray([int(np.clip(np.trunc(np.asarray(d, np.float64)), 0, 1.0e6).sum()) for d in d_per],
                                      np.int32)
                self._layout_frames(P)
        return P

    def _layout_frames(self, P: "Prepared") -> None:
        if int(P.frames.max()) < 1:
            raise ValueError("all durations are <= 0: nothing to synthesise")
        dev = self.device
        bucket, static = getattr(P, "bucket", None), getattr(P, "static", None)
        P.rf = make_rows(P.frames, P.sids, FRAME_GAP, dev, n_rows=bucket[1] if bucket else 0,
                         out=static.rf if static is not None else None)
        if P.noise is not None:
            fr = P.frames
            if isinstance(P.noise, torch.Tensor):
                per = [P.noise[b, :, :fr[b]].t().cpu().numpy() for b in range(P.B)]
            else:
                per = [P.noise[b][:, :fr[b]].t().cpu().numpy() for b in range(P.B)]
            P.eps = _upload(P.rf.scatter(per, np.float32, width=192), dev, static.eps if static is not None else None)

    @torch.no_grad()
    def run(self, P: "Prepared", outputs: str = "all", timings: Optional[dict] = None):
        """Device side of `infer`: kernels only (plus one small D2H of frame counts when durations are predicted).
        With `overlap_calls` the latent stages are enqueued on the side stream (see __init__); the decoder and the
        outputs stay on the caller's stream, which waits for the latents through an event."""
        lib, dev = self._lib, self.device
        B, Tp, rp = P.B, P.Tp, P.rp
        Rp = rp.n_rows
        ev = []
        main = torch.cuda.current_stream(dev)
        overlap = bool(self.overlap_calls)
        lat = self._side_stream() if overlap else main
        if overlap:
            # the host may run at most two calls ahead of the GPU: buffers handed between the two streams are recycled by
            # the caching allocator only once the other stream is done with them, so an unbounded run-ahead would make it
            # cudaMalloc (and thereby synchronise) for every call still in flight
            while len(self._in_flight) >= 2:
                self._in_flight.pop(0).synchronize()

        def mark(name, st):
            if timings is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                ev.append((name, e))

        with torch.cuda.device(dev), torch.cuda.stream(lat):
            stream = lat.cuda_stream
            rf_rows = P.rf.n_rows if P.rf is not None else 16
            ws = self._workspace_lat(Rp, rf_rows) if overlap else self._workspace(Rp, rf_rows)
            mark("start", lat)
            x = torch.empty(Rp, 192, dtype=torch.float32, device=dev)
            check(lib.vs_text_encode(self._model, ctypes.byref(rp.struct), ptr(P.ids_rows), ptr(x), ptr(ws), ws.numel(),
                                     stream), "vs_text_encode")
            mark("text_encoder", lat)
            dur = torch.empty(Rp, dtype=torch.float64, device=dev)
            f0 = torch.empty(Rp, dtype=torch.float32, device=dev)
            energy = torch.empty(Rp, dtype=torch.float32, device=dev)
            check(lib.vs_variance_adapter(self._model, ctypes.byref(rp.struct), ptr(x), P.d_mode, P.d_scale, ptr(P.d_ctrl),
                                          P.p_mode, P.p_scale, ptr(P.p_ctrl), P.e_mode, P.e_scale, ptr(P.e_ctrl),
                                          ptr(dur), ptr(f0), ptr(energy), ptr(ws), ws.numel(), stream),
                  "vs_variance_adapter")
            mark("variance_adapter", lat)
            cum = torch.empty(Rp, dtype=torch.int32, device=dev)
            frames_d = torch.empty(B, dtype=torch.int32, device=dev)
            check(lib.vs_length_regulate_count(ctypes.byref(rp.struct), ptr(dur), ptr(cum), ptr(frames_d), stream),
                  "vs_length_regulate_count")
            if P.rf is None:
                P.frames = frames_d.cpu().numpy()        # predicted durations: the one host sync of the path
                self._layout_frames(P)
                ws = self._workspace_lat(Rp, P.rf.n_rows) if overlap else self._workspace(Rp, P.rf.n_rows)
            rf, frames = P.rf, P.frames
            # bucketed layouts (infer_graphed): output shapes follow the bucket, not the utterance, so a captured graph fits
            # every utterance of the bucket; the caller slices
            Rf, Tf = rf.n_rows, (rf.n_rows if P.bucket else int(frames.max()))
            x_f = torch.empty(Rf, 192, dtype=torch.float32, device=dev)
            lr_index = torch.empty(Rf, dtype=torch.int32, device=dev)
            check(lib.vs_length_regulate_gather(ctypes.byref(rp.struct), ctypes.byref(rf.struct), ptr(x), ptr(cum),
                                                ptr(x_f), ptr(lr_index), stream), "vs_length_regulate_gather")
            mark("length_regulator", lat)
            # eps of models.py:718: injected (parity runs) or drawn inside the sampling kernel (Philox keyed by this seed);
            # seeds come from torch's default CPU generator, so torch.manual_seed() makes a run reproducible
            seed = 0 if P.eps is not None else self._next_seed()
            m_p = torch.empty(Rf, 192, dtype=torch.float32, device=dev)
            logs_p = torch.empty_like(m_p)
            z = torch.empty_like(m_p)
            check(lib.vs_frame_prior(self._model, ctypes.byref(rf.struct), ptr(x_f), ptr(P.eps), seed, P.noise_scale,
                                     ptr(x_f), ptr(m_p), ptr(logs_p), ptr(z), ptr(ws), ws.numel(), stream),
                  "vs_frame_prior")
            mark("frame_prior", lat)
            z_p = z.clone() if outputs == "all" else None
            check(lib.vs_flow_reverse(self._model, ctypes.byref(rf.struct), ptr(z), ptr(ws), ws.numel(), stream),
                  "vs_flow_reverse")
            mark("flow", lat)
        if overlap:
            done = torch.cuda.Event()
            done.record(lat)
            main.wait_event(done)
            for t in (z, z_p, m_p, logs_p, dur, f0, energy, lr_index, rf.row_utt, rp.row_utt):
                if t is not None:
                    t.record_stream(main)      # allocated on the side stream, read by the decoder / unpack kernels on `main`
        self.last_rows = (rp, rf)
        self.last_lr_index = lr_index
        if outputs == "latents":               # infer_stream: the decoder runs chunk by chunk on these rows
            if timings is not None:
                timings["_events"] = ev
            return z, rf

        with torch.cuda.device(dev):
            stream = main.cuda_stream
            ws = self._workspace(Rp, Rf)
            if overlap:
                mark("decoder_start", main)
            wave = torch.empty(Rf * self.hop_length, dtype=torch.float32, device=dev)
            ml = P.max_len
            check(lib.vs_hifigan_decode(self._model, ctypes.byref(rf.struct), ptr(z), ml, ptr(wave),
                                        int(self.decoder_precision), ptr(ws), ws.numel(), stream), "vs_hifigan_decode")
            mark("decoder", main)
            if overlap:
                dec_done = torch.cuda.Event()
                dec_done.record(main)
                self._in_flight.append(dec_done)

            def unpack(src, C, mul, t_max):
                out = torch.empty(B, C, t_max, dtype=torch.float32, device=dev)
                check(lib.vs_unpack_rows(ctypes.byref(rf.struct), ptr(src), C, mul, t_max, ptr(out), stream),
                      "vs_unpack_rows")
                return out

            t_dec = Tf if ml < 0 else max(0, min(Tf, ml))
            if t_dec == 0:
                o = torch.zeros(B, 1, 0, dtype=torch.float32, device=dev)
            else:
                o = unpack(wave, 1, self.hop_length, t_dec * self.hop_length)
            mark("unpack", main)
            if timings is not None:
                timings["_events"] = ev
            x_mask = (torch.arange(Tf, device=dev)[None, :] < P.frames_dev()[:, None])[:, None, :]
            if outputs != "all":
                return o, x_mask, None, None, None, None
            lat_out = tuple(unpack(t, 192, 1, Tf) for t in (z, z_p, m_p, logs_p))
            lens = P.lens

            def phon_out(t):                       # ragged [Rp] -> [B,Tp] (zero at pads)
                res = torch.zeros(B, Tp, dtype=t.dtype, device=dev)
                idx_b = torch.from_numpy(np.repeat(np.arange(B), lens)).to(dev)
                idx_t = torch.from_numpy(np.concatenate([np.arange(n) for n in lens])).to(dev)
                src = torch.from_numpy(np.concatenate([np.arange(s, s + n) for s, n in zip(rp.starts, lens)])).to(dev)
                res[idx_b, idx_t] = t[src]
                return res

            if isinstance(P.duration_control, torch.Tensor):
                duration = P.duration_control                     # returned verbatim (models.py:682, Q7)
            else:
                duration = phon_out(dur).to(torch.float32)[:, None, :]
            return o, x_mask, lat_out, duration, phon_out(f0), phon_out(energy)

    @staticmethod
    def _next_seed() -> int:
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    def _side_stream(self) -> torch.cuda.Stream:
        if self._lat_stream is None:
            self._lat_stream = torch.cuda.Stream(device=self.device)
        return self._lat_stream

    @staticmethod
    def resolve_timings(timings: dict) -> Dict[str, float]:
        """After a synchronize: milliseconds per subsystem from the CUDA events `run(timings=...)` recorded."""
        ev = timings.pop("_events")
        out = {}
        for (_, a), (name, b) in zip(ev[:-1], ev[1:]):
            if name == "decoder_start":            # overlap_calls: the hand-over between the two streams is not a stage
                continue
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    @torch.no_grad()
    def infer(self, phonemes, phonemes_lengths, sid=None, noise_scale=1, max_len=None, energy_control: Control = None,
              pitch_control: Control = None, duration_control: Control = None, noise=None, outputs: str = "all"):
        """Same arguments and return tuple as reference models.py:672-722:
        (o [B,1,hop*Tf'], x_mask bool [B,1,Tf], (z, z_p, m_p, logs_p) [B,192,Tf], duration, F0 [B,Tp], energy [B,Tp]).
        `noise`: optional eps, a [B,192,Tf] tensor or a list of [192,Tf_b] tensors.  `outputs="audio"` skips the
        latents (returns None in their place).
        """
        P = self.prepare(phonemes, phonemes_lengths, sid, noise_scale, max_len, energy_control, pitch_control,
                         duration_control, noise)
        return self.run(P, outputs)

    # ------------------------------------------------------------------------------------------------
    # Latency path: the whole device side of one call as ONE CUDA-graph launch per row bucket
    # ------------------------------------------------------------------------------------------------
    GRAPH_BUCKET_PHONEME_ROWS = 16
    GRAPH_BUCKET_FRAME_ROWS = 64

    @torch.no_grad()
    def infer_graphed(self, phonemes, phonemes_lengths, sid=None, noise_scale=1, energy_control: Control = None,
                      pitch_control: Control = None, duration_control: Control = None, noise=None):
        """`infer(..., outputs="audio")` for the latency path (every reference call site is one utterance per call,
        inference.py:40-44): the ~250 kernel launches of a call are captured once per ROW BUCKET (phoneme rows rounded up
        to 16, frame rows to 64; the tail of a bucket is gap rows, which every kernel already treats as padding) and
        replayed as a single graph launch - same kernels, same arithmetic as `infer` on the same layout.  Needs durations
        as a tensor (the frame layout must be known on the host before the launch).  Returns (o [B,1,hop*Tf], x_mask)."""
        if not isinstance(duration_control, torch.Tensor):
            raise ValueError("infer_graphed needs duration_control as a tensor (frame counts fix the graph's bucket)")
        from .layout import plan_starts
        from .sharding import frames_from_durations
        lens = phonemes_lengths.detach().cpu().numpy().astype(np.int32).reshape(-1)
        B = int(lens.shape[0])
        dc = duration_control.detach()
        dc = dc.reshape(dc.shape[0], -1) if dc.dim() == 3 else (dc[None] if dc.dim() == 1 else dc)
        frames = frames_from_durations([dc[b, :lens[b]].cpu().numpy() for b in range(B)])
        up = lambda n, m: -(-int(n) // m) * m
        bucket = (up(plan_starts(lens, PHONEME_GAP)[1], self.GRAPH_BUCKET_PHONEME_ROWS),
                  up(plan_starts(frames, FRAME_GAP)[1], self.GRAPH_BUCKET_FRAME_ROWS))
        sig = tuple(isinstance(c, torch.Tensor) or (1.0 if c is None else float(c)) for c in (pitch_control, energy_control))
        key = (B, int(phonemes.shape[-1]), bucket, sig, float(noise_scale), int(self.decoder_precision))
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        ent = self._graphs.get(key)
        dev = self.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            if ent is None:
                # eps always comes from a static buffer here (a kernel argument such as the Philox seed would be frozen)
                eps0 = torch.zeros(B, 192, int(frames.max()))
                P = self.prepare(phonemes, phonemes_lengths, sid, noise_scale, None, energy_control, pitch_control,
                                 duration_control, eps0, bucket=bucket)
                self.run(P, outputs="audio")                     # eager once: kernel attributes, allocator warm-up
                torch.cuda.synchronize(dev)
                saved = (self._ws, self.overlap_calls)
                self._ws, self.overlap_calls = None, False       # the graph owns its workspace (allocated from its private pool)
                g = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(g, stream=self._graph_stream()):
                        o, _, _, _, _, _ = self.run(P, outputs="audio")
                    ws = self._ws
                finally:
                    self._ws, self.overlap_calls = saved
                ent = self._graphs[key] = dict(graph=g, P=P, o=o, ws=ws)
            P = self.prepare(phonemes, phonemes_lengths, sid, noise_scale, None, energy_control, pitch_control,
                             duration_control, None, bucket=bucket, static=ent["P"])
            sP = ent["P"]
            if noise is not None:
                per = [(noise[b] if not isinstance(noise, torch.Tensor) else noise[b])[:, :frames[b]].t().cpu().numpy() for b in range(B)]
                _upload(P.rf.scatter(per, np.float32, width=192), dev, sP.eps)
            else:
                check(self._lib.vs_randn(sP.eps.data_ptr(), sP.eps.numel(), self._next_seed(), stream.cuda_stream), "vs_randn")
            ent["graph"].replay()
            tf = int(frames.max())
            o = ent["o"][:, :, : tf * self.hop_length]
            x_mask = (torch.arange(tf, device=dev)[None, :] < P.frames_dev()[:, None])[:, None, :]
            return o, x_mask

    def _graph_stream(self) -> torch.cuda.Stream:
        if getattr(self, "_gstream", None) is None:
            self._gstream = torch.cuda.Stream(device=self.device)
        return self._gstream

    # ------------------------------------------------------------------------------------------------
    # Long-form path (BASELINE.json configs[3]): chunked decoder with receptive-field overlap
    # ------------------------------------------------------------------------------------------------
    DECODER_HALO_FRAMES = 16     # >= the decoder's reach of +-12.33 frames, kept a multiple of 8

    @torch.no_grad()
    def infer_stream(self, phonemes, phonemes_lengths, sid=None, noise_scale=1, max_len=None,
                     energy_control: Control = None, pitch_control: Control = None, duration_control: Control = None,
                     noise=None, chunk_frames: int = 128, first_chunk_frames: Optional[int] = None,
                     halo_frames: Optional[int] = None):
        """Generator over waveform chunks of ONE utterance (every reference call site is batch 1: inference.py:44,
        inference_api.py:46): the latent path (text encoder ... flow, < 10 % of the work) runs once, then the HiFi-GAN
        decoder (models.py:271-290) runs on windows of `chunk_frames` frames extended by `halo_frames` of context on
        each side; the halo's samples are dropped.  The halo covers the decoder's whole receptive field, so the
        concatenated chunks are BIT-IDENTICAL to `infer(...)[0]` (tests/test_gpu_configs.py) while the first audio is
        available after one small decode instead of the whole utterance.
        Yields (first_sample, wave) with wave a float32 device tensor [n_samples]; the consumer may copy chunk i to the
        host while chunk i+1 is computed (each chunk owns its buffer)."""
        P = self.prepare(phonemes, phonemes_lengths, sid, noise_scale, None, energy_control, pitch_control,
                         duration_control, noise)
        if P.B != 1:
            raise ValueError("infer_stream synthesises one utterance per call (batch 1, like the reference call sites)")
        halo = self.DECODER_HALO_FRAMES if halo_frames is None else int(halo_frames)
        if chunk_frames < 1 or halo < 0:
            raise ValueError("chunk_frames must be >= 1 and halo_frames >= 0")
        z, rf = self.run(P, outputs="latents")
        lib, dev, hop = self._lib, self.device, self.hop_length
        tf = int(P.frames[0])
        if max_len is not None:
            tf = max(0, min(tf, int(max_len)))                       # models.py:720 truncates before the decoder
        z0 = int(rf.starts[0])
        s = 0
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            while s < tf:
                n = chunk_frames if (s > 0 or first_chunk_frames is None) else int(first_chunk_frames)
                e = min(tf, s + max(1, n))
                lo, hi = max(0, s - halo), min(tf, e + halo)
                rows = make_rows([hi - lo], P.sids, FRAME_GAP, dev)
                zc = torch.zeros(rows.n_rows, 192, dtype=torch.float32, device=dev)
                zc[: hi - lo] = z[z0 + lo: z0 + hi]
                ws = self._workspace(P.rp.n_rows, max(rows.n_rows, rf.n_rows))
                wave = torch.empty(rows.n_rows * hop, dtype=torch.float32, device=dev)
                check(lib.vs_hifigan_decode(self._model, ctypes.byref(rows.struct), ptr(zc), -1, ptr(wave),
                                            int(self.decoder_precision), ptr(ws), ws.numel(), stream), "vs_hifigan_decode")
                yield s * hop, wave[(s - lo) * hop: (e - lo) * hop]
                s = e

    # ------------------------------------------------------------------------------------------------
    # 8(f): voice_conversion (reference models.py:724-732)
    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def voice_conversion(self, y, y_lengths, sid_src, sid_tgt, noise=None):
        """y: linear spectrogram [B, spec_channels, T]; returns (o_hat [B,1,hop*T], y_mask [B,1,T] float,
        (z, z_p, z_hat) [B,192,T]) like the reference.  `noise`: optional eps of PosteriorEncoder (models.py:239),
        [B,192,T] tensor or list of [192,T_b].  Needs a checkpoint that still has enc_q.* (training checkpoints do)."""
        if not self._loaded:
            raise _lib.VsError("load_state_dict()/load_checkpoint() first")
        if "enc_q.pre.w" not in self._weights:
            raise _lib.VsError("voice_conversion needs the enc_q.* (posterior encoder) weights in the state dict")
        lib, dev = self._lib, self.device
        B, C, T = y.shape
        lens = y_lengths.detach().cpu().numpy().astype(np.int32).reshape(-1)
        src = sid_src.detach().cpu().numpy().astype(np.int32).reshape(-1)
        tgt = sid_tgt.detach().cpu().numpy().astype(np.int32).reshape(-1)
        if lens.shape[0] != B or src.shape[0] != B or tgt.shape[0] != B or lens.min() < 1 or lens.max() > T:
            raise ValueError("y_lengths / sid_src / sid_tgt must have one valid entry per utterance")
        c_pad = self._weights["enc_q.pre.w"].shape[0]
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rows_src = make_rows(lens, src, FRAME_GAP, dev)
            rows_tgt = make_rows(lens, tgt, FRAME_GAP, dev)       # same layout, target speakers
            R = rows_src.n_rows
            yh = y.detach().float().cpu().numpy()
            spec = torch.from_numpy(rows_src.scatter([np.pad(yh[b, :, :lens[b]].T, ((0, 0), (0, c_pad - C))) for b in range(B)],
                                                     np.float32, width=c_pad)).to(dev)
            if noise is None:
                eps = None
            else:
                per = [(noise[b] if not isinstance(noise, torch.Tensor) else noise[b])[:, :lens[b]].t().cpu().numpy()
                       for b in range(B)]
                eps = _upload(rows_src.scatter(per, np.float32, width=192), dev)
            ws = self._workspace(16, R)
            z = torch.empty(R, 192, dtype=torch.float32, device=dev)
            m_q, logs_q = torch.empty_like(z), torch.empty_like(z)
            check(lib.vs_posterior_encode(self._model, ctypes.byref(rows_src.struct), ptr(spec), ptr(eps),
                                          0 if eps is not None else self._next_seed(), ptr(z), ptr(m_q),
                                          ptr(logs_q), ptr(ws), ws.numel(), stream), "vs_posterior_encode")
            z_p = z.clone()
            check(lib.vs_flow_forward(self._model, ctypes.byref(rows_src.struct), ptr(z_p), ptr(ws), ws.numel(), stream),
                  "vs_flow_forward")
            z_hat = z_p.clone()
            check(lib.vs_flow_reverse(self._model, ctypes.byref(rows_tgt.struct), ptr(z_hat), ptr(ws), ws.numel(), stream),
                  "vs_flow_reverse")
            wave = torch.empty(R * self.hop_length, dtype=torch.float32, device=dev)
            check(lib.vs_hifigan_decode(self._model, ctypes.byref(rows_tgt.struct), ptr(z_hat), -1, ptr(wave),
                                        int(self.decoder_precision), ptr(ws), ws.numel(), stream), "vs_hifigan_decode")

            def unpack(srct, Cc, mul, t_max):
                out = torch.empty(B, Cc, t_max, dtype=torch.float32, device=dev)
                check(lib.vs_unpack_rows(ctypes.byref(rows_src.struct), ptr(srct), Cc, mul, t_max, ptr(out), stream),
                      "vs_unpack_rows")
                return out

            tm = int(lens.max())
            o = unpack(wave, 1, self.hop_length, tm * self.hop_length)
            y_mask = (torch.arange(tm, device=dev)[None, :] < torch.from_numpy(lens).to(dev)[:, None])[:, None, :].float()
            return o, y_mask, tuple(unpack(t, 192, 1, tm) for t in (z, z_p, z_hat))


class Prepared:
    """Device-resident inputs + row layouts of one `infer` call (see SynthesizerTrn.prepare)."""

    def frames_dev(self) -> torch.Tensor:
        return self.rf.utt_len                      # int32 [B] on the device: the frame counts of the row layout


def _upload(a: np.ndarray, dev, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    # always through pinned memory: a copy from pageable memory first synchronises the stream it is queued on, i.e. the
    # host would wait for whatever the previous call left there (see layout.make_rows)
    src = torch.from_numpy(a).pin_memory()
    if out is not None:                        # refill a static buffer (CUDA-graph replays read it)
        out.copy_(src, non_blocking=True)
        return out
    return src.to(dev, non_blocking=True)


def load_checkpoint(checkpoint_path: str, model: SynthesizerTrn, optimizer=None, skip_optimizer: bool = False):
    """utils.load_checkpoint (utils.py:21-51) for this model: reads ckpt['model'], tolerant of extra keys.
    Returns (model, optimizer, learning_rate, iteration) like the reference."""
    ckpt = torch.load(checkpoint_path, map_location="cpu")
    model.load_state_dict(ckpt["model"])
    return model, optimizer, ckpt.get("learning_rate"), ckpt.get("iteration")
