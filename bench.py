#!/usr/bin/env python
"""Benchmark of the SynthesizerTrn.infer hot path (BASELINE.json metric: synthesized audio-seconds per second at
44.1 kHz; p50 RTF per utterance).

  python bench.py --gpus N --steps K --warmup W            # our arm (under torchrun for N > 1: one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU oracle port on the host cores

Workload = BASELINE.json configs[1]: a batch of 64 synthetic ~5 s utterances (40 phonemes, given MFA-style durations,
predicted pitch/energy, noise_scale .667, random-init weights of configs/config.json, seed 1234), per GPU (weak scaling:
utterances are independent, sharded by rank, no collective on the data path).

Extra fields of the same JSON line (our arm): `c3` = BASELINE.json configs[2] (512 mixed-length utterances, 1-15 s) STRONG-scaled
over the N ranks through `vispeech_b200.batching.synthesize` (length-bucketed LPT plan, waveforms to pinned host memory) with
per-rank times and the plan's imbalance; `p50_rtf` / `rtf` = per-utterance real-time factor of the batch-1 latency path;
`gpu_eager_baseline` (N = 1) = the reference's eager op sequence (oracle port on cuda: ATen / cuDNN, fp32 with TF32 allowed,
and bf16 autocast) on the same box - a labelled comparator, not the reference arm.

A "step" = one pass of the hot path over the batch.  `value` times the device work with inputs resident in HBM
(`SynthesizerTrn.run`); `e2e` times the public call `SynthesizerTrn.infer` from pinned HOST tensors plus the D2H copy
of the waveforms.  `roofline` is the decoder's tcgen05 conv kernel (the dominant kernel): algorithmic decoder FLOPs
(815,300,608 per valid frame) / CUDA-event time of vs_hifigan_decode on its stream.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEC_FLOP_PER_FRAME = 815_300_608          # SURVEY.md App. C (2 * 407,650,304 MAC)
# DRAM bytes (read + write) of one decoder pass over the default workload (64 utterances, 27,591 frames): sum of
# dram__bytes_read.sum + dram__bytes_write.sum over the decoder's 61 launches, profiles/launches_r1_traffic.csv
DEC_DRAM_PROFILE = os.path.join(ROOT, "profiles", "decoder_traffic_r2.json")   # {"batch": 64, "dram_bytes": ..., "source": ...}
HOP, SR = 512, 44100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step (configs[1]: 64)")
    ap.add_argument("--cpu-utts", type=int, default=3, help="utterances in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2] strong-scaling leg")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the GPU-eager comparator (N = 1 only)")
    ap.add_argument("--c3-reps", type=int, default=2)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_batch(utts):
    import torch
    B = len(utts)
    tp = max(u["ids"].numel() for u in utts)
    ids = torch.zeros(B, tp, dtype=torch.long)
    dur = torch.zeros(B, tp, dtype=torch.long)
    for b, u in enumerate(utts):
        n = u["ids"].numel()
        ids[b, :n], dur[b, :n] = u["ids"], u["duration"]
    lens = torch.LongTensor([u["ids"].numel() for u in utts])
    sid = torch.LongTensor([u["sid"] for u in utts])
    return [t.pin_memory() for t in (ids, lens, sid, dur)]


def cpu_oracle_time(sd, utts, repeats=1):
    """Wall time of the CPU oracle port (per-utterance batch-1 loop, as every reference call site runs)."""
    import torch
    from oracle.vispeech_oracle import infer_one
    t0 = time.perf_counter()
    for _ in range(repeats):
        for u in utts:
            infer_one(sd, u["ids"], u["sid"], 0.667, None, duration_control=u["duration"])
    return (time.perf_counter() - t0) / repeats


def run_c3(net, rank, world, dev, reps, dist):
    """configs[2]: the SAME 512 utterances on every rank, each rank synthesises its shard of the LPT plan (no data-path
    collective), waveforms copied to pinned host memory; time = slowest rank, work = all 512 utterances."""
    import torch
    from oracle import inputs as oin
    from vispeech_b200.batching import synthesize
    utts = oin.c3(batch=512, seed=2)
    audio_s = oin.audio_seconds(utts)
    pool, stats = {}, {}
    synthesize(net, utts, rank=rank, world_size=world, max_frames_per_batch=32768, host_pool=pool)      # warm-up: workspaces, pinned buffers
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = synthesize(net, utts, rank=rank, world_size=world, max_frames_per_batch=32768, host_pool=pool, stats=stats)
    torch.cuda.synchronize()
    mine_s = (time.perf_counter() - t0) / reps
    t = torch.tensor([mine_s, float(stats["frames"]), float(stats["infer_calls"]), float(stats["d2h_bytes"])], dtype=torch.float64, device=dev)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    per = [[float(v) for v in x.cpu()] for x in allt]
    slowest = max(p[0] for p in per)
    return {"workload": "configs[2]: 512 utterances of 1-15 s (seed 2), given durations, length-bucketed (<= 32768 frames per call), "
                        "LPT-sharded over %d rank(s), waveforms fp32 to pinned host memory" % world,
            "scaling": "strong", "value": audio_s / slowest, "unit": "audio-s/s", "audio_s": audio_s, "seconds": slowest,
            "per_rank_seconds": [round(p[0], 4) for p in per], "per_rank_frames": [int(p[1]) for p in per],
            "per_rank_infer_calls": [int(p[2]) for p in per], "d2h_bytes_all_ranks": int(sum(p[3] for p in per)),
            "plan_imbalance": stats["imbalance"], "measured_imbalance": slowest / (sum(p[0] for p in per) / len(per)),
            "reps": reps}


def batch1_rtf(net, utts, n=8, graphed=False):
    """Per-utterance real-time factor of the latency path: public infer() on ONE utterance, host tensors in, waveform in
    pinned host memory, serialised (what every reference call site does, inference.py:40-44)."""
    import torch
    rtfs = []
    host = None
    for rep in range(3 if graphed else 2):             # first pass(es) = warm-up (graph capture per row bucket)
        rtfs = []
        for u in utts[:n]:
            ids, dur = u["ids"][None], u["duration"][None]
            lens, sid = torch.LongTensor([u["ids"].numel()]), torch.LongTensor([u["sid"]])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if graphed:
                o, _ = net.infer_graphed(ids, lens, sid=sid, noise_scale=0.667, duration_control=dur)
            else:
                o, *_ = net.infer(ids, lens, sid=sid, noise_scale=0.667, duration_control=dur, outputs="audio")
            if host is None or host.numel() < o.numel():
                host = torch.empty(o.numel() * 2, dtype=o.dtype).pin_memory()
                t0 = time.perf_counter()               # do not time the one-off pinned allocation
            host[: o.numel()].copy_(o.reshape(-1), non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            rtfs.append(dt / (int(u["duration"].sum()) * HOP / SR))
    rtfs.sort()
    return {"p50": rtfs[len(rtfs) // 2], "min": rtfs[0], "max": rtfs[-1], "n": len(rtfs),
            "path": ("batch-1 infer_graphed() (one CUDA-graph launch per call)" if graphed else "batch-1 infer()") +
                    ": host tensors in -> waveform in pinned host memory, one utterance (~5 s) per call"}


def gpu_eager(sd, utts, frames, dev, steps=3):
    """The reference's eager op sequence on THIS GPU (oracle port with its state dict on cuda: ATen conv1d / conv_transpose1d
    through cuDNN): (a) the HiFi-GAN decoder on the padded B x 192 x Tmax batch, fp32 with TF32 allowed and under bf16
    autocast; (b) the whole path as the per-utterance batch-1 loop every reference call site runs.  A comparator for our
    kernels, not the reference arm."""
    import torch
    from oracle.vispeech_oracle import generator, infer_one
    from oracle.weights import DEFAULT_CONFIG
    sd_d = {k: v.to(dev) for k, v in sd.items()}
    B, tmax = len(utts), max(frames)
    audio = sum(frames) * HOP / SR
    g = torch.Generator(device=dev).manual_seed(0)
    z = torch.randn(B, 192, tmax, device=dev, generator=g)
    for b, f in enumerate(frames):
        z[b, :, f:] = 0
    gvec = sd_d["emb_g.weight"][torch.tensor([u["sid"] for u in utts], device=dev)].unsqueeze(-1)
    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad():
            for name, ctx in (("decoder_fp32_tf32", None), ("decoder_bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
                def run():
                    if ctx is None:
                        return generator(sd_d, z, gvec, DEFAULT_CONFIG)
                    with ctx:
                        return generator(sd_d, z, gvec, DEFAULT_CONFIG)
                try:
                    run(); run()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        run()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    res[name] = {"ms_per_step": ms, "audio_s_per_s": audio / (ms * 1e-3),
                                 "tflops": DEC_FLOP_PER_FRAME * sum(frames) / (ms * 1e-3) / 1e12}
                except Exception as e:                  # e.g. out of memory on a smaller part: report, do not fail the bench
                    res[name] = {"error": str(e).splitlines()[0][:200]}
                torch.cuda.empty_cache()
            torch.backends.cudnn.benchmark = False         # every utterance has its own length: no per-shape autotuning
            sample = utts[:4]
            for u in sample[:1]:
                infer_one(sd_d, u["ids"].to(dev), u["sid"], 0.667, None, duration_control=u["duration"].to(dev))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for u in sample:
                infer_one(sd_d, u["ids"].to(dev), u["sid"], 0.667, None, duration_control=u["duration"].to(dev))["o"].cpu()
            dt = time.perf_counter() - t0
            a = sum(int(u["duration"].sum()) for u in sample) * HOP / SR
            res["full_path_batch1_loop"] = {"audio_s_per_s": a / dt, "utterances": len(sample),
                                            "note": "infer_one per utterance on cuda (fp32, TF32 allowed), waveform .cpu()"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    res["what"] = ("oracle port (the reference's eager ATen/cuDNN op sequence, weight norm folded per call as its forward hooks do) "
                   "on the same GPU, same batch (B=%d, %d valid frames, padded to %d frames per utterance)" % (B, sum(frames), tmax))
    return res


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The reference is pure Python/PyTorch and cannot
    travel to the GPU box, so this times the oracle port (pinned to the reference by tests/golden) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import inputs as oin
    from oracle.weights import make_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(1234)
    utts = oin.c2(batch=args.batch, seed=1)[: max(1, min(args.cpu_utts, 2))]
    audio = oin.audio_seconds(utts)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_time(sd, utts[:1])
    times = [cpu_oracle_time(sd, utts) for _ in range(args.steps)]
    total = sum(times)
    value = audio * args.steps / total
    line = {
        "impl": "reference", "metric": "synthesized audio-sec/sec at 44.1 kHz", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: B=64 x ~5 s utterances, 40 phonemes, given durations, random-init config.json",
                   "sample": "%d utterances of that batch per step (%.1f audio-s)" % (len(utts), audio)},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": "%d utterances (%.1f audio-s) per step, %d steps, per-utterance batch-1 loop" % (
                             len(utts), audio, args.steps)},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL's version banner, library chatter) has been
    diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)                   # keep stdout for the JSON line only
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it: NCCL's banner goes to fd 1, which main() has pointed at stderr, so the
        # process's real stdout still carries only the JSON line
        dist.init_process_group("nccl", device_id=dev)

    from oracle import inputs as oin                      # synthetic workload generator + CPU baseline only
    from oracle.weights import make_state_dict
    from vispeech_b200 import _lib, build_from_hparams, get_hparams_from_file

    sd = make_state_dict(1234)
    net = build_from_hparams(get_hparams_from_file(), device=dev)
    net.load_state_dict(sd)
    lib = _lib.load()
    for opt in ("fused_respair", "tf32_min_rows", "x3_min_rows", "tf32_prior", "decoder_streams", "attention_mma", "wn_fused", "tf32_cluster", "mrf_fused", "split16", "resblock_fused", "pair_conv", "pair_fused", "coupling_fused", "pdl", "tap_pairs", "conv_spread", "attention_small"):     # A/B knobs for profiling runs (defaults otherwise)
        if os.environ.get("VS_" + opt.upper()):
            _lib.check(lib.vs_set_option(opt.encode(), int(os.environ["VS_" + opt.upper()])))

    utts = oin.c2(batch=args.batch, seed=1 + rank)        # independent utterances per rank: no collective on the path
    frames = oin.frame_counts(utts)
    audio_s = oin.audio_seconds(utts)
    ids, lens, sid, dur = make_batch(utts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: inputs already in HBM when the timed region starts
    P = net.prepare(ids, lens, sid=sid, noise_scale=0.667, duration_control=dur)
    for _ in range(args.warmup):
        net.run(P, outputs="audio")
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.vs_launch_count()
    timing_list = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        t = {}
        net.run(P, outputs="audio", timings=t)
        timing_list.append(t)
    e1.record()
    barrier()
    launches = lib.vs_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    stages = {}
    for t in timing_list:
        for k, v in net.resolve_timings(t).items():
            stages[k] = stages.get(k, 0.0) + v / args.steps
    dec_ms = stages.get("decoder", float("nan"))

    # ---------------- end-to-end arm: public API from pinned host tensors, waveforms copied back to pinned host memory.
    # The engine runs in its throughput mode here (overlap_calls: the latent stages of call i+1 are enqueued on a second
    # stream and overlap the decoder of call i; VS_OVERLAP=0 turns it off for A/B runs).  The device-resident arm above
    # keeps the stages serialised on one stream so that the stage times and the roofline stay attributable.
    rtf = batch1_rtf(net, utts) if rank == 0 else None   # latency mode (no cross-call overlap), one utterance per call
    rtf_graph = batch1_rtf(net, utts, graphed=True) if rank == 0 else None
    net.overlap_calls = os.environ.get("VS_OVERLAP", "1") != "0"
    o, _, _, _, _, _ = net.infer(ids, lens, sid=sid, noise_scale=0.667, duration_control=dur, outputs="audio")
    host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for _ in range(2)]
    h2d = int(P.ids_rows.numel() * 4 + P.d_ctrl.numel() * 8 + 2 * (P.rp.n_rows + 3 * P.B) * 4 + P.rf.n_rows * 4)
    d2h = int(o.numel() * 4)
    copy_stream = torch.cuda.Stream(device=dev)

    in_flight = []                                        # copy-done events of the calls not yet retired

    def e2e_step(i):
        # the waveforms of step i travel to pinned host memory on a second stream while step i+1 computes.  The host may
        # run at most two calls ahead of the GPU (a bounded request queue): that keeps the device allocator in steady
        # state - an unbounded run-ahead makes it cudaMalloc, i.e. synchronise, for every call still in flight.
        if len(in_flight) >= 2:
            in_flight.pop(0).synchronize()
        o, *_ = net.infer(ids, lens, sid=sid, noise_scale=0.667, duration_control=dur, outputs="audio")
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            host_out[i % 2].copy_(o, non_blocking=True)
            o.record_stream(copy_stream)
            copied = torch.cuda.Event()
            copied.record(copy_stream)
        in_flight.append(copied)

    for i in range(max(3, args.warmup)):
        e2e_step(i)
    barrier()
    in_flight.clear()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0

    # the same loop with the waveforms converted to s16 on the device (vs_wave_pcm16) before the copy: half the D2H bytes
    from vispeech_b200.postprocess import to_pcm16
    n_valid = [int(f) * HOP for f in frames]
    host_pcm = [torch.empty(o.shape[0], o.shape[2], dtype=torch.int16).pin_memory() for _ in range(2)]

    def e2e_pcm_step(i):
        if len(in_flight) >= 2:
            in_flight.pop(0).synchronize()
        o, *_ = net.infer(ids, lens, sid=sid, noise_scale=0.667, duration_control=dur, outputs="audio")
        pcm = to_pcm16(o, n_valid, SR, SR)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            host_pcm[i % 2].copy_(pcm, non_blocking=True)
            pcm.record_stream(copy_stream)
            copied = torch.cuda.Event()
            copied.record(copy_stream)
        in_flight.append(copied)

    in_flight.clear()
    for i in range(3):
        e2e_pcm_step(i)
    barrier()
    in_flight.clear()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_pcm_step(i)
    barrier()
    e2e_pcm_s = time.perf_counter() - t0
    in_flight.clear()

    c3 = None if args.no_c3 else run_c3(net, rank, world, dev, args.c3_reps, dist)

    # ---------------- reduce over ranks: time = max, work = sum
    stats = torch.tensor([ms_total, e2e_s * 1000.0, dec_ms, e2e_pcm_s * 1000.0], dtype=torch.float64, device=dev)
    work = torch.tensor([audio_s, float(sum(frames))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, dec_ms_max, e2e_pcm_ms = [float(x) for x in stats.cpu()]
    audio_all, frames_all = [float(x) for x in work.cpu()]

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            pk = json.load(open(peaks_path))
            peak, peak_src = float(pk["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
        else:
            peak, peak_src = 1400.0, "fallback sustained figure of B200_PROFILING.md (of fallback)"
        # roofline of the dominant kernel, per GPU: this rank's decoder FLOPs / this rank's decoder time
        dec_tflops = DEC_FLOP_PER_FRAME * sum(frames) / (stages["decoder"] * 1e-3) / 1e12
        ms_per_step = ms_total / args.steps
        value = audio_all * args.steps / (ms_total * 1e-3)
        traffic, traffic_note = None, "no ncu capture of this commit's decoder for this batch size (profiles/decoder_traffic_r2.json)"
        if os.path.exists(DEC_DRAM_PROFILE):
            tp = json.load(open(DEC_DRAM_PROFILE))
            if int(tp.get("batch", -1)) == args.batch:
                traffic, traffic_note = float(tp["dram_bytes"]), tp.get("source", "")
        line = {
            "metric": "synthesized audio-sec/sec at 44.1 kHz", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "configs[1]: B=%d x ~5 s utterances per GPU (40 phonemes, given durations, predicted "
                                   "pitch/energy, noise_scale .667), random-init configs/config.json seed 1234" % args.batch,
                       "utterances_per_gpu": args.batch, "valid_frames_per_gpu": int(sum(frames)),
                       "audio_s_per_step_all_gpus": audio_all, "parallelism": "utterance-sharded x%d, no collectives" % world,
                       "precision": "decoder: fp16 operands, fp32 accumulate (tcgen05 kind::f16); the last MRF stage and the k=3 ResBlock of the C=64 stage keep their residual stream in fp32 in TMEM; "
                                    "flow GEMMs: TF32 (kind::tf32); frame prior, projection, text encoder and predictors: three-term fp16 hi/lo products on kind::f16 (fp32-level accuracy); "
                                    "attention: the same three-term form on tcgen05 at frame level, fp32 CUDA cores at phoneme level; layernorm fp32",
                       "pipelining": "value/roofline: stages serialised on one stream; e2e: public infer() in throughput mode "
                                     "(latent stages of call i+1 overlap the decoder of call i on a second stream, waveform "
                                     "D2H of call i overlaps call i+1)" if net.overlap_calls else "none",
                       "l2": "per-step activations (~%.1f GB) >> 126 MB L2; no explicit flush needed" % (
                           sum(frames) * 112 * 16384 * 2 / 1e9 / 16)},
            "p50_rtf": min(rtf["p50"], rtf_graph["p50"]), "rtf": rtf, "rtf_cuda_graph": rtf_graph,
            "e2e": {"value": audio_all * args.steps / (e2e_ms * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "pcm16": {"value": audio_all * args.steps / (e2e_pcm_ms * 1e-3), "d2h_bytes_per_step": d2h // 2,
                              "note": "same loop, waveform converted to s16 on the device (vs_wave_pcm16) before the D2H copy"}},
            "gpu_launches": int(launches),
            "stages_ms": {k: round(v, 3) for k, v in stages.items()},
            "roofline": {"bound": "tensor", "kernel": "the whole decoder, 47 launches per step: umma_conv1d_kernel (conv_pre, ups, C=256 stage), umma_pair_kernel + umma_pairfused_kernel (C=128 stage on CTA pairs, "
                                   "tcgen05 cta_group::2), umma_resblock_kernel + umma_respair_kernel (C=64 stage), umma_mrf_kernel (C=32 stage + conv_post + tanh)",
                         "achieved": dec_tflops, "peak": peak, "unit": "TFLOP/s", "frac": dec_tflops / peak,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": peak_src,
                         "ms_per_step": stages["decoder"], "share_of_step": stages["decoder"] / ms_per_step},
            "clocks": clocks,
        }
        if c3 is not None:
            line["c3"] = c3
        if world == 1 and not args.no_gpu_eager:
            net._ws = net._ws_lat = None                # hand the workspaces back before the eager decoder allocates ~20 GB
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager(sd, utts, frames, dev)
            ge = line["gpu_eager_baseline"].get("decoder_bf16_autocast", {})
            if "ms_per_step" in ge:
                line["gpu_eager_baseline"]["our_decoder_speedup_vs_bf16_autocast"] = ge["ms_per_step"] / stages["decoder"]
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sample = utts[: args.cpu_utts]
            cpu_oracle_time(sd, sample[:1])
            t_cpu = cpu_oracle_time(sd, sample)
            line["cpu_baseline"] = {"value": oin.audio_seconds(sample) / t_cpu, "unit": "audio-s/s", "cores": cores,
                                    "kind": "port",
                                    "sample": "first %d utterances of the batch (%.1f audio-s), oracle port, "
                                              "per-utterance batch-1 loop, 1 warm-up" % (len(sample), oin.audio_seconds(sample))}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
