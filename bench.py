#!/usr/bin/env python
"""Benchmark of the SynthesizerTrn.infer hot path (BASELINE.json metric: synthesized audio-seconds per second at
44.1 kHz; p50 RTF per utterance).

  python bench.py --gpus N --steps K --warmup W            # our arm (under torchrun for N > 1: one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU oracle port on the host cores

Workload = BASELINE.json configs[1]: a batch of 64 synthetic ~5 s utterances (40 phonemes, given MFA-style durations,
predicted pitch/energy, noise_scale .667, random-init weights of configs/config.json, seed 1234), per GPU (weak scaling:
utterances are independent, sharded by rank, no collective on the data path).

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
DEC_DRAM_BYTES_C2 = 65.63e9
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
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("VS_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its banner to stdout: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    from oracle import inputs as oin                      # synthetic workload generator + CPU baseline only
    from oracle.weights import make_state_dict
    from vispeech_b200 import _lib, build_from_hparams, get_hparams_from_file

    sd = make_state_dict(1234)
    net = build_from_hparams(get_hparams_from_file(), device=dev)
    net.load_state_dict(sd)
    lib = _lib.load()
    for opt in ("fused_respair", "tf32_min_rows", "x3_min_rows", "tf32_prior", "decoder_streams", "attention_mma", "wn_fused", "tf32_cluster"):     # A/B knobs for profiling runs (defaults otherwise)
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

    # ---------------- reduce over ranks: time = max, work = sum
    stats = torch.tensor([ms_total, e2e_s * 1000.0, dec_ms], dtype=torch.float64, device=dev)
    work = torch.tensor([audio_s, float(sum(frames))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, dec_ms_max = [float(x) for x in stats.cpu()]
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
        rtf = sorted((ms_per_step * 1e-3) / (f * HOP / SR) for f in frames)
        line = {
            "metric": "synthesized audio-sec/sec at 44.1 kHz", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: B=%d x ~5 s utterances per GPU (40 phonemes, given durations, predicted "
                                   "pitch/energy, noise_scale .667), random-init configs/config.json seed 1234" % args.batch,
                       "utterances_per_gpu": args.batch, "valid_frames_per_gpu": int(sum(frames)),
                       "audio_s_per_step_all_gpus": audio_all, "parallelism": "utterance-sharded x%d, no collectives" % world,
                       "precision": "decoder: bf16 operands, fp32 accumulate (tcgen05 kind::f16); flow GEMMs: TF32; frame prior, projection, "
                                    "text encoder and predictors: 3xTF32 (error-compensated, tcgen05 kind::tf32); attention / layernorm: fp32",
                       "pipelining": "value/roofline: stages serialised on one stream; e2e: public infer() in throughput mode "
                                     "(latent stages of call i+1 overlap the decoder of call i on a second stream, waveform "
                                     "D2H of call i overlaps call i+1)" if net.overlap_calls else "none",
                       "l2": "per-step activations (~%.1f GB) >> 126 MB L2; no explicit flush needed" % (
                           sum(frames) * 112 * 16384 * 2 / 1e9 / 16)},
            "p50_rtf": rtf[len(rtf) // 2],
            "e2e": {"value": audio_all * args.steps / (e2e_ms * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "stages_ms": {k: round(v, 3) for k, v in stages.items()},
            "roofline": {"bound": "tensor", "kernel": "umma_conv1d_kernel + umma_respair_kernel (the whole decoder: ~60 launches per step)",
                         "achieved": dec_tflops, "peak": peak, "unit": "TFLOP/s", "frac": dec_tflops / peak,
                         "traffic": DEC_DRAM_BYTES_C2 if (args.batch == 64) else None,
                         "traffic_note": "DRAM read+write bytes of one decoder pass (all 61 launches), ncu, "
                                         "profiles/launches_r1_traffic.csv; the unfused algorithmic minimum is 0.07 GB "
                                         "(z in, waveform out): the rest is the activation stream between the ~60 convs",
                         "peak_source": peak_src,
                         "ms_per_step": stages["decoder"], "share_of_step": stages["decoder"] / ms_per_step},
            "clocks": clocks,
        }
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
