#!/usr/bin/env python
"""Latency / RTF of the batch-1 call sites (BASELINE.json configs[0] and configs[3]) on one GPU.

  python tools/latency.py [--reps 20]

For C1 (40 phonemes, ~4.9 s) and C4 (one 60 s utterance): wall-clock latency of the public `infer` call from host tensors
to the waveform in pinned host memory, its RTF, and for C4 the time to the FIRST audio chunk and the total time of the
chunked decoder (`infer_stream`, receptive-field overlap).  Prints one JSON line per case.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--first", type=int, default=32)
    args = ap.parse_args()
    import torch
    from oracle import inputs as oin
    from oracle.weights import make_state_dict
    from vispeech_b200 import build_from_hparams, get_hparams_from_file

    net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
    net.load_state_dict(make_state_dict(1234))
    for opt in ("pdl", "pair_conv", "pair_fused", "resblock_fused", "coupling_fused", "coupling_min_rows", "x3_min_rows", "tf32_min_rows", "conv_spread", "attention_small", "decoder_streams", "attention_mma"):       # A/B knobs: VS_PDL=37 python tools/latency.py
        if os.environ.get("VS_" + opt.upper()):
            net.set_option(opt, int(os.environ["VS_" + opt.upper()]))

    def med(f, reps):
        for _ in range(3):
            f()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e3, min(ts) * 1e3

    for name, u in (("C1", oin.c1()[0]), ("C4", oin.c4()[0])):
        tf = oin.frame_counts([u])[0]
        audio_s = tf * 512 / 44100
        a = (u["ids"][None], torch.LongTensor([u["ids"].numel()]))
        kw = dict(sid=torch.LongTensor([u["sid"]]), noise_scale=0.667, duration_control=u["duration"][None])
        host = torch.empty(tf * 512, dtype=torch.float32).pin_memory()

        def full():
            o, *_ = net.infer(*a, outputs="audio", **kw)
            host.copy_(o[0, 0], non_blocking=True)
            torch.cuda.synchronize()

        first_ms = []

        def stream():
            t0 = time.perf_counter()
            for i, (start, w) in enumerate(net.infer_stream(*a, chunk_frames=args.chunk, first_chunk_frames=args.first, **kw)):
                host[start:start + w.numel()].copy_(w, non_blocking=True)
                if i == 0:
                    torch.cuda.synchronize()
                    first_ms.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()

        full_ms, full_min = med(full, args.reps)
        line = {"case": name, "frames": tf, "audio_s": round(audio_s, 3), "infer_ms_p50": round(full_ms, 3),
                "infer_ms_min": round(full_min, 3), "rtf_p50": full_ms * 1e-3 / audio_s}
        stream_ms, _ = med(stream, max(3, args.reps // 2))
        line.update({"stream_total_ms_p50": round(stream_ms, 3), "stream_first_chunk_ms_p50": round(statistics.median(first_ms), 3),
                     "chunk_frames": args.chunk, "first_chunk_frames": args.first})
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
