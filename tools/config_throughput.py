#!/usr/bin/env python
"""Throughput of BASELINE.json configs[2] (C3: 512 utterances of 1-15 s, length-bucketed) and configs[4] (C5: 256
manual-edit utterances) on ONE GPU through `vispeech_b200.batching.synthesize` (host tensors in, waveforms left on the
device), in the engine's throughput mode.  One JSON line per config.   python tools/config_throughput.py [--reps 3]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max-frames", type=int, default=32768)
    args = ap.parse_args()
    import torch
    from oracle import inputs as oin
    from oracle.weights import make_state_dict
    from vispeech_b200 import build_from_hparams, get_hparams_from_file
    from vispeech_b200.batching import synthesize
    from vispeech_b200.sharding import bucket_batches, frames_from_durations
    net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
    net.load_state_dict(make_state_dict(1234))
    net.overlap_calls = True
    for name, utts in (("C3", oin.c3()), ("C5", oin.c5())):
        frames = frames_from_durations([u["duration"] for u in utts])
        audio_s = float(frames.sum()) * 512 / 44100
        n_calls = len(bucket_batches(range(len(utts)), frames, args.max_frames))
        synthesize(net, utts, max_frames_per_batch=args.max_frames, keep_on_device=True)      # warm-up (workspace sizing)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            out = synthesize(net, utts, max_frames_per_batch=args.max_frames, keep_on_device=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        assert len(out) == len(utts)
        t = min(ts)
        print(json.dumps({"config": name, "utterances": len(utts), "frames_min": int(frames.min()), "frames_max": int(frames.max()),
                          "audio_s": round(audio_s, 1), "infer_calls": n_calls, "seconds": round(t, 4),
                          "audio_s_per_s": round(audio_s / t, 1)}), flush=True)


if __name__ == "__main__":
    main()
