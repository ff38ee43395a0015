"""Run `infer` on one utterance (C1 by default, C4 with argv[1] == c4) a few times - for ncu launch lists and host profiles:
   python tools/one_infer.py [c1|c4] [reps]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs as oin
from oracle.weights import make_state_dict
from vispeech_b200 import build_from_hparams, get_hparams_from_file
case = sys.argv[1] if len(sys.argv) > 1 else "c1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
for opt in ("pdl", "conv_spread", "attention_small", "coupling_min_rows", "x3_min_rows"):     # A/B knobs: VS_CONV_SPREAD=0 python tools/one_infer.py
    if os.environ.get("VS_" + opt.upper()):
        net.set_option(opt, int(os.environ["VS_" + opt.upper()]))
u = (oin.c4 if case == "c4" else oin.c1)()[0]
a = (u["ids"][None], torch.LongTensor([u["ids"].numel()]))
kw = dict(sid=torch.LongTensor([u["sid"]]), noise_scale=0.667, duration_control=u["duration"][None])
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    P = net.prepare(*a, **kw)
    t1 = time.perf_counter()
    o, *_ = net.run(P, outputs="audio")
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("rep %d: prepare %.3f ms  run(host) %.3f ms  drain %.3f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
if os.environ.get("VS_CPROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        o, *_ = net.infer(*a, outputs="audio", **kw)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
