"""One device-resident step of the bench workload (configs[1]: 64 x ~5 s) a few times - the target of ncu launch lists:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv python tools/one_step.py [reps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs as oin
from oracle.weights import make_state_dict
from vispeech_b200 import build_from_hparams, get_hparams_from_file
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
utts = oin.c2(batch=64, seed=1)
tp = max(u["ids"].numel() for u in utts)
ids = torch.stack([u["ids"] for u in utts]); dur = torch.stack([u["duration"] for u in utts])
P = net.prepare(ids, torch.LongTensor([tp] * 64), sid=torch.LongTensor([u["sid"] for u in utts]), noise_scale=0.667, duration_control=dur)
for _ in range(reps):
    net.run(P, outputs="audio")
torch.cuda.synchronize()
print("done")
