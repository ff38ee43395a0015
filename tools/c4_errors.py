"""Latent errors of the C4 case (one 60 s utterance, 5168 frames) against the oracle, for a few precision settings."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs as oin
from oracle.vispeech_oracle import infer_one
from oracle.weights import make_state_dict
from vispeech_b200 import _lib, build_from_hparams, get_hparams_from_file
sd = make_state_dict(1234)
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(sd)
lib = _lib.load()
u = oin.c4()[0]
tf = oin.frame_counts([u])[0]
eps = oin.draw_noise([tf], 31)[0]
ref = infer_one(sd, u["ids"], u["sid"], 0.667, eps, duration_control=u["duration"])
for name, opts in (("default", {}), ("tf32 prior", {"tf32_prior": 1}), ("all 3xTF32", {"tf32_min_rows": 1 << 30})):
    for k, v in opts.items():
        _lib.check(lib.vs_set_option(k.encode(), v))
    o, x_mask, (z, z_p, m_p, logs_p), *_ = net.infer(u["ids"][None], torch.LongTensor([u["ids"].numel()]), sid=torch.LongTensor([u["sid"]]),
                                                    noise_scale=0.667, duration_control=u["duration"][None], noise=[eps])
    torch.cuda.synchronize()
    print(name, " ".join("%s=%.2e" % (n, float((t[0].cpu() - ref[n]).abs().max())) for n, t in (("m_p", m_p), ("logs_p", logs_p), ("z_p", z_p), ("z", z))))
    _lib.check(lib.vs_set_option(b"tf32_min_rows", 4096))
    _lib.check(lib.vs_set_option(b"tf32_prior", 0))
