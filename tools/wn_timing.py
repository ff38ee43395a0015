"""Where does the WN-layer kernel's MMA warp wait?  Needs VS_UMMA_TIMING=1 python vispeech_b200/build.py --force.
Runs one C2-sized flow pass with the timing buffer armed for the LAST wn layer launch only (buffer is overwritten per launch)."""
import os, sys, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs as oin
from oracle.weights import make_state_dict
from vispeech_b200 import _lib, build_from_hparams, get_hparams_from_file
lib = _lib.load()
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
utts = oin.c2(batch=64, seed=1)
ids = torch.stack([u["ids"] for u in utts]); dur = torch.stack([u["duration"] for u in utts]); sid = torch.LongTensor([u["sid"] for u in utts])
P = net.prepare(ids, torch.LongTensor([40] * 64), sid=sid, noise_scale=0.667, duration_control=dur)
net.run(P, outputs="audio"); torch.cuda.synchronize()
buf = torch.zeros(148 * 16 + 148 * 8, dtype=torch.int64, device="cuda:0")
_lib.check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
z, rf = net.run(P, outputs="latents"); torch.cuda.synchronize()
_lib.check(lib.vs_set_option(b"umma_timing_buffer", 0))
t = buf[148 * 16:].view(-1, 8)[:148].double()
used = t[:, 0] > 0
two = t[used & (t[:, 5] == 2)].mean(0); one = t[used & (t[:, 5] == 1)].mean(0)
for name, r in (("CTAs with 2 tiles", two), ("CTAs with 1 tile", one)):
    tot = max(r[0].item(), 1)
    print("%s: total %.0f clk | wait a_full %.1f%%  b_full %.1f%%  acts_full %.1f%%  acc2_empty %.1f%%  -> issuing %.1f%%" % (
        name, tot, 100 * r[1] / tot, 100 * r[2] / tot, 100 * r[3] / tot, 100 * r[4] / tot, 100 * (tot - r[1:5].sum().item()) / tot))
