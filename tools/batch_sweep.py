"""Shape sweep: the same utterance inside batches of very different sizes (every precision regime and tile-count parity of the
kernels: fp32 < 256 rows, 3xTF32 < 4096, plain TF32 + one-kernel WN layers above; odd / even tile counts; > 1 wave).
Prints the difference of utterance 0's latent and waveform to its batch-1 result; asserts the parity bars."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs as oin
from oracle.weights import make_state_dict
from oracle.metrics import snr_db
from vispeech_b200 import build_from_hparams, get_hparams_from_file
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
utts = oin.c2(batch=150, seed=9)
frames = oin.frame_counts(utts)
noise = oin.draw_noise(frames, 77)
ref = None
for B in (1, 2, 3, 9, 10, 33, 64, 97, 150):
    ids = torch.stack([u["ids"] for u in utts[:B]]); dur = torch.stack([u["duration"] for u in utts[:B]])
    sid = torch.LongTensor([u["sid"] for u in utts[:B]])
    o, m, (z, *_), *_ = net.infer(ids, torch.LongTensor([40] * B), sid=sid, noise_scale=0.667, duration_control=dur, noise=noise[:B])
    torch.cuda.synchronize()
    assert torch.isfinite(o).all() and torch.isfinite(z).all(), B
    tf = frames[0]
    z0, o0 = z[0, :, :tf].cpu(), o[0, 0, :tf * 512].cpu()
    if ref is None:
        ref = (z0, o0)
    dz, s = float((z0 - ref[0]).abs().max()), snr_db(ref[1], o0) if B > 1 else float("inf")
    # last utterance of the batch against its own batch-1 run (catches tail-tile bugs)
    j = B - 1
    o1, _, (z1, *_), *_ = net.infer(ids[j:j + 1], torch.LongTensor([40]), sid=sid[j:j + 1], noise_scale=0.667, duration_control=dur[j:j + 1], noise=noise[j:j + 1])
    tfj = frames[j]
    dzl = float((z[j, :, :tfj].cpu() - z1[0, :, :tfj].cpu()).abs().max())
    sl = snr_db(o1[0, 0, :tfj * 512].cpu(), o[j, 0, :tfj * 512].cpu()) if B > 1 else float("inf")
    print("B=%3d rows=%6d  utt0: max|dz| %.2e  wave SNR %.1f dB   last utt: max|dz| %.2e  SNR %.1f dB" % (B, sum(frames[:B]), dz, s, dzl, sl))
    assert dz <= 1e-2 and dzl <= 1e-2 and s >= 30 and sl >= 30, B
print("ok")
