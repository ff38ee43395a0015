"""A small pass over every kernel family for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_run.py
Batch of 4 x ~3 s utterances (tensor-core paths: TF32 / 3xTF32 convs, MMA attention, all fused ResBlock forms), one batch-1
call (cluster split-K fp32 convs, CUDA-core attention), the chunked decoder, PCM16 and the GPU mel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs as oin
from oracle.weights import make_state_dict
from vispeech_b200 import build_from_hparams, get_hparams_from_file, mel_spectrogram_torch
from vispeech_b200.postprocess import to_pcm16
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
utts = oin.c2(batch=4, seed=3, target_frames=260)
ids = torch.stack([u["ids"] for u in utts]); dur = torch.stack([u["duration"] for u in utts])
sid = torch.LongTensor([u["sid"] for u in utts])
o, *_ = net.infer(ids, torch.LongTensor([40] * 4), sid=sid, noise_scale=0.667, duration_control=dur)
from vispeech_b200 import _lib
_lib.check(_lib.load().vs_set_option(b"tf32_min_rows", 1))          # plain-TF32 route: the one-kernel coupling layer (umma_coupling.cu) ...
o3, *_ = net.infer(ids, torch.LongTensor([40] * 4), sid=sid, noise_scale=0.667, duration_control=dur, outputs="audio")
net.set_option("coupling_fused", 0)                                 # ... and the one-kernel WN layer (umma_wn.cu)
net.set_option("pair_fused", 2)                                     # the CTA-pair fused iteration at C = 64 too
o4, *_ = net.infer(ids, torch.LongTensor([40] * 4), sid=sid, noise_scale=0.667, duration_control=dur, outputs="audio")
net.set_option("coupling_fused", 1); net.set_option("pair_fused", 1)
_lib.check(_lib.load().vs_set_option(b"tf32_min_rows", 4096))
net.overlap_calls = True
o2, *_ = net.infer(ids, torch.LongTensor([40] * 4), sid=sid, noise_scale=0.667, outputs="audio")        # predicted durations
net.overlap_calls = False
o1, *_ = net.infer(ids[:1], torch.LongTensor([40]), sid=sid[:1], noise_scale=0.667, duration_control=dur[:1], outputs="audio")
chunks = [w for _, w in net.infer_stream(ids[:1], torch.LongTensor([40]), sid=sid[:1], duration_control=dur[:1], chunk_frames=64)]
pcm = to_pcm16(o, [o.shape[2]] * 4, 44100, 22050)
mel = mel_spectrogram_torch(o[:, 0], 2048, 80, 44100, 512, 2048, 0, None)
torch.cuda.synchronize()
print("ok", tuple(o.shape), tuple(o2.shape), tuple(o1.shape), len(chunks), tuple(pcm.shape), tuple(mel.shape),
      bool(torch.isfinite(o).all()), bool(torch.isfinite(mel).all()))
