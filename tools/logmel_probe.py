"""Log-mel error of the C1 golden waveform under decoder option variants (where the max sits, how noisy the bar is):
   python tools/logmel_probe.py"""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.metrics import mel_spectrogram, snr_db
from oracle.weights import make_state_dict
from vispeech_b200 import build_from_hparams, get_hparams_from_file, _lib
from test_gpu_infer import run_golden
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
lib = _lib.load()
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    name = os.path.basename(path)[:-4]
    if name.startswith(("vc", "filelist")):
        continue
    d = dict(np.load(path))
    if "o" not in d:
        continue
    ref = torch.from_numpy(d["o"].astype(np.float32)) / (64 if int(d.get("o_is_f16x64", 0)) else 1)
    m_ref = mel_spectrogram(ref)
    for label, prec, opts in (("fp32", 1, {}), ("f16", 0, {}), ("f16 no resblock64", 0, {"resblock_fused": 0}), ("f16 no pair conv", 0, {"pair_conv": 0, "pair_fused": 0}),
                              ("f16 no-mrf", 0, {"mrf_fused": 0}), ("f16 unfused pairs", 0, {"fused_respair": 0}),
                              ("f16 one stream", 0, {"decoder_streams": 1}), ("f16 no conv spread", 0, {"conv_spread": 0}),
                              ("f16 attention 64", 0, {"attention_small": 0}),
                              ("f16 small-call paths off", 0, {"decoder_streams": 1, "conv_spread": 0, "attention_small": 0})):
        for k, v in opts.items():
            net.set_option(k, v)
        o = run_golden(net, d, prec)[0]
        for k in opts:
            net.set_option(k, {"resblock_fused": 1, "mrf_fused": 1, "fused_respair": 2, "pair_conv": 1, "pair_fused": 1, "decoder_streams": 0, "conv_spread": 1,
                               "attention_small": 1}[k])
        w = o[0, 0].cpu()
        n = min(w.numel(), ref.numel())
        m = mel_spectrogram(w)
        err = (m_ref - m).abs()
        i = int(err.argmax()); b, f = divmod(i, err.shape[1])
        top = torch.topk(err.reshape(-1), 5).values.tolist()
        print("%-10s %-18s max %.4f mean %.5f at bin %d frame %d (ref log-mel %.2f, max %.2f)  top5 %s  snr %.1f dB" %
              (name, label, float(err.max()), float(err.mean()), b, f, float(m_ref[b, f]), float(m_ref.max()),
               ["%.4f" % t for t in top], snr_db(ref[:n], w[:n])))
