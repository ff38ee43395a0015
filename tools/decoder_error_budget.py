"""Which fp16 roundings of the product decoder cost how much (CPU simulation on the C1 golden latent; test infrastructure):
the oracle's generator with selectable roundings - conv operands / weights per conv class, residual-stream storage per stage -
against the unrounded fp32 run.  python tools/decoder_error_budget.py"""
import glob, os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.metrics import mel_spectrogram, snr_db
from oracle.weights import make_state_dict
from oracle.vispeech_oracle import fold_weight_norm, DEFAULT_CONFIG as cfg, LRELU_SLOPE

h = lambda t: t.to(torch.float16).float()
def hl(t):                       # hi + lo split: ~22 significant bits
    hi = t.to(torch.float16).float()
    return hi + (t - hi).to(torch.float16).float()

def gen(sd, z, g, R):
    """R: set of rounding switches."""
    rw = lambda w, key: h(w) if key in R else w
    ro = lambda x, key: h(x) if key in R else x
    x = F.conv1d(ro(z, "pre_op"), rw(sd["dec.conv_pre.weight"], "pre_op"), sd["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, sd["dec.cond.weight"], sd["dec.cond.bias"])
    nk = 3
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(ro(x, "up_op%d" % i), rw(fold_weight_norm(sd, "dec.ups.%d" % i), "up_w%d" % i), sd["dec.ups.%d.bias" % i],
                               stride=u, padding=(k - u) // 2)
        x = ro(x, "up_out%d" % i)          # the stage's input stored in fp16
        xs = None
        for j in range(nk):
            p = "dec.resblocks.%d" % (i * nk + j)
            kk = cfg.resblock_kernel_sizes[j]
            y = x
            for m, d in enumerate(cfg.resblock_dilation_sizes[j]):
                xt = F.leaky_relu(y, LRELU_SLOPE)
                xt = F.conv1d(ro(xt, "rb_op%d" % i), rw(fold_weight_norm(sd, "%s.convs1.%d" % (p, m)), "rb_w%d" % i), sd["%s.convs1.%d.bias" % (p, m)],
                              padding=(kk * d - d) // 2, dilation=d)
                xt = F.leaky_relu(xt, LRELU_SLOPE)
                xt = F.conv1d(ro(xt, "rb_op%d" % i), rw(fold_weight_norm(sd, "%s.convs2.%d" % (p, m)), "rb_w%d" % i), sd["%s.convs2.%d.bias" % (p, m)],
                              padding=(kk - 1) // 2)
                y = xt + y
                if "store%d" % i in R:      # residual stream stored as fp16 a = lrelu(y), recovered as min(a, 10 a)
                    a = h(F.leaky_relu(y, LRELU_SLOPE)); y = torch.minimum(a, 10 * a)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)
    x = F.conv1d(x, sd["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)

torch.set_num_threads(16)
sd = make_state_dict(1234)
d = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1.npz")))
z = torch.from_numpy(d["z"]).float()[None]
g = sd["emb_g.weight"][int(d["sid"])].reshape(1, -1, 1)
with torch.no_grad():
    ref = gen(sd, z, g, set())[0, 0]
    m_ref = mel_spectrogram(ref)
    cases = [("pre_op", {"pre_op"})]
    for i in range(4):
        cases += [("up_op%d" % i, {"up_op%d" % i}), ("up_w%d" % i, {"up_w%d" % i}), ("up_out%d" % i, {"up_out%d" % i}), ("rb_op%d" % i, {"rb_op%d" % i}),
                  ("rb_w%d" % i, {"rb_w%d" % i}), ("store%d" % i, {"store%d" % i})]
    prod = {"pre_op"} | {"up_op%d" % i for i in range(4)} | {"up_w%d" % i for i in range(4)} | {"rb_op%d" % i for i in range(4)} | \
           {"rb_w%d" % i for i in range(4)} | {"store%d" % i for i in range(3)} | {"up_out%d" % i for i in range(3)}
    cases += [("product (all but store3 / up_out3)", prod), ("product minus up ops", prod - {"up_op%d" % i for i in range(4)} - {"up_w%d" % i for i in range(4)}),
              ("product minus store", prod - {"store%d" % i for i in range(3)} - {"up_out%d" % i for i in range(3)}),
              ("product minus rb_op3, rb_w3", prod - {"rb_op3", "rb_w3"})]
    for name, R in cases:
        o = gen(sd, z, g, R)[0, 0]
        err = (mel_spectrogram(o) - m_ref).abs()
        print("%-40s snr %6.1f dB   log-mel max %.4f mean %.5f" % (name, snr_db(ref, o), float(err.max()), float(err.mean())), flush=True)
