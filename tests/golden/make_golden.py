"""Generate tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference/models.py`).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does:
  1. imports `models.SynthesizerTrn`, `utils`, `text.symbols` from /root/reference;
  2. builds the model exactly as inference.py:26-34 does and loads the state dict of
     `oracle.weights.make_state_dict(1234)` into it (asserting that every on-path key of
     the real state dict is present with the right shape - this pins the schema);
  3. for each case patches `torch.randn_like` to return the injected eps and calls
     `net.infer(...)` with batch 1, exactly as every reference call site does;
  4. stores inputs + the returned tuple (+ enc_p / lr outputs via forward hooks).

The oracle restatement is then checked against these files by tests/test_oracle_golden.py.
"""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import inputs as oin                      # noqa: E402
from oracle.weights import make_state_dict, schema    # noqa: E402


def build_reference(sd, with_vc=False):
    import utils  # reference
    logging.disable(logging.CRITICAL)
    from models import SynthesizerTrn
    from text.symbols import symbols
    hps = utils.get_hparams_from_file("/root/reference/configs/config.json")
    torch.manual_seed(hps.train.seed)
    net = SynthesizerTrn(len(symbols), hps.data.filter_length // 2 + 1, hps.data.hop_length, hps.data.sampling_rate,
                         hps.train.segment_size // hps.data.hop_length, n_speakers=hps.data.n_speakers,
                         **hps.model).eval()
    ref_sd = net.state_dict()
    off_path = ("enc_p.proj.", "frame_prior_net.emb.", "energy_predictor.predictor.proj.") + (() if with_vc else ("enc_q.",))
    on_path = {k: tuple(v.shape) for k, v in ref_sd.items() if not k.startswith(off_path)}
    ours = {k: shp for k, shp, _ in schema(with_vc=with_vc)}
    assert set(on_path) == set(ours), (sorted(set(on_path) ^ set(ours))[:10])
    for k in on_path:
        assert on_path[k] == tuple(ours[k]) == tuple(sd[k].shape), k
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith(off_path) for m in missing)
    return net


ONLY = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]     # e.g. --only=c1 regenerates one fixture


def run_case(net, name, ids, sid, noise_scale, noise, max_len=None, energy_control=None, pitch_control=None,
             duration_control=None):
    if ONLY and name not in ONLY:
        return
    taps = {}
    h1 = net.enc_p.register_forward_hook(lambda m, i, o: taps.__setitem__("x_enc", o[0][0].clone()))
    h2 = net.lr.register_forward_hook(lambda m, i, o: taps.__setitem__("x_lr", o[0][0].clone()))
    real_randn_like = torch.randn_like

    def fake_randn_like(t, *a, **k):
        assert tuple(t.shape[1:]) == tuple(noise.shape), (t.shape, noise.shape)
        return noise.reshape(t.shape).to(t.dtype)

    torch.randn_like = fake_randn_like
    try:
        with torch.no_grad():
            o, x_mask, (z, z_p, m_p, logs_p), duration, F0, energy = net.infer(
                ids[None], torch.LongTensor([ids.numel()]), sid=torch.LongTensor([sid]), noise_scale=noise_scale,
                max_len=max_len, energy_control=None if energy_control is None else (
                    energy_control[None] if isinstance(energy_control, torch.Tensor) else energy_control),
                pitch_control=None if pitch_control is None else (
                    pitch_control[None] if isinstance(pitch_control, torch.Tensor) else pitch_control),
                duration_control=None if duration_control is None else (
                    duration_control[None] if isinstance(duration_control, torch.Tensor) else duration_control))
    finally:
        torch.randn_like = real_randn_like
        h1.remove(), h2.remove()
    assert x_mask.dtype == torch.bool

    def ctl(c):
        if c is None:
            return np.zeros(0, np.float32), 0
        if isinstance(c, torch.Tensor):
            return c.numpy(), 2
        return np.asarray([c], np.float64), 1

    dc, dk = ctl(duration_control)
    pc, pk = ctl(pitch_control)
    ec, ek = ctl(energy_control)
    big = o.numel() > 100_000          # C1: drop regenerable / redundant taps, keep the waveform in fp32
    out = dict(
        ids=ids.numpy(), sid=np.int64(sid), noise_scale=np.float64(noise_scale), noise=noise.numpy(),
        max_len=np.int64(-1 if max_len is None else max_len),
        duration_control=dc, duration_kind=np.int64(dk), pitch_control=pc, pitch_kind=np.int64(pk),
        energy_control=ec, energy_kind=np.int64(ek),
        x_enc=taps["x_enc"].numpy(), x_lr=taps["x_lr"].numpy(),
        x_mask=x_mask[0, 0].numpy(), z=z[0].numpy(), z_p=z_p[0].numpy(), m_p=m_p[0].numpy(), logs_p=logs_p[0].numpy(),
        duration=duration.reshape(-1).numpy(), F0=F0.reshape(-1).numpy(), energy=energy.reshape(-1).numpy(),
        # the waveform is stored in fp32 for every case: an fp16 copy (used in round 1 for C1) alone costs 8e-3 of the
        # 1e-2 log-mel budget of tests/test_gpu_infer.py::test_c1_log_mel_of_waveform
        o_is_f16x64=np.int64(0),
        o=o[0, 0].numpy(),
    )
    if big:   # regenerable (oracle.inputs.draw_noise([Tf], 100)) or redundant: keep the fixture small
        for k in ("noise", "x_lr", "logs_p", "z_p"):
            del out[k]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-14s Tp=%3d Tf=%4d samples=%7d  |o|max=%.4f  -> %s (%.0f KB)" % (
        name, ids.numel(), z.shape[2], o.numel(), float(o.abs().max()), os.path.basename(path),
        os.path.getsize(path) / 1024))


def run_vc_case(net, name, spec, sid_src, sid_tgt, noise):
    """voice_conversion (models.py:724-732), batch 1, eps of PosteriorEncoder (models.py:239) injected."""
    real_randn_like = torch.randn_like
    torch.randn_like = lambda t, *a, **k: noise.reshape(t.shape).to(t.dtype)
    try:
        with torch.no_grad():
            o_hat, y_mask, (z, z_p, z_hat) = net.voice_conversion(spec[None], torch.LongTensor([spec.shape[1]]),
                                                                   torch.LongTensor([sid_src]), torch.LongTensor([sid_tgt]))
    finally:
        torch.randn_like = real_randn_like
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, spec=spec.numpy(), sid_src=np.int64(sid_src), sid_tgt=np.int64(sid_tgt), noise=noise.numpy(),
                        z=z[0].numpy(), z_p=z_p[0].numpy(), z_hat=z_hat[0].numpy(), o=o_hat[0, 0].numpy(),
                        y_mask=y_mask[0, 0].numpy())
    print("%-14s T=%d samples=%d -> %s (%.0f KB)" % (name, spec.shape[1], o_hat.numel(), os.path.basename(path),
                                                      os.path.getsize(path) / 1024))


def main_vc():
    sd = make_state_dict(1234, with_vc=True)
    net = build_reference(sd, with_vc=True)
    g = torch.Generator().manual_seed(123)
    spec = torch.rand(1025, 23, generator=g) ** 2 * 4.0          # magnitude-spectrogram-like, non-negative
    run_vc_case(net, "vc1", spec, 7, 64, torch.randn(192, 23, generator=g))


def main():
    if "--vc" in sys.argv:
        return main_vc()
    sd = make_state_dict(1234)
    net = build_reference(sd)
    g = torch.Generator().manual_seed(99)

    def rint(lo, hi, n):
        return torch.randint(lo, hi + 1, (n,), generator=g)

    # g1: durations given as ints incl. a zero; everything else predicted
    ids = rint(1, 518, 12)
    dur = rint(1, 6, 12)
    dur[4] = 0
    tf = int(dur.sum())
    run_case(net, "g1_given_dur", ids, 3, 0.667, torch.randn(192, tf, generator=g), duration_control=dur)

    # g2: everything predicted, scalar controls.  Tf is only known after the predictor ran: probe first.
    ids = rint(1, 518, 9)
    from oracle.vispeech_oracle import infer_one
    probe = infer_one(sd, ids, 17, duration_control=1.3, pitch_control=1.1, energy_control=0.9, stop_after="lr")
    tf = int(probe["lr_index"].numel())
    run_case(net, "g2_predicted", ids, 17, 1.0, torch.randn(192, tf, generator=g), duration_control=1.3,
             pitch_control=1.1, energy_control=0.9)

    # g3: manual edit: fractional / negative / zero float durations, F0 in Hz with zeros, raw energy, max_len
    ids = rint(1, 518, 10)
    durf = torch.tensor([2.9, 0.0, -1.5, 3.2, 1.0, 4.999, 0.4, 2.0, 6.7, 1.5])
    f0 = 80 + 670 * torch.rand(10, generator=g)
    f0[2] = 0.0
    f0[7] = 0.0
    en = 150 * torch.rand(10, generator=g)
    tf = int(torch.trunc(durf).clamp_min(0).sum())
    run_case(net, "g3_manual_edit", ids, 66, 0.5, torch.randn(192, tf, generator=g), max_len=15,
             duration_control=durf, pitch_control=f0, energy_control=en)

    # g4: Tp=3 < window+1 (relative-embedding slicing edge, attentions.py:181-194); predicted path, Tp*Tf tiny
    ids = rint(1, 518, 3)
    dur = torch.tensor([2, 7, 3])
    run_case(net, "g4_short", ids, 0, 0.667, torch.randn(192, 12, generator=g), duration_control=dur)

    # C1 exactly as SURVEY 8d (B=1, Tp=40, U{3..18} durations, sid 0, noise_scale .667, seed 0)
    u = oin.c1()[0]
    tf = oin.frame_counts([u])[0]
    run_case(net, "c1", u["ids"], u["sid"], 0.667, oin.draw_noise([tf], 100)[0], duration_control=u["duration"])


if __name__ == "__main__":
    main()
