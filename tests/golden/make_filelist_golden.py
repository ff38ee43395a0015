"""Realism fixtures from the reference's own filelist ("also run the 38 real rows of filelists/train.list").

Run in the build container only:   python tests/golden/make_filelist_golden.py

  filelist_rows.npz  all 38 rows of /root/reference/filelists/train.list as model inputs: phoneme ids (through the
                     reference's text.symbols table), MFA durations, per-phoneme F0 (Hz) and energy, speaker id
                     (configs/config.json spk2id) - concatenated, with offsets;
  filelist_ref.npz   outputs of the UNMODIFIED reference (`models.SynthesizerTrn.infer`, batch 1, injected eps) for three of
                     them (the shortest, one with zero-length phonemes, a median one) driven by the row's own durations /
                     F0 / energy as duration_control / pitch_control / energy_control (the manual-edit path, models.py:681-706).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

from make_golden import build_reference                 # noqa: E402
from oracle.weights import make_state_dict             # noqa: E402


def main():
    import json
    from text.symbols import symbols
    sym = {s: i for i, s in enumerate(symbols)}
    spk2id = json.load(open("/root/reference/configs/config.json"))["data"]["spk2id"]
    rows = [l.rstrip("\n").split("|") for l in open("/root/reference/filelists/train.list", encoding="utf-8")]
    ids, dur, f0, en, sid, off = [], [], [], [], [], [0]
    for spk, _name, ph, d, f, e in rows:
        p = [sym[s] for s in ph.split(" ")]
        d, f, e = [int(x) for x in d.split()], [float(x) for x in f.split()], [float(x) for x in e.split()]
        assert len(p) == len(d) == len(f) == len(e)
        ids += p; dur += d; f0 += f; en += e
        sid.append(spk2id.get(spk, 0))
        off.append(len(ids))
    np.savez_compressed(os.path.join(HERE, "filelist_rows.npz"), ids=np.asarray(ids, np.int64), duration=np.asarray(dur, np.int64),
                        f0=np.asarray(f0, np.float32), energy=np.asarray(en, np.float32), sid=np.asarray(sid, np.int64),
                        offsets=np.asarray(off, np.int64))
    frames = [int(sum(dur[off[i]:off[i + 1]])) for i in range(len(rows))]
    order = np.argsort(frames)
    with_zero = [i for i in order if 0 in dur[off[i]:off[i + 1]]]
    picks = [int(order[0]), int(with_zero[0]) if with_zero else int(order[1]), int(order[len(order) // 2])]
    sd = make_state_dict(1234)
    net = build_reference(sd)
    out = {"picks": np.asarray(picks, np.int64)}
    real = torch.randn_like
    for n, i in enumerate(picks):
        a, b = off[i], off[i + 1]
        t_ids = torch.LongTensor(ids[a:b])
        t_dur = torch.LongTensor(dur[a:b])
        t_f0, t_en = torch.tensor(f0[a:b]), torch.tensor(en[a:b])
        g = torch.Generator().manual_seed(500 + i)
        eps = torch.randn(192, frames[i], generator=g)
        torch.randn_like = lambda t, *aa, **kk: eps.reshape(t.shape).to(t.dtype)
        try:
            with torch.no_grad():
                o, x_mask, (z, z_p, m_p, logs_p), duration, F0, energy = net.infer(
                    t_ids[None], torch.LongTensor([b - a]), sid=torch.LongTensor([sid[i]]), noise_scale=0.667,
                    duration_control=t_dur[None], pitch_control=t_f0[None], energy_control=t_en[None])
        finally:
            torch.randn_like = real
        assert z.shape[2] == frames[i] and x_mask.dtype == torch.bool
        out.update({"z%d" % n: z[0].numpy(), "F0_%d" % n: F0.reshape(-1).numpy(),
                    "energy%d" % n: energy.reshape(-1).numpy(), "o%d" % n: (o[0, 0] * 64).half().numpy(),
                    "eps_seed%d" % n: np.int64(500 + i)})
        print("row %2d: Tp=%d Tf=%d samples=%d |o|max=%.4f" % (i, b - a, frames[i], o.numel(), float(o.abs().max())))
    path = os.path.join(HERE, "filelist_ref.npz")
    np.savez_compressed(path, **out)
    print("-> filelist_rows.npz (%.0f KB), filelist_ref.npz (%.0f KB)" % (
        os.path.getsize(os.path.join(HERE, "filelist_rows.npz")) / 1024, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
