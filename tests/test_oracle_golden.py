"""Pin the oracle restatement against outputs of the unmodified reference (tests/golden/*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin
from oracle.vispeech_oracle import expansion_indices, infer_one

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("vc", "filelist")))
GOLDEN_VC = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vc*.npz")))


def _control(d, name):
    kind = int(d[name + "_kind"])
    if kind == 0:
        return None
    if kind == 1:
        return float(d[name + "_control"][0])
    return torch.from_numpy(d[name + "_control"])


def run_oracle_on_golden(sd, d):
    tf = d["z"].shape[1]
    noise = torch.from_numpy(d["noise"]) if "noise" in d else oin.draw_noise([tf], 100)[0]
    max_len = int(d["max_len"])
    return infer_one(sd, torch.from_numpy(d["ids"]), int(d["sid"]), float(d["noise_scale"]), noise,
                     None if max_len < 0 else max_len, _control(d, "energy"), _control(d, "pitch"),
                     _control(d, "duration"))


def snr_db(ref, x):
    ref, x = ref.double(), x.double()
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum().clamp_min(1e-300)))


def test_golden_files_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference(path, state_dict):
    d = dict(np.load(path))
    t = run_oracle_on_golden(state_dict, d)
    # integer work: bit exact
    assert t["x_mask"].numpy().reshape(-1).tolist() == d["x_mask"].reshape(-1).tolist()
    if int(d["duration_kind"]) == 2:
        dur = torch.from_numpy(d["duration_control"])
        assert torch.equal(expansion_indices(dur), t["lr_index"])
        assert np.array_equal(t["duration"].numpy(), d["duration"])
    else:
        assert np.array_equal(t["duration"].numpy(), d["duration"])        # ceil() of predictions: integral
    if "x_lr" in d:
        # the reference builds x_lr by per-phoneme expand+cat (models.py:418-427): same columns as our gather
        assert np.allclose(t["x_lr"].numpy(), d["x_lr"], atol=2e-5)
    # floating point: same torch build, same ops => tight tolerance (not bit exact: different op order in attention)
    for k, tol in (("x_enc", 2e-5), ("F0", 2e-3), ("energy", 2e-4), ("m_p", 5e-5), ("z", 1e-4)):
        err = float(np.abs(t[k].numpy() - d[k]).max())
        assert err <= tol, (k, err)
    for k in ("z_p", "logs_p"):
        if k in d:
            assert float(np.abs(t[k].numpy() - d[k]).max()) <= 1e-4, k
    o_ref = torch.from_numpy(d["o"].astype(np.float32))
    if int(d["o_is_f16x64"]):
        o_ref = o_ref / 64
        assert snr_db(o_ref, t["o"]) > 60.0          # fp16 storage of the fixture bounds this
    else:
        assert float((o_ref - t["o"]).abs().max()) <= 2e-5
        assert snr_db(o_ref, t["o"]) > 80.0
    assert t["o"].numel() == o_ref.numel()


@pytest.mark.parametrize("path", GOLDEN_VC, ids=[os.path.basename(p)[:-4] for p in GOLDEN_VC])
def test_oracle_voice_conversion_matches_reference(path):
    """8(f): voice_conversion (models.py:724-732) - posterior encoder, flow forward, flow reverse, decoder."""
    from oracle.vispeech_oracle import voice_conversion_one
    from oracle.weights import make_state_dict
    d = dict(np.load(path))
    sd = make_state_dict(1234, with_vc=True)
    t = voice_conversion_one(sd, torch.from_numpy(d["spec"]), int(d["sid_src"]), int(d["sid_tgt"]), torch.from_numpy(d["noise"]))
    for k in ("z", "z_p", "z_hat"):
        assert float(np.abs(t[k].numpy() - d[k]).max()) <= 2e-4, k
    assert float(np.abs(t["o"].numpy() - d["o"]).max()) <= 2e-5


def _filelist_rows():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "filelist_rows.npz"))
    off = d["offsets"]
    return [dict(ids=torch.from_numpy(d["ids"][a:b]), duration=torch.from_numpy(d["duration"][a:b]),
                 f0=torch.from_numpy(d["f0"][a:b]), energy=torch.from_numpy(d["energy"][a:b]), sid=int(d["sid"][i]))
            for i, (a, b) in enumerate(zip(off[:-1], off[1:]))]


def test_oracle_matches_reference_on_filelist_rows(state_dict):
    """Realism check: three real rows of the reference's filelists/train.list - MFA durations incl. zeros,
    measured F0 / energy as controls - run through the UNMODIFIED reference (tests/golden/make_filelist_golden.py)."""
    from oracle.vispeech_oracle import infer_one
    rows = _filelist_rows()
    assert len(rows) == 38
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "filelist_ref.npz"))
    for n, i in enumerate(ref["picks"]):
        u = rows[int(i)]
        tf = int(u["duration"].sum())
        eps = torch.randn(192, tf, generator=torch.Generator().manual_seed(int(ref["eps_seed%d" % n])))
        out = infer_one(state_dict, u["ids"], u["sid"], 0.667, eps, duration_control=u["duration"], pitch_control=u["f0"],
                        energy_control=u["energy"])
        assert out["z"].shape == ref["z%d" % n].shape
        assert float((out["z"] - torch.from_numpy(ref["z%d" % n])).abs().max()) <= 1e-4
        assert float((out["F0"].reshape(-1) - torch.from_numpy(ref["F0_%d" % n])).abs().max()) <= 1e-3
        assert float((out["energy"].reshape(-1) - torch.from_numpy(ref["energy%d" % n])).abs().max()) <= 1e-3
        o_ref = torch.from_numpy(ref["o%d" % n]).float() / 64
        assert snr_db(o_ref, out["o"]) >= 60.0
