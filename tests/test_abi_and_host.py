"""CPU-only checks: the C-ABI library builds, loads and exports every symbol of include/vispeech_b200.h; host-side
layout / packing / sharding logic.  No compute entry point is called (there is no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vispeech_b200.build import build
    build()
    from vispeech_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from vispeech_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vispeech_b200.h")).read()
    declared = set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vs_version() == 1


def test_sass_is_blackwell_native():
    """The decoder conv must be tcgen05 + TMA bulk copies + TMEM loads, not a legacy mma.sync path."""
    import subprocess
    from vispeech_b200.build import LIB_PATH
    sass = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass


def test_no_cpu_fallback_without_cuda():
    from vispeech_b200 import _lib, build_from_hparams, get_hparams_from_file
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.VsError):
        build_from_hparams(get_hparams_from_file())


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vispeech_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_row_layout_plan():
    from vispeech_b200.layout import plan_starts
    starts, n = plan_starts([5, 1, 9], gap=4, align=16)
    assert starts.tolist() == [0, 9, 14] and n == 32
    starts, n = plan_starts([431] * 64, gap=4, align=16)
    assert n % 16 == 0 and n >= 64 * 435 and all(b - a == 435 for a, b in zip(starts[:-1], starts[1:]))


def test_packing_shapes_and_flow_flip(state_dict):
    from vispeech_b200.packing import pack_state_dict, pack_umma, ups_union_taps
    p = pack_state_dict(state_dict)
    assert p["enc_p.encoder.0.wqkv"].shape == (192, 576)
    assert p["dec.ups.0.w"].shape == (8, 2, 512, 256) and p["dec.ups.2.w"].shape == (4, 1, 128, 64)
    assert [ups_union_taps(i) for i in range(4)] == [(3, 1), (3, 1), (1, 0), (3, 1)]
    assert p["dec16.ups.0.w"].numel() == 3 * 512 * 2048 and p["dec16.ups.0.w"].dtype == torch.float16
    # flow 1 and 3 run on a flipped tensor: pre reads reversed upper-half channels, post writes reversed lower half
    w = state_dict["flow.flows.2.pre.weight"][:, :, 0]            # [192 co, 96 ci]
    assert torch.equal(p["flow.1.pre.w"], torch.flip(w.t(), [0]))
    assert torch.equal(p["flow.0.pre.w"], state_dict["flow.flows.0.pre.weight"][:, :, 0].t())
    # slab layout: element (nb,t,kc,p,n,e) = W[t][kc*KC+8p+e][nb*Nblk+n]
    wt = torch.arange(2 * 128 * 512, dtype=torch.float32).reshape(2, 128, 512) % 251
    sl = pack_umma(wt).float().reshape(2, 2, 2, 8, 256, 8)
    assert sl[1, 1, 1, 3, 17, 5] == wt[1, 64 + 24 + 5, 256 + 17]


def test_polyphase_equals_conv_transpose(state_dict):
    """The polyphase rewrite used by both decoder paths equals F.conv_transpose1d (pure host math check)."""
    from vispeech_b200.packing import UP_KERNELS, UP_RATES, pack_state_dict, ups_phase_range
    p = pack_state_dict(state_dict)
    from oracle.weights import fold_weight_norm
    for i in (2, 3):
        s, K = UP_RATES[i], UP_KERNELS[i]
        wt = fold_weight_norm(state_dict, "dec.ups.%d" % i)
        cin, cout, _ = wt.shape
        x = torch.randn(1, cin, 11)
        ref = torch.nn.functional.conv_transpose1d(x, wt, None, stride=s, padding=(K - s) // 2)[0]
        per = p["dec.ups.%d.w" % i]
        out = torch.zeros(cout, 11 * s)
        for ph in range(s):
            lo, hi = ups_phase_range(i, ph)
            for t, d in enumerate(range(lo, hi + 1)):
                for q in range(11):
                    if 0 <= q + d < 11:
                        out[:, s * q + ph] += per[ph, t].t() @ x[0, :, q + d]
        assert torch.allclose(out, ref, atol=1e-5)


def test_mel_filterbank_matches_torchaudio_slaney():
    """vispeech_b200.mel.mel_filterbank restates librosa.filters.mel (what mel_processing.py:78 calls); torchaudio's
    Slaney-scale / Slaney-norm filterbank is the same definition and is the independent check available here."""
    import torchaudio
    from vispeech_b200.mel import mel_filterbank
    for sr, n_fft, n_mels, fmin, fmax in [(44100, 2048, 80, 0.0, None), (22050, 1024, 80, 0.0, 8000.0), (44100, 2048, 64, 50.0, 16000.0)]:
        ours = torch.from_numpy(mel_filterbank(sr, n_fft, n_mels, fmin, fmax))
        ref = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, float(fmax or sr / 2), n_mels, sr, norm="slaney",
                                                    mel_scale="slaney").t()
        assert ours.shape == ref.shape and float((ours - ref).abs().max()) <= 1e-5 * float(ref.abs().max())   # fp32 rounding of two float64 recipes


def test_dft_basis_reproduces_torch_stft():
    """The 4-tap conv with the windowed DFT basis over hop-sized rows (what the GPU GEMM computes) equals torch.stft with
    the reference's padding (mel_processing.py:64-68), in float64 on the CPU."""
    from vispeech_b200.mel import dft_basis
    n_fft, hop = 2048, 512
    w, half = dft_basis(n_fft, hop)
    g = torch.Generator().manual_seed(0)
    y = torch.randn(1, 5 * hop + 77, generator=g, dtype=torch.float64)
    pad = (n_fft - hop) // 2
    yp = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    ref = torch.stft(yp, n_fft, hop_length=hop, win_length=n_fft, window=torch.hann_window(n_fft, dtype=torch.float64),
                     center=False, return_complex=True)[0]                                    # [1025, frames]
    n_frames = y.shape[1] // hop
    assert ref.shape[1] == n_frames
    rows = torch.zeros(n_frames + 3, w.shape[1], dtype=torch.float64)
    flat = yp[0, : (n_frames + 3) * hop]
    rows[: flat.numel() // hop, :hop] = flat[: flat.numel() // hop * hop].reshape(-1, hop)
    out = sum(rows[t: t + n_frames] @ w[t].double() for t in range(4))                       # [frames, 2 * half]
    re, im = out[:, :1025].t(), out[:, half:half + 1025].t()
    assert float((re - ref.real).abs().max()) <= 1e-4 and float((im - ref.imag).abs().max()) <= 1e-4


def test_coupling_pack_reproduces_oracle_coupling_layer(state_dict):
    """Host logic of the one-kernel coupling layer (packing.pack_coupling / pack_coupling_bias; csrc/umma_coupling.cu): a numpy
    statement of exactly what the kernel computes from the packed slab stream and bias table - post folded into the skip half of
    res_skip, cumulative h biases, gate96 column order, the Flip folded into the odd layers - must reproduce the oracle's coupling
    layers (modules.py:324-343 + WN 148-176) on a random input, for a flipped and an unflipped layer, up to the fp16 rounding of
    the weights."""
    import torch.nn.functional as F
    from oracle.vispeech_oracle import DEFAULT_CONFIG as cfg, wn
    from vispeech_b200.packing import COUPLING_SLAB, pack_state_dict
    packed = pack_state_dict(state_dict)
    g = torch.Generator().manual_seed(3)
    T, sid = 50, 17
    gvec = state_dict["emb_g.weight"][sid].reshape(1, -1, 1)
    for f in (0, 1):
        blob = packed["c16.flow.%d.w" % f].float().reshape(92, COUPLING_SLAB)
        bias = packed["c16.flow.%d.b" % f]
        def slab(i, k):                                                     # -> [K][96]
            return blob[i][: (k // 8) * 96 * 8].reshape(k // 8, 96, 8).permute(0, 2, 1).reshape(k, 96).double()
        hb, mb = bias[: 4 * 192].reshape(4, 192).double(), bias[768:864].double()
        cg = bias[864:].reshape(-1, 4, 384)[sid].double()
        z = torch.randn(T, 192, generator=g).double()                        # PHYSICAL layout (the flips are folded into the weights)
        flipped = f % 2 == 1
        x0 = z[:, 96:] if flipped else z[:, :96]
        # ---- what the kernel computes
        h = x0 @ torch.cat([slab(0, 96), slab(1, 96)], 1)                    # pre (bias in hb[0])
        m = torch.zeros(T, 96, dtype=torch.float64)
        s = 2
        for l in range(4):
            hin = F.pad((h + hb[l]).t()[None], (2, 2))[0].t()                # [T + 4][192], zero padding
            acts = torch.zeros(T, 192, dtype=torch.float64)
            for nb in range(4):
                a = sum(hin[t:t + T] @ slab(s + t, 192) for t in range(5)) + cg[l, 96 * nb: 96 * nb + 96]
                acts[:, 48 * nb: 48 * nb + 48] = torch.tanh(a[:, :48]) * torch.sigmoid(a[:, 48:])
                s += 5
            if l < 3:
                h = h + acts @ torch.cat([slab(s, 192), slab(s + 1, 192)], 1)
                s += 2
            m = m + acts @ slab(s, 192)
            s += 1
        assert s == 92
        m = m + mb
        # ---- the oracle on the logical (flipped) tensor
        src = "flow.flows.%d" % (2 * f)
        x = torch.flip(z, [1]) if flipped else z
        xl = x.t()[None].float()
        hh = F.conv1d(xl[:, :96], state_dict[src + ".pre.weight"], state_dict[src + ".pre.bias"])
        hh = wn(state_dict, src + ".enc", hh, gvec, cfg)
        m_ref = F.conv1d(hh, state_dict[src + ".post.weight"], state_dict[src + ".post.bias"])[0].t().double()   # logical channel order
        if flipped:
            m_ref = torch.flip(m_ref, [1])                                  # m_phys[p] = m[95 - p]
        scale = float(m_ref.abs().max())
        err = float((m - m_ref).abs().max())
        assert err <= 5e-3 * max(scale, 1.0), (f, err, scale)
