"""GPU parity, op level: every CUDA kernel family against a CPU fp64/fp32 statement of the same op.
All calls go through the C ABI (ctypes).  Tolerances are stated per test."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util
    assert torch.cuda.is_available()
    return gpu_util


UMMA_CASES = [
    # (R, Cin, N, taps, dil) - the decoder's shapes plus ragged row counts
    (1000, 32, 32, 3, 1), (517, 32, 32, 11, 5), (640, 64, 64, 7, 3), (384, 128, 128, 11, 5), (300, 128, 128, 3, 1),
    (256, 256, 256, 3, 3), (200, 256, 256, 11, 5), (130, 192, 512, 7, 1),
]


@pytest.mark.parametrize("R,cin,n,taps,dil", UMMA_CASES)
def test_umma_conv_matches_cpu(G, R, cin, n, taps, dil):
    """fp16 operands, fp32 accumulate: compare against an fp64 conv of the fp16-rounded operands.
    Tolerance: fp16 output rounding (2^-11 relative) + fp32 accumulation noise."""
    g = torch.Generator().manual_seed(R + cin + taps)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(taps, cin, n, generator=g) / (cin * taps) ** 0.5
    b = torch.randn(n, generator=g) * 0.1
    pad_l = (taps - 1) // 2
    raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), dil=dil, pad_l=pad_l, act_slope=0.1)
    ref = G.ref_conv_rows(G.f16_round(x), G.f16_round(w), b, dil=dil, pad_l=pad_l)
    err = (raw.cpu().double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
    ref_act = torch.where(ref > 0, ref, 0.1 * ref)
    assert (act.cpu().double() - ref_act).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("R,cin,n,taps,dil", [(370, 192, 512, 7, 1), (2960, 256, 256, 11, 5), (300, 512, 2048, 2, 1), (2960, 256, 1024, 2, 1),
                                              (130, 256, 256, 3, 1), (9000, 256, 256, 7, 3)])
def test_umma_conv_spread_over_n_blocks_is_bit_identical(G, R, cin, n, taps, dil):
    """Small calls cut the streamed-weight convs along N too (option conv_spread: n-blocks narrowed to >= 64 columns, fetched as column
    ranges of the packed slabs, one CTA per (row tile, n-block group)): the batch-1 shapes of conv_pre, the first two ConvTranspose and the
    C = 256 ResBlock convs (models.py:250-285), a masked gap and a ragged tail.  Same MMAs per output element -> identical bits; the
    last case has too many tiles to spread and checks the dispatch leaves it alone."""
    from vispeech_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(R + n)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(taps, cin, n, generator=g) / (cin * taps) ** 0.5
    b = torch.randn(n, generator=g) * 0.1
    res = G.f16_round(torch.randn(R, n, generator=g))
    row_utt = torch.zeros(R, dtype=torch.int32)
    row_utt[50:61] = -1
    x[50:61] = 0
    pad_l = (taps - 1) // 2
    outs = []
    for spread in (1, 0):
        _lib.check(lib.vs_set_option(b"conv_spread", spread))
        try:
            outs.append(G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), res=res.to(G.DEV), dil=dil, pad_l=pad_l, act_slope=0.1,
                                    row_utt=row_utt.to(G.DEV)))
        finally:
            _lib.check(lib.vs_set_option(b"conv_spread", 1))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = G.ref_conv_rows(G.f16_round(x), G.f16_round(w), b, dil=dil, pad_l=pad_l) + res.double()
    ref[50:61] = 0
    assert (outs[0][0].cpu().double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def test_umma_conv_residual_mask_and_scale(G):
    R, C = 700, 64
    g = torch.Generator().manual_seed(5)
    x, res = torch.randn(R, C, generator=g), torch.randn(R, C, generator=g)
    w = torch.randn(7, C, C, generator=g) / (7 * C) ** 0.5
    b = torch.randn(C, generator=g) * 0.1
    row_utt = torch.zeros(R // 4, dtype=torch.int32)
    row_utt[40:50] = -1                                            # rows 160..199 invalid
    raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), res=res.to(G.DEV), dil=1, pad_l=3, act_slope=0.01,
                           act_scale=1 / 3, row_utt=row_utt.to(G.DEV), row_div=4)
    ref = G.ref_conv_rows(G.f16_round(x), G.f16_round(w), b, pad_l=3) + G.f16_round(res).double()
    ref[160:200] = 0
    assert (raw.cpu().double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    ra = ref / 3
    ra = torch.where(ra > 0, ra, 0.01 * ra)
    assert (act.cpu().double() - ra).abs().max().item() <= 2e-3 * ra.abs().max().item()
    assert raw[160:200].abs().max().item() == 0 and act[160:200].abs().max().item() == 0


@pytest.mark.parametrize("R,C,k,dil,form", [(300, 128, 3, 1, "c1"), (384, 128, 11, 5, "c2_mid"), (256 * 150 + 77, 128, 7, 3, "c2_sum"),
                                           (256 * 75 + 130, 128, 11, 5, "c2_last"), (129, 128, 7, 5, "c1"), (100, 128, 11, 1, "c2_sum0"),
                                           (5000, 256, 3, 5, "c2_mid")])
def test_pair_conv_matches_cpu(G, R, C, k, dil, form):
    """The decoder's wide convs on a CTA pair (tcgen05 cta_group::2, csrc/umma_pair.cu) in each of the decoder's epilogue forms
    (modules.py:211-220, models.py:280-285) vs an fp64 conv of the fp16-rounded operands; incl. a masked gap, a ragged tail
    whose second CTA's tile is entirely out of range, and more units than CTA pairs."""
    from vispeech_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(R + k)
    x = torch.randn(R, C, generator=g)
    w = torch.randn(k, C, C, generator=g) / (C * k) ** 0.5
    b = torch.randn(C, generator=g) * 0.1
    res = G.f16_round(torch.randn(R, C, generator=g))
    res2 = G.f16_round(torch.randn(R, C, generator=g))
    row_utt = torch.zeros(R, dtype=torch.int32)
    lo, hi = min(40, R // 2), min(90, R // 2 + 10)
    row_utt[lo:hi] = -1
    x[lo:hi] = 0
    kw = dict(dil=dil, pad_l=(k - 1) // 2 * 1, row_utt=row_utt.to(G.DEV))
    ref = G.ref_conv_rows(G.f16_round(x), G.f16_round(w), b, dil=dil, pad_l=(k - 1) // 2)
    resd = res.double()
    x_rec = torch.minimum(resd, resd * float(np.float32(10.0)))
    _lib.check(lib.vs_set_option(b"pair_conv", 2))
    try:
        if form == "c1":
            raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), act_slope=0.1, want_raw=False, **kw)
            want_raw, want_act = None, torch.where(ref > 0, ref, 0.1 * ref)
        elif form == "c2_mid":
            raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), res=res.to(G.DEV), res_inv_slope=10.0, act_slope=0.1, want_raw=False, **kw)
            y = ref + x_rec
            want_raw, want_act = None, torch.where(y > 0, y, 0.1 * y)
        elif form == "c2_sum0":
            raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), res=res.to(G.DEV), res_inv_slope=10.0, want_act=False, **kw)
            want_raw, want_act = ref + x_rec, None
        elif form == "c2_sum":
            raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), res=res.to(G.DEV), res2=res2.to(G.DEV), res_inv_slope=10.0, want_act=False, **kw)
            want_raw, want_act = ref + x_rec + res2.double(), None
        else:
            raw, act = G.umma_conv(x.to(G.DEV), w, b.to(G.DEV), res=res.to(G.DEV), res2=res2.to(G.DEV), res_inv_slope=10.0, act_slope=0.1,
                                   act_scale=1 / 3, want_raw=False, **kw)
            y = (ref + x_rec + res2.double()) / 3
            want_raw, want_act = None, torch.where(y > 0, y, 0.1 * y)
    finally:
        _lib.check(lib.vs_set_option(b"pair_conv", 1))
    for got, want in ((raw, want_raw), (act, want_act)):
        if want is None:
            continue
        want = want.clone()
        want[lo:hi] = 0
        got = got.cpu().double()
        assert torch.isfinite(got).all()
        assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-3
        assert got[lo:hi].abs().max().item() == 0


@pytest.mark.parametrize("R,C,k,dil", [(1000, 32, 3, 1), (777, 32, 11, 5), (1500, 32, 7, 3), (600, 64, 3, 5), (129, 32, 11, 1),
                                      (246 * 3, 32, 11, 3), (5000, 64, 7, 5), (118 * 40, 32, 11, 1), (70000, 32, 7, 1),
                                      (5000, 64, 11, 5), (118 * 9 + 5, 64, 11, 1), (40000, 64, 11, 3)])
def test_fused_resblock_pair_matches_cpu(G, R, C, k, dil):
    _check_respair(G, R, C, k, dil)


@pytest.mark.parametrize("R,C,k,dil", [(600, 64, 3, 5), (5000, 64, 7, 5), (5000, 64, 11, 5), (118 * 9 + 5, 64, 11, 1), (236 * 80 + 3, 64, 11, 3),
                                      (122 * 2 * 90 + 7, 64, 7, 1), (300, 128, 3, 1), (126 * 2 * 80 + 100, 128, 3, 5), (127, 128, 3, 3)])
def test_pair_fused_resblock_iteration_matches_cpu(G, R, C, k, dil):
    """The same fused iteration on a CTA pair (tcgen05 cta_group::2, csrc/umma_pairfused.cu): C = 64 at every k and C = 128 at
    k = 3, incl. single-tile inputs (the pair's second CTA has no rows), ragged tails and more units than CTA pairs."""
    from vispeech_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.vs_set_option(b"pair_fused", 2))
    try:
        _check_respair(G, R, C, k, dil)
    finally:
        _lib.check(lib.vs_set_option(b"pair_fused", 1))


@pytest.mark.parametrize("R,C,k,dil", [(600, 64, 3, 5), (5000, 64, 7, 5), (5000, 64, 11, 5), (118 * 9 + 5, 64, 11, 1), (40000, 64, 11, 3), (777, 64, 7, 1)])
def test_tap_paired_resblock_iteration_matches_cpu(G, R, C, k, dil):
    """The same fused iteration with the conv taps issued in pairs as N = 128 MMAs (option tap_pairs; csrc/umma_respair.cu): the odd
    taps' half of the accumulator is re-aligned in the epilogue (lane shuffle + exchange between lane quarters) - covers both convs
    paired (k = 3, 7; k = 11 at d = 1, 3) and conv2 only (k = 11, d = 5: no room for conv1's exchange buffer)."""
    from vispeech_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.vs_set_option(b"tap_pairs", 2))
    try:
        _check_respair(G, R, C, k, dil)
    finally:
        _lib.check(lib.vs_set_option(b"tap_pairs", 0))


def _check_respair(G, R, C, k, dil):
    """y = c2(lrelu(c1(lrelu(x)))) + x (modules.py:211-220) in one kernel.  The kernel takes a = lrelu(x) and recovers
    the residual as min(a, a/slope); compared with an fp64 chain using the same fp16 roundings (a, the intermediate),
    incl. a masked gap (rows that must act as zero padding for BOTH convs).  C = 64, k = 11 takes the kernel's TIGHT form
    (single input stage / single intermediate buffer, residual re-read from global memory)."""
    g = torch.Generator().manual_seed(R + k)
    x = torch.randn(R, C, generator=g)
    row_utt = torch.zeros(R, dtype=torch.int32)
    row_utt[300:340] = -1
    x[300:340] = 0                                       # ragged-rows invariant: gap rows of the input are zero
    w1 = torch.randn(k, C, C, generator=g) / (k * C) ** 0.5
    w2 = torch.randn(k, C, C, generator=g) / (k * C) ** 0.5
    b1, b2 = torch.randn(C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    res2 = torch.randn(R, C, generator=g)
    a = G.f16_round(torch.where(x > 0, x, 0.1 * x))     # what the previous kernel would have stored
    raw, act = G.respair(a.to(G.DEV), w1, w2, b1.to(G.DEV), b2.to(G.DEV), dil, res2=res2.to(G.DEV), act_slope=0.01,
                         act_scale=1 / 3, row_utt=row_utt.to(G.DEV))
    ad = a.double()
    x_rec = torch.minimum(ad, ad * float(np.float32(1.0) / np.float32(0.1)))
    c1 = G.ref_conv_rows(a, G.f16_round(w1), b1, dil=dil, pad_l=(k - 1) // 2)
    t = G.f16_round(torch.where(c1 > 0, c1, 0.1 * c1).float())
    t[300:340] = 0
    y = G.ref_conv_rows(t, G.f16_round(w2), b2, dil=1, pad_l=(k - 1) // 2) + x_rec + G.f16_round(res2).double()
    y[300:340] = 0
    scale = y.abs().max().item()
    assert (raw.cpu().double() - y).abs().max().item() <= 2e-3 * scale
    ya = y / 3
    ya = torch.where(ya > 0, ya, 0.01 * ya)
    assert (act.cpu().double() - ya).abs().max().item() <= 2e-3 * scale
    if R > 340:
        assert raw[300:340].abs().max().item() == 0


@pytest.mark.parametrize("R,row_div,gaps", [(100, 1, []), (640 * 3 + 77, 1, [(900, 960)]), (4096, 4, [(0, 8), (2000, 2100), (4000, 4096)]),
                                              (123456, 512, [(51200, 52224)])])
def test_mrf_stage_matches_cpu(G, R, row_div, gaps):
    """The decoder's last MRF stage as one kernel (csrc/umma_mrf.cu): 3 x ResBlock1 (k = 3, 7, 11; d = 1, 3, 5) + sum / 3 +
    leaky_relu(0.01) + conv_post + tanh (models.py:279-288, modules.py:210-223) against an fp64 chain in which only the
    conv OPERANDS are rounded to fp16, as in the kernel (x, the MRF sum and the conv_post input stay in full precision).
    Covers: one partial super tile, several super tiles with a ragged tail, gaps at both ends, more super tiles than CTAs."""
    g = torch.Generator().manual_seed(R)
    ks, ds = (3, 7, 11), (1, 3, 5)
    valid = torch.ones(R, dtype=torch.bool)
    for lo, hi in gaps:
        valid[lo:hi] = False
    row_utt = torch.where(valid[::row_div], 0, -1).to(torch.int32)
    valid = (row_utt >= 0).repeat_interleave(row_div)[:R]
    x0 = torch.randn(R, 32, generator=g) * 0.5
    x0[~valid] = 0                                        # ragged-rows invariant
    W = [[[torch.randn(k, 32, 32, generator=g) / (32 * k) ** 0.5 for _ in range(2)] for _ in ds] for k in ks]
    B = [[[torch.randn(32, generator=g) * 0.1 for _ in range(2)] for _ in ds] for _ in ks]
    post_w = torch.randn(7, 32, generator=g) / (7 * 32) ** 0.5
    wave = G.mrf32(x0.to(G.DEV), W, B, post_w, row_utt.to(G.DEV), row_div).cpu().double()

    lrelu = lambda t, s=0.1: torch.where(t > 0, t, s * t)
    hi16 = x0.to(torch.float16)
    xin = hi16.double() + (x0 - hi16.float()).to(torch.float16).double()
    vm = valid.double()[:, None]
    total = torch.zeros(R, 32, dtype=torch.float64)
    for j, k in enumerate(ks):
        x = xin.clone()
        for m, d in enumerate(ds):
            a = G.f16_round(lrelu(x).float()).double() * vm
            c1 = G.ref_conv_rows(a, G.f16_round(W[j][m][0]), B[j][m][0], dil=d, pad_l=(k - 1) // 2)
            t = G.f16_round(lrelu(c1).float()).double() * vm
            x = x + G.ref_conv_rows(t, G.f16_round(W[j][m][1]), B[j][m][1], dil=1, pad_l=(k - 1) // 2)
        total += x
    v = lrelu(total / 3, 0.01) * vm
    ref = torch.tanh(G.ref_conv_rows(v, post_w[:, :, None], None, pad_l=3)[:, 0]) * valid.double()
    err = (wave - ref).abs().max().item()
    print("R=%d  max|wave - ref| = %.3e  (|ref| max %.3f)" % (R, err, ref.abs().max().item()))
    assert torch.isfinite(wave).all()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item())
    assert wave[~valid].abs().max().item() == 0 if (~valid).any() else True


@pytest.mark.parametrize("R,row_div,gaps", [(100, 1, []), (480 * 3 + 77, 1, [(900, 960)]), (4096, 4, [(0, 8), (2000, 2100), (4000, 4096)]),
                                              (480 * 160 + 31, 256, [(25600, 26112)])])
def test_resblock64_matches_cpu(G, R, row_div, gaps):
    """The k = 3 ResBlock1 of the C = 64 stage as one kernel (csrc/umma_resblock.cu; modules.py:210-223) against an fp64 chain
    in which only the conv OPERANDS are rounded to fp16 (the residual stream stays in full precision, as it does in TMEM).
    Covers a partial super tile, ragged tails, gaps at both ends and more super tiles than CTAs."""
    g = torch.Generator().manual_seed(R + 1)
    ds = (1, 3, 5)
    valid = torch.ones(R, dtype=torch.bool)
    for lo, hi in gaps:
        valid[lo:hi] = False
    row_utt = torch.where(valid[::row_div], 0, -1).to(torch.int32)
    valid = (row_utt >= 0).repeat_interleave(row_div)[:R]
    x0 = torch.randn(R, 64, generator=g) * 0.7
    x0[~valid] = 0
    W = [[torch.randn(3, 64, 64, generator=g) / (64 * 3) ** 0.5 for _ in range(2)] for _ in ds]
    B = [[torch.randn(64, generator=g) * 0.1 for _ in range(2)] for _ in ds]
    lrelu = lambda t, s=0.1: torch.where(t > 0, t, s * t)
    a0 = G.f16_round(lrelu(x0))                            # what the previous kernel stores
    out = G.resblock64(a0.to(G.DEV), W, B, row_utt.to(G.DEV), row_div).cpu().double()

    vm = valid.double()[:, None]
    ad = a0.double()
    x = torch.minimum(ad, ad * float(np.float32(10.0)))   # the kernel's x0 = min(a, 10 a) in fp32
    for m, d in enumerate(ds):
        a = (ad if m == 0 else G.f16_round(lrelu(x).float()).double()) * vm
        c1 = G.ref_conv_rows(a, G.f16_round(W[m][0]), B[m][0], dil=d, pad_l=1)
        t = G.f16_round(lrelu(c1).float()).double() * vm
        x = x + G.ref_conv_rows(t, G.f16_round(W[m][1]), B[m][1], dil=1, pad_l=1)
    ref = x * vm
    err = (out - ref).abs().max().item()
    print("R=%d  max|out - ref| = %.3e  (|ref| max %.3f)" % (R, err, ref.abs().max().item()))
    assert torch.isfinite(out).all()
    assert err <= 2e-3 * ref.abs().max().item()
    assert out[~valid].abs().max().item() == 0 if (~valid).any() else True


@pytest.mark.parametrize("stage", [0, 2, 3])
def test_umma_conv_transpose_polyphase(G, stage):
    """ConvTranspose1d (models.py:257-259) as a polyphase UMMA conv vs F.conv_transpose1d."""
    from vispeech_b200.packing import UP_KERNELS, UP_RATES, up_columns, ups_phase_range, ups_union_taps
    s, K = UP_RATES[stage], UP_KERNELS[stage]
    cin, cout = 512 >> stage, 256 >> stage
    R = 260
    g = torch.Generator().manual_seed(stage)
    x = torch.randn(R, cin, generator=g)
    wt = torch.randn(cin, cout, K, generator=g) / (cin * K / s) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    taps, pad_l = ups_union_taps(stage)
    uni = torch.zeros(taps, cin, s * cout)
    pad = (K - s) // 2
    cols = up_columns(cout, s)
    for ph in range(s):
        lo, hi = ups_phase_range(stage, ph)
        for d in range(lo, hi + 1):
            uni[d + pad_l, :, cols[ph * cout:(ph + 1) * cout]] = wt[:, :, ph + pad - s * d]
    raw, _ = G.umma_conv(x.to(G.DEV), uni, b.to(G.DEV), pad_l=pad_l, up=s, want_act=False)
    ref = torch.nn.functional.conv_transpose1d(G.f16_round(x).t()[None].double(), G.f16_round(wt).double(), b.double(),
                                               stride=s, padding=pad)[0].t()
    assert raw.shape == ref.shape
    assert (raw.cpu().double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("R,cin,cout,k,dil", [(100, 192, 576, 1, 1), (333, 192, 768, 3, 1), (77, 768, 192, 3, 1),
                                             (500, 96, 192, 1, 1), (260, 32, 1, 7, 1), (190, 64, 64, 11, 5)])
def test_conv_f32_matches_cpu(G, R, cin, cout, k, dil):
    """fp32 CUDA-core conv vs fp64: tolerance = fp32 accumulation error."""
    g = torch.Generator().manual_seed(k + cout)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(k, cin, cout, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    out = G.conv_f32(x.to(G.DEV), w.to(G.DEV), b.to(G.DEV), dil=dil, pad_l=(k - 1) // 2)
    ref = G.ref_conv_rows(x, w, b, dil=dil, pad_l=(k - 1) // 2)
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-4


@pytest.mark.parametrize("R,cin,cout,k,dil,act", [(300, 192, 576, 1, 1, 0), (1000, 192, 768, 3, 1, 1), (260, 768, 192, 3, 1, 0),
                                                 (500, 96, 192, 1, 1, 0), (700, 192, 384, 5, 1, 0), (129, 192, 96, 1, 1, 0),
                                                 (4100, 192, 192, 1, 1, 0)])
def test_conv_tf32_matches_cpu(G, R, cin, cout, k, dil, act):
    """TF32 tensor-core conv (operands rounded to 10-bit mantissa, fp32 accumulate) vs fp64 on the same rounded
    operands (tolerance: fp32 accumulation), and vs un-rounded fp64 (tolerance: TF32 operand rounding, 2^-11 relative
    per operand -> ~1e-3 of the output scale)."""
    from vispeech_b200.packing import round_tf32
    g = torch.Generator().manual_seed(k * 7 + cout)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(k, cin, cout, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    row_utt = torch.zeros(R, dtype=torch.int32)
    row_utt[5:9] = -1
    out = G.conv_tf32(x.to(G.DEV), w, b.to(G.DEV), dil=dil, pad_l=(k - 1) // 2, act=act, row_utt=row_utt.to(G.DEV)).cpu().double()
    ref_r = G.ref_conv_rows(round_tf32(x), round_tf32(w), b, dil=dil, pad_l=(k - 1) // 2)
    ref = G.ref_conv_rows(x, w, b, dil=dil, pad_l=(k - 1) // 2)
    if act:
        ref_r, ref = ref_r.clamp_min(0), ref.clamp_min(0)
    ref_r[5:9] = 0
    ref[5:9] = 0
    assert (out - ref_r).abs().max().item() <= 2e-4
    assert (out - ref).abs().max().item() <= 5e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("R,cin,cout,k,split3", [(300, 192, 576, 1, 0), (1000, 192, 768, 3, 1), (700, 192, 384, 5, 0), (129, 96, 192, 1, 1)])
def test_conv_tf32_cluster_multicast_matches_single_cta(G, R, cin, cout, k, split3):
    """The optional 2-CTA cluster form of the TF32 conv (each CTA fetches half of every weight slab and multicasts it; ring
    slots released by a multicast tcgen05.commit; odd tile counts leave one phantom tile) must be bit-identical to the
    single-CTA form: same MMAs in the same order per tile."""
    from vispeech_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(R + cout)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(k, cin, cout, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    outs = []
    for cl in (1, 2):
        _lib.check(lib.vs_set_option(b"tf32_cluster", cl))
        try:
            outs.append(G.conv_tf32(x.to(G.DEV), w, b.to(G.DEV), dil=1, pad_l=(k - 1) // 2, act=0, row_utt=None,
                                    split3=bool(split3)).cpu())
        finally:
            _lib.check(lib.vs_set_option(b"tf32_cluster", 1))
    assert torch.isfinite(outs[1]).all() and torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("R,cin,cout,k,act", [(300, 192, 576, 1, 0), (1000, 192, 768, 3, 1), (260, 768, 192, 3, 0),
                                             (500, 96, 192, 1, 0), (700, 192, 384, 5, 0), (2816, 192, 256, 3, 1)])
def test_conv_3xtf32_is_fp32_accurate(G, R, cin, cout, k, act):
    """Error-compensated 3xTF32 (a_hi w_hi + a_lo w_hi + a_hi w_lo) vs fp64: the tolerance is that of an fp32 conv
    (1e-4 absolute on O(1) outputs), i.e. ~50x tighter than plain TF32."""
    g = torch.Generator().manual_seed(k * 11 + cout)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(k, cin, cout, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    out = G.conv_tf32(x.to(G.DEV), w, b.to(G.DEV), pad_l=(k - 1) // 2, act=act, split3=True).cpu().double()
    ref = G.ref_conv_rows(x, w, b, pad_l=(k - 1) // 2)
    if act:
        ref = ref.clamp_min(0)
    assert (out - ref).abs().max().item() <= 1e-4


@pytest.mark.parametrize("R,cin,cout,k,act", [(300, 192, 576, 1, 0), (1000, 192, 768, 3, 1), (260, 768, 192, 3, 0), (2560, 768, 192, 3, 0),
                                             (500, 96, 192, 1, 0), (700, 192, 384, 5, 0), (2816, 192, 256, 3, 1), (27840, 192, 192, 1, 0),
                                             (129, 192, 96, 1, 0)])
def test_conv_split16_is_fp32_accurate(G, R, cin, cout, k, act):
    """fp16 hi/lo three-term conv (a_hi w_hi + a_lo w_hi + a_hi w_lo on tcgen05 kind::f16, csrc/umma_split.cu) vs fp64: the same
    fp32-level bar as 3xTF32 (1e-4 absolute on O(1) outputs), including K-slices (Cin = 768), n-blocks spread over CTAs (few
    row tiles), ragged row counts and gap rows."""
    g = torch.Generator().manual_seed(k * 13 + cout + R)
    x = torch.randn(R, cin, generator=g)
    w = torch.randn(k, cin, cout, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    row_utt = torch.zeros(R, dtype=torch.int32)
    row_utt[R // 3: R // 3 + 4] = -1                      # a gap: the conv must write exact zeros there
    x[R // 3: R // 3 + 4] = 0
    out = G.conv_split16(x.to(G.DEV), w, b.to(G.DEV), pad_l=(k - 1) // 2, act=act, row_utt=row_utt.to(G.DEV)).cpu().double()
    ref = G.ref_conv_rows(x, w, b, pad_l=(k - 1) // 2)
    if act:
        ref = ref.clamp_min(0)
    ref[R // 3: R // 3 + 4] = 0
    assert bool(torch.isfinite(out).all())
    assert float(out[R // 3: R // 3 + 4].abs().max()) == 0.0
    assert (out - ref).abs().max().item() <= 1e-4


def test_layernorm_rows(G):
    from vispeech_b200 import _lib
    lib = _lib.load()
    for C in (192, 256, 768):
        g = torch.Generator().manual_seed(C)
        a, b = torch.randn(50, C, generator=g), torch.randn(50, C, generator=g)
        gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
        row_utt = torch.zeros(50, dtype=torch.int32)
        row_utt[7] = -1
        out = torch.empty(50, C, device=G.DEV)
        args = [t.to(G.DEV) for t in (a, b, gamma, beta)]
        _lib.check(lib.vs_op_layernorm(*[t.data_ptr() for t in args], out.data_ptr(), 50, C,
                                       row_utt.to(G.DEV).data_ptr(), G.stream()))
        ref = torch.nn.functional.layer_norm(a + b, (C,), gamma, beta, 1e-5)
        ref[7] = 0
        assert (out.cpu() - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("kernel", [0, 2, 4])
@pytest.mark.parametrize("lengths", [[40], [3, 1, 70, 130], [431, 200], [129, 128, 127, 640]])
def test_rel_attention_matches_oracle(G, lengths, state_dict, kernel):
    """Banded relative attention (attentions.py:148-179) vs the oracle's restatement, several ragged lengths
    including T < window+1 and T > one key tile, for all three kernels (0 = CUDA-core fp32, 2 = 3xTF32 mma.sync, 4 = tcgen05
    with fp16 hi/lo operands, which the model uses from 128 rows per utterance up).  fp32 tolerance 2e-5 on O(1) outputs."""
    from oracle.vispeech_oracle import DEFAULT_CONFIG, relative_attention
    from vispeech_b200 import _lib
    lib = _lib.load()
    sd = state_dict
    p = "enc_p.encoder.attn_layers.1"
    rows = G.make_rows(lengths, [0] * len(lengths), 4, G.DEV)
    g = torch.Generator().manual_seed(sum(lengths))
    xs = [torch.randn(1, 192, n, generator=g) for n in lengths]
    qkv_rows = np.zeros((rows.n_rows, 576), np.float32)
    refs = []
    F = torch.nn.functional
    for b, x in enumerate(xs):
        q = F.conv1d(x, sd[p + ".conv_q.weight"], sd[p + ".conv_q.bias"])
        k = F.conv1d(x, sd[p + ".conv_k.weight"], sd[p + ".conv_k.bias"])
        v = F.conv1d(x, sd[p + ".conv_v.weight"], sd[p + ".conv_v.bias"])
        s = rows.starts[b]
        qkv_rows[s:s + lengths[b]] = torch.cat([q, k, v], 1)[0].t().numpy()
        # oracle output BEFORE conv_o: rerun with identity conv_o
        sd2 = dict(sd)
        sd2[p + ".conv_o.weight"] = torch.eye(192)[:, :, None]
        sd2[p + ".conv_o.bias"] = torch.zeros(192)
        refs.append(relative_attention(sd2, p, x, DEFAULT_CONFIG)[0].t())
    out = torch.empty(rows.n_rows, 192, device=G.DEV)
    d_qkv = torch.from_numpy(qkv_rows).to(G.DEV)
    d_ek = sd[p + ".emb_rel_k"][0].contiguous().to(G.DEV)
    d_ev = sd[p + ".emb_rel_v"][0].contiguous().to(G.DEV)
    _lib.check(lib.vs_set_option(b"attention_mma", kernel))
    try:
        ws = torch.empty(rows.n_rows * 8192 + len(lengths) * 200000 + 65536, dtype=torch.uint8, device=G.DEV)
        _lib.check(lib.vs_op_rel_attention(ctypes.byref(rows.struct), d_qkv.data_ptr(), d_ek.data_ptr(), d_ev.data_ptr(),
                                           out.data_ptr(), ws.data_ptr(), ws.numel(), G.stream()))
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.vs_set_option(b"attention_mma", 1))
    for b, ref in enumerate(refs):
        s = rows.starts[b]
        assert (out[s:s + lengths[b]].cpu() - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("lengths", [[70], [3, 1, 70, 100], [127, 64]])
def test_rel_attention_small_batch_variant_is_bit_identical(G, lengths, state_dict):
    """The CUDA-core attention takes 16 queries per CTA when 64-query CTAs would leave most SMs idle (option attention_small, the
    batch-1 path): same dot-product, softmax and P.V association per row, so an utterance's bits do not depend on the batch it is in."""
    from vispeech_b200 import _lib
    lib = _lib.load()
    p = "enc_p.encoder.attn_layers.0"
    rows = G.make_rows(lengths, [0] * len(lengths), 4, G.DEV)
    g = torch.Generator().manual_seed(len(lengths))
    d_qkv = torch.randn(rows.n_rows, 576, generator=g).to(G.DEV)
    d_ek = state_dict[p + ".emb_rel_k"][0].contiguous().to(G.DEV)
    d_ev = state_dict[p + ".emb_rel_v"][0].contiguous().to(G.DEV)
    outs = []
    _lib.check(lib.vs_set_option(b"attention_mma", 0))
    try:
        for small in (1, 0):
            _lib.check(lib.vs_set_option(b"attention_small", small))
            out = torch.zeros(rows.n_rows, 192, device=G.DEV)
            _lib.check(lib.vs_op_rel_attention(ctypes.byref(rows.struct), d_qkv.data_ptr(), d_ek.data_ptr(), d_ev.data_ptr(),
                                               out.data_ptr(), 0, 0, G.stream()))
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        _lib.check(lib.vs_set_option(b"attention_mma", 1))
        _lib.check(lib.vs_set_option(b"attention_small", 1))
    assert torch.equal(outs[0], outs[1]) and outs[0].abs().max().item() > 0


@pytest.mark.parametrize("R,first,last", [(300, 1, 0), (1000, 0, 0), (4229, 0, 1), (129, 1, 1)])
def test_wn_layer_kernel_matches_cpu(G, R, first, last):
    """The one-kernel WN layer (modules.py:153-176: in_layer k5 -> + cond(g) -> tanh.sigmoid -> res_skip 1x1 -> h / skip
    update) vs an fp64 statement with the kernel's operand roundings (h truncated to TF32 by the MMA, weights and the gate
    output rounded to TF32), incl. gap rows, per-speaker cond rows, first / last layer forms and odd tile counts."""
    from vispeech_b200 import _lib
    from vispeech_b200.packing import gate_columns, pack_tf32, round_tf32
    lib = _lib.load()
    H = 192
    g = torch.Generator().manual_seed(R + 2 * first + last)
    h = torch.randn(R, H, generator=g)
    skip0 = torch.randn(R, H, generator=g)
    row_utt = torch.zeros(R, dtype=torch.int32)
    row_utt[R // 2:] = 1
    row_utt[40:47] = -1
    h[40:47] = 0                                                 # ragged-rows invariant: gap rows of the input are zero
    w_in = torch.randn(5, H, 2 * H, generator=g) / (5 * H) ** 0.5
    b_in = torch.randn(2 * H, generator=g) * 0.1
    cond = torch.randn(3, 2 * H, generator=g) * 0.3             # 3 "speakers", this layer's columns
    sid = torch.tensor([2, 0], dtype=torch.int32)
    n_rs = H if last else 2 * H
    w_rs = torch.randn(1, H, n_rs, generator=g) / H ** 0.5
    b_rs = torch.randn(n_rs, generator=g) * 0.1
    perm = gate_columns(H)
    dev = G.DEV
    args = dict(w_in=pack_tf32(w_in[:, :, perm].contiguous()).to(dev), b_in=b_in[perm].contiguous().to(dev),
                cond=cond[:, perm].contiguous().to(dev), w_rs=pack_tf32(w_rs).to(dev), b_rs=b_rs.to(dev))
    d_h, d_skip = h.to(dev), skip0.clone().to(dev)
    d_sid, d_row_utt = sid.to(dev), row_utt.to(dev)              # keep the device copies alive across the call
    d_out = torch.full((R, H), float("nan"), device=dev)
    ws = torch.empty(3 * R * H * 4 + 4096, dtype=torch.uint8, device=dev)
    _lib.check(lib.vs_op_wn_layer(d_h.data_ptr(), args["w_in"].data_ptr(), args["b_in"].data_ptr(), args["cond"].data_ptr(), 2 * H,
                                  d_sid.data_ptr(), args["w_rs"].data_ptr(), args["b_rs"].data_ptr(),
                                  d_row_utt.data_ptr(), R, first, last, None if last else d_out.data_ptr(),
                                  d_skip.data_ptr(), ws.data_ptr(), ws.numel(), G.stream()), "vs_op_wn_layer")
    torch.cuda.synchronize()
    # reference
    h_tr = (h.view(torch.int32) & ~0x1FFF).view(torch.float32)   # the MMA reads the top 19 bits of the fp32 words
    a = G.ref_conv_rows(h_tr, round_tf32(w_in), b_in, dil=1, pad_l=2)
    valid = row_utt >= 0
    a = a + cond[sid.long()[row_utt.clamp_min(0).long()]].double()
    acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
    acts[~valid] = 0
    rs = round_tf32(acts.float()).double() @ round_tf32(w_rs)[0].double() + b_rs.double()
    skip_ref = (torch.zeros(R, H, dtype=torch.float64) if first else skip0.double()) + (rs if last else rs[:, H:])
    if first:
        skip_ref[~valid] = 0
    else:
        skip_ref[~valid] = skip0[~valid].double()
    assert (d_skip.cpu().double() - skip_ref).abs().max().item() <= 2e-3
    if not last:
        h_ref = h.double() + rs[:, :H]
        h_ref[~valid] = 0
        assert (d_out.cpu().double() - h_ref).abs().max().item() <= 2e-3
