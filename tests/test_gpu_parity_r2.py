"""GPU parity at the sizes BASELINE.json names (round-1 review: the bench batch itself, C5 at 256 utterances and the C3
latents were never compared with the oracle; the serving test compared the server with `infer` itself)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(state_dict):
    from vispeech_b200 import build_from_hparams, get_hparams_from_file
    n = build_from_hparams(get_hparams_from_file(), device="cuda:0")
    n.load_state_dict(state_dict)
    return n


def snr_db(ref, x):
    ref, x = ref.double().reshape(-1), x.double().reshape(-1)
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum().clamp_min(1e-300)))


def _pad(utts, key, dtype):
    tp = max(u["ids"].numel() for u in utts)
    out = torch.zeros(len(utts), tp, dtype=dtype)
    for b, u in enumerate(utts):
        out[b, : u[key].numel()] = u[key].to(dtype)
    return out


def test_bench_batch_b64_matches_oracle(net, state_dict):
    """The exact configs[1] batch bench.py times on rank 0 (oracle.inputs.c2(batch=64, seed=1): 64 utterances, 27.6 k frame
    rows, i.e. every precision regime the bench runs in: TF32 flow, 3xTF32 frame prior / phoneme level, fp16 decoder with the
    fp32-residual last stage), with injected noise, against per-utterance oracle runs of three of its utterances."""
    from oracle import inputs as oin
    from oracle.vispeech_oracle import infer_one
    utts = oin.c2(batch=64, seed=1)
    frames = oin.frame_counts(utts)
    noises = oin.draw_noise(frames, 77)
    o, x_mask, (z, z_p, m_p, logs_p), duration, f0, energy = net.infer(
        _pad(utts, "ids", torch.long), torch.LongTensor([u["ids"].numel() for u in utts]),
        sid=torch.LongTensor([u["sid"] for u in utts]), noise_scale=0.667, duration_control=_pad(utts, "duration", torch.long),
        noise=noises)
    torch.cuda.synchronize()
    assert int(x_mask.sum()) == sum(frames)
    worst = {"z": 0.0, "m_p": 0.0, "logs_p": 0.0, "f0": 0.0, "snr": 1e9}
    for b in (0, 31, 63):
        u, tf = utts[b], frames[b]
        ref = infer_one(state_dict, u["ids"], u["sid"], 0.667, noises[b], duration_control=u["duration"])
        worst["z"] = max(worst["z"], float((z[b, :, :tf].cpu() - ref["z"]).abs().max()))
        worst["m_p"] = max(worst["m_p"], float((m_p[b, :, :tf].cpu() - ref["m_p"]).abs().max()))
        worst["logs_p"] = max(worst["logs_p"], float((logs_p[b, :, :tf].cpu() - ref["logs_p"]).abs().max()))
        worst["f0"] = max(worst["f0"], float((f0[b, : u["ids"].numel()].cpu() - ref["F0"]).abs().max()))
        worst["snr"] = min(worst["snr"], snr_db(ref["o"], o[b, 0, : tf * 512].cpu()))
        assert float(o[b, 0, tf * 512:].abs().max() if tf * 512 < o.shape[2] else 0) == 0
    print("bench batch (B=64) worst case over 3 utterances:", worst)
    assert worst["z"] <= 1e-2 and worst["m_p"] <= 1e-2 and worst["logs_p"] <= 1e-2 and worst["f0"] <= 5e-2 and worst["snr"] >= 30.0, worst


def test_c5_b256_all_indices_exact(net, state_dict):
    """configs[4] at the size BASELINE.json names: 256 manual-edit utterances (float / zero / negative / very long durations,
    F0 and energy given).  Every utterance's expansion indices must equal the reference rule exactly (models.py:418-427);
    latents of two utterances against the oracle."""
    from oracle import inputs as oin
    from oracle.vispeech_oracle import expansion_indices, infer_one
    utts = oin.c5(batch=256, seed=4)
    frames = [int(expansion_indices(u["duration"]).numel()) for u in utts]
    noises = oin.draw_noise(frames, 55)
    P = net.prepare(_pad(utts, "ids", torch.long), torch.LongTensor([u["ids"].numel() for u in utts]),
                    sid=torch.LongTensor([u["sid"] for u in utts]), noise_scale=0.667,
                    duration_control=_pad(utts, "duration", torch.float64), pitch_control=_pad(utts, "f0", torch.float32),
                    energy_control=_pad(utts, "energy", torch.float32), noise=noises)
    z, rf = net.run(P, outputs="latents")                    # everything up to the decoder, one call for all 256
    torch.cuda.synchronize()
    idx = net.last_lr_index.cpu().numpy()
    covered = np.zeros(rf.n_rows, bool)
    for b, u in enumerate(utts):
        want = expansion_indices(u["duration"]).numpy()
        s, n = int(rf.starts[b]), int(rf.lengths[b])
        assert n == want.size and np.array_equal(idx[s:s + n], want), b
        covered[s:s + n] = True
    assert (idx[~covered] == -1).all() and int(covered.sum()) == sum(frames)
    zc = z.cpu()
    for b in (17, 200):
        u = utts[b]
        ref = infer_one(state_dict, u["ids"], u["sid"], 0.667, noises[b], duration_control=u["duration"], pitch_control=u["f0"],
                        energy_control=u["energy"], stop_after="flow")
        s, n = int(rf.starts[b]), int(rf.lengths[b])
        assert float((zc[s:s + n].t() - ref["z"]).abs().max()) <= 1e-2, b


def test_c3_sampled_latents_match_oracle(net, state_dict):
    """configs[2]: five of the 512 mixed-length (1-15 s) utterances as ONE ragged call - m_p, z and the waveform of each
    against its batch-1 oracle run (the round-1 C3 test only checked waveforms)."""
    from oracle import inputs as oin
    from oracle.vispeech_oracle import infer_one
    utts_all = oin.c3(batch=512, seed=2)
    pick = [3, 77, 200, 311, 508]
    utts = [utts_all[i] for i in pick]
    frames = oin.frame_counts(utts)
    noises = [oin.draw_noise([frames[k]], 1000 + i)[0] for k, i in enumerate(pick)]
    o, x_mask, (z, z_p, m_p, logs_p), duration, f0, energy = net.infer(
        _pad(utts, "ids", torch.long), torch.LongTensor([u["ids"].numel() for u in utts]),
        sid=torch.LongTensor([u["sid"] for u in utts]), noise_scale=0.667,
        duration_control=_pad(utts, "duration", torch.float64 if any(u["duration"].is_floating_point() for u in utts) else torch.long),
        noise=noises)
    torch.cuda.synchronize()
    for b, u in enumerate(utts):
        tf = frames[b]
        ref = infer_one(state_dict, u["ids"], u["sid"], 0.667, noises[b], duration_control=u["duration"])
        assert float((m_p[b, :, :tf].cpu() - ref["m_p"]).abs().max()) <= 1e-2, b
        assert float((z[b, :, :tf].cpu() - ref["z"]).abs().max()) <= 1e-2, b
        assert snr_db(ref["o"], o[b, 0, : tf * 512].cpu()) >= 30.0, b


def test_serving_queue_matches_oracle_and_numpy_pcm(net, state_dict):
    """8(f) ranks 1-2 against the ORACLE: requests through the batching server (text route included) come back as the
    s16 / 22.05 kHz stream that the oracle's waveform gives under a float64 numpy statement of the post-processing."""
    from oracle.vispeech_oracle import infer_one
    from vispeech_b200.postprocess import decimate_reference, default_fir
    from vispeech_b200.serving import BatchingSynthesizer
    from vispeech_b200.text import cleaned_text_to_sequence
    g = torch.Generator().manual_seed(12)
    srv = BatchingSynthesizer(net, max_batch=16, max_wait_ms=20)
    reqs = [(torch.randint(1, 500, (8 + 3 * i,), generator=g), 10 * i, torch.randint(2, 9, (8 + 3 * i,), generator=g)) for i in range(5)]
    futs = [srv.submit(ids, sid, duration=dur, noise_scale=0.0) for ids, sid, dur in reqs]
    phones = "n i3 h ao3 sp sh iii4 j ie4".split()
    t_dur = torch.randint(3, 9, (len(phones),), generator=g)
    futs.append(srv.submit_text(" ".join(phones), 1, duration=t_dur, noise_scale=0.0))
    reqs.append((torch.LongTensor(cleaned_text_to_sequence(phones)), 1, t_dur))
    outs = [f.result(timeout=120) for f in futs]
    srv.close()
    h = default_fir()
    for (ids, sid, dur), got in zip(reqs, outs):
        tf = int(dur.sum())
        ref = infer_one(state_dict, ids, sid, 0.0, torch.zeros(192, tf), duration_control=dur)["o"].double().numpy().reshape(-1)
        t_out = (ref.size + 1) // 2
        want = np.clip(np.rint(decimate_reference(ref, h) * 32768.0), -32768, 32767)
        assert got.dtype == np.int16 and got.shape[0] == t_out
        assert snr_db(torch.from_numpy(want), torch.from_numpy(got.astype(np.float64))) >= 30.0


def test_in_kernel_philox_noise_is_standard_normal_and_seeded(net):
    """a16: without an injected eps the sampling kernel draws it itself (Philox4x32-10 + Box-Muller, csrc/ops_misc.cu).
    Recovered from the outputs, eps = (z_p - m_p) / (exp(logs_p) noise_scale) must be N(0, 1), reproducible under
    torch.manual_seed and different from call to call."""
    from oracle import inputs as oin
    utts = oin.c2(batch=8, seed=3)
    args = (_pad(utts, "ids", torch.long), torch.LongTensor([u["ids"].numel() for u in utts]))
    kw = dict(sid=torch.LongTensor([u["sid"] for u in utts]), noise_scale=0.667, duration_control=_pad(utts, "duration", torch.long))

    def eps_of(seed):
        torch.manual_seed(seed)
        o, x_mask, (z, z_p, m_p, logs_p), *_ = net.infer(*args, **kw)
        m = x_mask.expand_as(z_p)
        return ((z_p - m_p) / (torch.exp(logs_p) * 0.667))[m].double().cpu()

    a, b, c = eps_of(123), eps_of(123), eps_of(124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    n = a.numel()
    assert n > 500000
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    assert abs(float((a ** 4).mean()) - 3) < 0.05 and abs(float((a ** 3).mean())) < 0.02           # kurtosis 3, no skew
    assert 4.0 < float(a.abs().max()) < 7.0
    assert abs(float((a[:-1] * a[1:]).mean())) < 5e-3                                                # neighbours uncorrelated


def test_cuda_graph_latency_path_matches_infer(net, state_dict):
    """infer_graphed (one CUDA-graph launch per row bucket) against `infer` and the oracle: utterances of different lengths
    replay the SAME captured graph (bucket = frame rows rounded up to 64) - the second and third calls below never capture."""
    from oracle import inputs as oin
    from oracle.vispeech_oracle import infer_one
    g = torch.Generator().manual_seed(21)
    base = oin.c1()[0]
    tp = base["ids"].numel()
    cases = []
    for k in range(4):
        dur = base["duration"].clone()
        dur[k] += 3 * k                                    # 0 .. 9 more frames: same 64-row bucket for the first three
        if k == 3:
            dur[5] += 70                                   # another bucket
        cases.append(dict(ids=torch.randint(1, 400, (tp,), generator=g), sid=7 * k, duration=dur))
    n_graphs = []
    for u in cases:
        tf = int(u["duration"].sum())
        eps = oin.draw_noise([tf], 400 + tf)[0]
        args = (u["ids"][None], torch.LongTensor([tp]))
        kw = dict(sid=torch.LongTensor([u["sid"]]), noise_scale=0.667, duration_control=u["duration"][None], noise=[eps])
        o_g, mask_g = net.infer_graphed(*args, **kw)
        o_g = o_g.clone()
        n_graphs.append(len(net._graphs))
        o_e, mask_e, *_ = net.infer(*args, outputs="audio", **kw)
        torch.cuda.synchronize()
        assert o_g.shape == o_e.shape == (1, 1, tf * 512) and torch.equal(mask_g, mask_e)
        assert snr_db(o_e, o_g) >= 60.0
        ref = infer_one(state_dict, u["ids"], u["sid"], 0.667, eps, duration_control=u["duration"])
        assert snr_db(ref["o"], o_g[0, 0].cpu()) >= 30.0
    assert n_graphs == [1, 1, 1, 2], n_graphs
    # without injected noise: eps is drawn by the library (vs_randn) into the graph's static buffer, fresh for every call
    u = cases[0]
    a = net.infer_graphed(u["ids"][None], torch.LongTensor([tp]), sid=torch.LongTensor([0]), noise_scale=0.667,
                          duration_control=u["duration"][None])[0].clone()
    b = net.infer_graphed(u["ids"][None], torch.LongTensor([tp]), sid=torch.LongTensor([0]), noise_scale=0.667,
                          duration_control=u["duration"][None])[0].clone()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(a).all()) and not torch.equal(a, b)


@pytest.mark.parametrize("direction", ["reverse", "forward"])
def test_coupling_layer_kernel_matches_oracle_flow(net, state_dict, direction):
    """The flow with every coupling layer as ONE kernel (csrc/umma_coupling.cu; modules.py:324-343 + its WN, 148-176; models.py:202-209):
    13 utterances of ragged lengths (5,300 frame rows: the regime the kernel serves, >= tf32_min_rows), different speakers, gaps
    between utterances, tiles that straddle utterance boundaries - against the fp32 oracle per utterance (bar: latent <= 1e-2)
    and against the layer-by-layer TF32 path it replaces."""
    import ctypes
    from oracle.vispeech_oracle import DEFAULT_CONFIG, flow_forward, flow_reverse
    from vispeech_b200 import _lib
    from vispeech_b200._lib import check, ptr
    from vispeech_b200.layout import make_rows
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    lens = [431, 17, 900, 112, 113, 640, 5, 777, 300, 224, 1000, 450, 391]
    sids = [int(x) for x in torch.randint(0, 200, (len(lens),), generator=g)]
    rows = make_rows(lens, sids, 4, "cuda:0")
    R = rows.n_rows
    assert R >= 4096
    zs = [torch.randn(192, n, generator=g) for n in lens]
    z0 = torch.zeros(R, 192)
    row_utt = rows.row_utt.cpu().numpy()
    for b, n in enumerate(lens):
        s = int(np.nonzero(row_utt == b)[0][0])
        z0[s:s + n] = zs[b].t()
    ws = torch.empty(int(lib.vs_workspace_bytes_latent(net._model, R, R)), dtype=torch.uint8, device="cuda:0")
    fn = lib.vs_flow_reverse if direction == "reverse" else lib.vs_flow_forward
    outs = {}
    for fused in (1, 0):
        net.set_option("coupling_fused", fused)
        z = z0.to("cuda:0")
        check(fn(net._model, ctypes.byref(rows.struct), ptr(z), ptr(ws), ws.numel(), torch.cuda.current_stream().cuda_stream), "flow")
        torch.cuda.synchronize()
        outs[fused] = z.cpu()
    net.set_option("coupling_fused", 1)
    worst = 0.0
    for b, n in enumerate(lens):
        s = int(np.nonzero(row_utt == b)[0][0])
        gvec = state_dict["emb_g.weight"][sids[b]].reshape(1, -1, 1)
        ref = (flow_reverse if direction == "reverse" else flow_forward)(state_dict, zs[b][None], gvec, DEFAULT_CONFIG)[0].t()
        err = float((outs[1][s:s + n] - ref).abs().max())
        err0 = float((outs[0][s:s + n] - ref).abs().max())
        worst = max(worst, err)
        assert err <= 1e-2, (b, n, err, err0)
    print("coupling kernel, %s: worst |z - oracle| = %.2e; vs the layer-by-layer TF32 path %.2e" %
          (direction, worst, float((outs[1] - outs[0]).abs().max())))
    assert float(outs[1][row_utt < 0].abs().max()) == 0.0          # gap rows stay zero


def test_coupling_kernel_on_the_latency_path_meets_the_bars(net):
    """The one-kernel coupling layer on small calls (option coupling_min_rows = 1, the default: 76 -> 4 launches per flow pass) AND the
    fp32 / three-term kernels it replaces there (coupling_min_rows = 4096) on the golden utterances: latent z within 1e-2 of the
    reference, waveform SNR >= 30 dB."""
    import glob, os
    paths = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                   if not os.path.basename(p).startswith(("vc", "filelist")))
    from test_gpu_infer import run_golden
    try:
        for min_rows in (1, 4096):
            net.set_option("coupling_min_rows", min_rows)
            for path in paths:
                d = dict(np.load(path))
                o, x_mask, (z, z_p, m_p, logs_p), duration, f0, energy = run_golden(net, d, 0)
                err = float(np.abs(z[0].cpu().numpy() - d["z"]).max())
                ref = torch.from_numpy(d["o"].astype(np.float32)) / (64 if int(d["o_is_f16x64"]) else 1)
                snr = snr_db(ref, o[0, 0].cpu())
                print(min_rows, os.path.basename(path), "z err %.2e  snr %.1f dB" % (err, snr))
                assert err <= 1e-2 and snr >= 30.0, (path, err, snr)
    finally:
        net.set_option("coupling_min_rows", 1)
