"""CPU checks of the text -> ids bridge (reference text/symbols.py:1-22 and text/__init__.py:9-17),
the serving queue's text route and per-request validation, the utterance-list front end's control grouping and HParams."""
import os
import sys

import numpy as np
import pytest
import torch

from vispeech_b200 import text as T
from vispeech_b200.config import HParams, N_SYMBOLS, get_hparams_from_file
from vispeech_b200.serving import BatchingSynthesizer, create_app

REFERENCE = "/root/reference"


def test_symbol_table_shape_and_blocks():
    s = T.symbols
    assert len(s) == 519 == N_SYMBOLS and len(set(s)) == 519 and s[0] == "_"
    assert s[1:402] == sorted(s[1:402]) and "zh" in s[1:402] and "iiir5" in s[1:402]      # 401 zh
    assert s[402] == "ts." and s[402 + 41] == "dy." and s[444:513] == sorted(s[444:513])    # 42 ja, 69 en
    assert s[513:] == ["!", "?", "…", ",", ".", "sp"]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_symbol_table_equals_the_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_symbols", os.path.join(REFERENCE, "text", "symbols.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert T.symbols == m.symbols
    ref_map = {s: i for i, s in enumerate(m.symbols)}                  # text/__init__.py:7
    phones = ["n", "i3", "h", "ao3", "sp", "HH", "AH0", "pau", "."]
    assert T.cleaned_text_to_sequence(phones) == [ref_map[p] for p in phones]


def test_sequence_round_trip_and_unknown_symbol():
    phones = ["sh", "uo1", "h", "ua4", ",", "sp"]
    ids = T.cleaned_text_to_sequence(phones)
    assert T.sequence_to_cleaned_text(ids) == phones
    with pytest.raises(KeyError):
        T.cleaned_text_to_sequence(["not-a-phone"])
    assert T.remove_invalid_phonemes(["n", "-", "??", "--", "i3"]) == ["n", "sp", "sp", "i3"]     # cleaner.py:11-13,23-32


def test_frontend_cache_and_pluggable_g2p():
    calls = []

    def g2p(text):
        calls.append(text)
        return ["n", "i3", "-", "bogus", "h", "ao3"]

    fe = T.TextFrontend(g2p=g2p, cache_size=2)
    a = fe("你好")
    assert a == T.cleaned_text_to_sequence(["n", "i3", "sp", "h", "ao3"]) and calls == ["你好。"]     # inference_api.py:17
    assert fe("你好") == a and len(calls) == 1 and fe.hits == 1                                       # cached
    fe("早"), fe("安")
    fe("你好")                                                                                         # evicted (LRU of 2)
    assert len(calls) == 4
    assert fe("n i3 h ao3 sp") == T.cleaned_text_to_sequence("n i3 h ao3 sp".split()) and len(calls) == 4   # phoneme string: no G2P
    no_g2p = T.TextFrontend()
    no_g2p.g2p = None                                              # what a box without the reference's G2P stack sees
    with pytest.raises(ValueError):
        no_g2p("raw text")


class _FakeNet:
    n_vocab, n_speakers = 519, 200


def test_submit_validates_per_request_and_text_route():
    seen = []

    def fake(reqs):
        seen.append([r.sid for r in reqs])
        return [np.full(int(r.ids.numel()), 7, np.int16) for r in reqs]

    s = BatchingSynthesizer(net=_FakeNet(), synth_batch=fake, max_batch=8, max_wait_ms=20)
    good = [s.submit([1, 2, 3], 5, duration=[2, 2, 2]) for _ in range(3)]
    bad = [s.submit([1, 2, 3], 999, duration=[2, 2, 2]),            # sid out of range
           s.submit([1, 2, 600], 1, duration=[2, 2, 2]),            # id >= n_vocab (nn.Embedding would raise)
           s.submit([1, 2, 3], 1, duration=[2, 2]),                 # control shorter than ids
           s.submit([1, 2, 3], 1, f0=[100.0] * 4)]
    t = s.submit_text("n i3 h ao3 sp", 1)
    t_bad = s.submit_text("n i3 xyz", 1)
    assert all(f.result(timeout=10).shape[0] == 3 for f in good)
    assert all(isinstance(f.exception(timeout=10), ValueError) for f in bad + [t_bad])
    assert t.result(timeout=10).shape[0] == 5
    s.close()
    assert all(sid != 999 for batch in seen for sid in batch)       # the malformed requests never reached a batch


def test_failed_batch_is_retried_per_request():
    def flaky(reqs):
        if len(reqs) > 1:
            raise RuntimeError("batch failed")
        if reqs[0].sid == 3:
            raise RuntimeError("this one is bad")
        return [np.zeros(2, np.int16)]

    s = BatchingSynthesizer(synth_batch=flaky, max_batch=8, max_wait_ms=50)
    futs = [s.submit([1, 2], sid, duration=[1, 1]) for sid in (1, 2, 3, 4)]
    res = [f.exception(timeout=10) for f in futs]
    s.close()
    assert [r is None for r in res] == [True, True, False, True]


def test_asgi_route_text_and_ids():
    pytest.importorskip("fastapi")
    from starlette.testclient import TestClient
    s = BatchingSynthesizer(net=_FakeNet(), synth_batch=lambda reqs: [np.arange(int(r.ids.numel()), dtype=np.int16) for r in reqs],
                            max_wait_ms=1)
    try:
        c = TestClient(create_app(s))
    except Exception as e:                                         # httpx missing in this image: the route logic is covered above
        s.close()
        pytest.skip("starlette TestClient unavailable: %s" % e)
    r = c.get("/tts", params={"text": "n i3 h ao3 sp"})
    assert r.status_code == 200 and r.content[:4] == b"RIFF" and len(r.content) == 44 + 2 * 5
    assert c.get("/tts", params={"ids": "1,2,3", "sid": 2}).status_code == 200
    assert c.get("/tts", params={"text": "n i3 qqq"}).status_code == 400
    assert c.get("/tts").status_code == 400
    s.close()


def test_batching_groups_by_control_signature():
    """a bucket that mixes utterances with and without f0 / energy must not drop the supplied controls."""
    from vispeech_b200 import batching
    calls = []

    class Net:
        hop_length, sampling_rate, device = 512, 44100, torch.device("cpu")

        def infer(self, ids, lens, sid=None, noise_scale=1, duration_control=None, outputs="all", pitch_control=None,
                  energy_control=None, noise=None):
            calls.append((ids.shape[0], pitch_control is not None, energy_control is not None))
            tf = int(duration_control.sum(1).max())
            return (torch.zeros(ids.shape[0], 1, tf * 512), None)

    utts = []
    for i in range(6):
        u = {"ids": torch.arange(1, 5), "sid": 0, "duration": torch.tensor([3, 4, 5, 6])}
        if i % 2 == 0:
            u["f0"] = torch.full((4,), 200.0)
        if i % 3 == 0:
            u["energy"] = torch.full((4,), 50.0)
        utts.append(u)
    stats = {}
    out = batching.synthesize(Net(), utts, keep_on_device=True, stats=stats)
    assert stats["infer_calls"] == 4 and stats["utterances"] == 6 and abs(stats["imbalance"] - 1) < 1e-9
    assert sorted(out) == list(range(6))
    assert sorted(calls) == sorted([(1, True, True), (2, True, False), (1, False, True), (2, False, False)])


def test_hparams_access_styles():
    hps = get_hparams_from_file()
    assert hps.data.hop_length == hps["data"]["hop_length"] == 512 and "model" in hps and len(hps.model.keys()) > 5
    h = HParams(a={"b": 1}, c=2)
    h.d = {"e": 3}
    assert h.a.b == 1 and h["d"].e == 3 and dict(h.items())["c"] == 2
    with pytest.raises(AttributeError):
        h.nope
    assert "hidden_channels" in dict(**hps.model)                  # inference.py:34 splats **hps.model into the ctor
