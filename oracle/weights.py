"""Deterministic synthetic state dict with the reference's inference-path schema.

Test infrastructure (see oracle/__init__.py).  The reference ships no checkpoint
(SURVEY.md 8c), so every parity test runs on random weights.  `/root/reference` does
not exist on the GPU box, so the weights cannot come from constructing the reference
model there; instead this module draws every tensor of the schema in SURVEY.md
Appendix B (checked key-for-key against the real `SynthesizerTrn.state_dict()` by
`tests/golden/make_golden.py`) from a CPU `torch.Generator` seeded per key, with the
scale PyTorch's default initialisers would give, so activations stay well conditioned.

Deliberate departures from the reference's *initial* values (all are legal checkpoint
contents, chosen so that bugs cannot hide):
  * `flow.flows.{0,2,4,6}.post.{weight,bias}` are random, not zero (reference
    zero-inits them, modules.py:320-322, which would make the flow an identity);
  * LayerNorm gamma/beta are 1+0.1n / 0.1n instead of 1 / 0;
  * `weight_g` is ||weight_v|| * (1 + 0.1n) so that the weight-norm fold matters.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class ModelConfig:
    """The constants of configs/config.json:39-90 plus the ctor args of inference.py:26-33."""
    n_vocab: int = 519            # len(text/symbols.py:22)
    inter_channels: int = 192
    hidden_channels: int = 192
    filter_channels: int = 768
    n_heads: int = 2
    n_layers: int = 4             # text encoder and frame prior; pitch net hard-codes 6 (models.py:498)
    pitch_layers: int = 6
    kernel_size: int = 3
    window_size: int = 4          # attentions.py:14 default
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    upsample_rates: Tuple[int, ...] = (8, 8, 4, 2)
    upsample_initial_channel: int = 512
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4)
    n_speakers: int = 200
    gin_channels: int = 256
    flow_layers: int = 4          # WN n_layers (models.py:597)
    flow_kernel: int = 5
    n_flows: int = 4
    dp_filter: int = 256          # DurationPredictor(192, 256, 3, .5) models.py:599
    ep_filter: int = 768          # frame_prior_network.py:65
    sampling_rate: int = 44100
    hop_length: int = 512
    spec_channels: int = 1025     # filter_length // 2 + 1 (inference.py:28)

    @property
    def head_dim(self) -> int:
        return self.hidden_channels // self.n_heads


DEFAULT_CONFIG = ModelConfig()


def schema(cfg: ModelConfig = DEFAULT_CONFIG, with_vc: bool = False) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every tensor `infer` reads (with_vc: plus enc_q.* for `voice_conversion`)."""
    H, F, G = cfg.hidden_channels, cfg.filter_channels, cfg.gin_channels
    dk, W = cfg.head_dim, 2 * cfg.window_size + 1
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(prefix, co, ci, k, bias=True):
        out.append((prefix + ".weight", (co, ci, k), "conv"))
        if bias:
            out.append((prefix + ".bias", (co,), "bias:%d" % (ci * k)))

    def wn_conv(prefix, d0, d1, k, fan_in, bias_len):
        out.append((prefix + ".bias", (bias_len,), "bias:%d" % fan_in))
        out.append((prefix + ".weight_g", (d0, 1, 1), "wn_g"))
        out.append((prefix + ".weight_v", (d0, d1, k), "conv_fan:%d" % fan_in))

    def encoder(prefix, n_layers):
        for i in range(n_layers):
            a = "%s.attn_layers.%d" % (prefix, i)
            out.append((a + ".emb_rel_k", (1, W, dk), "rel"))
            out.append((a + ".emb_rel_v", (1, W, dk), "rel"))
            for n in "qkvo":
                conv("%s.conv_%s" % (a, n), H, H, 1)
            out.append(("%s.norm_layers_1.%d.gamma" % (prefix, i), (H,), "gamma"))
            out.append(("%s.norm_layers_1.%d.beta" % (prefix, i), (H,), "beta"))
            conv("%s.ffn_layers.%d.conv_1" % (prefix, i), F, H, cfg.kernel_size)
            conv("%s.ffn_layers.%d.conv_2" % (prefix, i), H, F, cfg.kernel_size)
            out.append(("%s.norm_layers_2.%d.gamma" % (prefix, i), (H,), "gamma"))
            out.append(("%s.norm_layers_2.%d.beta" % (prefix, i), (H,), "beta"))

    out.append(("emb_g.weight", (cfg.n_speakers, G), "normal:1.0"))
    out.append(("enc_p.symbol_emb.weight", (cfg.n_vocab, H), "normal:%r" % (H ** -0.5)))
    encoder("enc_p.encoder", cfg.n_layers)
    encoder("pitch_predictor.pitch_net", cfg.pitch_layers)
    encoder("frame_prior_net.fft_block", cfg.n_layers)

    D = cfg.dp_filter
    conv("duration_predictor.conv_1", D, H, 3)
    out.append(("duration_predictor.norm_1.gamma", (D,), "gamma"))
    out.append(("duration_predictor.norm_1.beta", (D,), "beta"))
    conv("duration_predictor.conv_2", D, D, 3)
    out.append(("duration_predictor.norm_2.gamma", (D,), "gamma"))
    out.append(("duration_predictor.norm_2.beta", (D,), "beta"))
    conv("duration_predictor.proj", 1, D, 1)
    conv("duration_predictor.cond", H, G, 1)

    conv("pitch_predictor.proj_f0", 1, H, 1)
    conv("pitch_predictor.cond", H, G, 1)

    E = cfg.ep_filter
    conv("energy_predictor.cond", H, G, 1)
    p = "energy_predictor.predictor.conv_layer"
    conv(p + ".conv_1.conv", E, H, 3)
    out.append((p + ".layer_norm_1.weight", (E,), "gamma"))
    out.append((p + ".layer_norm_1.bias", (E,), "beta"))
    conv(p + ".conv_2.conv", E, E, 3)
    out.append((p + ".layer_norm_2.weight", (E,), "gamma"))
    out.append((p + ".layer_norm_2.bias", (E,), "beta"))
    out.append(("energy_predictor.predictor.linear_layer.weight", (1, E), "conv_fan:%d" % E))
    out.append(("energy_predictor.predictor.linear_layer.bias", (1,), "bias:%d" % E))

    conv("pitch_prenet", H, 1, 3)
    conv("energy_prenet", H, 1, 3)
    conv("project.proj", 2 * cfg.inter_channels, H, 1)

    half = cfg.inter_channels // 2
    for f in range(0, 2 * cfg.n_flows, 2):
        p = "flow.flows.%d" % f
        conv(p + ".pre", H, half, 1)
        conv(p + ".post", half, H, 1)
        for i in range(cfg.flow_layers):
            wn_conv("%s.enc.in_layers.%d" % (p, i), 2 * H, H, cfg.flow_kernel, H * cfg.flow_kernel, 2 * H)
            rs = 2 * H if i < cfg.flow_layers - 1 else H
            wn_conv("%s.enc.res_skip_layers.%d" % (p, i), rs, H, 1, H, rs)
        wn_conv(p + ".enc.cond_layer", 2 * H * cfg.flow_layers, G, 1, G, 2 * H * cfg.flow_layers)

    C0 = cfg.upsample_initial_channel
    conv("dec.conv_pre", C0, cfg.inter_channels, 7)
    conv("dec.cond", C0, G, 1)
    ch = C0
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = C0 // (2 ** i), C0 // (2 ** (i + 1))
        # ConvTranspose1d weight is [Cin, Cout, k]; weight-norm dim 0 = in-channels
        wn_conv("dec.ups.%d" % i, cin, cout, k, cin * k // u, cout)
        ch = cout
        for j, kk in enumerate(cfg.resblock_kernel_sizes):
            r = "dec.resblocks.%d" % (i * len(cfg.resblock_kernel_sizes) + j)
            for m in range(len(cfg.resblock_dilation_sizes[j])):
                wn_conv("%s.convs1.%d" % (r, m), ch, ch, kk, ch * kk, ch)
                wn_conv("%s.convs2.%d" % (r, m), ch, ch, kk, ch * kk, ch)
    conv("dec.conv_post", 1, ch, 7, bias=False)
    if with_vc:   # PosteriorEncoder(spec_channels, inter, hidden, 5, 1, 16, gin) models.py:595-596
        conv("enc_q.pre", H, cfg.spec_channels, 1)
        conv("enc_q.proj", 2 * cfg.inter_channels, H, 1)
        for i in range(16):
            wn_conv("enc_q.enc.in_layers.%d" % i, 2 * H, H, 5, H * 5, 2 * H)
            rs = 2 * H if i < 15 else H
            wn_conv("enc_q.enc.res_skip_layers.%d" % i, rs, H, 1, H, rs)
        wn_conv("enc_q.enc.cond_layer", 2 * H * 16, G, 1, G, 2 * H * 16)
    return out


def _gen_for(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(("%d:%s" % (seed, key)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def make_state_dict(seed: int = 1234, cfg: ModelConfig = DEFAULT_CONFIG, with_vc: bool = False) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors keyed exactly like the reference `state_dict()` (un-folded weight norm)."""
    sd: Dict[str, torch.Tensor] = {}
    entries = schema(cfg, with_vc)
    for key, shape, kind in entries:
        if kind == "wn_g":
            continue
        g = _gen_for(key, seed)
        if kind == "conv":
            fan_in = shape[1] * shape[2]
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind.startswith("conv_fan:"):
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind.startswith("bias:"):
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind.startswith("normal:"):
            t = torch.randn(shape, generator=g) * float(kind.split(":")[1])
        elif kind == "rel":
            t = torch.randn(shape, generator=g) * (shape[-1] ** -0.5)
        elif kind == "gamma":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "beta":
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            raise ValueError(kind)
        sd[key] = t.float().contiguous()
    for key, shape, kind in entries:
        if kind != "wn_g":
            continue
        v = sd[key[: -len("weight_g")] + "weight_v"]
        g = _gen_for(key, seed)
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
        sd[key] = (norm * (1.0 + 0.1 * torch.randn(shape, generator=g))).float().contiguous()
    return sd


def fold_weight_norm(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """w = g * v / ||v||, norm over every dim except 0 (torch.nn.utils.weight_norm, dim=0).

    Used by the reference implicitly through the weight-norm forward pre-hook on
    dec.ups.*, dec.resblocks.*, flow.*.enc.* (models.py:257-259, modules.py:130,137,146,191-208).
    """
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)
