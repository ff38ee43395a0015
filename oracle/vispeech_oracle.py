"""CPU fp32 restatement of `SynthesizerTrn.infer` (reference models.py:672-722).  Device-agnostic: with a state dict on
`cuda` the same eager op sequence runs on the GPU - bench.py's clearly-labelled `gpu_eager_baseline` (what the reference's
own ATen / cuDNN path would do on this box), never the product path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path and the
`cpu_baseline` of bench.py.  It is a *restatement*, not an import: `/root/reference` does
not exist on the GPU box.  It is pinned against the unmodified reference through
`tests/golden/*.npz` (made by tests/golden/make_golden.py, checked by
tests/test_oracle_golden.py).

Semantics: one utterance at a time, exactly what every reference call site does
(inference.py:41-44, inference_api.py:41-47, gui.py:96-100, train.py:289-301 are all
batch 1).  A batch is a Python loop over utterances (`infer_batch`); this is the
"per-utterance batch-1" parity definition of SURVEY.md Appendix D, Q1.

Functions take the reference's own state dict (weight-norm un-folded) and plain tensors.
Layout is the reference's [1, C, T].
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn.functional as F

from .weights import DEFAULT_CONFIG, ModelConfig, fold_weight_norm

LRELU_SLOPE = 0.1  # modules.py:17
Control = Union[None, float, int, torch.Tensor]


# ----------------------------------------------------------------------------- blocks
def channel_layer_norm(x, gamma, beta, eps=1e-5):
    """modules.py:29-32 - LayerNorm over the channel dim of [1, C, T]."""
    return F.layer_norm(x.transpose(1, 2), (x.shape[1],), gamma, beta, eps).transpose(1, 2)


def relative_attention(sd, p, x, cfg: ModelConfig):
    """attentions.py:138-179 (MultiHeadAttention.forward/attention), self-attention, no pads.

    The reference pads the 9 relative embeddings out to 2T-1 and uses pad/reshape tricks
    (attentions.py:199-243); on an un-padded utterance that equals the banded form used
    here: scores[i,j] += q_i . Ek[j-i+w] and out_i += sum_d p[i,i+d] Ev[d+w] for |d| <= w.
    """
    H, nh, dk, w = cfg.hidden_channels, cfg.n_heads, cfg.head_dim, cfg.window_size
    T = x.shape[2]
    q = F.conv1d(x, sd[p + ".conv_q.weight"], sd[p + ".conv_q.bias"])
    k = F.conv1d(x, sd[p + ".conv_k.weight"], sd[p + ".conv_k.bias"])
    v = F.conv1d(x, sd[p + ".conv_v.weight"], sd[p + ".conv_v.bias"])
    q = q.view(nh, dk, T).transpose(1, 2) / math.sqrt(dk)      # [nh, T, dk]   (:151,155)
    k = k.view(nh, dk, T).transpose(1, 2)
    v = v.view(nh, dk, T).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(1, 2))                # [nh, T, T]
    ek = sd[p + ".emb_rel_k"][0]                               # [2w+1, dk], shared by heads (:125-128)
    ev = sd[p + ".emb_rel_v"][0]
    rel = torch.matmul(q, ek.t())                              # [nh, T, 2w+1]
    idx = torch.arange(T, device=x.device)
    band = idx[None, :] - idx[:, None]                         # j - i
    inband = band.abs() <= w
    bidx = (band + w).clamp(0, 2 * w)
    scores = scores + torch.where(inband[None], torch.gather(rel, 2, bidx[None].expand(nh, T, T)),
                                  torch.zeros((), dtype=x.dtype, device=x.device))
    pr = F.softmax(scores, dim=-1)                             # (:171)
    out = torch.matmul(pr, v)                                  # [nh, T, dk]
    # relative values: weights on the band only (:174-177)
    pb = torch.zeros(nh, T, 2 * w + 1, dtype=x.dtype, device=x.device)
    for d in range(-w, w + 1):
        lo, hi = max(0, -d), min(T, T - d)
        if hi > lo:
            ii = torch.arange(lo, hi, device=x.device)
            pb[:, ii, d + w] = pr[:, ii, ii + d]
    out = out + torch.matmul(pb, ev)
    out = out.transpose(1, 2).contiguous().view(1, H, T)       # (:178)
    return F.conv1d(out, sd[p + ".conv_o.weight"], sd[p + ".conv_o.bias"])


def ffn(sd, p, x, cfg: ModelConfig):
    """attentions.py:277-285 with _same_padding (:296-303); masks are all-ones here."""
    k = cfg.kernel_size
    pl, pr = (k - 1) // 2, k // 2
    h = F.conv1d(F.pad(x, (pl, pr)), sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"])
    h = torch.relu(h)
    return F.conv1d(F.pad(h, (pl, pr)), sd[p + ".conv_2.weight"], sd[p + ".conv_2.bias"])


def encoder(sd, p, x, n_layers, cfg: ModelConfig):
    """attentions.py:35-47 - post-LN transformer blocks."""
    for i in range(n_layers):
        y = relative_attention(sd, "%s.attn_layers.%d" % (p, i), x, cfg)
        x = channel_layer_norm(x + y, sd["%s.norm_layers_1.%d.gamma" % (p, i)], sd["%s.norm_layers_1.%d.beta" % (p, i)])
        y = ffn(sd, "%s.ffn_layers.%d" % (p, i), x, cfg)
        x = channel_layer_norm(x + y, sd["%s.norm_layers_2.%d.gamma" % (p, i)], sd["%s.norm_layers_2.%d.beta" % (p, i)])
    return x


def text_encoder(sd, ids, cfg: ModelConfig):
    """models.py:168-174.  enc_p.proj exists but is never applied."""
    x = F.embedding(ids, sd["enc_p.symbol_emb.weight"]) * math.sqrt(cfg.hidden_channels)
    x = x.t().unsqueeze(0)
    return encoder(sd, "enc_p.encoder", x, cfg.n_layers, cfg)


def duration_predictor(sd, x, g):
    """models.py:119-133."""
    p = "duration_predictor"
    h = x + F.conv1d(g, sd[p + ".cond.weight"], sd[p + ".cond.bias"])
    h = torch.relu(F.conv1d(h, sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"], padding=1))
    h = channel_layer_norm(h, sd[p + ".norm_1.gamma"], sd[p + ".norm_1.beta"])
    h = torch.relu(F.conv1d(h, sd[p + ".conv_2.weight"], sd[p + ".conv_2.bias"], padding=1))
    h = channel_layer_norm(h, sd[p + ".norm_2.gamma"], sd[p + ".norm_2.beta"])
    return F.conv1d(h, sd[p + ".proj.weight"], sd[p + ".proj.bias"])        # logw [1,1,T]


def pitch_predictor(sd, x, g, cfg: ModelConfig):
    """models.py:505-514."""
    p = "pitch_predictor"
    h = x + F.conv1d(g, sd[p + ".cond.weight"], sd[p + ".cond.bias"])
    h = encoder(sd, p + ".pitch_net", h, cfg.pitch_layers, cfg)
    return F.conv1d(h, sd[p + ".proj_f0.weight"], sd[p + ".proj_f0.bias"]).squeeze(1)   # [1,T]


def energy_predictor(sd, x, g):
    """frame_prior_network.py:119-124, 104-109, 50-55 (channel-last nn.LayerNorm(768))."""
    p = "energy_predictor"
    c = p + ".predictor.conv_layer"
    h = x + F.conv1d(g, sd[p + ".cond.weight"], sd[p + ".cond.bias"])
    for n in ("1", "2"):
        h = torch.relu(F.conv1d(h, sd["%s.conv_%s.conv.weight" % (c, n)], sd["%s.conv_%s.conv.bias" % (c, n)], padding=1))
        h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), sd["%s.layer_norm_%s.weight" % (c, n)],
                         sd["%s.layer_norm_%s.bias" % (c, n)], 1e-5).transpose(1, 2)
    out = F.linear(h.transpose(1, 2), sd[p + ".predictor.linear_layer.weight"], sd[p + ".predictor.linear_layer.bias"])
    return out.squeeze(-1)                                                   # [1,T]


def expansion_counts(duration: torch.Tensor) -> torch.Tensor:
    """models.py:421-423: n_i = max(int(d_i.item()), 0) - truncation toward zero, then clamp."""
    d = duration.reshape(-1)
    if d.is_floating_point():
        d = torch.trunc(d.double())
    return d.to(torch.int64).clamp_min(0)


def expansion_indices(duration: torch.Tensor) -> torch.Tensor:
    """Frame t copies phoneme idx[t] (models.py:418-427 `expand` + cat)."""
    n = expansion_counts(duration)
    return torch.repeat_interleave(torch.arange(n.numel(), device=n.device), n)


def wn(sd, p, x, g, cfg: ModelConfig, n_layers=None):
    """modules.py:148-176 (WN.forward) with the fused gate of commons.py:100-107."""
    H, L, k = cfg.hidden_channels, (n_layers or cfg.flow_layers), cfg.flow_kernel
    out = torch.zeros_like(x)
    gc = F.conv1d(g, fold_weight_norm(sd, p + ".cond_layer"), sd[p + ".cond_layer.bias"])
    for i in range(L):
        x_in = F.conv1d(x, fold_weight_norm(sd, "%s.in_layers.%d" % (p, i)), sd["%s.in_layers.%d.bias" % (p, i)],
                        padding=(k - 1) // 2)
        a = x_in + gc[:, 2 * H * i: 2 * H * (i + 1)]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
        rs = F.conv1d(acts, fold_weight_norm(sd, "%s.res_skip_layers.%d" % (p, i)),
                      sd["%s.res_skip_layers.%d.bias" % (p, i)])
        if i < L - 1:
            x = x + rs[:, :H]
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out


def flow_reverse(sd, z_p, g, cfg: ModelConfig):
    """models.py:202-209 reverse branch; layers modules.py:324-343 (mean_only) and Flip :270-277."""
    half = cfg.inter_channels // 2
    x = z_p
    for f in reversed(range(0, 2 * cfg.n_flows, 2)):
        x = torch.flip(x, [1])
        p = "flow.flows.%d" % f
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, sd[p + ".pre.weight"], sd[p + ".pre.bias"])
        h = wn(sd, p + ".enc", h, g, cfg)
        m = F.conv1d(h, sd[p + ".post.weight"], sd[p + ".post.bias"])
        x = torch.cat([x0, x1 - m], 1)
    return x


def flow_forward(sd, z, g, cfg: ModelConfig):
    """models.py:203-205 forward branch: RCL0, Flip, ..., RCL3, Flip with x1 = m + x1 (mean_only, modules.py:334-338)."""
    half = cfg.inter_channels // 2
    x = z
    for f in range(0, 2 * cfg.n_flows, 2):
        p = "flow.flows.%d" % f
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, sd[p + ".pre.weight"], sd[p + ".pre.bias"])
        h = wn(sd, p + ".enc", h, g, cfg)
        m = F.conv1d(h, sd[p + ".post.weight"], sd[p + ".post.bias"])
        x = torch.flip(torch.cat([x0, m + x1], 1), [1])
    return x


def posterior_encoder(sd, spec, g, noise, cfg: ModelConfig):
    """models.py:233-241 for one un-padded utterance: spec [1, 1025, T] -> z, m, logs."""
    x = F.conv1d(spec, sd["enc_q.pre.weight"], sd["enc_q.pre.bias"])
    x = wn(sd, "enc_q.enc", x, g, cfg, n_layers=16)
    stats = F.conv1d(x, sd["enc_q.proj.weight"], sd["enc_q.proj.bias"])
    m, logs = stats[:, :cfg.inter_channels], stats[:, cfg.inter_channels:]
    return m + noise.reshape(m.shape) * torch.exp(logs), m, logs


@torch.no_grad()
def voice_conversion_one(sd, spec: torch.Tensor, sid_src: int, sid_tgt: int, noise: torch.Tensor,
                         cfg: ModelConfig = DEFAULT_CONFIG) -> Dict[str, torch.Tensor]:
    """models.py:724-732 for one utterance: spec [1025, T] (linear spectrogram), noise [192, T]."""
    g_src = sd["emb_g.weight"][int(sid_src)].reshape(1, -1, 1)
    g_tgt = sd["emb_g.weight"][int(sid_tgt)].reshape(1, -1, 1)
    z, m_q, logs_q = posterior_encoder(sd, spec[None].float(), g_src, noise, cfg)
    z_p = flow_forward(sd, z, g_src, cfg)
    z_hat = flow_reverse(sd, z_p, g_tgt, cfg)
    o = generator(sd, z_hat, g_tgt, cfg)
    return dict(z=z[0], m_q=m_q[0], logs_q=logs_q[0], z_p=z_p[0], z_hat=z_hat[0], o=o[0, 0])


def resblock1(sd, p, x, k, dilations):
    """modules.py:210-223 with x_mask=None (the decoder passes no mask)."""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, fold_weight_norm(sd, "%s.convs1.%d" % (p, m)), sd["%s.convs1.%d.bias" % (p, m)],
                      padding=(k * d - d) // 2, dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, fold_weight_norm(sd, "%s.convs2.%d" % (p, m)), sd["%s.convs2.%d.bias" % (p, m)],
                      padding=(k - 1) // 2)
        x = xt + x
    return x


def generator(sd, z, g, cfg: ModelConfig):
    """models.py:271-290 (HiFi-GAN Generator.forward)."""
    x = F.conv1d(z, sd["dec.conv_pre.weight"], sd["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, sd["dec.cond.weight"], sd["dec.cond.bias"])
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, fold_weight_norm(sd, "dec.ups.%d" % i), sd["dec.ups.%d.bias" % i],
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            y = resblock1(sd, "dec.resblocks.%d" % (i * nk + j), x, cfg.resblock_kernel_sizes[j],
                          cfg.resblock_dilation_sizes[j])
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                      # default slope 0.01 (models.py:286, quirk Q3)
    x = F.conv1d(x, sd["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


# ----------------------------------------------------------------------------- the path
@torch.no_grad()
def infer_one(sd: Dict[str, torch.Tensor], ids: torch.Tensor, sid: int, noise_scale: float = 1.0,
              noise: Optional[torch.Tensor] = None, max_len: Optional[int] = None,
              energy_control: Control = None, pitch_control: Control = None, duration_control: Control = None,
              cfg: ModelConfig = DEFAULT_CONFIG, stop_after: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """One utterance through models.py:672-722.  `ids` is int64 [Tp]; controls are None /
    scalar / [Tp] tensors as in the reference.  `noise` is the eps of models.py:718,
    [192, Tf]; if None it is drawn with torch.randn.  Returns every stage tap.
    """
    ids = ids.reshape(-1).long()
    Tp = ids.numel()
    g = sd["emb_g.weight"][int(sid)].reshape(1, -1, 1)                        # :674
    x = text_encoder(sd, ids, cfg)                                            # :678
    taps = {"x_enc": x[0].clone()}

    if isinstance(duration_control, torch.Tensor):                            # :681-688
        duration = duration_control.reshape(-1)
    else:
        ctrl = 1 if duration_control is None else duration_control
        logw = duration_predictor(sd, x, g)
        taps["logw"] = logw.reshape(-1)
        duration = torch.ceil((torch.exp(logw) - 1) * ctrl).reshape(-1)
    taps["duration"] = duration

    if isinstance(pitch_control, torch.Tensor):                               # :691-698
        lf0 = (2595.0 * torch.log10(1.0 + pitch_control.reshape(1, -1).float() / 700.0)) / 500
    else:
        ctrl = 1 if pitch_control is None else pitch_control
        lf0 = pitch_predictor(sd, x, g, cfg) * ctrl
    x = x + F.conv1d(lf0.unsqueeze(1), sd["pitch_prenet.weight"], sd["pitch_prenet.bias"], padding=1)
    taps["lf0"] = lf0.reshape(-1)
    taps["F0"] = ((torch.pow(10, lf0 * 500 / 2590) - 1) * 700).reshape(-1)   # 2590: reference typo, kept (Q2)

    if isinstance(energy_control, torch.Tensor):                              # :701-708
        norm_e = (energy_control.reshape(1, -1).float() - 60) / 36
    else:
        ctrl = 1 if energy_control is None else energy_control
        norm_e = (((energy_predictor(sd, x, g) * 36 + 60) * ctrl) - 60) / 36
    x = x + F.conv1d(norm_e.unsqueeze(1), sd["energy_prenet.weight"], sd["energy_prenet.bias"], padding=1)
    taps["energy"] = (norm_e * 36 + 60).reshape(-1)
    taps["x_var"] = x[0].clone()
    if stop_after == "variance":
        return taps

    idx = expansion_indices(duration)                                         # :711 LengthRegulator
    taps["lr_index"] = idx
    Tf = idx.numel()
    x_frame = x[:, :, idx]
    taps["x_lr"] = x_frame[0].clone()
    taps["x_mask"] = torch.ones(1, Tf, dtype=torch.bool, device=x.device)                      # :713 (bool, Q5)
    if stop_after == "lr":
        return taps

    x_frame = encoder(sd, "frame_prior_net.fft_block", x_frame, cfg.n_layers, cfg)   # :715-716
    taps["x_frame"] = x_frame[0].clone()
    stats = F.conv1d(x_frame, sd["project.proj.weight"], sd["project.proj.bias"])     # :717
    m_p, logs_p = stats[:, :cfg.inter_channels], stats[:, cfg.inter_channels:]
    if noise is None:
        noise = torch.randn(cfg.inter_channels, Tf, device=x.device)
    z_p = m_p + noise.reshape(1, cfg.inter_channels, Tf) * torch.exp(logs_p) * noise_scale   # :718
    taps.update(m_p=m_p[0], logs_p=logs_p[0], z_p=z_p[0])
    if stop_after == "prior":
        return taps
    z = flow_reverse(sd, z_p, g, cfg)                                         # :719
    taps["z"] = z[0]
    if stop_after == "flow":
        return taps
    o = generator(sd, z[:, :, :max_len], g, cfg)                              # :720
    taps["o"] = o[0, 0]
    return taps


@torch.no_grad()
def infer_batch(sd, ids_list: Sequence[torch.Tensor], sids: Sequence[int], noise_scale=1.0,
                noises: Optional[Sequence[torch.Tensor]] = None, max_len=None,
                energy_controls=None, pitch_controls=None, duration_controls=None,
                cfg: ModelConfig = DEFAULT_CONFIG, stop_after=None) -> List[Dict[str, torch.Tensor]]:
    """Per-utterance batch-1 loop (the parity definition).  Control lists may hold None/scalars/tensors."""
    B = len(ids_list)

    def pick(c, b):
        if isinstance(c, (list, tuple)):
            return c[b]
        return c

    return [infer_one(sd, ids_list[b], sids[b], noise_scale, None if noises is None else noises[b], max_len,
                      pick(energy_controls, b), pick(pitch_controls, b), pick(duration_controls, b), cfg, stop_after)
            for b in range(B)]
