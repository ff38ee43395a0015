"""Reference state dict -> packed device tensors for libvispeech_b200 (load-time work, not the hot path).

Replaces what the reference gets implicitly from `utils.load_checkpoint` (utils.py:21-51) plus the
weight-norm forward pre-hooks (checkpoints hold un-folded weight_g/weight_v):
  * weight norm folded: w = g * v / ||v|| over every dim but 0 (ConvTranspose1d: per INPUT channel);
  * conv weights [Cout,Cin,k] -> [k][Cin][Cout] fp32 (coalesced over Cout);
  * conv_q|k|v fused into one [192][576] matrix;
  * speaker conditioning (emb_g -> 1x1 conv, models.py:674-675 + every `.cond`/`cond_layer`) evaluated once
    for all 200 speakers into lookup tables;
  * the Flip() layers of the flow (modules.py:270-277) folded into channel-permuted pre/post weights of the
    odd coupling layers (csrc/model.cu vs_flow_reverse);
  * ConvTranspose1d rewritten as polyphase taps;
  * decoder weights additionally in the f16 slab layout of csrc/umma_conv.cuh.
Names on the right-hand side are the keys `vs_model_set_tensor` expects (csrc/model.cu, csrc/decoder*.cu).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

UP_RATES = (8, 8, 4, 2)
UP_KERNELS = (16, 16, 4, 4)
RES_KERNELS = (3, 7, 11)
N_DIL = 3


def fold(sd, prefix: str) -> torch.Tensor:
    v, g = sd[prefix + ".weight_v"].float(), sd[prefix + ".weight_g"].float()
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)


def tcn(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, k] -> [k][Cin][Cout]."""
    return w.permute(2, 1, 0).contiguous()


def ups_phase_range(i: int, ph: int) -> Tuple[int, int]:
    s, K, p = UP_RATES[i], UP_KERNELS[i], (UP_KERNELS[i] - UP_RATES[i]) // 2
    dmax = (ph + p) // s
    return dmax - K // s + 1, dmax


def ups_union_taps(i: int) -> Tuple[int, int]:
    """(taps, pad_l) of the union over phases - must match csrc/decoder_umma.cu ups_taps()."""
    lo = min(ups_phase_range(i, ph)[0] for ph in range(UP_RATES[i]))
    hi = max(ups_phase_range(i, ph)[1] for ph in range(UP_RATES[i]))
    return hi - lo + 1, -lo


def up_columns(cout: int, s: int) -> torch.Tensor:
    """GEMM column of (phase ph, channel co) for a polyphase ConvTranspose1d in csrc/umma_conv.cu:
    col = (co // 8) * 8 * s + ph * 8 + co % 8, so that one epilogue thread writes the s phases of a channel
    group to s consecutive output rows (s * 16 contiguous bytes).  Returns index [s * cout] (phase-major input)."""
    ph = torch.arange(s).repeat_interleave(cout)
    co = torch.arange(cout).repeat(s)
    return (co // 8) * 8 * s + ph * 8 + co % 8


def pack_umma(w: torch.Tensor) -> torch.Tensor:
    """[taps][Cin][N] fp32 -> f16 [NB][taps][Cin/KC][KC/8][Nblk][8] (csrc/umma_conv.cuh)."""
    taps, cin, n = w.shape
    kc = min(cin, 64)
    nblk = min(n, 256)
    assert cin % kc == 0 and n % nblk == 0 and kc % 16 == 0
    x = w.reshape(taps, cin // kc, kc // 8, 8, n // nblk, nblk)
    x = x.permute(4, 0, 1, 2, 5, 3).contiguous()
    return x.to(torch.float16).reshape(-1)


def split16_slice(cin: int) -> int:
    """Channels per K-slice of the fp16 hi/lo conv (csrc/umma_split.cu): the whole K up to 192, else slices of 192."""
    return cin if cin <= 192 else 192


def pack_split16(w: torch.Tensor) -> torch.Tensor:
    """[taps][Cin][N] fp32 -> fp16 slabs of csrc/umma_split.cu: w = hi + lo (each fp16); per K-slice of cs channels and per
    KC-channel chunk the slabs alternate [w_hi, w_lo] (the w_hi slab serves the products a_hi w_hi and a_lo w_hi, the w_lo slab
    a_hi w_lo; all three accumulate into one accumulator), laid out [slice][NB][tap][cs / KC][hi|lo][KC/8][Nblk][8] with KC = 64
    (32 when cs is not a multiple of 64)."""
    taps, cin, n = w.shape
    cs = split16_slice(cin)
    assert cin % cs == 0 and cs % 32 == 0
    kc = 64 if cs % 64 == 0 else 32
    nblk = next(n // nb for nb in range(1, 17) if n % nb == 0 and n // nb <= 256 and (n // nb) % 32 == 0)
    hi = w.to(torch.float16)
    lo = (w - hi.float()).to(torch.float16)
    out = []
    for s in range(cin // cs):
        h, l = hi[:, s * cs:(s + 1) * cs], lo[:, s * cs:(s + 1) * cs]
        k2 = torch.stack([h.reshape(taps, cs // kc, kc, n), l.reshape(taps, cs // kc, kc, n)], dim=2)   # [taps][cs/kc][hi|lo][kc][N]
        x = k2.reshape(taps, 2 * cs // kc, kc // 8, 8, n // nblk, nblk)
        #              t     kc             p       e  nb         n
        out.append(x.permute(4, 0, 1, 2, 5, 3).contiguous().reshape(-1))   # nb t kc p n e
    return torch.cat(out)


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32 (10-bit mantissa), kept in fp32 words (what tcgen05 kind::tf32 reads)."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_tf32(w: torch.Tensor, split3: bool = False) -> torch.Tensor:
    """[taps][Cin][N] fp32 -> TF32 slabs (csrc/umma_tf32.cuh).
    plain : [NB][Cin/KA][taps][KA/32][8 planes][Nblk][4], KA = min(Cin, 96), values rounded to TF32;
    split3: [NB][Cin/KA][taps][KA/16][hi|lo][4 planes][Nblk][4], KA = min(Cin, 48), w = hi + lo (each TF32)."""
    taps, cin, n = w.shape
    ka_max, slab = (48, 16) if split3 else (96, 32)
    ka = min(cin, ka_max)
    assert cin % ka == 0 and ka % slab == 0
    nblk = next(n // nb for nb in range(1, 17) if n % nb == 0 and n // nb <= 256 and (n // nb) % 32 == 0)

    def slabs(x):
        x = x.reshape(taps, cin // ka, ka // slab, slab // 4, 4, n // nblk, nblk)
        #             t     ka         j           p          e  nb         n
        return x.permute(5, 1, 0, 2, 3, 6, 4).contiguous()       # nb ka t j p n e

    hi = round_tf32(w)
    if not split3:
        return slabs(hi).reshape(-1)
    lo = round_tf32(w - hi)
    return torch.stack([slabs(hi), slabs(lo)], dim=4).contiguous().reshape(-1)   # nb ka t j [hi|lo] p n e


def gate_columns(h: int = 192) -> torch.Tensor:
    """Column order of a WN in_layer for the fused gate epilogue (csrc/umma_tf32.cuh epi=1): blocks of
    [16 tanh channels | the same 16 sigmoid channels].  Returns perm with new[:, i] = old[:, perm[i]], len 2h."""
    c = torch.arange(h).reshape(h // 16, 16)
    return torch.cat([c, c + h], dim=1).reshape(-1)


COUPLING_SLAB = 24 * 96 * 8          # halves per weight slab of csrc/umma_coupling.cu: [K = 192 as 24 planes][N = 96][8]


def gate96_columns(h: int = 192) -> torch.Tensor:
    """Column order of a WN in_layer for csrc/umma_coupling.cu: four blocks of [48 tanh channels | the same 48 sigmoid channels]."""
    c = torch.arange(h).reshape(h // 48, 48)
    return torch.cat([c, c + h], dim=1).reshape(-1)


def pack_coupling(pre, in_w, rs_w, post) -> torch.Tensor:
    """All GEMM operands of one mean-only coupling layer (modules.py:324-343 with its WN, modules.py:148-176) as the fp16 slab
    stream csrc/umma_coupling.cu consumes, in consumption order; every slab is [K/8 planes][96 columns][8] padded to COUPLING_SLAB:
      pre (K = 96) columns 0-95, 96-191;  then per WN layer l: in_layer n-block nb = 0..3 (gate96 column order) x tap 0..4,
      res_skip -> h columns 0-95, 96-191 (not in the last layer), res_skip -> m: the skip half of res_skip folded with `post`
      (m = post(sum_l skip_l) = sum_l acts_l (W_skip_l W_post): the skip tensor itself is never formed).
    pre [96][192], in_w[l] [5][192][384], rs_w[l] [192][384 | 192], post [192][96] (all fp32, K-major rows = input channels)."""
    def slab(w):                                       # w [K][96] -> planes
        k = w.shape[0]
        x = w.reshape(k // 8, 8, 96).permute(0, 2, 1).contiguous().reshape(-1)
        return torch.cat([x, torch.zeros(COUPLING_SLAB - x.numel())])
    out = [slab(pre[:, :96]), slab(pre[:, 96:])]
    gp = gate96_columns(192)
    L = len(in_w)
    for l in range(L):
        wg = in_w[l][:, :, gp]                         # [5][192][384]
        for nb in range(4):
            for t in range(5):
                out.append(slab(wg[t][:, 96 * nb:96 * nb + 96]))
        last = l == L - 1
        if not last:
            out += [slab(rs_w[l][:, :96]), slab(rs_w[l][:, 96:192])]
        skip_w = rs_w[l] if last else rs_w[l][:, 192:]
        out.append(slab((skip_w.double() @ post.double()).float()))
    return torch.cat(out).to(torch.float16).contiguous()


def pack_coupling_bias(pre_b, in_b, cond_tab, rs_b, post, post_b) -> torch.Tensor:
    """fp32 side table of pack_coupling: [hb: L x 192 cumulative biases of h before layer l][mb: 96][per speaker: L x 384 =
    in_layer bias + cond_layer(emb_g) in gate96 order].  cond_tab [n_spk][L * 384]."""
    L = len(in_b)
    gp = gate96_columns(192)
    hb, acc = [], pre_b.clone()
    skip_b = torch.zeros(192)
    for l in range(L):
        hb.append(acc.clone())
        if l < L - 1:
            acc = acc + rs_b[l][:192]
            skip_b = skip_b + rs_b[l][192:]
        else:
            skip_b = skip_b + rs_b[l]
    mb = post_b + (skip_b.double() @ post.double()).float()
    n_spk = cond_tab.shape[0]
    cg = (cond_tab.reshape(n_spk, L, 384) + torch.stack(in_b)[None])[:, :, gp]
    return torch.cat([torch.cat(hb), mb, cg.reshape(-1)]).float().contiguous()


def pack_state_dict(sd: Dict[str, torch.Tensor], n_layers=4, pitch_layers=6, n_flows=4, flow_layers=4) -> Dict[str, torch.Tensor]:
    """Returns {packed name: CPU tensor (fp32 or f16, contiguous)}."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    out: Dict[str, torch.Tensor] = {}
    emb_g = sd["emb_g.weight"]                                     # [200, 256]

    def cond_table(w, b):                                          # Conv1d(gin, C, 1) applied to every speaker
        return (emb_g @ w[:, :, 0].t() + b).contiguous()

    out["emb"] = sd["enc_p.symbol_emb.weight"].contiguous()

    def encoder(prefix, layers):
        for i in range(layers):
            a = "%s.attn_layers.%d" % (prefix, i)
            q = "%s.%d." % (prefix, i)
            out[q + "wqkv"] = torch.cat([tcn(sd["%s.conv_%s.weight" % (a, n)])[0] for n in "qkv"], dim=1).contiguous()
            out[q + "bqkv"] = torch.cat([sd["%s.conv_%s.bias" % (a, n)] for n in "qkv"]).contiguous()
            out[q + "wo"] = tcn(sd[a + ".conv_o.weight"])[0].contiguous()
            out[q + "bo"] = sd[a + ".conv_o.bias"]
            out[q + "ek"] = sd[a + ".emb_rel_k"][0].contiguous()
            out[q + "ev"] = sd[a + ".emb_rel_v"][0].contiguous()
            out[q + "g1"] = sd["%s.norm_layers_1.%d.gamma" % (prefix, i)]
            out[q + "b1"] = sd["%s.norm_layers_1.%d.beta" % (prefix, i)]
            out[q + "g2"] = sd["%s.norm_layers_2.%d.gamma" % (prefix, i)]
            out[q + "b2"] = sd["%s.norm_layers_2.%d.beta" % (prefix, i)]
            out[q + "w1"] = tcn(sd["%s.ffn_layers.%d.conv_1.weight" % (prefix, i)])
            out[q + "bf1"] = sd["%s.ffn_layers.%d.conv_1.bias" % (prefix, i)]
            out[q + "w2"] = tcn(sd["%s.ffn_layers.%d.conv_2.weight" % (prefix, i)])
            out[q + "bf2"] = sd["%s.ffn_layers.%d.conv_2.bias" % (prefix, i)]

    encoder("enc_p.encoder", n_layers)
    encoder("pitch_predictor.pitch_net", pitch_layers)
    encoder("frame_prior_net.fft_block", n_layers)

    p = "duration_predictor"
    out["dp.cond_tab"] = cond_table(sd[p + ".cond.weight"], sd[p + ".cond.bias"])
    out["dp.w1"], out["dp.b1"] = tcn(sd[p + ".conv_1.weight"]), sd[p + ".conv_1.bias"]
    out["dp.g1"], out["dp.be1"] = sd[p + ".norm_1.gamma"], sd[p + ".norm_1.beta"]
    out["dp.w2"], out["dp.b2"] = tcn(sd[p + ".conv_2.weight"]), sd[p + ".conv_2.bias"]
    out["dp.g2"], out["dp.be2"] = sd[p + ".norm_2.gamma"], sd[p + ".norm_2.beta"]
    out["dp.wp"], out["dp.bp"] = sd[p + ".proj.weight"].reshape(-1).contiguous(), sd[p + ".proj.bias"]

    p = "pitch_predictor"
    out["pp.cond_tab"] = cond_table(sd[p + ".cond.weight"], sd[p + ".cond.bias"])
    out["pp.wf0"], out["pp.bf0"] = sd[p + ".proj_f0.weight"].reshape(-1).contiguous(), sd[p + ".proj_f0.bias"]

    p = "energy_predictor"
    c = p + ".predictor.conv_layer"
    out["ep.cond_tab"] = cond_table(sd[p + ".cond.weight"], sd[p + ".cond.bias"])
    out["ep.w1"], out["ep.b1"] = tcn(sd[c + ".conv_1.conv.weight"]), sd[c + ".conv_1.conv.bias"]
    out["ep.g1"], out["ep.be1"] = sd[c + ".layer_norm_1.weight"], sd[c + ".layer_norm_1.bias"]
    out["ep.w2"], out["ep.b2"] = tcn(sd[c + ".conv_2.conv.weight"]), sd[c + ".conv_2.conv.bias"]
    out["ep.g2"], out["ep.be2"] = sd[c + ".layer_norm_2.weight"], sd[c + ".layer_norm_2.bias"]
    out["ep.wl"] = sd[p + ".predictor.linear_layer.weight"].reshape(-1).contiguous()
    out["ep.bl"] = sd[p + ".predictor.linear_layer.bias"]

    for n in ("pitch_prenet", "energy_prenet"):
        out[n + ".w"] = sd[n + ".weight"].reshape(-1, 3).contiguous()         # [192][3]
        out[n + ".b"] = sd[n + ".bias"]
    out["proj.w"], out["proj.b"] = tcn(sd["project.proj.weight"])[0].contiguous(), sd["project.proj.bias"]

    for f in range(n_flows):
        src = "flow.flows.%d" % (2 * f)
        dst = "flow.%d." % f
        flipped = (f % 2) == 1       # reversed(flows) starts with a Flip: layers 3 and 1 see a flipped tensor
        pre = tcn(sd[src + ".pre.weight"])[0]            # [96 ci][192 co]
        post = tcn(sd[src + ".post.weight"])[0]          # [192 ci][96 co]
        post_b = sd[src + ".post.bias"]
        if flipped:
            pre = torch.flip(pre, [0])                   # x0[c] = phys[191-c]
            post = torch.flip(post, [1])                 # m_phys[p] = m[95-p]
            post_b = torch.flip(post_b, [0])
        out[dst + "pre.w"], out[dst + "pre.b"] = pre.contiguous(), sd[src + ".pre.bias"]
        out[dst + "post.w"], out[dst + "post.b"] = post.contiguous(), post_b.contiguous()
        out[dst + "cond_tab"] = cond_table(fold(sd, src + ".enc.cond_layer"), sd[src + ".enc.cond_layer.bias"])
        for l in range(flow_layers):
            out["%s%d.in.w" % (dst, l)] = tcn(fold(sd, "%s.enc.in_layers.%d" % (src, l)))
            out["%s%d.in.b" % (dst, l)] = sd["%s.enc.in_layers.%d.bias" % (src, l)]
            out["%s%d.rs.w" % (dst, l)] = tcn(fold(sd, "%s.enc.res_skip_layers.%d" % (src, l)))[0].contiguous()
            out["%s%d.rs.b" % (dst, l)] = sd["%s.enc.res_skip_layers.%d.bias" % (src, l)]
        # the whole coupling layer as one kernel (csrc/umma_coupling.cu): fp16 slab stream + fp32 bias / per-speaker table
        out["c16." + dst + "w"] = pack_coupling(out[dst + "pre.w"], [out["%s%d.in.w" % (dst, l)] for l in range(flow_layers)],
                                               [out["%s%d.rs.w" % (dst, l)] for l in range(flow_layers)], out[dst + "post.w"])
        out["c16." + dst + "b"] = pack_coupling_bias(out[dst + "pre.b"], [out["%s%d.in.b" % (dst, l)] for l in range(flow_layers)],
                                                    out[dst + "cond_tab"], [out["%s%d.rs.b" % (dst, l)] for l in range(flow_layers)],
                                                    out[dst + "post.w"], out[dst + "post.b"])

    # ---- posterior encoder (only voice_conversion uses it; skipped when the checkpoint has no enc_q.*)
    wn_prefixes = [("flow.%d." % f, flow_layers) for f in range(n_flows)]
    if "enc_q.pre.weight" in sd:
        pre = tcn(sd["enc_q.pre.weight"])[0]                      # [1025][192]
        c_pad = -(-pre.shape[0] // 96) * 96                       # 1056: multiple of 96 (TF32 stage) and 48 (3xTF32)
        out["enc_q.pre.w"] = torch.cat([pre, torch.zeros(c_pad - pre.shape[0], pre.shape[1])], 0).contiguous()
        out["enc_q.pre.b"] = sd["enc_q.pre.bias"]
        out["enc_q.proj.w"], out["enc_q.proj.b"] = tcn(sd["enc_q.proj.weight"])[0].contiguous(), sd["enc_q.proj.bias"]
        n_q = 16
        out["enc_q.cond_tab"] = cond_table(fold(sd, "enc_q.enc.cond_layer"), sd["enc_q.enc.cond_layer.bias"])
        for l in range(n_q):
            out["enc_q.%d.in.w" % l] = tcn(fold(sd, "enc_q.enc.in_layers.%d" % l))
            out["enc_q.%d.in.b" % l] = sd["enc_q.enc.in_layers.%d.bias" % l]
            out["enc_q.%d.rs.w" % l] = tcn(fold(sd, "enc_q.enc.res_skip_layers.%d" % l))[0].contiguous()
            out["enc_q.%d.rs.b" % l] = sd["enc_q.enc.res_skip_layers.%d.bias" % l]
        wn_prefixes.append(("enc_q.", n_q))

    # ---- decoder
    out["dec.pre.w"], out["dec.pre.b"] = tcn(sd["dec.conv_pre.weight"]), sd["dec.conv_pre.bias"]
    out["dec16.pre.w"] = pack_umma(out["dec.pre.w"])
    out["dec.cond_tab"] = cond_table(sd["dec.cond.weight"], sd["dec.cond.bias"])
    out["dec.post.w"] = tcn(sd["dec.conv_post.weight"]).reshape(-1).contiguous()        # [7][32]
    for i, (s, K) in enumerate(zip(UP_RATES, UP_KERNELS)):
        wt = fold(sd, "dec.ups.%d" % i)                   # [Cin, Cout, K]
        cin, cout, _ = wt.shape
        pad = (K - s) // 2
        per = torch.zeros(s, K // s, cin, cout)
        taps, pad_l = ups_union_taps(i)
        uni = torch.zeros(taps, cin, s * cout)
        cols = up_columns(cout, s)
        for ph in range(s):
            lo, hi = ups_phase_range(i, ph)
            for t, d in enumerate(range(lo, hi + 1)):
                k = ph + pad - s * d
                assert 0 <= k < K
                per[ph, t] = wt[:, :, k]
                uni[d + pad_l, :, cols[ph * cout:(ph + 1) * cout]] = wt[:, :, k]
        out["dec.ups.%d.w" % i] = per.contiguous()
        out["dec.ups.%d.b" % i] = sd["dec.ups.%d.bias" % i]
        out["dec16.ups.%d.w" % i] = pack_umma(uni)
        for j in range(len(RES_KERNELS)):
            n = i * len(RES_KERNELS) + j
            for m in range(N_DIL):
                for cname, key in (("c1", "convs1"), ("c2", "convs2")):
                    w = tcn(fold(sd, "dec.resblocks.%d.%s.%d" % (n, key, m)))
                    out["dec.rb.%d.%s.%d.w" % (n, cname, m)] = w
                    out["dec.rb.%d.%s.%d.b" % (n, cname, m)] = sd["dec.resblocks.%d.%s.%d.bias" % (n, key, m)]
                    out["dec16.rb.%d.%s.%d.w" % (n, cname, m)] = pack_umma(w)
    # WN in_layers for the fused tanh*sigmoid epilogue: gate-interleaved columns (weights, bias, per-speaker cond)
    gperm = gate_columns(192)
    for pre_, nl in wn_prefixes:
        tab = out[pre_ + "cond_tab"].reshape(emb_g.shape[0], nl, 384)
        out[pre_ + "cond_tab_gate"] = tab[:, :, gperm].reshape(emb_g.shape[0], -1).contiguous()
        for l in range(nl):
            out["%s%d.in_gate.w" % (pre_, l)] = out["%s%d.in.w" % (pre_, l)][:, :, gperm].contiguous()
            out["%s%d.in_gate.b" % (pre_, l)] = out["%s%d.in.b" % (pre_, l)][gperm].contiguous()
    # tensor-core copies of the GEMM-shaped conv weights: "tf32." = plain TF32 (frame level, >= 4096 rows),
    # "x3." = error-compensated 3xTF32 (fp32-level accuracy: phoneme level and small frame-level calls)
    def is_gemm(name):
        if name.startswith(("tf32.", "x3.", "s16.", "dec", "emb")) or not name.endswith((".wqkv", ".wo", ".w1", ".w2", ".w")):
            return False
        return name.startswith(("enc_p.", "pitch_predictor.", "frame_prior_net.", "dp.w", "ep.w", "proj.", "flow.", "enc_q."))
    for name in [k for k in out if is_gemm(k)]:
        w = out[name] if out[name].dim() == 3 else out[name][None]
        if w.shape[1] % 32 == 0 and (w.shape[1] < 96 or w.shape[1] % 96 == 0) and w.shape[2] % 32 == 0:
            out["tf32." + name] = pack_tf32(w)
        if w.shape[1] % 16 == 0 and (w.shape[1] < 48 or w.shape[1] % 48 == 0) and w.shape[2] % 32 == 0:
            out["x3." + name] = pack_tf32(w, split3=True)
        cs = split16_slice(w.shape[1])
        if w.shape[1] % cs == 0 and cs % 32 == 0 and w.shape[2] % 32 == 0:
            out["s16." + name] = pack_split16(w)
    for pre_, nl in wn_prefixes:
        for l in range(nl):
            del out["%s%d.in_gate.w" % (pre_, l)]              # only its packed copies are used
    return {k: v.contiguous() for k, v in out.items()}
