"""Helpers shared by the GPU parity tests (op-level calls through the C ABI)."""
import ctypes

import numpy as np
import torch

from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
from vispeech_b200.layout import make_rows
from vispeech_b200.packing import pack_tf32, pack_umma, round_tf32

DEV = "cuda:0"


def stream():
    return torch.cuda.current_stream().cuda_stream


def to_planar(x: torch.Tensor) -> torch.Tensor:
    """[R][C] fp32 -> planar f16 [C/8][R][8] (on the same device)."""
    R, C = x.shape
    return x.to(torch.float16).reshape(R, C // 8, 8).permute(1, 0, 2).contiguous()


def from_planar(p: torch.Tensor) -> torch.Tensor:
    """planar f16 [C/8][R][8] -> [R][C] fp32."""
    P, R, _ = p.shape
    return p.permute(1, 0, 2).reshape(R, P * 8).float()


def f16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).float()


def umma_conv(x, w_tcn, bias=None, res=None, dil=1, pad_l=0, up=1, act_slope=1.0, act_scale=1.0, row_utt=None,
              row_div=1, want_raw=True, want_act=True, res2=None, res_inv_slope=0.0):
    """x [R][Cin] fp32 (device), w_tcn [taps][Cin][N] fp32.  Returns (raw, act) as [R*up][N/up] fp32."""
    lib = _lib.load()
    R, cin = x.shape
    taps, _, n = w_tcn.shape
    cout = n // up
    xin = to_planar(x)
    wp = pack_umma(w_tcn.cpu()).to(x.device)
    raw = torch.full((cout // 8, R * up, 8), float("nan"), dtype=torch.float16, device=x.device) if want_raw else None
    act = torch.full((cout // 8, R * up, 8), float("nan"), dtype=torch.float16, device=x.device) if want_act else None
    resp = to_planar(res) if res is not None else None
    res2p = to_planar(res2) if res2 is not None else None
    check(lib.vs_op_conv1d_umma2(ptr(xin), ptr(wp), ptr(bias), ptr(resp), ptr(res2p), float(res_inv_slope), ptr(raw), ptr(act), R, cin, n,
                                 taps, dil, pad_l, up, float(act_slope), float(act_scale), ptr(row_utt), row_div, stream()),
          "vs_op_conv1d_umma2")
    torch.cuda.synchronize()
    return (from_planar(raw) if want_raw else None), (from_planar(act) if want_act else None)


def respair(x, w1, w2, b1, b2, dil, res2=None, act_slope=1.0, act_scale=1.0, row_utt=None, row_div=1):
    """Fused ResBlock1 iteration on planar f16 rows.  x = the ACTIVATED input a = lrelu(x_raw) [R][C] fp32 device,
    w [k][C][C].  Returns (raw, act) [R][C]."""
    lib = _lib.load()
    R, C = x.shape
    k = w1.shape[0]
    xin = to_planar(x)
    w1p, w2p = pack_umma(w1.cpu()).to(x.device), pack_umma(w2.cpu()).to(x.device)
    r2 = to_planar(res2) if res2 is not None else None
    raw = torch.full((C // 8, R, 8), float("nan"), dtype=torch.float16, device=x.device)
    act = torch.full((C // 8, R, 8), float("nan"), dtype=torch.float16, device=x.device)
    check(lib.vs_op_respair(ptr(xin), ptr(w1p), ptr(w2p), ptr(b1), ptr(b2), ptr(r2), ptr(raw), ptr(act), R, C, k, dil,
                            float(act_slope), float(act_scale), ptr(row_utt), row_div, stream()), "vs_op_respair")
    torch.cuda.synchronize()
    return from_planar(raw), from_planar(act)


def mrf32(x0, W, B, post_w, row_utt, row_div=1):
    """Whole last MRF stage + conv_post + tanh (csrc/umma_mrf.cu).  x0 [R][32] fp32 (device), W[j][m][c] = [k][32][32] fp32,
    B[j][m][c] = [32] fp32 (CPU), post_w [7][32] fp32 (CPU).  x0 is handed over as fp16 hi + lo planes.  Returns wave [R]."""
    import ctypes
    lib = _lib.load()
    R = x0.shape[0]
    hi = x0.to(torch.float16)
    lo = (x0 - hi.float()).to(torch.float16)
    hi_p, lo_p = to_planar(hi.float()), to_planar(lo.float())
    wp = [pack_umma(W[j][m][c].cpu()).to(x0.device) for j in range(3) for m in range(3) for c in range(2)]
    bh = [B[j][m][c].float().contiguous().cpu() for j in range(3) for m in range(3) for c in range(2)]
    w_arr = (ctypes.c_void_p * 18)(*[t.data_ptr() for t in wp])
    b_arr = (ctypes.c_void_p * 18)(*[t.data_ptr() for t in bh])
    pw = post_w.float().contiguous().cpu()
    wave = torch.full((R,), float("nan"), dtype=torch.float32, device=x0.device)
    check(lib.vs_op_mrf32(ptr(hi_p), ptr(lo_p), ctypes.cast(w_arr, ctypes.c_void_p), ctypes.cast(b_arr, ctypes.c_void_p), ptr(pw),
                          ptr(row_utt), row_div, R, ptr(wave), stream()), "vs_op_mrf32")
    torch.cuda.synchronize()
    return wave


def resblock64(a, W, B, row_utt, row_div=1):
    """Whole k = 3 ResBlock1 of the C = 64 stage (csrc/umma_resblock.cu).  a = lrelu(x0) [R][64] fp32 (device), W[m][c] = [3][64][64],
    B[m][c] = [64] (CPU).  Returns the ResBlock's output [R][64] fp32."""
    lib = _lib.load()
    R = a.shape[0]
    ap = to_planar(a)
    wp = [pack_umma(W[m][c].cpu()).to(a.device) for m in range(3) for c in range(2)]
    bh = [B[m][c].float().contiguous().cpu() for m in range(3) for c in range(2)]
    w_arr = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in wp])
    b_arr = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in bh])
    out = torch.full((8, R, 8), float("nan"), dtype=torch.float16, device=a.device)
    check(lib.vs_op_resblock64(ptr(ap), ctypes.cast(w_arr, ctypes.c_void_p), ctypes.cast(b_arr, ctypes.c_void_p), ptr(row_utt), row_div,
                               R, ptr(out), stream()), "vs_op_resblock64")
    torch.cuda.synchronize()
    return from_planar(out)


def conv_f32(x, w_tcn, bias=None, dil=1, pad_l=0, in_slope=1.0, act=0, row_utt=None, row_div=1):
    lib = _lib.load()
    R, cin = x.shape
    k, _, cout = w_tcn.shape
    out = torch.full((R, cout), float("nan"), dtype=torch.float32, device=x.device)
    check(lib.vs_op_conv1d_f32(ptr(x), cin, ptr(w_tcn), ptr(bias), ptr(out), cout, R, cin, cout, k, dil, pad_l,
                               float(in_slope), act, ptr(row_utt), row_div, stream()), "vs_op_conv1d_f32")
    torch.cuda.synchronize()
    return out


def conv_tf32(x, w_tcn, bias=None, dil=1, pad_l=0, act=0, row_utt=None, split3=False):
    lib = _lib.load()
    R, cin = x.shape
    k, _, cout = w_tcn.shape
    wp = pack_tf32(w_tcn.cpu(), split3=split3).to(x.device)
    out = torch.full((R, cout), float("nan"), dtype=torch.float32, device=x.device)
    check(lib.vs_op_conv1d_tf32(ptr(x), cin, ptr(wp), ptr(bias), ptr(out), cout, R, cin, cout, k, dil, pad_l, act,
                                1 if split3 else 0, ptr(row_utt), stream()), "vs_op_conv1d_tf32")
    torch.cuda.synchronize()
    return out


def conv_split16(x, w_tcn, bias=None, dil=1, pad_l=0, act=0, row_utt=None):
    """fp16 hi/lo three-term conv on tcgen05 kind::f16 (csrc/umma_split.cu) through its op-level hook."""
    from vispeech_b200.packing import pack_split16
    lib = _lib.load()
    R, cin = x.shape
    k, _, cout = w_tcn.shape
    wp = pack_split16(w_tcn.cpu()).to(x.device)
    out = torch.full((R, cout), float("nan"), dtype=torch.float32, device=x.device)
    ws = torch.empty(R * (4 * cin + 4 * cout * max(1, cin // 192)) + (1 << 16), dtype=torch.uint8, device=x.device)
    check(lib.vs_op_conv1d_split(ptr(x), cin, ptr(wp), ptr(bias), ptr(out), cout, R, cin, cout, k, dil, pad_l, act,
                                 ptr(row_utt), ptr(ws), ws.numel(), stream()), "vs_op_conv1d_split")
    torch.cuda.synchronize()
    return out


def ref_conv_rows(x, w_tcn, bias=None, dil=1, pad_l=0):
    """CPU reference of the row conv: out[r] = sum_t in[r + (t-pad_l)*dil] @ W[t] (+bias); zero outside [0,R)."""
    x = x.double().cpu()
    w = w_tcn.double().cpu()
    R = x.shape[0]
    out = torch.zeros(R, w.shape[2], dtype=torch.float64)
    for t in range(w.shape[0]):
        sh = (t - pad_l) * dil
        lo, hi = max(0, -sh), min(R, R - sh)
        if hi > lo:
            out[lo:hi] += x[lo + sh:hi + sh] @ w[t]
    if bias is not None:
        out += bias.double().cpu()
    return out


def snr_db(ref, x):
    ref, x = ref.double().cpu().reshape(-1), x.double().cpu().reshape(-1)
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum().clamp_min(1e-300)))
