"""A/B of the relative-attention kernels (0 CUDA-core fp32, 2 3xTF32 mma.sync, 3 plain-TF32 mma.sync, 4 tcgen05 fp16 hi/lo) at a given shape:
   python tools/attention_timing.py [n_utt=64] [T=431]
Prints ms per launch (CUDA events, 20 launches) for both and the max abs difference between their outputs."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check
from vispeech_b200.layout import make_rows
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 431
rows = make_rows([T] * B, [0] * B, 4, dev)
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(rows.n_rows, 576, device=dev, generator=g)
qkv *= (rows.row_utt >= 0).float()[:, None]
ek = torch.randn(9, 96, device=dev, generator=g) * 0.1
ev = torch.randn(9, 96, device=dev, generator=g) * 0.1
outs = {}
ws = torch.empty(rows.n_rows * 8192 + B * 200000 + 65536, dtype=torch.uint8, device=dev)
for mode in (0, 2, 3, 4):
    check(lib.vs_set_option(b"attention_mma", mode))
    out = torch.empty(rows.n_rows, 192, device=dev)
    for _ in range(3):
        check(lib.vs_op_rel_attention(ctypes.byref(rows.struct), qkv.data_ptr(), ek.data_ptr(), ev.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        check(lib.vs_op_rel_attention(ctypes.byref(rows.struct), qkv.data_ptr(), ek.data_ptr(), ev.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st))
    e1.record(); torch.cuda.synchronize()
    flop = 2 * B * T * T * 96 * 4
    ms = e0.elapsed_time(e1) / 20
    print("attention_mma=%d  B=%d T=%d  %.3f ms  %.1f TFLOP/s (algorithmic)" % (mode, B, T, ms, flop / ms / 1e9))
    outs[mode] = out
for m in (2, 3, 4):
    print("mode %d: max |mma - fp32| = %.3e   (max |out| = %.3f)" % (m, float((outs[0] - outs[m]).abs().max()), float(outs[0].abs().max())))

if os.environ.get("VS_LIB_DIR"):      # diagnostics build (VS_UMMA_TIMING=1 VS_LIB_DIR=... python vispeech_b200/build.py): clock stamps of the tcgen05 kernel
    n_cta = ((T + 127) // 128) * 2 * B
    buf = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
    check(lib.vs_set_option(b"attention_mma", 4))
    check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
    check(lib.vs_op_rel_attention(ctypes.byref(rows.struct), qkv.data_ptr(), ek.data_ptr(), ev.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    t = buf.view(n_cta, 16).double()
    t = t[t[:, 12] > 0].mean(0)
    names = ["setup done", "Q+K0 landed", "PV0 issued", "PV1", "PV2", "PV3", "PV4", "PV5", "P(g0,t0) stored", "P(g0,t1)", "P(g0,t2)", "last O_DONE", "output stored"]
    print("attention_umma_kernel, mean clocks since CTA start: " + "  ".join("%s %.0f" % (n, v) for n, v in zip(names, t.tolist())))
