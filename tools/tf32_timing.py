"""Where do the TF32 conv's warps wait?  Needs VS_UMMA_TIMING=1 python vispeech_b200/build.py --force."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
from vispeech_b200.packing import pack_tf32
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)


def run(name, R, cin, n, taps, split3):
    x = torch.randn(R, cin, device=dev)
    w = torch.randn(taps, cin, n) / (cin * taps) ** 0.5
    wp = pack_tf32(w, split3=bool(split3)).to(dev)
    b = torch.randn(n, device=dev)
    o = torch.empty(R, n, device=dev)

    def call():
        check(lib.vs_op_conv1d_tf32(ptr(x), cin, ptr(wp), ptr(b), ptr(o), n, R, cin, n, taps, 1, (taps - 1) // 2, 0, split3, None, st))
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    call(); call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf.zero_()
    check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
    call()
    torch.cuda.synchronize()
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    t = buf.view(148, 4, 4).double()
    used = t[:, 2, 0] > 0
    t = t[used].mean(0)
    print("%-26s %.1f us  ctas=%d" % (name, ms * 1e3, int(used.sum())))
    for r, (role, names) in enumerate((("loader", ("a_empty", "-", "-")), ("w producer", ("b_empty", "-", "-")),
                                        ("mma", ("a_full", "acc_empty", "b_full")), ("epilogue", ("acc_full", "-", "-")))):
        tot = max(t[r, 0].item(), 1)
        print("    %-10s total %8.0f clk  " % (role, tot) + "  ".join("%s %4.1f%%" % (nm, 100 * t[r, 1 + i].item() / tot) for i, nm in enumerate(names) if nm != "-"))


run("phoneme qkv x3", 2810, 192, 576, 1, 1)
run("phoneme ffn1 x3", 2810, 192, 768, 3, 1)
run("phoneme ffn2 x3", 2810, 768, 192, 3, 1)
run("frame ffn2 x3", 27840, 768, 192, 3, 1)
run("frame ffn1 x3", 27840, 192, 768, 3, 1)
run("frame wn in (tf32)", 27840, 192, 384, 5, 0)
run("frame wn rs (tf32)", 27840, 192, 384, 1, 0)
run("frame qkv x3", 27840, 192, 576, 1, 1)
run("frame o x3", 27840, 192, 192, 1, 1)
