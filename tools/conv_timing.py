"""Where do the tcgen05 conv's warps wait?  Runs decoder conv shapes with the kernel's wait-clock counters switched on
(vs_set_option("umma_timing_buffer"); needs a library built with VS_UMMA_TIMING=1 python vispeech_b200/build.py --force) and prints, per role, total clocks and the share spent in each mbarrier wait."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
FRAMES = 27840
buf = torch.zeros(296 * 16, dtype=torch.int64, device=dev)


def run(name, R, cin, n, taps, dil):
    x = (torch.randn(cin // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w = (torch.randn(taps * cin * n, device=dev) / (cin * taps) ** 0.5).to(torch.float16)
    b = torch.randn(n, device=dev)
    o1 = torch.empty(n // 8, R, 8, device=dev, dtype=torch.float16)

    def call():
        check(lib.vs_op_conv1d_umma(ptr(x), ptr(w), ptr(b), None, None, ptr(o1), R, cin, n, taps, dil, (taps - 1) // 2, 1, 0.1, 1.0, None, 1, st))
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
    t = buf[: 296 * 12].view(296, 3, 4).double()
    used = t[:, 1, 0] > 0
    t = t[used].mean(0)
    n_mma = (R / 128) * taps * (cin / 16) * (n / min(n, 256))
    print("%-14s %.3f ms  ctas=%d  clk/MMA/SM=%.0f" % (name, ms, int(used.sum()), ms * 1e-3 * 1.9e9 / (n_mma / 148)))
    for r, (role, names) in enumerate((("producer", ("a_empty", "b_empty", "-")), ("mma", ("a_full", "acc_empty", "b_full")),
                                        ("epilogue", ("acc_full", "-", "-")))):
        tot = t[r, 0].item()
        print("    %-9s total %8.0f clk  " % (role, tot) + "  ".join("%s %4.1f%%" % (nm, 100 * t[r, 1 + i].item() / tot) for i, nm in enumerate(names) if nm != "-"))


def run_pair(name, R, C, k, d):
    x = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w1 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    w2 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    b1, b2 = torch.randn(C, device=dev), torch.randn(C, device=dev)
    o = torch.empty_like(x)

    def call():
        check(lib.vs_op_respair(ptr(x), ptr(w1), ptr(w2), ptr(b1), ptr(b2), None, None, ptr(o), R, C, k, d, 0.1, 1.0, None, 1, st))
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
    t = buf[: 296 * 16].view(296, 4, 4).double()
    used = t[:, 1, 0] > 0
    t = t[used].mean(0)
    print("%-14s %.3f ms  ctas=%d" % (name, ms, int(used.sum())))
    for r, (role, names) in enumerate((("producer", ("xa_empty", "-", "-")), ("mma", ("xa_full+acc1_empty", "a2_full", "acc2_empty")),
                                        ("epilogue1", ("acc1_full", "a2_empty", "-")), ("epilogue2", ("acc2_full", "-", "-")))):
        tot = t[r, 0].item()
        print("    %-9s total %8.0f clk  " % (role, tot) + "  ".join("%s %4.1f%%" % (nm, 100 * t[r, 1 + i].item() / tot) for i, nm in enumerate(names) if nm != "-"))


if os.environ.get("PAIRS", "1") == "1":
    run_pair("s3 pair k3", FRAMES * 512, 32, 3, 1)
    run_pair("s3 pair k11", FRAMES * 512, 32, 11, 1)
    run_pair("s2 pair k3", FRAMES * 256, 64, 3, 1)
    run_pair("s2 pair k7", FRAMES * 256, 64, 7, 1)
    sys.exit(0)
run("s0 c1 k3 d1", FRAMES * 8, 256, 256, 3, 1)
run("s0 c1 k11 d1", FRAMES * 8, 256, 256, 11, 1)
run("s0 c1 k11 d5", FRAMES * 8, 256, 256, 11, 5)
run("s1 c1 k7 d1", FRAMES * 64, 128, 128, 7, 1)
run("s1 c1 k11 d5", FRAMES * 64, 128, 128, 11, 5)
run("s1 c1 k3 d1", FRAMES * 64, 128, 128, 3, 1)
run("s1 c1 k11 d1", FRAMES * 64, 128, 128, 11, 1)
run("s2 c1 k7 d1", FRAMES * 256, 64, 64, 7, 1)
run("s2 c1 k11 d1", FRAMES * 256, 64, 64, 11, 1)
run("s2 c1 k11 d5", FRAMES * 256, 64, 64, 11, 5)
run("s3 c1 k7 d1", FRAMES * 512, 32, 32, 7, 1)
run("s3 c1 k11 d1", FRAMES * 512, 32, 32, 11, 1)
