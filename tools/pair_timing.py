"""The CTA-pair conv (csrc/umma_pair.cu) at the C2 size of the C = 128 stage: time per launch against the single-CTA kernel and,
with a diagnostics build (VS_UMMA_TIMING=1 VS_LIB_DIR=vispeech_b200/lib_timing python vispeech_b200/build.py --force; run this
tool with the same VS_LIB_DIR), where its roles wait.   python tools/pair_timing.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
FRAMES = 28800
buf = torch.zeros(296 * 16, dtype=torch.int64, device=dev)
if os.environ.get("VS_TAP_PAIRS"):
    check(lib.vs_set_option(b"tap_pairs", int(os.environ["VS_TAP_PAIRS"])))


def run(name, R, C, taps, dil, res):
    x = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w = (torch.randn(taps * C * C, device=dev) / (C * taps) ** 0.5).to(torch.float16)
    b = torch.randn(C, device=dev)
    r1 = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16) if res else None
    o1 = torch.empty(C // 8, R, 8, device=dev, dtype=torch.float16)

    def call():
        check(lib.vs_op_conv1d_umma2(ptr(x), ptr(w), ptr(b), ptr(r1), None, 10.0 if res else 0.0, None, ptr(o1), R, C, C, taps, dil,
                                     (taps - 1) // 2, 1, 0.1, 1.0, None, 1, st))

    def timed():
        call(); call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            call()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    check(lib.vs_set_option(b"pair_conv", 0)); ms0 = timed()
    check(lib.vs_set_option(b"pair_conv", 2)); ms1 = timed()
    n_mma = (R / 128) * taps * (C / 16)
    print("%-16s single-CTA %.3f ms   pair %.3f ms   (tensor floor %.3f ms at 1.75 GHz)" % (name, ms0, ms1, n_mma * (C / 2) / 148 / 1.75e6))
    buf.zero_()
    check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
    call()
    torch.cuda.synchronize()
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    t = buf.view(296, 4, 4).double()
    used = t[:, 0, 0] > 0
    if not used.any():
        return
    lead = t[0::2][used[0::2]].mean(0)
    for r, (role, names) in enumerate((("producer", ("a_empty",)), ("mma", ("a_full", "acc_empty", "commits")), ("relay", ("a_land",)),
                                        ("epilogue", ("acc_full", "tcgen05.ld")))):
        tot = lead[r, 0].item()
        if tot > 0:
            print("    %-9s total %8.0f clk  " % (role, tot) + "  ".join("%s %4.1f%%" % (nm, 100 * lead[r, 1 + i].item() / tot) for i, nm in enumerate(names)))


def run_fused(name, R, C, k, d):
    x = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w1 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    w2 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    b1, b2 = torch.randn(C, device=dev), torch.randn(C, device=dev)
    o = torch.empty_like(x)

    def call():
        check(lib.vs_op_respair(ptr(x), ptr(w1), ptr(w2), ptr(b1), ptr(b2), None, None, ptr(o), R, C, k, d, 0.1, 1.0, None, 1, st))

    def timed():
        call(); call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            call()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    ms0 = float("nan")
    if C <= 64:
        check(lib.vs_set_option(b"pair_fused", 0)); ms0 = timed()
    check(lib.vs_set_option(b"pair_fused", 2)); ms1 = timed()
    n_mma = 2 * (R / (128 - (k - 1))) * k * (C / 16)
    print("%-16s single-CTA %.3f ms   pair %.3f ms   (operand-fetch floor of the pair form %.3f ms at 1.75 GHz)" %
          (name, ms0, ms1, n_mma * max(C / 2, (4096 + 16 * C) / 128) / 148 / 1.75e6))
    buf.zero_()
    check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
    call()
    torch.cuda.synchronize()
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    check(lib.vs_set_option(b"pair_fused", 1))
    t = buf[: 148 * 20].view(148, 5, 4).double()
    used = t[:, 0, 0] > 0
    if not used.any():
        return
    lead = t[0::2][used[0::2]].mean(0)
    for r, (role, names) in enumerate((("producer", ("a_empty",)), ("mma", ("conv1 issue", "conv2 issue", "commit")), ("relay", ()),
                                        ("epilogue1", ("acc1_full", "mid_empty")), ("epilogue2", ("acc2_full",)))):
        tot = lead[r, 0].item()
        if tot > 0 and names:
            print("    %-9s total %8.0f clk  " % (role, tot) + "  ".join("%s %4.1f%%" % (nm, 100 * lead[r, 1 + i].item() / tot) for i, nm in enumerate(names)))


if os.environ.get("FUSED", "1") == "1":
    for k in (3, 7, 11):
        run_fused("s2 pair k%d d3" % k, FRAMES * 256, 64, k, 3)
    run_fused("s1 pair k3 d3", FRAMES * 64, 128, 3, 3)
    sys.exit(0)
for k in (3, 7, 11):
    run("s1 c1 k%d d1" % k, FRAMES * 64, 128, k, 1, False)
    run("s1 c2 k%d +res" % k, FRAMES * 64, 128, k, 1, True)
run("s0 c1 k3 d5", FRAMES * 8, 256, 3, 5, False)
run("s0 c2 k3 +res", FRAMES * 8, 256, 3, 1, True)
