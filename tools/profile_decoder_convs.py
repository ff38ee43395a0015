"""Per-shape timing of the tcgen05 decoder conv (CUDA events, op-level C-ABI calls) at the C2 workload size.
Prints, for every distinct conv shape of the HiFi-GAN decoder, time, TFLOP/s and algorithmic GB/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib                     # noqa: E402
from vispeech_b200._lib import check, ptr          # noqa: E402

FRAMES = int(os.environ.get("FRAMES", 27840))
REPS = 5
lib = _lib.load()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream


def run(name, R, cin, n, taps, dil, up=1, res=False, two_out=False, count=1):
    cout = n // up
    x = (torch.randn(cin // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w = (torch.randn(taps * cin * n, device=dev) / (cin * taps) ** 0.5).to(torch.float16)
    b = torch.randn(cout, device=dev)
    r = (torch.randn(cout // 8, R * up, 8, device=dev)).to(torch.float16) if res else None
    o1 = torch.empty(cout // 8, R * up, 8, device=dev, dtype=torch.float16)
    o2 = torch.empty_like(o1) if two_out else None
    pad_l = (taps - 1) // 2

    def call():
        check(lib.vs_op_conv1d_umma(ptr(x), ptr(w), ptr(b), ptr(r), ptr(o2), ptr(o1), R, cin, n, taps, dil, pad_l, up, 0.1, 1.0,
                                    None, 1, st))
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    flop = 2.0 * R * cin * n * taps
    if up > 1:
        flop = 2.0 * R * cin * n * 2 * (taps > 1) + 2.0 * R * cin * n * (taps == 1)   # algorithmic: K/s taps per phase
    byts = 2.0 * R * cin + 2.0 * R * up * cout * ((1 if res else 0) + 1 + (1 if two_out else 0))
    print("%-22s R=%9d Cin=%3d N=%4d k=%2d d=%d  %8.3f ms  %7.1f TFLOP/s  %7.1f GB/s   x%d = %7.2f ms" % (
        name, R, cin, n, taps, dil, ms, flop / ms / 1e9, byts / ms / 1e6, count, ms * count))
    return ms * count


total = 0.0
def all_convs():
    global total
    total = 0.0
    total += run("conv_pre", FRAMES, 192, 512, 7, 1)
    mul = 1
    for i, (s, K) in enumerate(zip((8, 8, 4, 2), (16, 16, 4, 4))):
        cin, cout = 512 >> i, 256 >> i
        taps = 1 if K == s else 3
        total += run("ups%d" % i, FRAMES * mul, cin, cout * s, taps, 1, up=s)
        mul *= s
        R = FRAMES * mul
        for k in (3, 7, 11):
            for d in (1, 3, 5):
                total += run("s%d c1 k%d d%d" % (i, k, d), R, cout, cout, k, d)
            total += run("s%d c2 k%d (res)" % (i, k), R, cout, cout, k, 1, res=True, count=3)


if os.environ.get("CONVS", "1") == "1":
    all_convs()
print("sum of decoder convs: %.2f ms" % total)


def run_pair(name, R, C, k, d, count=1):
    x = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w1 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    w2 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    b1, b2 = torch.randn(C, device=dev), torch.randn(C, device=dev)
    o = torch.empty_like(x)

    def call():
        check(lib.vs_set_option(b"fused_respair", 2))
        check(lib.vs_op_respair(ptr(x), ptr(w1), ptr(w2), ptr(b1), ptr(b2), None, ptr(o), None, R, C, k, d, 1.0, 1.0, None, 1, st))
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    flop = 2 * 2.0 * R * C * C * k
    byts = 2.0 * R * C * 2
    print("%-22s R=%9d C=%3d k=%2d d=%d  %8.3f ms  %7.1f TFLOP/s  %7.1f GB/s" % (name, R, C, k, d, ms, flop / ms / 1e9, byts / ms / 1e6))


if os.environ.get("QUICK") == "k11":
    for d in (1, 3, 5):
        run_pair("s2 pair k11 d%d" % d, FRAMES * 256, 64, 11, d)
elif os.environ.get("QUICK"):
    run_pair("s3 pair k3 d1", FRAMES * 512, 32, 3, 1)
    run_pair("s3 pair k11 d1", FRAMES * 512, 32, 11, 1)
    run_pair("s2 pair k3 d1", FRAMES * 256, 64, 3, 1)
    run_pair("s2 pair k7 d1", FRAMES * 256, 64, 7, 1)
elif os.environ.get("PAIRS", "1") == "1":
    for k in (3, 7, 11):
        for d in (1, 3, 5):
            run_pair("s3 pair k%d d%d" % (k, d), FRAMES * 512, 32, k, d)
    for k in (3, 7, 11):
        for d in (1, 3, 5):
            run_pair("s2 pair k%d d%d" % (k, d), FRAMES * 256, 64, k, d)
