"""Probe: does the tap row-shift (descriptor start address not a multiple of 8 rows) slow the tcgen05 operand fetch?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PAIRS"] = "0"
import importlib.util
spec = importlib.util.spec_from_file_location("prof", os.path.join(ROOT, "tools", "profile_decoder_convs.py"))
# reuse run() without executing the table: copy the function here
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream

def run(name, R, cin, n, taps, dil):
    x = (torch.randn(cin // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w = (torch.randn(taps * cin * n, device=dev) / (cin * taps) ** 0.5).to(torch.float16)
    b = torch.randn(n, device=dev)
    o1 = torch.empty(n // 8, R, 8, device=dev, dtype=torch.float16)
    pad_l = (taps - 1) // 2
    def call():
        check(lib.vs_op_conv1d_umma(ptr(x), ptr(w), ptr(b), None, None, ptr(o1), R, cin, n, taps, dil, pad_l, 1, 0.1, 1.0, None, 1, st))
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-18s R=%8d C=%3d k=%2d d=%2d  %7.3f ms  %7.1f TFLOP/s" % (name, R, cin, taps, dil, ms, 2.0 * R * cin * n * taps / ms / 1e9))

F = 27840
for d in (1, 2, 4, 8, 16):
    run("s0 k11", F * 8, 256, 256, 11, d)
for d in (1, 2, 4, 8, 16):
    run("s1 k11", F * 64, 128, 128, 11, d)
for d in (1, 2, 4, 8):
    run("s2 k11", F * 256, 64, 64, 11, d)
