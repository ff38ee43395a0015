"""Experiment: do an epilogue-bound (k=3) and an MMA-bound (k=11) fused ResBlock kernel overlap when they share the SMs?
Runs the two C=32 pair kernels back to back on one stream, then concurrently on two streams with half the CTA slots each."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"
R, C = 27840 * 512, 32
def mk(k):
    x = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16)
    w1 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    w2 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
    return x, w1, w2, torch.randn(C, device=dev), torch.randn(C, device=dev), torch.empty_like(x)
A, B = mk(int(sys.argv[1]) if len(sys.argv) > 1 else 3), mk(int(sys.argv[2]) if len(sys.argv) > 2 else 11)
ka, kb = A[1].numel() // (C * C), B[1].numel() // (C * C)
def call(t, k, stream):
    x, w1, w2, b1, b2, o = t
    check(lib.vs_op_respair(ptr(x), ptr(w1), ptr(w2), ptr(b1), ptr(b2), None, ptr(o), None, R, C, k, 1, 1.0, 1.0, None, 1, stream.cuda_stream))
s0 = torch.cuda.current_stream(); s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for div, conc in ((1, False), (2, True), (1, True)):
    check(lib.vs_set_option(b"respair_grid_div", div))
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s0)
        if conc:
            s1.wait_stream(s0); s2.wait_stream(s0)
            for _ in range(3):
                call(A, ka, s1); call(B, kb, s2)
            s0.wait_stream(s1); s0.wait_stream(s2)
        else:
            for _ in range(3):
                call(A, ka, s0); call(B, kb, s0)
        e1.record(s0); torch.cuda.synchronize()
    print("k=%d + k=%d  grid_div=%d  %s: %.3f ms per (A,B)" % (ka, kb, div, "two streams" if conc else "one stream ", e0.elapsed_time(e1) / 3))
check(lib.vs_set_option(b"respair_grid_div", 1))
