"""Run one fused ResBlock pair a few times (for ncu captures): python tools/one_pair.py R C taps dil"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
R, C, k, d = [int(a) for a in sys.argv[1:5]]
check(lib.vs_set_option(b"fused_respair", 2))
x = (torch.randn(C // 8, R, 8, device=dev) * 0.5).to(torch.float16)
w1 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
w2 = (torch.randn(k * C * C, device=dev) / (C * k) ** 0.5).to(torch.float16)
b1, b2 = torch.randn(C, device=dev), torch.randn(C, device=dev)
o = torch.empty_like(x)
for _ in range(3):
    check(lib.vs_op_respair(ptr(x), ptr(w1), ptr(w2), ptr(b1), ptr(b2), None, None, ptr(o), R, C, k, d, 0.1, 1.0, None, 1, st))
torch.cuda.synchronize()
