"""Run one tcgen05 conv shape a few times (for ncu captures): python tools/one_conv.py R Cin N taps dil [res] [two_out]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
R, cin, n, taps, dil = [int(a) for a in sys.argv[1:6]]
res = len(sys.argv) > 6 and sys.argv[6] == "1"
two = len(sys.argv) > 7 and sys.argv[7] == "1"
x = (torch.randn(cin // 8, R, 8, device=dev) * 0.5).to(torch.float16)
w = (torch.randn(taps * cin * n, device=dev) / (cin * taps) ** 0.5).to(torch.float16)
b = torch.randn(n, device=dev)
r = torch.randn(n // 8, R, 8, device=dev).to(torch.float16) if res else None
o1 = torch.empty(n // 8, R, 8, device=dev, dtype=torch.float16)
o2 = torch.empty_like(o1) if two else None
for _ in range(4):
    check(lib.vs_op_conv1d_umma(ptr(x), ptr(w), ptr(b), ptr(r), ptr(o2), ptr(o1), R, cin, n, taps, dil, (taps - 1) // 2, 1, 0.1, 1.0, None, 1, st))
torch.cuda.synchronize()
