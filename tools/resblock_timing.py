"""The whole-ResBlock kernel of the C = 64 stage (csrc/umma_resblock.cu) at the C2 size: time per launch against the three fused
conv-pair launches it replaces (VS_LIB_DIR selects the build).   python tools/resblock_timing.py"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
R = 28800 * 256
g = torch.Generator(device=dev).manual_seed(0)
a = (torch.randn(8, R, 8, device=dev, generator=g) * 0.5).to(torch.float16)
w = [(torch.randn(3 * 64 * 64, device=dev, generator=g) / (64 * 3) ** 0.5).to(torch.float16) for _ in range(6)]
bh = [torch.randn(64) * 0.05 for _ in range(6)]
bd = [b.to(dev) for b in bh]
w_arr = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in w])
b_arr = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in bh])
row_utt = torch.zeros(28800, dtype=torch.int32, device=dev)
out = torch.empty_like(a); t1 = torch.empty_like(a)


def whole():
    check(lib.vs_op_resblock64(ptr(a), ctypes.cast(w_arr, ctypes.c_void_p), ctypes.cast(b_arr, ctypes.c_void_p), ptr(row_utt), 256, R,
                               ptr(out), st), "vs_op_resblock64")


def three():
    src = a
    for m, d in enumerate((1, 3, 5)):
        dst = t1 if m % 2 == 0 else out
        check(lib.vs_op_respair(ptr(src), ptr(w[2 * m]), ptr(w[2 * m + 1]), ptr(bd[2 * m]), ptr(bd[2 * m + 1]), None, None, ptr(dst), R, 64, 3, d,
                                0.1, 1.0, ptr(row_utt), 256, st), "vs_op_respair")
        src = dst


for name, fn in (("whole ResBlock, one kernel", whole), ("three fused conv-pair launches", three)):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-34s %.3f ms" % (name, e0.elapsed_time(e1) / 5))
