"""The one-kernel last MRF stage (csrc/umma_mrf.cu) at the C2 size: time per launch and, with a diagnostics build
(VS_UMMA_TIMING=1 VS_LIB_DIR=vispeech_b200/lib_timing python vispeech_b200/build.py --force; run this tool with the same
VS_LIB_DIR), where its MMA warp and the two epilogue crews wait.   python tools/mrf_timing.py [frames]"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispeech_b200 import _lib
from vispeech_b200._lib import check, ptr
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 27840
R = frames * 512
g = torch.Generator(device=dev).manual_seed(0)
hi = (torch.randn(4, R, 8, device=dev, generator=g) * 0.3).to(torch.float16)
lo = (torch.randn(4, R, 8, device=dev, generator=g) * 1e-4).to(torch.float16)
ks = (3, 7, 11)
wp = [(torch.randn(ks[j] * 32 * 32, device=dev, generator=g) / (32 * ks[j]) ** 0.5 * 0.5).to(torch.float16) for j in range(3) for m in range(3) for c in range(2)]
bh = [torch.randn(32) * 0.05 for _ in range(18)]
w_arr = (ctypes.c_void_p * 18)(*[t.data_ptr() for t in wp])
b_arr = (ctypes.c_void_p * 18)(*[t.data_ptr() for t in bh])
pw = torch.randn(7, 32) * 0.05
row_utt = torch.zeros(frames, dtype=torch.int32, device=dev)
wave = torch.empty(R, device=dev)


def call():
    check(lib.vs_op_mrf32(ptr(hi), ptr(lo), ctypes.cast(w_arr, ctypes.c_void_p), ctypes.cast(b_arr, ctypes.c_void_p), ptr(pw),
                          ptr(row_utt), 512, R, ptr(wave), st), "vs_op_mrf32")


check(lib.vs_set_option(b"umma_timing_buffer", 0))
call(); call(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flop = 2 * 6 * 32 * 32 * 21 * R
print("umma_mrf  frames=%d rows=%d  %.3f ms/launch  %.0f TFLOP/s (useful)  finite=%s" % (frames, R, ms, flop / ms * 1e-9, bool(torch.isfinite(wave).all())))
if os.environ.get("VS_LIB_DIR"):
    buf = torch.zeros(148 * 4 * 5, dtype=torch.int64, device=dev)
    check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
    call(); torch.cuda.synchronize()
    check(lib.vs_set_option(b"umma_timing_buffer", 0))
    t = buf.view(148, 4, 5).double().mean(0)
    names = (("conv1 MMA", ("a_ready", "acc1_empty", "issue", "w_full")), ("conv2 MMA", ("mid_ready", "-", "issue", "w_full")), ("epilogue 1", ("acc1_full", "-", "-", "-")),
             ("epilogue 2", ("x_full", "bar.sync", "-", "-")))
    for r, (role, nm) in enumerate(names):
        tot = t[r, 0].item()
        print("    %-10s total %9.0f clk  " % (role, tot) + "  ".join("%s %4.1f%%" % (n, 100 * t[r, 1 + i].item() / tot) for i, n in enumerate(nm) if n != "-"))
