"""The one-kernel coupling layer (csrc/umma_coupling.cu) at the C2 size: time per flow pass with the kernel on / off and, with a
diagnostics build (VS_UMMA_TIMING=1 VS_LIB_DIR=vispeech_b200/lib_timing python vispeech_b200/build.py --force; run this tool with
the same VS_LIB_DIR), where its MMA issuer waits.   python tools/coupling_timing.py"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.weights import make_state_dict
from vispeech_b200 import _lib, build_from_hparams, get_hparams_from_file
from vispeech_b200._lib import check, ptr
from vispeech_b200.layout import make_rows
lib = _lib.load()
net = build_from_hparams(get_hparams_from_file(), device="cuda:0")
net.load_state_dict(make_state_dict(1234))
lens = [431 + (i % 7) for i in range(64)]
rows = make_rows(lens, [i % 200 for i in range(64)], 4, "cuda:0")
R = rows.n_rows
z0 = torch.randn(R, 192, device="cuda:0") * (rows.row_utt >= 0)[:, None]
ws = torch.empty(int(lib.vs_workspace_bytes_latent(net._model, R, R)), dtype=torch.uint8, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda:0")


def run():
    z = z0.clone()
    check(lib.vs_flow_reverse(net._model, ctypes.byref(rows.struct), ptr(z), ptr(ws), ws.numel(), st), "flow")


for fused in (0, 1):
    net.set_option("coupling_fused", fused)
    run(); run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    print("coupling_fused=%d: %.3f ms per flow pass (%d rows, 4 coupling layers)" % (fused, e0.elapsed_time(e1) / 5, R))
check(lib.vs_set_option(b"umma_timing_buffer", buf.data_ptr()))
run()
torch.cuda.synchronize()
check(lib.vs_set_option(b"umma_timing_buffer", 0))
t = buf.view(148, 8).double()
used = t[:, 0] > 0
if used.any():
    m = t[used].mean(0)
    print("MMA issuer (last coupling layer, mean over %d CTAs): total %.0f clk; waits: weights %.1f%%  h16 %.1f%%  acc_empty %.1f%%  acts %.1f%%  x0 %.1f%%" %
          (int(used.sum()), m[0], *(100 * m[i] / m[0] for i in range(1, 6))))
