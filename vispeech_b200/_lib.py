"""ctypes binding of include/vispeech_b200.h.  No fallback: if the CUDA library is missing this raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.environ.get("VS_LIB_DIR") or os.path.join(HERE, "lib"), "libvispeech_b200.so")

VS_DTYPE_F32, VS_DTYPE_F16, VS_DTYPE_I64, VS_DTYPE_F64 = 0, 1, 2, 3


class VsConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("n_vocab", "hidden", "filter", "n_heads", "n_layers", "pitch_layers", "window",
                                       "gin", "n_speakers", "flow_layers", "n_flows", "upsample_initial", "hop")]


class VsRows(ctypes.Structure):
    _fields_ = [("n_utt", c_int32), ("n_rows", c_int32), ("max_len", c_int32), ("reserved", c_int32),
                ("row_utt", c_void_p), ("utt_start", c_void_p), ("utt_len", c_void_p), ("sid", c_void_p)]


class VsError(RuntimeError):
    pass


_SIGNATURES = {
    "vs_last_error": (c_char_p, []),
    "vs_version": (c_int32, []),
    "vs_launch_count": (c_int64, []),
    "vs_set_option": (c_int32, [c_char_p, c_int64]),
    "vs_model_set_option": (c_int32, [c_void_p, c_char_p, c_int64]),
    "vs_model_create": (c_int32, [POINTER(VsConfig), POINTER(c_void_p)]),
    "vs_model_destroy": (None, [c_void_p]),
    "vs_model_set_tensor": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_int32]),
    "vs_model_finalize": (c_int32, [c_void_p]),
    "vs_workspace_bytes": (c_int64, [c_void_p, c_int32, c_int32]),
    "vs_workspace_bytes_latent": (c_int64, [c_void_p, c_int32, c_int32]),
    "vs_workspace_bytes_decoder": (c_int64, [c_void_p, c_int32, c_int32]),
    "vs_text_encode": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_variance_adapter": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_int32, c_float, c_void_p, c_int32, c_float,
                                      c_void_p, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int64, c_void_p]),
    "vs_length_regulate_count": (c_int32, [POINTER(VsRows), c_void_p, c_void_p, c_void_p, c_void_p]),
    "vs_length_regulate_gather": (c_int32, [POINTER(VsRows), POINTER(VsRows), c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "vs_frame_prior": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_void_p, c_uint64, c_float, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_randn": (c_int32, [c_void_p, c_int64, c_uint64, c_void_p]),
    "vs_flow_reverse": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_flow_forward": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_posterior_encode": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int64, c_void_p]),
    "vs_hifigan_decode": (c_int32, [c_void_p, POINTER(VsRows), c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int64,
                                    c_void_p]),
    "vs_wave_pcm16": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p]),
    "vs_mel_spectrogram": (c_int32, [POINTER(VsRows), c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                     c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_unpack_rows": (c_int32, [POINTER(VsRows), c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "vs_op_conv1d_f32": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                   c_int32, c_int32, c_int32, c_float, c_int32, c_void_p, c_int32, c_void_p]),
    "vs_op_layernorm": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "vs_op_rel_attention": (c_int32, [POINTER(VsRows), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_op_conv1d_tf32": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                    c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "vs_op_conv1d_split": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_op_wn_layer": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                 c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vs_op_conv1d_umma": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                    c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_void_p, c_int32, c_void_p]),
    "vs_op_conv1d_umma2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_void_p, c_int32, c_void_p]),
    "vs_op_respair": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                c_int32, c_int32, c_float, c_float, c_void_p, c_int32, c_void_p]),
    "vs_op_mrf32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "vs_op_resblock64": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load() -> ctypes.CDLL:
    """Load libvispeech_b200.so (built in-tree by vispeech_b200/build.py).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VsError("%s is missing - build it with `python -m vispeech_b200.build` (or __graft_entry__.build()); "
                      "there is no CPU or PyTorch fallback for this path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().vs_last_error()
        raise VsError("%s failed (status %d): %s" % (what or "vispeech_b200 call", status,
                                                     msg.decode() if msg else "?"))


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, None -> NULL."""
    if t is None:
        return None
    return t.data_ptr()
