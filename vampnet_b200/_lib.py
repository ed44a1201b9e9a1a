"""ctypes binding of the C ABI declared in include/vampnet_b200.h.

The library is built in-tree by ``python -m vampnet_b200.build`` (nvcc, sm_100a).  There is no
CPU fallback: if the shared object is missing or cannot be loaded, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvampnet_b200.so")

EPI_BF16, EPI_QKV, EPI_RESID, EPI_GEGLU, EPI_BIAS_F32 = range(5)
FAMILIES = ("embed", "rmsnorm", "gemm_qkv", "attention", "gemm_attn_out", "gemm_ffn_up", "gemm_ffn_down",
            "gemm_classifier", "sample_remask", "state")


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_heads", "n_layers", "n_codebooks", "n_conditioning_codebooks", "latent_dim", "d_model", "vocab_size")]


class Weights(C.Structure):
    _fields_ = [
        ("emb_table", C.c_void_p), ("emb_w3", C.c_void_p), ("emb_b", C.c_void_p), ("norm1", C.c_void_p),
        ("wqkv", C.c_void_p), ("wo", C.c_void_p), ("norm3", C.c_void_p), ("w1", C.c_void_p), ("w2", C.c_void_p),
        ("norm_f", C.c_void_p), ("wcls", C.c_void_p), ("bcls", C.c_void_p), ("rel_bias", C.c_void_p),
        ("rel_sat", C.c_int32),
    ]


class GenParams(C.Structure):
    _fields_ = [
        ("sampling_steps", C.c_int32), ("temperature", C.c_float), ("gamma", C.POINTER(C.c_float)),
        ("temp_eff", C.POINTER(C.c_float)), ("do_sample", C.POINTER(C.c_int32)), ("seed_lo", C.c_uint32),
        ("seed_hi", C.c_uint32), ("use_graph", C.c_int32), ("top_p", C.c_float),
    ]


_SIGS = {
    "vnb_abi_version": (C.c_int32, []),
    "vnb_last_error": (C.c_char_p, []),
    "vnb_model_create": (C.c_int32, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(C.c_void_p)]),
    "vnb_model_destroy": (None, [C.c_void_p]),
    "vnb_forward_codes": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vnb_forward_latents": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vnb_forward_latents_acts": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "vnb_get_hidden": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vnb_generate": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(GenParams),
                                 C.c_void_p, C.c_void_p]),
    "vnb_launch_count": (C.c_uint64, []),
    "vnb_graph_capture_count": (C.c_uint64, []),
    "vnb_set_option": (C.c_int32, [C.c_char_p, C.c_int32]),
    "vnb_get_option": (C.c_int32, [C.c_char_p, C.POINTER(C.c_int32)]),
    "vnb_profile_begin": (C.c_int32, [C.c_void_p]),
    "vnb_profile_end": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32]),
    "vnb_sample_step": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                    C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]),
    "vnb_op_gemm": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vnb_op_attention": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "vnb_codec_conv1d": (C.c_int32, [C.c_void_p] * 6 + [C.c_int32] * 13 + [C.c_void_p]),
    "vnb_codec_rvq": (C.c_int32, [C.c_int32] + [C.c_void_p] * 11 + [C.c_int32] * 6 + [C.c_void_p] * 3),
    "vnb_codec_conv_tc": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "vnb_codec_conv_in": (C.c_int32, [C.c_void_p] * 7 + [C.c_int32] * 5 + [C.c_void_p]),
    "vnb_codec_conv_out": (C.c_int32, [C.c_void_p] * 5 + [C.c_int32] * 5 + [C.c_void_p]),
    "vnb_set_error_cuda": (C.c_int32, [C.c_char_p, C.c_int32]),
    "vnb_dbg_gemm_ref": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """Load (once) and return the shared library.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # not a fallback: the only way to get the kernels is to compile them (nvcc cross-compiles sm_100a anywhere)
        try:
            from . import build as _build
            _build.build()
        except Exception as e:
            raise RuntimeError(
                f"{LIB_PATH} not found and building it failed ({e}); build it with `python -m vampnet_b200.build` "
                "(nvcc, sm_100a). vampnet_b200 has no CPU or PyTorch fallback.") from e
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.vnb_abi_version() != 2:
        raise RuntimeError("ABI version mismatch between _lib.py and libvampnet_b200.so")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise RuntimeError("vampnet_b200: " + lib().vnb_last_error().decode(errors="replace"))


def ptr(t):
    """Raw device/host address of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
