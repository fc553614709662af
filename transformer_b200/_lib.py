"""ctypes binding of libb200xf.so (C ABI in include/b200xf.h).

There is NO CPU fallback: if the shared library (built in-tree by `__graft_entry__.build()` /
`make -C transformer_b200/csrc`) is missing, importing a compute entry point raises immediately.
torch is used only as the buffer layer (device memory, streams); every FLOP runs in the library's kernels.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200XF_LIB_PATH") or os.path.join(_HERE, "libb200xf.so")      # (override: A/B builds in one GPU run)

_lib = None


class B200xfError(RuntimeError):
    pass


class ModelConfig(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("d_model", C.c_int32), ("num_heads", C.c_int32), ("dff", C.c_int32),
                ("input_vocab_size", C.c_int32), ("target_vocab_size", C.c_int32),
                ("max_batch", C.c_int32), ("max_src_len", C.c_int32), ("max_tgt_len", C.c_int32),
                ("pos_rows", C.c_int32), ("dropout_rate", C.c_float), ("ln_eps", C.c_float)]


_P = C.c_void_p
_I64 = C.c_int64
_I = C.c_int
_F = C.c_float
_U64 = C.c_uint64
_U32 = C.c_uint32

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGS = {
    "b200xf_gemm_bf16": [_P, _I64, _I, _P, _I64, _I, _P, _I64, _I, _P, _P, _I64, _I, _I, _I64, _I64, _I64, _P],
    "b200xf_gemm_bf16_colsum": [_P, _I64, _I, _P, _I64, _I, _P, _I64, _P, _P, _I64, _I, _I64, _I64, _I64, _P, _P],
    "b200xf_gemm_bias_dropout_residual_ln": [_P, _I64, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _F, _F, _U64, _U32, _P],
    "b200xf_attn_fwd": [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _I, _P, _I64, _I64, _I, _I, _I, _I, _I, _P],
    "b200xf_attn_bwd": [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _P, _I64, _I, _P, _I64, _P, _I64, _P, _I, _P, _I64,
                        _I64, _I, _I, _I, _I, _I, _P],
    "b200xf_attn_delta": [_P, _I64, _P, _I64, _P, _I, _I, _I, _I, _P],
    "b200xf_attn_fwd_residual": [_P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P, _P, _I, _P, _I64, _I64, _I, _I, _I, _I, _I, _P],
    "b200xf_attn_delta_residual": [_P, _P, _I64, _P, _I64, _P, _I, _I, _I, _I, _P],
    "b200xf_attn_probs": [_P, _I64, _P, _I64, _P, _P, _P, _I, _P, _I64, _I64, _I, _I, _I, _I, _I, _P],
    "b200xf_embed_fwd": [_P, _I, _I, _P, _P, _P, _P, _I64, _I, _I, _I64, _F, _U64, _U32, _P],
    "b200xf_embed_bwd": [_P, _I, _I, _P, _P, _P, _I64, _I, _I, _I64, _F, _U64, _U32, _P],
    "b200xf_residual_ln_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _F, _F, _U64, _U32, _P],
    "b200xf_residual_ln_bwd": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _F, _U64, _U32, _P],
    "b200xf_xent": [_P, _I, _I64, _P, _I, _I, _I, _P, _I64, _P, _P, _I64, _I, _F, _P],
    "b200xf_colsum_bf16": [_P, _I64, _P, _I64, _I, _P],
    "b200xf_adam_step": [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _P],
    "b200xf_adam_step_dev": [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _P],
    "b200xf_step_params": [_P, _U64, _F, _F, _F, _F, _F, _P, _P, _P, _P],
    "b200xf_model_set_seed_ptr": [_P, _P],
    "b200xf_cast_f32_bf16": [_P, _P, _I64, _P],
    "b200xf_cast_bf16_f32": [_P, _P, _I64, _P],
    "b200xf_cast_f32_bf16_strided": [_P, _P, _I64, _I, _I64, _P],
    "b200xf_pad_mask": [_P, _I, _I, _P, _I64, _I, _P],
    "b200xf_model_param_count": [C.POINTER(ModelConfig)],
    "b200xf_model_workspace_bytes": [C.POINTER(ModelConfig)],
    "b200xf_model_variable": [C.POINTER(ModelConfig), _I, C.c_char_p, _I, C.POINTER(_I64), C.POINTER(_I64),
                              C.POINTER(_I64), C.POINTER(_I64)],
    "b200xf_model_create": [C.POINTER(ModelConfig), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)],
    "b200xf_model_destroy": [_P],
    "b200xf_model_set_concurrency": [_P, _I],
    "b200xf_model_sync_shadow": [_P, _P],
    "b200xf_model_forward": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _U64, _P, _P, _I64, _P, _P, _I64, _P],
    "b200xf_model_loss": [_P, _P, _I, _I, _F, _P, _I, _P],
    "b200xf_model_backward": [_P, _I, _P],
    "b200xf_model_grad_range": [C.POINTER(ModelConfig), _I, C.POINTER(_I64), C.POINTER(_I64)],
    "b200xf_model_backward_part": [_P, _I, _I, _P],
    "b200xf_model_grad_range_part": [C.POINTER(ModelConfig), _I, _I, C.POINTER(_I64), C.POINTER(_I64)],
    "b200xf_model_set_valid_target_len_ptr": [_P, _P],
    "b200xf_model_decode_begin": [_P, _P, _I, _I, _P],
    "b200xf_model_decode_step": [_P, _P, _P, _I64, _P],
    "b200xf_model_attention_weights": [_P, _I, _I, _P, _P],
    "b200xf_model_tap": [_P, C.c_char_p, _P, _P],
    "b200xf_last_error": [],
    "b200xf_version": [],
    "b200xf_launch_count": [],
    "b200xf_prof_enable": [_I],
    "b200xf_debug_set_timeline": [_P],
    "b200xf_gemm_force_cta_group": [_I],
    "b200xf_gemm_set_chunks_per_unit": [_I],
    "b200xf_prof_collect": [C.POINTER(C.c_double), _I],
}
_RESTYPES = {"b200xf_last_error": C.c_char_p, "b200xf_launch_count": C.c_longlong,
             "b200xf_model_param_count": _I64, "b200xf_model_workspace_bytes": _I64,
             "b200xf_model_destroy": None, "b200xf_model_set_concurrency": None, "b200xf_model_set_valid_target_len_ptr": None, "b200xf_model_set_seed_ptr": None, "b200xf_prof_enable": None, "b200xf_gemm_force_cta_group": None, "b200xf_gemm_set_chunks_per_unit": None}

EXPORTED_SYMBOLS = tuple(_SIGS)


def load():
    """dlopen the library once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200xfError(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built (run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C transformer_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().b200xf_last_error()
        raise B200xfError(f"{what or 'b200xf call'} failed (status {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)


def launch_count():
    return int(load().b200xf_launch_count())
