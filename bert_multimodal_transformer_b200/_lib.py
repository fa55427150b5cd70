"""ctypes binding of libmagb200.so (the C ABI declared in include/magb200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every
entry point raises.  The library is located in-tree (next to this file) so the GPU box's
process list shows it loaded.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libmagb200.so")

F32, F16, BF16 = 0, 1, 2
EPI_STORE_F32, EPI_STORE_F16, EPI_STORE_BF16, EPI_GELU_F16, EPI_DGELU_F16, EPI_ATOMIC_F32, \
    EPI_STORE_F32_F16, EPI_RELU_F32 = range(8)

MAX_LAYERS = 48
# offsets table indices (include/magb200.h: MAGB200_OFF_*)
(OFF_WORD, OFF_POS, OFF_TYPE, OFF_EMB_LN_W, OFF_EMB_LN_B, OFF_MAG_WHV, OFF_MAG_WHA, OFF_MAG_WV, OFF_MAG_WA,
 OFF_MAG_BIAS4, OFF_MAG_LN_W, OFF_MAG_LN_B, OFF_POOL_W, OFF_POOL_B, OFF_CLS_W, OFF_CLS_B, OFF_MAG_PACKED16,
 OFF_TOTAL, OFF_RESERVED0, OFF_RESERVED1) = range(20)
OFF_LAYER0 = 20
(LOFF_QKV_W, LOFF_QKV_B, LOFF_AO_W, LOFF_AO_B, LOFF_LN1_W, LOFF_LN1_B, LOFF_F1_W, LOFF_F1_B, LOFF_F2_W, LOFF_F2_B,
 LOFF_LN2_W, LOFF_LN2_B) = range(12)
OFF_PER_LAYER = 12
PEEK_EMB, PEEK_MAG, PEEK_LAYER_OUT, PEEK_POOLED, PEEK_GATES = range(5)


# XLNet offsets (include/magb200.h: MAGB200_XOFF_* / MAGB200_XLOFF_*)
(XOFF_WORD, XOFF_MAG_WHV, XOFF_MAG_WHA, XOFF_MAG_WV, XOFF_MAG_WA, XOFF_MAG_BIAS4, XOFF_MAG_LN_W, XOFF_MAG_LN_B,
 XOFF_SUM_W, XOFF_SUM_B, XOFF_CLS_W, XOFF_CLS_B, XOFF_MAG_PACKED16, XOFF_TOTAL) = range(14)
XOFF_LAYER0 = 16
(XLOFF_Q, XLOFF_K, XLOFF_V, XLOFF_O, XLOFF_R, XLOFF_RR_BIAS, XLOFF_RS_BIAS, XLOFF_RW_BIAS, XLOFF_SEG, XLOFF_LN1_W,
 XLOFF_LN1_B, XLOFF_F1_W, XLOFF_F1_B, XLOFF_F2_W, XLOFF_F2_B, XLOFF_LN2_W, XLOFF_LN2_B, XLOFF_QKV_PACKED16) = range(18)
XOFF_PER_LAYER = 18


class XlnetDims(Structure):
    """struct magb200_xlnet_dims (include/magb200.h)."""
    _fields_ = [(n, c_int32) for n in ("batch", "seq", "hidden", "heads", "inter", "layers", "vocab", "visual_dim",
                                       "acoustic_dim", "num_labels", "ka", "injection_index")] + \
               [(n, c_float) for n in ("ln_eps", "mag_ln_eps", "beta_shift", "p_drop", "p_mag", "p_summary",
                                       "grad_scale")] + [("seed", c_uint64), ("seed_dev", c_void_p)]


class BertDims(Structure):
    """struct magb200_bert_dims (include/magb200.h)."""
    _fields_ = [(n, c_int32) for n in ("batch", "seq", "hidden", "heads", "inter", "layers", "vocab", "max_pos",
                                       "type_vocab", "visual_dim", "acoustic_dim", "num_labels", "ka")] + \
               [(n, c_float) for n in ("ln_eps", "mag_ln_eps", "beta_shift", "p_hidden", "p_attn", "p_mag", "p_cls",
                                       "grad_scale")] + [("seed", c_uint64), ("seed_dev", c_void_p)]


_lib = None

# name -> (restype, argtypes); kept in one table so tests can check it against include/magb200.h
SIGNATURES = {
    "magb200_last_error": (c_char_p, []),
    "magb200_version": (c_int, []),
    "magb200_launch_count": (c_uint64, []),
    "magb200_device_check": (c_int, []),
    "magb200_gemm_set_debug": (None, [c_void_p, c_int]),
    "magb200_gemm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int,
                             c_void_p]),
    "magb200_dropout_mask": (c_int, [c_void_p, c_int64, c_float, c_uint64, c_uint32, c_void_p]),
    "magb200_cast_f32_f16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "magb200_embed_ln_fwd": (c_int, [c_void_p] * 9 + [c_float, c_float, c_uint64, c_uint32] + [c_int] * 6
                             + [c_void_p, c_void_p, c_void_p]),
    "magb200_embed_ln_bwd": (c_int, [c_void_p] * 7 + [c_float, c_float, c_uint64, c_uint32, c_int, c_int, c_int,
                                                      c_float] + [c_void_p] * 6),
    "magb200_mag_pack_weights": (c_int, [c_void_p] * 4 + [c_int, c_int, c_int, c_void_p, c_void_p]),
    "magb200_embed_mag_scratch_bytes": (c_size_t, [c_int]),
    "magb200_embed_mag_fwd": (c_int, [c_void_p] * 9 + [c_float, c_float, c_uint64, c_uint32] + [c_int] * 6
                              + [c_void_p] * 4 + [c_float, c_float, c_float, c_uint32] + [c_void_p] * 4
                              + [c_size_t, c_void_p, c_void_p, c_void_p]),
    "magb200_mag_fwd_exact": (c_int, [c_void_p] * 10 + [c_float, c_float, c_float, c_uint64, c_uint32, c_int, c_int,
                                                        c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "magb200_mag_fwd": (c_int, [c_void_p] * 7 + [c_float, c_float, c_float, c_uint64, c_uint32, c_int, c_int, c_int,
                                                 c_int] + [c_void_p] * 5),
    "magb200_mag_bwd": (c_int, [c_void_p] * 6 + [c_float, c_float, c_float, c_uint64, c_uint32, c_int, c_int, c_int,
                                                 c_int, c_float] + [c_void_p] * 11),
    "magb200_dropout_residual_ln_fwd": (c_int, [c_void_p] * 4 + [c_float, c_float, c_uint64, c_uint32, c_int]
                                        + [c_void_p] * 4),
    "magb200_linear_ln_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int] + [c_void_p] * 4
                              + [c_float, c_float, c_uint64, c_uint32] + [c_void_p] * 4),
    "magb200_ln_bwd": (c_int, [c_void_p] * 5 + [c_float, c_uint64, c_uint32, c_int, c_float] + [c_void_p] * 6),
    "magb200_colsum_f16": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "magb200_sdpa_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_uint64, c_uint32, c_void_p,
                                 c_void_p, c_void_p]),
    "magb200_sdpa_bwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_float, c_uint64, c_uint32, c_float,
                                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "magb200_sdpa_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "magb200_cls_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_uint64, c_uint32, c_void_p,
                                c_void_p, c_void_p]),
    "magb200_cls_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_uint64, c_uint32, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "magb200_mse_loss": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "magb200_adamw_step": (c_int, [c_void_p] * 6 + [c_int64, c_float, c_float, c_float, c_float, c_float, c_int,
                                                    c_float, c_int, c_void_p, c_void_p]),
    "magb200_peer_barrier": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "magb200_adamw_peer_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                        c_int64, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_void_p, c_int,
                                        c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "magb200_peer_mark_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "magb200_peer_combine_marks": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "magb200_bert_workspace_bytes": (c_size_t, [POINTER(BertDims)]),
    "magb200_bert_refresh_operands": (c_int, [POINTER(BertDims), POINTER(c_int64), c_void_p, c_void_p, c_void_p]),
    "magb200_bert_forward": (c_int, [POINTER(BertDims), POINTER(c_int64)] + [c_void_p] * 8 + [c_size_t, c_void_p,
                                                                                              c_void_p]),
    "magb200_bert_backward": (c_int, [POINTER(BertDims), POINTER(c_int64)] + [c_void_p] * 6 + [c_size_t, c_void_p,
                                                                                               c_void_p, c_void_p]),
    "magb200_bert_backward_part": (c_int, [POINTER(BertDims), POINTER(c_int64)] + [c_void_p] * 6 + [c_size_t, c_void_p,
                                                                                                    c_void_p, c_int, c_int,
                                                                                                    c_int, c_void_p]),
    "magb200_bert_backward_adamw": (c_int, [POINTER(BertDims), POINTER(c_int64)] + [c_void_p] * 6 +
                                    [c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                     c_float, c_float, c_int, c_float, c_int, c_void_p, c_void_p]),
    "magb200_bert_attention_probs": (c_int, [POINTER(BertDims), c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "magb200_bert_peek": (c_int, [POINTER(BertDims), c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "magb200_xlnet_workspace_bytes": (c_size_t, [POINTER(XlnetDims)]),
    "magb200_xlnet_refresh_operands": (c_int, [POINTER(XlnetDims), POINTER(c_int64), c_void_p, c_void_p, c_void_p]),
    "magb200_xlnet_forward": (c_int, [POINTER(XlnetDims), POINTER(c_int64)] + [c_void_p] * 8 + [c_size_t, c_void_p,
                                                                                                c_void_p]),
    "magb200_xlnet_backward": (c_int, [POINTER(XlnetDims), POINTER(c_int64)] + [c_void_p] * 6 + [c_size_t, c_void_p,
                                                                                                 c_void_p, c_void_p]),
    "magb200_xlnet_peek": (c_int, [POINTER(XlnetDims), c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "magb200_xlnet_rel_attn_fwd": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 6 + [c_int] * 4
                                   + [c_float, c_uint64, c_uint32, c_void_p, c_void_p]),
    "magb200_xlnet_rel_attn_bwd": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 7 + [c_int] * 4
                                   + [c_float, c_uint64, c_uint32, c_float] + [c_void_p] * 7),
    "magb200_xlnet_pos_emb": (c_int, [c_int, c_int, c_float, c_uint64, c_uint32, c_void_p, c_void_p]),
}


class MagB200Error(RuntimeError):
    pass


def load(path: str | None = None):
    """Load (once) and return the ctypes handle.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("MAGB200_LIB_PATH") or LIB_PATH      # the override serves A/B builds (tools/)
    if not os.path.exists(path):
        raise MagB200Error(
            f"{path} not found: build it with `python -m bert_multimodal_transformer_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().magb200_last_error()
        raise MagB200Error(f"libmagb200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream(device=None):
    """torch's current stream of `device` (default: the current device).  The native schedules also use the CUDA
    current device for their per-device state, so callers whose tensors may live on another device wrap the call in
    `torch.cuda.device(dev)` (the model classes do)."""
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
