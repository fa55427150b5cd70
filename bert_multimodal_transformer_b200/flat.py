"""Flat parameter storage for the native schedules.

libmagb200's whole-model entry points (include/magb200.h: magb200_bert_forward/backward) take ONE
fp32 buffer with every parameter at a 256-element-aligned offset, an fp16 mirror (the GEMM operand
copy) and a gradient buffer of the same shape.  The nn.Parameters the reference's driver sees
(multimodal_driver.py:328-343 groups them by name) are views into that buffer, so an external
optimizer that updates `p.data` in place updates the flat buffer too.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import torch

from . import _lib as L

ALIGN = 256


def _align(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def bert_param_order(num_layers: int) -> List[str]:
    """State-dict keys of MAG_BertForSequenceClassification in flat order.
    query|key|value (weights, then biases) and the four MAG biases are adjacent by construction."""
    names = ["bert.embeddings.word_embeddings.weight", "bert.embeddings.position_embeddings.weight",
             "bert.embeddings.token_type_embeddings.weight", "bert.embeddings.LayerNorm.weight",
             "bert.embeddings.LayerNorm.bias"]
    for i in range(num_layers):
        p = f"bert.encoder.layer.{i}."
        names += [p + "attention.self.query.weight", p + "attention.self.key.weight", p + "attention.self.value.weight",
                  p + "attention.self.query.bias", p + "attention.self.key.bias", p + "attention.self.value.bias",
                  p + "attention.output.dense.weight", p + "attention.output.dense.bias",
                  p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
                  p + "intermediate.dense.weight", p + "intermediate.dense.bias",
                  p + "output.dense.weight", p + "output.dense.bias",
                  p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]
    names += ["bert.pooler.dense.weight", "bert.pooler.dense.bias",
              "bert.MAG.W_hv.weight", "bert.MAG.W_ha.weight", "bert.MAG.W_v.weight", "bert.MAG.W_a.weight",
              "bert.MAG.W_hv.bias", "bert.MAG.W_ha.bias", "bert.MAG.W_v.bias", "bert.MAG.W_a.bias",
              "bert.MAG.LayerNorm.weight", "bert.MAG.LayerNorm.bias", "classifier.weight", "classifier.bias"]
    return names


class FlatLayout:
    """name -> (offset, shape) with 256-element alignment, plus the C offsets table."""

    def __init__(self, named_shapes: List[Tuple[str, Tuple[int, ...]]]):
        self.entries: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape in named_shapes:
            n = 1
            for s in shape:
                n *= s
            self.entries[name] = (off, tuple(shape))
            off = _align(off + n)
        self.total = off

    def offset(self, name: str) -> int:
        return self.entries[name][0]

    def numel(self, name: str) -> int:
        n = 1
        for s in self.entries[name][1]:
            n *= s
        return n

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.entries[name]
        n = self.numel(name)
        return flat[off:off + n].view(shape)

    def decay_chunks(self, no_decay=("bias", "LayerNorm.bias", "LayerNorm.weight")) -> torch.Tensor:
        """uint8 flag per 256-element chunk: 1 = weight decay applies.  Same substring rule as
        multimodal_driver.py:328-343."""
        flags = torch.zeros(self.total // ALIGN, dtype=torch.uint8)
        for name, (off, _shape) in self.entries.items():
            if not any(nd in name for nd in no_decay):
                flags[off // ALIGN:_align(off + self.numel(name)) // ALIGN] = 1
        return flags


def bert_offsets_table(layout: FlatLayout, num_layers: int, hidden: int, ka: int):
    """The int64[] consumed by magb200_bert_* (index scheme: MAGB200_OFF_* in include/magb200.h)."""
    n = L.OFF_LAYER0 + L.OFF_PER_LAYER * num_layers
    t = (ctypes.c_int64 * n)()
    o = layout.offset
    t[L.OFF_WORD] = o("bert.embeddings.word_embeddings.weight")
    t[L.OFF_POS] = o("bert.embeddings.position_embeddings.weight")
    t[L.OFF_TYPE] = o("bert.embeddings.token_type_embeddings.weight")
    t[L.OFF_EMB_LN_W] = o("bert.embeddings.LayerNorm.weight")
    t[L.OFF_EMB_LN_B] = o("bert.embeddings.LayerNorm.bias")
    t[L.OFF_MAG_WHV] = o("bert.MAG.W_hv.weight")
    t[L.OFF_MAG_WHA] = o("bert.MAG.W_ha.weight")
    t[L.OFF_MAG_WV] = o("bert.MAG.W_v.weight")
    t[L.OFF_MAG_WA] = o("bert.MAG.W_a.weight")
    t[L.OFF_MAG_BIAS4] = o("bert.MAG.W_hv.bias")
    for k, nm in enumerate(("W_hv", "W_ha", "W_v", "W_a")):
        assert o(f"bert.MAG.{nm}.bias") == t[L.OFF_MAG_BIAS4] + k * hidden
    t[L.OFF_MAG_LN_W] = o("bert.MAG.LayerNorm.weight")
    t[L.OFF_MAG_LN_B] = o("bert.MAG.LayerNorm.bias")
    t[L.OFF_POOL_W] = o("bert.pooler.dense.weight")
    t[L.OFF_POOL_B] = o("bert.pooler.dense.bias")
    t[L.OFF_CLS_W] = o("classifier.weight")
    t[L.OFF_CLS_B] = o("classifier.bias")
    t[L.OFF_MAG_PACKED16] = layout.total          # packed gate matrix lives after the mirrored region
    t[L.OFF_TOTAL] = layout.total
    for i in range(num_layers):
        p = f"bert.encoder.layer.{i}."
        b = L.OFF_LAYER0 + L.OFF_PER_LAYER * i
        t[b + L.LOFF_QKV_W] = o(p + "attention.self.query.weight")
        assert o(p + "attention.self.key.weight") == t[b + L.LOFF_QKV_W] + hidden * hidden
        assert o(p + "attention.self.value.weight") == t[b + L.LOFF_QKV_W] + 2 * hidden * hidden
        t[b + L.LOFF_QKV_B] = o(p + "attention.self.query.bias")
        assert o(p + "attention.self.key.bias") == t[b + L.LOFF_QKV_B] + hidden
        assert o(p + "attention.self.value.bias") == t[b + L.LOFF_QKV_B] + 2 * hidden
        t[b + L.LOFF_AO_W] = o(p + "attention.output.dense.weight")
        t[b + L.LOFF_AO_B] = o(p + "attention.output.dense.bias")
        t[b + L.LOFF_LN1_W] = o(p + "attention.output.LayerNorm.weight")
        t[b + L.LOFF_LN1_B] = o(p + "attention.output.LayerNorm.bias")
        t[b + L.LOFF_F1_W] = o(p + "intermediate.dense.weight")
        t[b + L.LOFF_F1_B] = o(p + "intermediate.dense.bias")
        t[b + L.LOFF_F2_W] = o(p + "output.dense.weight")
        t[b + L.LOFF_F2_B] = o(p + "output.dense.bias")
        t[b + L.LOFF_LN2_W] = o(p + "output.LayerNorm.weight")
        t[b + L.LOFF_LN2_B] = o(p + "output.LayerNorm.bias")
    return t


def packed16_numel(hidden: int, ka: int) -> int:
    return 4 * hidden * ka


# ------------------------------------------------------------------------------------------------ XLNet
def xlnet_param_order(num_layers: int) -> List[str]:
    """State-dict keys of MAG_XLNetForSequenceClassification in flat order.
    `transformer.mask_emb` is not part of the flat buffer: the content-stream path never reads it
    (xlnet.py:306-313) and it receives no gradient."""
    names = ["transformer.word_embedding.weight"]
    for i in range(num_layers):
        p = f"transformer.layer.{i}."
        names += [p + "rel_attn.q", p + "rel_attn.k", p + "rel_attn.v", p + "rel_attn.o", p + "rel_attn.r",
                  p + "rel_attn.r_r_bias", p + "rel_attn.r_s_bias", p + "rel_attn.r_w_bias", p + "rel_attn.seg_embed",
                  p + "rel_attn.layer_norm.weight", p + "rel_attn.layer_norm.bias",
                  p + "ff.layer_1.weight", p + "ff.layer_1.bias", p + "ff.layer_2.weight", p + "ff.layer_2.bias",
                  p + "ff.layer_norm.weight", p + "ff.layer_norm.bias"]
    names += ["transformer.MAG.W_hv.weight", "transformer.MAG.W_ha.weight", "transformer.MAG.W_v.weight",
              "transformer.MAG.W_a.weight", "transformer.MAG.W_hv.bias", "transformer.MAG.W_ha.bias",
              "transformer.MAG.W_v.bias", "transformer.MAG.W_a.bias", "transformer.MAG.LayerNorm.weight",
              "transformer.MAG.LayerNorm.bias", "sequence_summary.summary.weight", "sequence_summary.summary.bias",
              "logits_proj.weight", "logits_proj.bias"]
    return names


def xlnet_offsets_table(layout: FlatLayout, num_layers: int, hidden: int, ka: int):
    """int64[] for magb200_xlnet_* (MAGB200_XOFF_* / MAGB200_XLOFF_*).  The fp16 buffer holds, after the
    mirrored region: the packed MAG gate matrix, then one packed [768][2304] q|k|v operand per layer."""
    n = L.XOFF_LAYER0 + L.XOFF_PER_LAYER * num_layers
    t = (ctypes.c_int64 * n)()
    o = layout.offset
    t[L.XOFF_WORD] = o("transformer.word_embedding.weight")
    t[L.XOFF_MAG_WHV] = o("transformer.MAG.W_hv.weight")
    t[L.XOFF_MAG_WHA] = o("transformer.MAG.W_ha.weight")
    t[L.XOFF_MAG_WV] = o("transformer.MAG.W_v.weight")
    t[L.XOFF_MAG_WA] = o("transformer.MAG.W_a.weight")
    t[L.XOFF_MAG_BIAS4] = o("transformer.MAG.W_hv.bias")
    for k, nm in enumerate(("W_hv", "W_ha", "W_v", "W_a")):
        assert o(f"transformer.MAG.{nm}.bias") == t[L.XOFF_MAG_BIAS4] + k * hidden
    t[L.XOFF_MAG_LN_W] = o("transformer.MAG.LayerNorm.weight")
    t[L.XOFF_MAG_LN_B] = o("transformer.MAG.LayerNorm.bias")
    t[L.XOFF_SUM_W] = o("sequence_summary.summary.weight")
    t[L.XOFF_SUM_B] = o("sequence_summary.summary.bias")
    t[L.XOFF_CLS_W] = o("logits_proj.weight")
    t[L.XOFF_CLS_B] = o("logits_proj.bias")
    t[L.XOFF_MAG_PACKED16] = layout.total
    t[L.XOFF_TOTAL] = layout.total
    extra = layout.total + packed16_numel(hidden, ka)
    for i in range(num_layers):
        p = f"transformer.layer.{i}."
        b = L.XOFF_LAYER0 + L.XOFF_PER_LAYER * i
        for slot, nm in ((L.XLOFF_Q, "rel_attn.q"), (L.XLOFF_K, "rel_attn.k"), (L.XLOFF_V, "rel_attn.v"),
                         (L.XLOFF_O, "rel_attn.o"), (L.XLOFF_R, "rel_attn.r"), (L.XLOFF_RR_BIAS, "rel_attn.r_r_bias"),
                         (L.XLOFF_RS_BIAS, "rel_attn.r_s_bias"), (L.XLOFF_RW_BIAS, "rel_attn.r_w_bias"),
                         (L.XLOFF_SEG, "rel_attn.seg_embed"), (L.XLOFF_LN1_W, "rel_attn.layer_norm.weight"),
                         (L.XLOFF_LN1_B, "rel_attn.layer_norm.bias"), (L.XLOFF_F1_W, "ff.layer_1.weight"),
                         (L.XLOFF_F1_B, "ff.layer_1.bias"), (L.XLOFF_F2_W, "ff.layer_2.weight"),
                         (L.XLOFF_F2_B, "ff.layer_2.bias"), (L.XLOFF_LN2_W, "ff.layer_norm.weight"),
                         (L.XLOFF_LN2_B, "ff.layer_norm.bias")):
            t[b + slot] = o(p + nm)
        t[b + L.XLOFF_QKV_PACKED16] = extra + i * hidden * 3 * hidden
    return t


def xlnet_fp16_numel(layout: FlatLayout, num_layers: int, hidden: int, ka: int) -> int:
    return layout.total + packed16_numel(hidden, ka) + num_layers * hidden * 3 * hidden
