"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Compatibility shim that lets the *verbatim* reference files
(/root/reference/{modeling,bert,xlnet}.py) import and run on CPU against the
installed transformers 5.5.0 although they were written for transformers==3.0.2
(requirements.txt:348).  It is used in exactly two places:

  * oracle/make_golden.py  -- mints tests/golden/*.npz from the real reference
    (only possible in the build container, where /root/reference exists);
  * tests/test_oracle_vs_reference.py -- pins oracle/restatement.py against the
    real reference when /root/reference is present (skipped on the GPU box).

What is patched (reference call sites in parentheses):
  1. sys.modules['global_configs'] <- stub (modeling.py:4, bert.py:26, xlnet.py:6).
     The real file pins CUDA_VISIBLE_DEVICES and DEVICE=cuda:0 (global_configs.py:4,7).
  2. transformers.activations.swish                          (bert.py:14)
  3. modeling_utils.{apply_chunking_to_forward, find_pruneable_heads_and_indices,
     prune_linear_layer}                                      (bert.py:18-23)
  4. PreTrainedModel.get_head_mask                            (bert.py:208)
  5. PreTrainedModel.get_extended_attention_mask, 3.0.2 arithmetic
     (1 - m[:,None,None,:]) * -10000.0                        (bert.py:180-182)
  6. PreTrainedModel.init_weights before post_init            (bert.py:90,249; xlnet.py:40,441)
  7. transformers.modeling_bert  (BertEncoder with the 3.0.2 call signature,
     tuple return)                                            (bert.py:13,17,221-230)
  8. transformers.configuration_bert                          (bert.py:15)
  9. transformers.modeling_xlnet (XLNetLayer accepting head_mask=,
     SequenceSummary alias)                                   (xlnet.py:8-12,374-385)
"""
from __future__ import annotations

import sys
import types

import torch

REFERENCE_DIR = "/root/reference"
_installed = False


def make_global_configs(visual_dim=47, acoustic_dim=74, device="cpu"):
    g = types.ModuleType("global_configs")
    import os

    g.os = os
    g.torch = torch
    g.DEVICE = torch.device(device)
    g.ACOUSTIC_DIM = acoustic_dim
    g.VISUAL_DIM = visual_dim
    g.TEXT_DIM = 768
    g.XLNET_INJECTION_INDEX = 1
    return g


def install(visual_dim=47, acoustic_dim=74, device="cpu"):
    """Install the shim and put the reference directory on sys.path."""
    global _installed
    import transformers
    import transformers.activations as act
    import transformers.modeling_utils as mu
    from transformers import pytorch_utils as pu

    # (1) stub global_configs -- always refreshed so V/A can change between calls
    sys.modules["global_configs"] = make_global_configs(visual_dim, acoustic_dim, device)
    for name in ("modeling", "bert", "xlnet"):
        # modeling.py star-imports the dims at import time (modeling.py:4,15-19)
        sys.modules.pop(name, None)

    if _installed:
        return
    # (2)
    if not hasattr(act, "swish"):
        act.swish = torch.nn.functional.silu
    if not hasattr(act, "gelu_new"):
        act.gelu_new = act.ACT2FN["gelu_new"]
    if not hasattr(act, "gelu"):
        act.gelu = torch.nn.functional.gelu
    # (3)
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = pu.prune_linear_layer
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = getattr(
            pu, "find_pruneable_heads_and_indices", lambda *a, **k: (set(), None)
        )
    PTM = mu.PreTrainedModel

    # (4)
    def get_head_mask(self, head_mask, num_hidden_layers, *_a, **_k):
        assert head_mask is None, "shim supports head_mask=None only"
        return [None] * num_hidden_layers

    PTM.get_head_mask = get_head_mask

    # (5) transformers 3.0.2 modeling_utils.get_extended_attention_mask (2-D mask branch)
    def get_extended_attention_mask(self, attention_mask, input_shape, device=None, dtype=None):
        m = attention_mask[:, None, None, :].to(torch.float32)
        return (1.0 - m) * -10000.0

    PTM.get_extended_attention_mask = get_extended_attention_mask

    # (6)
    _orig_init_weights = PTM.init_weights

    def init_weights(self):
        if not hasattr(self, "all_tied_weights_keys"):
            self.post_init()
        else:
            _orig_init_weights(self)

    PTM.init_weights = init_weights

    # (7)
    from transformers.models.bert import modeling_bert as mb

    class BertEncoder(mb.BertEncoder):
        def forward(
            self,
            hidden_states,
            attention_mask=None,
            head_mask=None,
            encoder_hidden_states=None,
            encoder_attention_mask=None,
            output_attentions=False,
            output_hidden_states=False,
        ):
            out = super().forward(hidden_states, attention_mask=attention_mask)
            return (out.last_hidden_state,)

    m_bert = types.ModuleType("transformers.modeling_bert")
    m_bert.BertPreTrainedModel = mb.BertPreTrainedModel
    m_bert.BertEmbeddings = mb.BertEmbeddings
    m_bert.BertEncoder = BertEncoder
    m_bert.BertPooler = mb.BertPooler
    sys.modules["transformers.modeling_bert"] = m_bert
    transformers.modeling_bert = m_bert

    # (8)
    m_cfg = types.ModuleType("transformers.configuration_bert")
    m_cfg.BertConfig = transformers.BertConfig
    sys.modules["transformers.configuration_bert"] = m_cfg
    transformers.configuration_bert = m_cfg

    # (9)
    from transformers.models.xlnet import modeling_xlnet as mx

    class XLNetLayer(mx.XLNetLayer):
        def forward(self, output_h, output_g, attn_mask_h=None, attn_mask_g=None, r=None,
                    seg_mat=None, mems=None, target_mapping=None, head_mask=None,
                    output_attentions=False):
            assert head_mask is None
            return super().forward(output_h, output_g, attn_mask_h, attn_mask_g, r, seg_mat,
                                   mems=mems, target_mapping=target_mapping,
                                   output_attentions=output_attentions)

    m_xl = types.ModuleType("transformers.modeling_xlnet")
    m_xl.XLNetPreTrainedModel = mx.XLNetPreTrainedModel
    m_xl.XLNetLayer = XLNetLayer
    m_xl.SequenceSummary = mx.XLNetSequenceSummary
    sys.modules["transformers.modeling_xlnet"] = m_xl
    transformers.modeling_xlnet = m_xl

    _installed = True


def import_reference(visual_dim=47, acoustic_dim=74, device="cpu"):
    """Return the reference's (modeling, bert, xlnet) modules, imported verbatim."""
    import importlib
    import os

    if not os.path.isdir(REFERENCE_DIR):
        raise FileNotFoundError(f"{REFERENCE_DIR} is not present (only in the build container)")
    install(visual_dim, acoustic_dim, device)
    # The repo root holds drop-in modules with the same bare names; make sure the
    # reference directory wins for this import and is removed afterwards.
    saved = list(sys.path)
    try:
        sys.path.insert(0, REFERENCE_DIR)
        mods = []
        for name in ("modeling", "bert", "xlnet"):
            sys.modules.pop(name, None)
        for name in ("modeling", "bert", "xlnet"):
            mods.append(importlib.import_module(name))
    finally:
        sys.path[:] = saved
    # keep the bare names out of sys.modules so the repo's own drop-ins can be imported
    # later; transformers looks classes up through sys.modules[cls.__module__], so
    # re-register every reference module under an alias and point its classes there.
    for name, mod in zip(("modeling", "bert", "xlnet"), mods):
        alias = "_mag_reference_" + name
        sys.modules[alias] = mod
        for obj in vars(mod).values():
            if isinstance(obj, type) and getattr(obj, "__module__", None) == name:
                obj.__module__ = alias
    for name in ("modeling", "bert", "xlnet", "global_configs"):
        sys.modules.pop(name, None)
    return tuple(mods)


class MultimodalConfig:
    """Same duck type as multimodal_driver.py:76-79."""

    def __init__(self, beta_shift, dropout_prob):
        self.beta_shift = beta_shift
        self.dropout_prob = dropout_prob


def bert_config(num_labels=1, **kw):
    from transformers import BertConfig

    cfg = BertConfig(num_labels=num_labels, **kw)
    cfg._attn_implementation = "eager"
    return cfg


def xlnet_config(num_labels=1, **kw):
    from transformers import XLNetConfig

    base = dict(vocab_size=32000, d_model=768, n_layer=12, n_head=12, d_inner=3072,
                ff_activation="gelu", untie_r=True, attn_type="bi", mem_len=None,
                reuse_len=None, bi_data=False, clamp_len=-1, same_length=False,
                dropout=0.1, layer_norm_eps=1e-12, summary_type="last",
                summary_use_proj=True, summary_activation="tanh",
                summary_last_dropout=0.1, pad_token_id=5, num_labels=num_labels)
    base.update(kw)
    cfg = XLNetConfig(**base)
    cfg._attn_implementation = "eager"
    return cfg
