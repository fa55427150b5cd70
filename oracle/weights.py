"""TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Deterministic, construction-order-independent parameter sets for MAG, MAG-BERT and
MAG-XLNet, keyed by the reference's state-dict names.

The reference initialises every Linear/Embedding weight ~N(0, 0.02), biases 0 and
LayerNorm (1, 0) (transformers 3.0.2 `_init_weights`, applied through
bert.py:90,249 / xlnet.py:40,441).  Reproducing *those exact draws* would tie the
goldens to torch's module construction order, so instead every tensor is drawn from
its own numpy PCG64 stream seeded by crc32(key) ^ seed: the same key always gets the
same values on any machine, with or without torch.  Biases and LayerNorm parameters
are drawn non-trivially (small noise around 0 / 1) so that parity tests exercise
them; `reference_like=True` restores the reference's exact-zero biases, which is what
makes h_m == 0 on CLS/SEP/pad rows and hits the guard at modeling.py:35-36.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rng(key: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF))


def _normal(key, shape, seed, std):
    return torch.from_numpy((_rng(key, seed).standard_normal(shape, dtype=np.float32) * std).astype(np.float32))


def _fill(sd, key, shape, seed, kind, reference_like):
    if kind == "w":
        sd[key] = _normal(key, shape, seed, 0.02)
    elif kind == "b":
        sd[key] = torch.zeros(shape) if reference_like else _normal(key, shape, seed, 0.02)
    elif kind == "ln_w":
        sd[key] = torch.ones(shape) if reference_like else 1.0 + _normal(key, shape, seed, 0.05)
    elif kind == "ln_b":
        sd[key] = torch.zeros(shape) if reference_like else _normal(key, shape, seed, 0.05)
    else:
        raise ValueError(kind)


def mag_state_dict(prefix="", hidden=768, visual_dim=47, acoustic_dim=74, seed=0, reference_like=False):
    """Keys of modeling.MAG (modeling.py:15-23)."""
    sd = {}
    _fill(sd, prefix + "W_hv.weight", (hidden, visual_dim + hidden), seed, "w", reference_like)
    _fill(sd, prefix + "W_hv.bias", (hidden,), seed, "b", reference_like)
    _fill(sd, prefix + "W_ha.weight", (hidden, acoustic_dim + hidden), seed, "w", reference_like)
    _fill(sd, prefix + "W_ha.bias", (hidden,), seed, "b", reference_like)
    _fill(sd, prefix + "W_v.weight", (hidden, visual_dim), seed, "w", reference_like)
    _fill(sd, prefix + "W_v.bias", (hidden,), seed, "b", reference_like)
    _fill(sd, prefix + "W_a.weight", (hidden, acoustic_dim), seed, "w", reference_like)
    _fill(sd, prefix + "W_a.bias", (hidden,), seed, "b", reference_like)
    _fill(sd, prefix + "LayerNorm.weight", (hidden,), seed, "ln_w", reference_like)
    _fill(sd, prefix + "LayerNorm.bias", (hidden,), seed, "ln_b", reference_like)
    return sd


def bert_state_dict(num_layers=12, hidden=768, inter=3072, vocab=30522, max_pos=512, type_vocab=2,
                    num_labels=1, visual_dim=47, acoustic_dim=74, seed=0, reference_like=False):
    """Keys of bert.MAG_BertForSequenceClassification (bert.py:240-249, 76-90)."""
    sd = {}
    f = lambda k, s, kind: _fill(sd, k, s, seed, kind, reference_like)
    f("bert.embeddings.word_embeddings.weight", (vocab, hidden), "w")
    f("bert.embeddings.position_embeddings.weight", (max_pos, hidden), "w")
    f("bert.embeddings.token_type_embeddings.weight", (type_vocab, hidden), "w")
    f("bert.embeddings.LayerNorm.weight", (hidden,), "ln_w")
    f("bert.embeddings.LayerNorm.bias", (hidden,), "ln_b")
    for i in range(num_layers):
        p = f"bert.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            f(p + f"attention.self.{n}.weight", (hidden, hidden), "w")
            f(p + f"attention.self.{n}.bias", (hidden,), "b")
        f(p + "attention.output.dense.weight", (hidden, hidden), "w")
        f(p + "attention.output.dense.bias", (hidden,), "b")
        f(p + "attention.output.LayerNorm.weight", (hidden,), "ln_w")
        f(p + "attention.output.LayerNorm.bias", (hidden,), "ln_b")
        f(p + "intermediate.dense.weight", (inter, hidden), "w")
        f(p + "intermediate.dense.bias", (inter,), "b")
        f(p + "output.dense.weight", (hidden, inter), "w")
        f(p + "output.dense.bias", (hidden,), "b")
        f(p + "output.LayerNorm.weight", (hidden,), "ln_w")
        f(p + "output.LayerNorm.bias", (hidden,), "ln_b")
    f("bert.pooler.dense.weight", (hidden, hidden), "w")
    f("bert.pooler.dense.bias", (hidden,), "b")
    sd.update(mag_state_dict("bert.MAG.", hidden, visual_dim, acoustic_dim, seed, reference_like))
    f("classifier.weight", (num_labels, hidden), "w")
    f("classifier.bias", (num_labels,), "b")
    return sd


def xlnet_state_dict(num_layers=12, hidden=768, inter=3072, vocab=32000, n_head=12, d_head=64,
                     num_labels=1, visual_dim=47, acoustic_dim=74, seed=0, reference_like=False):
    """Keys of xlnet.MAG_XLNetForSequenceClassification (xlnet.py:432-441, 15-40)."""
    sd = {}
    f = lambda k, s, kind: _fill(sd, k, s, seed, kind, reference_like)
    f("transformer.word_embedding.weight", (vocab, hidden), "w")
    # mask_emb is uninitialised memory in the reference; pin it to 0.
    sd["transformer.mask_emb"] = torch.zeros(1, 1, hidden)
    for i in range(num_layers):
        p = f"transformer.layer.{i}."
        for n in ("q", "k", "v", "o", "r"):
            f(p + f"rel_attn.{n}", (hidden, n_head, d_head), "w")
        for n in ("r_r_bias", "r_s_bias", "r_w_bias"):
            f(p + f"rel_attn.{n}", (n_head, d_head), "w")
        f(p + "rel_attn.seg_embed", (2, n_head, d_head), "w")
        f(p + "rel_attn.layer_norm.weight", (hidden,), "ln_w")
        f(p + "rel_attn.layer_norm.bias", (hidden,), "ln_b")
        f(p + "ff.layer_norm.weight", (hidden,), "ln_w")
        f(p + "ff.layer_norm.bias", (hidden,), "ln_b")
        f(p + "ff.layer_1.weight", (inter, hidden), "w")
        f(p + "ff.layer_1.bias", (inter,), "b")
        f(p + "ff.layer_2.weight", (hidden, inter), "w")
        f(p + "ff.layer_2.bias", (hidden,), "b")
    sd.update(mag_state_dict("transformer.MAG.", hidden, visual_dim, acoustic_dim, seed, reference_like))
    f("sequence_summary.summary.weight", (hidden, hidden), "w")
    f("sequence_summary.summary.bias", (hidden,), "b")
    f("logits_proj.weight", (num_labels, hidden), "w")
    f("logits_proj.bias", (num_labels,), "b")
    return sd
