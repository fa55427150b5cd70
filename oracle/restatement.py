"""TEST INFRASTRUCTURE ONLY (see oracle/README.md) -- the CPU oracle.

A plain torch-fp32 restatement of the reference's hot path,
written functionally over a state dict keyed by the reference's parameter names.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product package never does.

Pinning: tests/test_oracle_vs_reference.py runs these functions against the
*verbatim* reference (oracle/hf302_shim.py) in the build container, and
tests/golden/*.npz (made by oracle/make_golden.py from the verbatim reference)
pins them everywhere else.

Every function cites the reference lines it restates.  The third-party arithmetic
(transformers==3.0.2, not vendored in /root/reference) is restated from its call
sites in bert.py / xlnet.py and SURVEY.md Appendix A.

Dropout: the reference applies nn.Dropout at the sites listed in DROPOUT_SITES.  The
oracle takes an optional `drop(site, x)` callable so a test can inject the exact
keep-masks the CUDA kernels generated; the default is eval mode (identity).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

DROPOUT_SITES = (
    "emb",            # BertEmbeddings.dropout(0.1) after LayerNorm            (bert.py:211-216)
    "mag",            # MAG.dropout(p=multimodal dropout_prob)                 (modeling.py:47-49)
    "L{i}.attn_prob",  # BertSelfAttention dropout on softmax probabilities     (App. A.2)
    "L{i}.attn_out",   # BertSelfOutput.dropout before residual+LN              (App. A.2)
    "L{i}.ffn_out",    # BertOutput.dropout before residual+LN                  (App. A.2)
    "pooled",         # MAG_BertForSequenceClassification.dropout               (bert.py:304-306)
)


def _identity(site, x):
    return x


# --------------------------------------------------------------------------- MAG
def mag_forward(sd, prefix, text, visual, acoustic, beta_shift, drop=_identity, want=None):
    """modeling.py:25-51 (MAG.forward).  Any leading dims."""
    eps = 1e-6                                                                   # modeling.py:26
    W_hv, b_hv = sd[prefix + "W_hv.weight"], sd[prefix + "W_hv.bias"]
    W_ha, b_ha = sd[prefix + "W_ha.weight"], sd[prefix + "W_ha.bias"]
    W_v, b_v = sd[prefix + "W_v.weight"], sd[prefix + "W_v.bias"]
    W_a, b_a = sd[prefix + "W_a.weight"], sd[prefix + "W_a.bias"]
    weight_v = F.relu(F.linear(torch.cat((visual, text), dim=-1), W_hv, b_hv))   # :27
    weight_a = F.relu(F.linear(torch.cat((acoustic, text), dim=-1), W_ha, b_ha))  # :28
    h_m = weight_v * F.linear(visual, W_v, b_v) + weight_a * F.linear(acoustic, W_a, b_a)  # :30
    em_norm = text.norm(2, dim=-1)                                               # :32
    hm_norm = h_m.norm(2, dim=-1)                                                # :33
    hm_norm = torch.where(hm_norm == 0, torch.ones_like(hm_norm), hm_norm)       # :35-36
    thresh = (em_norm / (hm_norm + eps)) * beta_shift                            # :38
    alpha = torch.min(thresh, torch.ones_like(thresh)).unsqueeze(-1)             # :40-43
    shifted = alpha * h_m + text                                                 # :45,48
    out = F.layer_norm(shifted, (text.shape[-1],), sd[prefix + "LayerNorm.weight"],
                       sd[prefix + "LayerNorm.bias"], 1e-5)                      # :22,48 (nn.LayerNorm default eps)
    if want is not None:
        want["mag.alpha"] = alpha.squeeze(-1)
        want["mag.h_m"] = h_m
    return drop("mag", out)                                                      # :47-49


# --------------------------------------------------------------------------- BERT
def bert_embeddings(sd, input_ids, token_type_ids, drop=_identity, eps=1e-12):
    """BertEmbeddings.forward (transformers 3.0.2; called at bert.py:211-216; App. A.1)."""
    S = input_ids.shape[1]
    pos = torch.arange(S, device=input_ids.device)
    e = (sd["bert.embeddings.word_embeddings.weight"][input_ids]
         + sd["bert.embeddings.position_embeddings.weight"][pos][None]
         + sd["bert.embeddings.token_type_embeddings.weight"][token_type_ids])
    e = F.layer_norm(e, (e.shape[-1],), sd["bert.embeddings.LayerNorm.weight"],
                     sd["bert.embeddings.LayerNorm.bias"], eps)
    return drop("emb", e)


def bert_layer(sd, i, x, ext_mask, n_head, drop=_identity, eps=1e-12, want=None):
    """BertLayer.forward (transformers 3.0.2; the loop inside bert.py:221-229; App. A.2)."""
    p = f"bert.encoder.layer.{i}."
    B, S, H = x.shape
    d = H // n_head

    def heads(t):
        return t.view(B, S, n_head, d).permute(0, 2, 1, 3)

    q = heads(F.linear(x, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"]))
    k = heads(F.linear(x, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"]))
    v = heads(F.linear(x, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"]))
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d) + ext_mask
    pr = drop(f"L{i}.attn_prob", F.softmax(s, dim=-1))
    if want is not None:
        want[f"L{i}.probs"] = pr                     # what BertSelfAttention returns with output_attentions
    ctx = torch.matmul(pr, v).permute(0, 2, 1, 3).reshape(B, S, H)
    a = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    a = F.layer_norm(drop(f"L{i}.attn_out", a) + x, (H,), sd[p + "attention.output.LayerNorm.weight"],
                     sd[p + "attention.output.LayerNorm.bias"], eps)
    f = F.gelu(F.linear(a, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    y = F.linear(f, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    y = F.layer_norm(drop(f"L{i}.ffn_out", y) + a, (H,), sd[p + "output.LayerNorm.weight"],
                     sd[p + "output.LayerNorm.bias"], eps)
    if want is not None:
        want[f"L{i}.ctx"] = ctx
        want[f"L{i}.attn_ln"] = a
        want[f"L{i}.out"] = y
    return y


def mag_bert_forward(sd, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None,
                     beta_shift=1.0, n_layer=12, n_head=12, drop=_identity, want=None):
    """MAG_BertForSequenceClassification.forward (bert.py:251-324) through
    MAG_BertModel.forward (bert.py:106-237).  Returns logits [B, num_labels]."""
    B, S = input_ids.shape
    if attention_mask is None:                                                   # bert.py:172-173
        attention_mask = torch.ones(B, S, device=input_ids.device)
    if token_type_ids is None:                                                   # bert.py:174-176
        token_type_ids = torch.zeros(B, S, dtype=torch.long, device=input_ids.device)
    ext = (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0  # bert.py:180-182 (3.0.2)
    emb = bert_embeddings(sd, input_ids, token_type_ids, drop)                   # bert.py:211-216
    x = mag_forward(sd, "bert.MAG.", emb, visual, acoustic, beta_shift, drop, want)  # bert.py:219
    if want is not None:
        want["emb"] = emb
        want["mag"] = x
    for i in range(n_layer):                                                     # bert.py:221-229
        x = bert_layer(sd, i, x, ext, n_head, drop, want=want)
    pooled = torch.tanh(F.linear(x[:, 0], sd["bert.pooler.dense.weight"], sd["bert.pooler.dense.bias"]))  # bert.py:231
    if want is not None:
        want["sequence_output"] = x
        want["pooled"] = pooled
    logits = F.linear(drop("pooled", pooled), sd["classifier.weight"], sd["classifier.bias"])  # bert.py:304-307
    return logits


# --------------------------------------------------------------------------- XLNet
def xlnet_pos_emb(S, hidden, device=None):
    """relative_positional_encoding / positional_embedding for attn_type 'bi',
    bi_data False, clamp_len -1 (xlnet.py:93-146).  Returns [2S, hidden]."""
    freq_seq = torch.arange(0, hidden, 2.0, dtype=torch.float32, device=device)   # xlnet.py:106
    inv_freq = 1 / torch.pow(10000, freq_seq / hidden)                            # xlnet.py:107
    pos_seq = torch.arange(S, -S, -1.0, dtype=torch.float32, device=device)       # xlnet.py:111,139 (beg=klen, end=-qlen)
    sinusoid = torch.einsum("i,d->id", pos_seq, inv_freq)                         # xlnet.py:94
    return torch.cat([torch.sin(sinusoid), torch.cos(sinusoid)], dim=-1)          # xlnet.py:95-96


def xlnet_layer(sd, i, h, pos_emb, key_pad, tt, n_head, drop=_identity, eps=1e-12, want=None):
    """XLNetLayer.forward, single content stream (g is None): XLNetRelativeAttention
    (rel_attn_core, rel_shift_bnij, post_attention) + XLNetFeedForward
    (transformers 3.0.2; called at xlnet.py:374-385; App. A.3).

    Batch-first restatement: h [B,S,H]; pos_emb [2S,H]; key_pad [B,S] bool (True = pad);
    tt [B,S] segment ids.  Equivalent to the reference's seq-first einsums because every
    contraction is per batch element.
    """
    p = f"transformer.layer.{i}.rel_attn."
    B, S, H = h.shape
    d = H // n_head
    Wq, Wk, Wv, Wo, Wr = (sd[p + n].reshape(H, H) for n in ("q", "k", "v", "o", "r"))
    q = (h @ Wq).view(B, S, n_head, d)
    k = (h @ Wk).view(B, S, n_head, d)
    v = (h @ Wv).view(B, S, n_head, d)
    # k_head_r: the reference projects pos_emb expanded over the batch (xlnet.py:332-333); the
    # projection does not depend on b (in eval mode), so project once.
    # In training mode the reference drops the batch-expanded pos_emb (xlnet.py:332-333), so every batch element
    # sees its own mask: pos_emb may then be [B, 2S, H] (tests inject the kernels' masks that way).
    ac = torch.einsum("bind,bjnd->bnij", q + sd[p + "r_w_bias"], k)
    if pos_emb.dim() == 3:
        kr = (pos_emb @ Wr).view(B, 2 * S, n_head, d)
        bd_raw = torch.einsum("bind,bpnd->bnip", q + sd[p + "r_r_bias"], kr)     # [B,n,S,2S]
    else:
        kr = (pos_emb @ Wr).view(2 * S, n_head, d)
        bd_raw = torch.einsum("bind,pnd->bnip", q + sd[p + "r_r_bias"], kr)      # [B,n,S,2S]
    # rel_shift_bnij: shifted[i, j] = raw[i, j + S - i]   (verified against HF)
    ii = torch.arange(S, device=h.device)[:, None]
    jj = torch.arange(S, device=h.device)[None, :]
    idx = (jj + S - ii).expand(B, n_head, S, S)
    bd = torch.gather(bd_raw, 3, idx)
    ef2 = torch.einsum("bind,snd->bnis", q + sd[p + "r_s_bias"], sd[p + "seg_embed"])  # [B,n,S,2]
    diff = (tt[:, :, None] != tt[:, None, :])                                    # xlnet.py:325 seg_mat
    ef = torch.where(diff[:, None], ef2[..., 1:2], ef2[..., 0:1])
    score = (ac + bd + ef) * (1.0 / math.sqrt(d))
    # non_tgt_mask (xlnet.py:288-296): key j masked for query i iff pad_j and i != j
    eye = torch.eye(S, dtype=torch.bool, device=h.device)
    masked = key_pad[:, None, None, :] & ~eye[None, None]
    score = score - 1e30 * masked.to(score.dtype)
    pr = drop(f"L{i}.attn_prob", F.softmax(score, dim=3))
    vec = torch.einsum("bnij,bjnd->bind", pr, v).reshape(B, S, H)
    attn_out = drop(f"L{i}.attn_out", vec @ Wo.t())                               # einsum(ibnd,hnd->ibh)
    a = F.layer_norm(attn_out + h, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], eps)
    f = f"transformer.layer.{i}.ff."
    u = F.gelu(F.linear(a, sd[f + "layer_1.weight"], sd[f + "layer_1.bias"]))
    u = drop(f"L{i}.ffn_mid", u)
    y = drop(f"L{i}.ffn_out", F.linear(u, sd[f + "layer_2.weight"], sd[f + "layer_2.bias"]))
    y = F.layer_norm(y + a, (H,), sd[f + "layer_norm.weight"], sd[f + "layer_norm.bias"], eps)
    if want is not None:
        want[f"L{i}.out"] = y
    return y


def mag_xlnet_forward(sd, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None,
                      beta_shift=1.0, n_layer=12, n_head=12, injection_index=1, drop=_identity, want=None):
    """MAG_XLNetForSequenceClassification.forward (xlnet.py:443-527) through
    MAG_XLNetModel.forward (xlnet.py:148-429), content stream only (target_mapping None,
    mems None, perm_mask None).  Returns logits [B, num_labels]."""
    B, S = input_ids.shape
    H = sd["transformer.word_embedding.weight"].shape[1]
    if attention_mask is None:
        key_pad = torch.zeros(B, S, dtype=torch.bool, device=input_ids.device)
    else:
        key_pad = (1.0 - attention_mask.to(torch.float32)) > 0                    # xlnet.py:263-264,283
    if token_type_ids is None:
        token_type_ids = torch.zeros(B, S, dtype=torch.long, device=input_ids.device)
    h = drop("emb", sd["transformer.word_embedding.weight"][input_ids])          # xlnet.py:304-305
    pos_emb = drop("pos", xlnet_pos_emb(S, H, input_ids.device))                  # xlnet.py:332-333
    for i in range(n_layer):                                                     # xlnet.py:362-388
        if i == injection_index:                                                 # xlnet.py:371-372
            h = mag_forward(sd, "transformer.MAG.", h, visual, acoustic, beta_shift, drop, want)
            if want is not None:
                want["mag"] = h
        h = xlnet_layer(sd, i, h, pos_emb, key_pad, token_type_ids, n_head, drop, want=want)
    out = drop("final", h)                                                       # xlnet.py:396
    summary = torch.tanh(F.linear(out[:, -1], sd["sequence_summary.summary.weight"],
                                  sd["sequence_summary.summary.bias"]))           # xlnet.py:506 (summary 'last')
    if want is not None:
        want["sequence_output"] = out
        want["summary"] = summary
    return F.linear(drop("pooled", summary), sd["logits_proj.weight"], sd["logits_proj.bias"])  # xlnet.py:507-509


def mse_loss(logits, labels):
    """multimodal_driver.py:371-373: MSELoss()(logits.view(-1), label_ids.view(-1))."""
    return F.mse_loss(logits.view(-1), labels.view(-1))
