"""-m gpu: MAG-XLNet (BASELINE.json configs[3]) -- relative-attention operator against a torch fp32
restatement, and the whole model through the reference-shaped API against the golden vectors minted
from the VERBATIM reference (logits within 1e-3) and the CPU oracle (gradients)."""
import contextlib
import io
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bert_multimodal_transformer_b200 import _lib as L
from oracle import inputs as gen
from oracle import restatement as R
from oracle import weights
from _gpu_util import keep_mask, lib, rel_err, stream, to_cuda

pytestmark = pytest.mark.gpu


def make_xlnet(layers=12, visual_dim=47, acoustic_dim=74, seed=0):
    from oracle.hf302_shim import xlnet_config
    from bert_multimodal_transformer_b200.xlnet import MAG_XLNetForSequenceClassification
    import global_configs
    global_configs.VISUAL_DIM, global_configs.ACOUSTIC_DIM = visual_dim, acoustic_dim

    class MC:
        beta_shift, dropout_prob = 1.0, 0.5
    cfg = xlnet_config(n_layer=layers)
    with contextlib.redirect_stdout(io.StringIO()):
        model = MAG_XLNetForSequenceClassification(cfg, MC())
    sd = weights.xlnet_state_dict(num_layers=layers, visual_dim=visual_dim, acoustic_dim=acoustic_dim, seed=seed)
    model.load_state_dict(sd, strict=True)
    return model.to("cuda"), sd


def test_pos_emb_matches_oracle():
    S = 50
    pos = torch.empty(2 * S, 768, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_xlnet_pos_emb(S, 1, 0.0, 0, 4, L.ptr(pos), stream()))
    ref = R.xlnet_pos_emb(S, 768)
    assert (pos.float().cpu() - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("S,p", [(50, 0.0), (50, 0.1), (23, 0.0)])
def test_rel_attn_fwd_bwd(S, p):
    g = torch.Generator().manual_seed(11)
    B, heads, d = 3, 12, 64
    Hd = heads * d
    qkv = (torch.randn(B * S, 3 * Hd, generator=g) * 0.6).half()
    kr = (torch.randn(2 * S, Hd, generator=g) * 0.6).half()
    rw, rr, rs = (torch.randn(heads, d, generator=g) * 0.3 for _ in range(3))
    seg = torch.randn(2, heads, d, generator=g) * 0.3
    mask = torch.ones(B, S, dtype=torch.int64)
    mask[1, :S // 3] = 0                       # XLNet pads on the left
    tt = torch.zeros(B, S, dtype=torch.int64)
    tt[:, -1] = 2
    tt[1, :S // 3] = 3
    seed, site = 5, 16
    x = qkv.float().view(B, S, 3, heads, d).requires_grad_(True)
    krf = kr.float().view(2 * S, heads, d).requires_grad_(True)
    rwp, rrp, rsp, segp = (t.clone().requires_grad_(True) for t in (rw, rr, rs, seg))
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]                                     # [B,S,n,d]
    ac = torch.einsum("bind,bjnd->bnij", q + rwp, k)
    bd_raw = torch.einsum("bind,pnd->bnip", q + rrp, krf)
    ii, jj = torch.arange(S)[:, None], torch.arange(S)[None, :]
    bd = torch.gather(bd_raw, 3, (jj + S - ii).expand(B, heads, S, S))
    ef2 = torch.einsum("bind,snd->bnis", q + rsp, segp)
    diff = (tt[:, :, None] != tt[:, None, :])
    ef = torch.where(diff[:, None], ef2[..., 1:2], ef2[..., 0:1])
    score = (ac + bd + ef) * 0.125
    masked = (mask == 0)[:, None, None, :] & ~torch.eye(S, dtype=torch.bool)[None, None]
    score = score - 1e30 * masked.float()
    pr = F.softmax(score, dim=3)
    if p > 0:
        pr = pr * keep_mask((B, heads, S, S), p, seed, site).cpu() / (1 - p)
    ctx_ref = torch.einsum("bnij,bjnd->bind", pr, v).reshape(B * S, Hd)
    dctx = torch.randn(B * S, Hd, generator=g).half()
    ctx_ref.backward(dctx.float())
    dev = [t.cuda().contiguous() for t in (qkv, kr, rw, rr, rs, seg, mask, tt, dctx)]
    qkv_d, kr_d, rw_d, rr_d, rs_d, seg_d, mask_d, tt_d, dctx_d = dev
    ctx = torch.empty(B * S, Hd, device="cuda", dtype=torch.float16)
    L.check(lib().magb200_xlnet_rel_attn_fwd(L.ptr(qkv_d), L.ptr(kr_d), 1, L.ptr(rw_d), L.ptr(rr_d), L.ptr(rs_d),
                                             L.ptr(seg_d), L.ptr(mask_d), L.ptr(tt_d), B, S, heads, d, p, seed, site,
                                             L.ptr(ctx), stream()))
    assert (ctx.float().cpu() - ctx_ref.detach()).abs().max().item() < 5e-3
    dqkv = torch.zeros(B * S, 3 * Hd, device="cuda", dtype=torch.float16)
    dkr = torch.zeros(2 * S, Hd, device="cuda")
    d_rw, d_rr, d_rs = (torch.zeros(heads, d, device="cuda") for _ in range(3))
    d_seg = torch.zeros(2, heads, d, device="cuda")
    L.check(lib().magb200_xlnet_rel_attn_bwd(L.ptr(qkv_d), L.ptr(kr_d), 1, L.ptr(rw_d), L.ptr(rr_d), L.ptr(rs_d),
                                             L.ptr(seg_d), L.ptr(mask_d), L.ptr(tt_d), L.ptr(dctx_d), B, S, heads, d, p, seed,
                                             site, 1.0, L.ptr(dqkv), L.ptr(dkr), L.ptr(d_rw), L.ptr(d_rr), L.ptr(d_rs),
                                             L.ptr(d_seg), stream()))
    dqkv_ref = x.grad.reshape(B * S, 3 * Hd)
    assert rel_err(dqkv.float(), dqkv_ref) < 6e-3
    assert rel_err(dkr, krf.grad.reshape(2 * S, Hd)) < 6e-3
    for got, want in ((d_rw, rwp.grad), (d_rr, rrp.grad), (d_rs, rsp.grad), (d_seg, segp.grad)):
        assert rel_err(got, want) < 6e-3


def _forward(model, cb, **kw):
    return model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"],
                 token_type_ids=cb["token_type_ids"], **kw)


@pytest.mark.parametrize("name", ["xlnet_mosi_b8", "xlnet_mosi_b48"])
def test_xlnet_logits_match_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    model, _ = make_xlnet(12, V, A, wseed)
    model.eval()
    cb = to_cuda(gen.xlnet_batch(B, 50, V, A, seed=iseed))
    with torch.no_grad():
        out = _forward(model, cb)
        seq = model.transformer(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"],
                                token_type_ids=cb["token_type_ids"])[0]
    assert isinstance(out, tuple) and len(out) == 1 and out[0].shape == (B, 1)
    assert np.abs(out[0].cpu().numpy() - g["logits"]).max() < 1e-3
    assert np.abs(seq[:2].cpu().numpy() - g["sequence_output"]).max() < 2e-2
    dims, ws = model._last
    assert np.abs(model._engine.peek(dims, ws, L.PEEK_MAG)[:2].cpu().numpy() - g["mag"]).max() < 1e-2


def test_xlnet_gradients_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "xlnet_mosi_b8.npz"))
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    model, _ = make_xlnet(12, V, A, wseed)
    model.eval()
    cb = to_cuda(gen.xlnet_batch(B, 50, V, A, seed=iseed))
    out = _forward(model, cb, labels=cb["labels"])
    assert len(out) == 2 and abs(out[0].item() - float(g["loss"])) < 2e-3
    out[0].backward()
    bad = []
    for k, p in model.named_parameters():
        if k == "transformer.mask_emb":
            assert p.grad is None
            continue
        ref_norm = float(g["gnorm/" + k])
        if ref_norm < 1e-7:
            assert p.grad.norm().item() < 1e-4, k
            continue
        if abs(p.grad.norm().item() - ref_norm) > 3e-2 * ref_norm:
            bad.append((k, p.grad.norm().item(), ref_norm))
    assert not bad, bad[:6]


def test_xlnet_full_gradient_against_oracle_and_train_mode_runs():
    layers = 3
    model, sd = make_xlnet(layers, seed=2)
    model.eval()
    b = gen.xlnet_batch(5, 50, seed=5)
    cb = to_cuda(b)
    loss = _forward(model, cb, labels=cb["labels"])[0]
    loss.backward()
    ref_sd = {k: (v.clone().requires_grad_(True) if k != "transformer.mask_emb" else v) for k, v in sd.items()}
    ref = R.mag_xlnet_forward(ref_sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"],
                              n_layer=layers)
    R.mse_loss(ref, b["labels"]).backward()
    assert abs(loss.item() - R.mse_loss(ref, b["labels"]).item()) < 2e-3
    for k, p in model.named_parameters():
        if k == "transformer.mask_emb" or ref_sd[k].grad.norm().item() < 1e-7:
            continue
        assert rel_err(p.grad, ref_sd[k].grad) < 3e-2, (k, rel_err(p.grad, ref_sd[k].grad))
    # training mode: dropout on every site, position dropout per batch element; finite and different from eval
    model.train()
    model.zero_grad()
    l2 = _forward(model, cb, labels=cb["labels"])[0]
    l2.backward()
    assert math.isfinite(l2.item()) and abs(l2.item() - loss.item()) > 1e-6
    assert all(torch.isfinite(p.grad).all() for k, p in model.named_parameters() if p.grad is not None)


def test_xlnet_training_mode_matches_oracle_with_the_kernels_dropout_masks():
    """Dropout on at every site of MAG_XLNetModel.forward (word embeddings xlnet.py:304-305, the batch-expanded position
    embeddings :332-333, MAG, attention probabilities / attention output / FFN activation / FFN output per layer, the
    final output :396, the summary): the oracle is fed the exact keep-masks the kernels generate."""
    layers, B, S = 2, 4, 50
    model, sd = make_xlnet(layers, seed=3)
    model.train()
    b = gen.xlnet_batch(B, S, seed=12)
    cb = to_cuda(b)
    torch.manual_seed(123)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    torch.manual_seed(123)
    loss = _forward(model, cb, labels=cb["labels"])[0]
    loss.backward()
    sites = {"emb": (1, (B, S, 768), 0.1), "mag": (2, (B, S, 768), 0.5), "pooled": (3, (B, 768), 0.1),
             "pos": (4, (B, 2 * S, 768), 0.1), "final": (5, (B, S, 768), 0.1)}
    for l in range(layers):
        sites[f"L{l}.attn_prob"] = (16 + 4 * l, (B, 12, S, S), 0.1)
        sites[f"L{l}.attn_out"] = (17 + 4 * l, (B, S, 768), 0.1)
        sites[f"L{l}.ffn_out"] = (18 + 4 * l, (B, S, 768), 0.1)
        sites[f"L{l}.ffn_mid"] = (19 + 4 * l, (B, S, 3072), 0.1)
    masks = {k: (keep_mask(shape, p, seed, site).cpu(), p) for k, (site, shape, p) in sites.items()}

    def drop(site, x):
        m, p = masks[site]
        return x * m / (1 - p)            # "pos": [2S, H] * [B, 2S, H] -> one dropped copy per batch element

    ref_sd = {k: (v.clone().requires_grad_(True) if k != "transformer.mask_emb" else v) for k, v in sd.items()}
    ref = R.mag_xlnet_forward(ref_sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"],
                              n_layer=layers, drop=drop)
    ref_loss = R.mse_loss(ref, b["labels"])
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 5e-3 * max(1.0, abs(ref_loss.item()))
    named = dict(model.named_parameters())
    for k in ("logits_proj.weight", "transformer.layer.1.ff.layer_2.weight", "transformer.layer.0.rel_attn.q",
              "transformer.layer.1.rel_attn.r", "transformer.layer.0.rel_attn.r_r_bias", "transformer.MAG.W_hv.weight",
              "transformer.layer.1.ff.layer_1.weight", "transformer.word_embedding.weight"):
        assert rel_err(named[k].grad, ref_sd[k].grad) < 4e-2, (k, rel_err(named[k].grad, ref_sd[k].grad))


def test_xlnet_api_contract():
    model, _ = make_xlnet(2, seed=1)
    model.eval()
    cb = to_cuda(gen.xlnet_batch(2, 50, seed=3))
    with pytest.raises(ValueError, match="You have to specify either input_ids or inputs_embeds"):
        model(None, cb["visual"], cb["acoustic"])
    with pytest.raises(NotImplementedError):
        model(cb["input_ids"], cb["visual"], cb["acoustic"], perm_mask=torch.zeros(2, 50, 50))
    with torch.no_grad():
        a = _forward(model, cb)[0]
        b = model(cb["input_ids"], cb["visual"], cb["acoustic"], token_type_ids=cb["token_type_ids"],
                  input_mask=1.0 - cb["attention_mask"].float())[0]       # input_mask: 1 = padding (xlnet.py:263)
    assert torch.equal(a, b)
