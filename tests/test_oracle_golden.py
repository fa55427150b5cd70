"""CPU: the oracle restatement (oracle/restatement.py) against the golden vectors minted
from the VERBATIM reference (oracle/make_golden.py).  This is what pins the oracle on
machines without /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as gen
from oracle import restatement as R
from oracle import weights

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order (split concat GEMMs, batch-first XLNet)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _check_grads(g, named, rtol=2e-3):
    for k, t in named.items():
        ref_norm = float(g["gnorm/" + k])
        got = t.detach().double().flatten()
        assert abs(got.norm().item() - ref_norm) <= rtol * max(ref_norm, 1e-6) + 1e-7, k
        np.testing.assert_allclose(got[:8].float().numpy(), g["ghead/" + k], rtol=5e-3,
                                   atol=5e-3 * max(ref_norm / max(got.numel(), 1) ** 0.5, 1e-9) + 1e-9, err_msg=k)


@pytest.mark.parametrize("name", ["mag_mosi", "mag_mosi_zero_rows_beta1e-3", "mag_mosei"])
def test_mag_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    V, A, wseed, iseed, zero_rows, reflike = [int(x) for x in g["args"]]
    sd = {k: v.requires_grad_(True) for k, v in weights.mag_state_dict("", 768, V, A, wseed, bool(reflike)).items()}
    text, vis, aco = gen.mag_case(4, 50, 768, V, A, iseed, bool(zero_rows))
    text.requires_grad_(True)
    out = R.mag_forward(sd, "", text, vis, aco, float(g["beta_shift"]))
    assert np.abs(out.detach().numpy() - g["out"]).max() < 1e-5   # BASELINE.json configs[0] tolerance
    (out * torch.from_numpy(g["probe"])).sum().backward()
    assert np.abs(text.grad.numpy() - g["dtext"]).max() < 1e-4
    _check_grads(g, {k: v.grad for k, v in sd.items()})


@pytest.mark.parametrize("name,with_grads", [("bert_mosi_b8", True), ("bert_mosi_b8_reflike", True),
                                             ("bert_mosei_b8", False), ("bert_mosi_b48", False)])
def test_bert_matches_reference_golden(golden_dir, name, with_grads):
    g = _load(golden_dir, name)
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    sd = weights.bert_state_dict(visual_dim=V, acoustic_dim=A, seed=wseed, reference_like=bool(reflike))
    if with_grads:
        sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    b = gen.bert_batch(B, 50, V, A, seed=iseed)
    want = {}
    with torch.set_grad_enabled(with_grads):
        logits = R.mag_bert_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                    b["token_type_ids"], want=want)
        loss = R.mse_loss(logits, b["labels"])
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < TOL
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert np.abs(want["pooled"].detach().numpy() - g["pooled"]).max() < TOL
    assert np.abs(want["mag"][:2].detach().numpy() - g["mag"]).max() < TOL
    assert np.abs(want["emb"][:2].detach().numpy() - g["emb"]).max() < TOL
    assert np.abs(want["sequence_output"][:2].detach().numpy() - g["sequence_output"]).max() < 1e-4
    if with_grads:
        loss.backward()
        _check_grads(g, {k: v.grad for k, v in sd.items()})


@pytest.mark.parametrize("name,with_grads", [("xlnet_mosi_b8", True), ("xlnet_mosi_b48", False)])
def test_xlnet_matches_reference_golden(golden_dir, name, with_grads):
    g = _load(golden_dir, name)
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    sd = weights.xlnet_state_dict(visual_dim=V, acoustic_dim=A, seed=wseed, reference_like=bool(reflike))
    if with_grads:
        sd = {k: (v.requires_grad_(True) if k != "transformer.mask_emb" else v) for k, v in sd.items()}
    b = gen.xlnet_batch(B, 50, V, A, seed=iseed)
    want = {}
    with torch.set_grad_enabled(with_grads):
        logits = R.mag_xlnet_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                     b["token_type_ids"], want=want)
        loss = R.mse_loss(logits, b["labels"])
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < TOL
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert np.abs(want["mag"][:2].detach().numpy() - g["mag"]).max() < 1e-4
    assert np.abs(want["sequence_output"][:2].detach().numpy() - g["sequence_output"]).max() < 1e-4
    if with_grads:
        loss.backward()
        _check_grads(g, {k: v.grad for k, v in sd.items() if v.grad is not None})


@pytest.mark.parametrize("name,with_grads", [("bert_mosi_minlen_b4", True), ("bert_mosi_b1", True),
                                             ("bert_mosi_s512_b2", False)])
def test_bert_edge_cases_match_reference_golden(golden_dir, name, with_grads):
    """Extremes of the reference's data pipeline: sentences with no word tokens (CLS + SEP only, every modality row
    zero -> the hm_norm == 0 guard of modeling.py:35-36 on whole sequences), batch 1, and max_position_embeddings."""
    g = _load(golden_dir, name)
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    S, lens = int(g["seq"]), (g["lens"].tolist() or None)
    sd = weights.bert_state_dict(visual_dim=V, acoustic_dim=A, seed=wseed, reference_like=bool(reflike))
    if with_grads:
        sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    b = gen.bert_batch(B, S, V, A, seed=iseed, lens=lens)
    want = {}
    with torch.set_grad_enabled(with_grads):
        logits = R.mag_bert_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                    b["token_type_ids"], want=want)
        loss = R.mse_loss(logits, b["labels"])
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < TOL
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    assert np.abs(want["pooled"].detach().numpy() - g["pooled"]).max() < TOL
    if g["mag"].size:
        k = g["mag"].shape[0]
        assert np.abs(want["mag"][:k].detach().numpy() - g["mag"]).max() < TOL
        assert np.abs(want["sequence_output"][:k].detach().numpy() - g["sequence_output"]).max() < 1e-4
    if with_grads:
        loss.backward()
        _check_grads(g, {k: v.grad for k, v in sd.items()})


def test_xlnet_edge_case_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "xlnet_mosi_minlen_b4")
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    S, lens = int(g["seq"]), g["lens"].tolist()
    sd = weights.xlnet_state_dict(visual_dim=V, acoustic_dim=A, seed=wseed, reference_like=bool(reflike))
    sd = {k: (v.requires_grad_(True) if k != "transformer.mask_emb" else v) for k, v in sd.items()}
    b = gen.xlnet_batch(B, S, V, A, seed=iseed, lens=lens)
    want = {}
    logits = R.mag_xlnet_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                 b["token_type_ids"], want=want)
    loss = R.mse_loss(logits, b["labels"])
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < TOL
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert np.abs(want["mag"][:2].detach().numpy() - g["mag"]).max() < 1e-4
    loss.backward()
    _check_grads(g, {k: v.grad for k, v in sd.items() if v.grad is not None})


def test_oracle_full_tensor_gradient_error_is_at_rounding_level(golden_dir):
    """The restatement's autograd gradients against the verbatim reference through the random-projection fingerprints
    (oracle/fingerprint.py): a full-tensor relative error estimate for all 211 tensors of MAG-BERT."""
    from oracle.fingerprint import projection_error
    g = np.load(os.path.join(golden_dir, "bert_mosi_b8.npz"))
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    sd = weights.bert_state_dict(visual_dim=V, acoustic_dim=A, seed=wseed, reference_like=bool(reflike))
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    b = gen.bert_batch(B, 50, V, A, seed=iseed)
    out = R.mag_bert_forward(ref_sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"])
    R.mse_loss(out, b["labels"]).backward()
    for k, v in ref_sd.items():
        ref_norm = float(g["gnorm/" + k])
        if ref_norm < 1e-7:
            continue
        assert projection_error(k, v.grad.numpy(), g["gproj/" + k], ref_norm) < 2e-4, k
