"""GPU: parity on the edge-case goldens minted from the verbatim reference (oracle/make_golden.py::edge_cases):
sentences with no word tokens (CLS + SEP only; every modality row zero -> MAG's hm_norm == 0 guard on a whole sequence),
batch 1, and BERT's maximum sequence length (forward only: attention backward is limited to S <= 128)."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as gen
from _gpu_util import make_model, to_cuda

pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-3      # BASELINE.json north_star


def _case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, V, A, wseed, iseed, reflike = [int(x) for x in g["args"]]
    return g, B, V, A, wseed, iseed, bool(reflike), int(g["seq"]), (g["lens"].tolist() or None)


@pytest.mark.parametrize("name", ["bert_mosi_minlen_b4", "bert_mosi_b1", "bert_mosi_s512_b2"])
def test_bert_edge_case_logits(golden_dir, name):
    g, B, V, A, wseed, iseed, reflike, S, lens = _case(golden_dir, name)
    model, _ = make_model(12, V, A, wseed, reflike)
    model.eval()
    cb = to_cuda(gen.bert_batch(B, S, V, A, seed=iseed, lens=lens))
    with torch.no_grad():
        out = model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"],
                    token_type_ids=cb["token_type_ids"])
    assert out[0].shape == (B, 1)
    assert np.abs(out[0].cpu().numpy() - g["logits"]).max() < LOGIT_TOL


@pytest.mark.parametrize("name", ["bert_mosi_minlen_b4", "bert_mosi_b1"])
def test_bert_edge_case_gradients(golden_dir, name):
    g, B, V, A, wseed, iseed, reflike, S, lens = _case(golden_dir, name)
    model, _ = make_model(12, V, A, wseed, reflike)
    model.eval()
    cb = to_cuda(gen.bert_batch(B, S, V, A, seed=iseed, lens=lens))
    out = model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"],
                token_type_ids=cb["token_type_ids"], labels=cb["labels"])
    assert abs(out[0].item() - float(g["loss"])) < 2e-3 * max(1.0, abs(float(g["loss"])))
    out[0].backward()
    bad = []
    for k, p in model.named_parameters():
        ref_norm = float(g["gnorm/" + k])
        if ref_norm < 1e-7:
            assert p.grad.norm().item() < 1e-4, k
            continue
        if abs(p.grad.norm().item() - ref_norm) > 3e-2 * ref_norm:
            bad.append((k, p.grad.norm().item(), ref_norm))
    assert not bad, bad[:5]
