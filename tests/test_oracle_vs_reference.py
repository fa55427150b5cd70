"""CPU, build container only: the oracle restatement against the VERBATIM reference run live
through oracle/hf302_shim.py on a fresh seed (not one of the committed goldens).  Skipped
where /root/reference does not exist (the GPU box)."""
import os

import pytest
import torch

from oracle import hf302_shim as shim
from oracle import inputs as gen
from oracle import restatement as R
from oracle import weights

pytestmark = pytest.mark.skipif(not os.path.isdir(shim.REFERENCE_DIR), reason="reference tree not present")


def test_mag_live():
    modeling, _, _ = shim.import_reference(47, 74)
    m = modeling.MAG(768, 0.5, 0.5).eval()
    sd = weights.mag_state_dict("", seed=5)
    m.load_state_dict(sd)
    text, vis, aco = gen.mag_case(seed=21, zero_rows=True)
    with torch.no_grad():
        ref = m(text, vis, aco)
        got = R.mag_forward(sd, "", text, vis, aco, 0.5)
    assert (ref - got).abs().max().item() < 1e-5


def test_bert_live():
    _, bert, _ = shim.import_reference(47, 74)
    m = bert.MAG_BertForSequenceClassification(shim.bert_config(), shim.MultimodalConfig(1.0, 0.5)).eval()
    sd = weights.bert_state_dict(seed=3)
    m.load_state_dict(sd, strict=False)
    b = gen.bert_batch(4, 50, seed=77)
    with torch.no_grad():
        ref = m(b["input_ids"], b["visual"], b["acoustic"], attention_mask=b["attention_mask"],
                token_type_ids=b["token_type_ids"])[0]
        got = R.mag_bert_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                 b["token_type_ids"])
    assert (ref - got).abs().max().item() < 2e-5


def test_xlnet_live():
    _, _, xlnet = shim.import_reference(47, 74)
    m = xlnet.MAG_XLNetForSequenceClassification(shim.xlnet_config(), shim.MultimodalConfig(1.0, 0.5)).eval()
    sd = weights.xlnet_state_dict(seed=3)
    m.load_state_dict(sd)
    b = gen.xlnet_batch(4, 50, seed=78)
    with torch.no_grad():
        ref = m(b["input_ids"], b["visual"], b["acoustic"], attention_mask=b["attention_mask"].float(),
                token_type_ids=b["token_type_ids"])[0]
        got = R.mag_xlnet_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                  b["token_type_ids"])
    assert (ref - got).abs().max().item() < 2e-5
