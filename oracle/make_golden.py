"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.npz from the VERBATIM reference.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

The reference ships no golden vectors, KATs or fixtures, so
parity is pinned by running the reference's own modeling.py / bert.py / xlnet.py
(imported unmodified through oracle/hf302_shim.py) on seeded inputs
(oracle/inputs.py) and seeded parameters (oracle/weights.py) and committing what it
returns.  The GPU box has no /root/reference; it only reads the .npz files.

Each file stores the generator arguments it was made with, so a test can rebuild the
identical inputs/parameters without this script.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import hf302_shim as shim
from . import inputs as gen
from . import weights
from .fingerprint import project

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def grad_fingerprint(named_grads):
    """Per-parameter (l2 norm, sum, first 8 elements, 4 random projections) -- small but position-sensitive."""
    out = {}
    for k, g in named_grads.items():
        g = g.detach().double().flatten()
        out["gnorm/" + k] = np.float64(g.norm().item())
        out["gsum/" + k] = np.float64(g.sum().item())
        out["ghead/" + k] = g[:8].float().numpy()
        out["gproj/" + k] = project(k, g.numpy())
    return out


def make_mag(name, visual_dim, acoustic_dim, beta_shift, zero_rows, wseed, iseed, reference_like):
    modeling, _, _ = shim.import_reference(visual_dim, acoustic_dim)
    m = modeling.MAG(768, beta_shift, 0.5).eval()
    sd = weights.mag_state_dict("", 768, visual_dim, acoustic_dim, wseed, reference_like)
    m.load_state_dict(sd, strict=True)
    text, vis, aco = gen.mag_case(4, 50, 768, visual_dim, acoustic_dim, iseed, zero_rows)
    text.requires_grad_(True)
    out = m(text, vis, aco)
    probe = torch.from_numpy(np.random.Generator(np.random.PCG64(99)).standard_normal(out.shape, dtype=np.float32))
    (out * probe).sum().backward()
    blob = dict(out=out.detach().numpy(), dtext=text.grad.numpy(), probe=probe.numpy(),
                args=np.array([visual_dim, acoustic_dim, wseed, iseed, int(zero_rows), int(reference_like)]),
                beta_shift=np.float64(beta_shift))
    blob.update(grad_fingerprint({k: p.grad for k, p in m.named_parameters()}))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **blob)
    print(name, "out mean abs", float(out.abs().mean()))


def make_bert(name, batch, visual_dim, acoustic_dim, wseed, iseed, reference_like, with_grads, keep_rows=2, seq=50,
              lens=None):
    _, bert, _ = shim.import_reference(visual_dim, acoustic_dim)
    cfg = shim.bert_config(num_labels=1)
    m = bert.MAG_BertForSequenceClassification(cfg, shim.MultimodalConfig(1.0, 0.5)).eval()
    sd = weights.bert_state_dict(visual_dim=visual_dim, acoustic_dim=acoustic_dim, seed=wseed,
                                 reference_like=reference_like)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    b = gen.bert_batch(batch, seq, visual_dim, acoustic_dim, seed=iseed, lens=lens)
    feats = {}
    hooks = [m.bert.MAG.register_forward_hook(lambda _m, _i, o: feats.__setitem__("mag", o.detach())),
             m.bert.embeddings.register_forward_hook(lambda _m, _i, o: feats.__setitem__("emb", o.detach()))]
    with torch.set_grad_enabled(with_grads):
        enc = m.bert(b["input_ids"], b["visual"], b["acoustic"], attention_mask=b["attention_mask"],
                     token_type_ids=b["token_type_ids"])
        logits = m(b["input_ids"], b["visual"], b["acoustic"], attention_mask=b["attention_mask"],
                   token_type_ids=b["token_type_ids"])[0]
        loss = torch.nn.functional.mse_loss(logits.view(-1), b["labels"].view(-1))
    for h in hooks:
        h.remove()
    blob = dict(logits=logits.detach().numpy(), loss=np.float64(loss.item()),
                pooled=enc[1].detach().numpy(), sequence_output=enc[0][:keep_rows].detach().numpy(),
                mag=feats["mag"][:keep_rows].numpy(), emb=feats["emb"][:keep_rows].numpy(),
                args=np.array([batch, visual_dim, acoustic_dim, wseed, iseed, int(reference_like)]),
                seq=np.int64(seq), lens=np.asarray([] if lens is None else lens, np.int64))
    if with_grads:
        m.zero_grad()
        loss.backward()
        blob.update(grad_fingerprint({k: p.grad for k, p in m.named_parameters()}))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **blob)
    print(name, "logits", logits.detach().flatten()[:4].tolist(), "loss", loss.item())


def make_xlnet(name, batch, visual_dim, acoustic_dim, wseed, iseed, reference_like, with_grads, keep_rows=2, seq=50,
               lens=None):
    _, _, xlnet = shim.import_reference(visual_dim, acoustic_dim)
    cfg = shim.xlnet_config(num_labels=1)
    m = xlnet.MAG_XLNetForSequenceClassification(cfg, shim.MultimodalConfig(1.0, 0.5)).eval()
    sd = weights.xlnet_state_dict(visual_dim=visual_dim, acoustic_dim=acoustic_dim, seed=wseed,
                                  reference_like=reference_like)
    m.load_state_dict(sd, strict=True)
    b = gen.xlnet_batch(batch, seq, visual_dim, acoustic_dim, seed=iseed, lens=lens)
    feats = {}
    hook = m.transformer.MAG.register_forward_hook(lambda _m, _i, o: feats.__setitem__("mag", o.detach()))
    with torch.set_grad_enabled(with_grads):
        # the driver passes the int mask as attention_mask (multimodal_driver.py:363-370);
        # xlnet.py:264 computes 1.0 - attention_mask, so a float mask is equivalent.
        enc = m.transformer(b["input_ids"], b["visual"], b["acoustic"], attention_mask=b["attention_mask"].float(),
                            token_type_ids=b["token_type_ids"])
        logits = m(b["input_ids"], b["visual"], b["acoustic"], attention_mask=b["attention_mask"].float(),
                   token_type_ids=b["token_type_ids"])[0]
        loss = torch.nn.functional.mse_loss(logits.view(-1), b["labels"].view(-1))
    hook.remove()
    blob = dict(logits=logits.detach().numpy(), loss=np.float64(loss.item()),
                sequence_output=enc[0][:keep_rows].detach().numpy(),
                mag=feats["mag"].transpose(0, 1)[:keep_rows].numpy(),
                args=np.array([batch, visual_dim, acoustic_dim, wseed, iseed, int(reference_like)]),
                seq=np.int64(seq), lens=np.asarray([] if lens is None else lens, np.int64))
    if with_grads:
        m.zero_grad()
        loss.backward()
        blob.update(grad_fingerprint({k: p.grad for k, p in m.named_parameters() if p.grad is not None}))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **blob)
    print(name, "logits", logits.detach().flatten()[:4].tolist(), "loss", loss.item())


def edge_cases():
    """Shapes the reference's data pipeline can produce at the extremes (multimodal_driver.py:130-205): a sentence
    with no word tokens (CLS + SEP only: every modality row is zero -> MAG's hm_norm == 0 guard on the whole
    sequence, XLNet rows that are all padding except two), batch 1, and BERT's maximum sequence length."""
    make_bert("bert_mosi_minlen_b4", 4, 47, 74, 0, 77, True, True, lens=[2, 3, 50, 2])
    make_bert("bert_mosi_b1", 1, 47, 74, 0, 78, False, True)
    make_bert("bert_mosi_s512_b2", 2, 47, 74, 0, 79, False, True, keep_rows=0, seq=512, lens=[512, 300])    # configs[4] length
    make_bert("bert_mosi_s128_b4", 4, 47, 74, 0, 81, False, True, keep_rows=0, seq=128, lens=[128, 65, 100, 7])
    make_xlnet("xlnet_mosi_minlen_b4", 4, 47, 74, 0, 80, False, True, lens=[2, 3, 50, 2])


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if os.environ.get("MAKE_GOLDEN_ONLY") == "edge":
        edge_cases()
        return
    # BASELINE.json configs[0]: MAG(hidden=768, visual=47, acoustic=74), bs=4 seq=50
    make_mag("mag_mosi", 47, 74, 1.0, False, 0, 7, False)
    make_mag("mag_mosi_zero_rows_beta1e-3", 47, 74, 1e-3, True, 0, 8, True)   # alpha<1 branch + hm_norm==0 guard
    make_mag("mag_mosei", 35, 74, 1.0, True, 1, 9, False)
    # configs[1] shape (MOSI) at bs=8 with grads, and at the full bs=48 forward-only
    make_bert("bert_mosi_b8", 8, 47, 74, 0, 1234, False, True)
    make_bert("bert_mosi_b48", 48, 47, 74, 0, 1234, False, True)               # configs[1] itself, with gradients
    make_bert("bert_mosi_b8_reflike", 8, 47, 74, 0, 4321, True, True)          # zero biases -> guard rows
    # configs[2] shape (MOSEI, V=35)
    make_bert("bert_mosei_b8", 8, 35, 74, 1, 1234, False, True)
    # configs[3]: MAG-XLNet
    make_xlnet("xlnet_mosi_b8", 8, 47, 74, 0, 1234, False, True)
    make_xlnet("xlnet_mosi_b48", 48, 47, 74, 0, 1234, False, False)
    edge_cases()


if __name__ == "__main__":
    main()
