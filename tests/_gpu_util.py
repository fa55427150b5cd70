"""Helpers shared by the -m gpu tests (which call the C ABI of libmagb200 through ctypes)."""
import contextlib
import io

import torch

from bert_multimodal_transformer_b200 import _lib as L


def lib():
    handle = L.load()
    L.check(handle.magb200_device_check())
    return handle


def stream():
    return L.current_stream()


def keep_mask(shape, p, seed, site):
    """The keep-mask (float 0/1) libmagb200 uses for (seed, site) over a tensor of `shape`."""
    n = 1
    for s in shape:
        n *= s
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    L.check(lib().magb200_dropout_mask(L.ptr(out), n, p, seed, site, stream()))
    return out.view(shape).float()


def make_model(layers=12, visual_dim=47, acoustic_dim=74, seed=0, reference_like=False, num_labels=1,
               beta_shift=1.0, dropout_prob=0.5):
    from transformers import BertConfig
    from bert_multimodal_transformer_b200.bert import MAG_BertForSequenceClassification
    from bert_multimodal_transformer_b200.modeling import MAG  # noqa: F401
    from oracle import weights
    import global_configs
    global_configs.VISUAL_DIM, global_configs.ACOUSTIC_DIM = visual_dim, acoustic_dim

    class MC:
        pass
    MC.beta_shift, MC.dropout_prob = beta_shift, dropout_prob
    cfg = BertConfig(num_labels=num_labels, num_hidden_layers=layers)
    with contextlib.redirect_stdout(io.StringIO()):
        model = MAG_BertForSequenceClassification(cfg, MC())
    sd = weights.bert_state_dict(num_layers=layers, visual_dim=visual_dim, acoustic_dim=acoustic_dim, seed=seed,
                                 reference_like=reference_like, num_labels=num_labels)
    model.load_state_dict(sd, strict=True)
    return model.to("cuda"), sd


def to_cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
