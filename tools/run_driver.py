#!/usr/bin/env python
"""Run the reference's UNMODIFIED training driver (multimodal_driver.py) against this package's classes.

    python tools/run_driver.py /path/to/multimodal_driver.py [--workdir DIR] [--make-fixtures] [--full-geometry]
                               [--train N --dev N --test N] -- [driver arguments, e.g. --n_epochs 1 --model bert-base-uncased]

What it does (nothing of the reference is copied or edited):

 1. puts the repo root first on sys.path, so the driver's bare imports `bert`, `xlnet`, `modeling`, `global_configs`,
    `argparse_utils` (multimodal_driver.py:29-33) resolve to the drop-in modules of this repository;
 2. installs `transformers.optimization.AdamW` (multimodal_driver.py:28; removed from transformers 5.x) as
    `bert_multimodal_transformer_b200.optim.AdamW` -- transformers-3.0.2 semantics, one fused launch per step when the
    parameters are a bound model's flat buffer;
 3. disables wandb network traffic (`WANDB_MODE=disabled`; a stub module if wandb is not importable);
 4. with --make-fixtures writes, into the work directory the driver runs in, what the driver expects to find on disk
    and what cannot be downloaded here: `datasets/<name>.pkl` in the documented format (README.md:129-159, synthetic
    MOSI/MOSEI-shaped data) and local `bert-base-uncased/` / `xlnet-base-cased/` directories (config, vocabulary /
    sentencepiece model built from the synthetic corpus, randomly initialised weights) that
    `from_pretrained("<name>")` resolves offline;
 5. sets sys.argv and executes the driver with runpy as `__main__` (it parses its arguments at import time,
    multimodal_driver.py:35-57).

On a box without a GPU the driver runs through set_up_data_loader() and prep_for_training()
(multimodal_driver.py:249-351) and stops in the first forward with MagB200Error (there is no CPU path); on a B200 it
trains.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORDS = ("the a of and to in is it that was for on are as with his they at be this from or one had by word but not what "
         "all were we when your can said there use an each which she do how their if will up other about out many then "
         "them these so some her would make like him into time has look two more write go see number no way could "
         "people my than first water been call who oil its now find long down day did get come made may part movie "
         "film great bad good really very boring amazing terrible loved hated acting plot scene actor director").split()


def install_shims():
    if ROOT not in sys.path or sys.path[0] != ROOT:
        if ROOT in sys.path:
            sys.path.remove(ROOT)
        sys.path.insert(0, ROOT)
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    try:
        import wandb  # noqa: F401
    except Exception:
        stub = types.ModuleType("wandb")
        stub.init = lambda *a, **k: None
        stub.log = lambda *a, **k: None
        stub.config = types.SimpleNamespace(update=lambda *a, **k: None)
        sys.modules["wandb"] = stub
    import transformers.optimization as topt
    from bert_multimodal_transformer_b200.optim import AdamW
    topt.AdamW = AdamW            # multimodal_driver.py:28  `from transformers.optimization import AdamW`


def _synthetic_split(n, visual_dim, acoustic_dim, rng, max_words):
    import numpy as np
    out = []
    for i in range(n):
        k = int(rng.integers(1, max_words + 1))
        words = [WORDS[int(j)] for j in rng.integers(0, len(WORDS), k)]
        visual = rng.standard_normal((k, visual_dim)).astype(np.float32)
        acoustic = rng.standard_normal((k, acoustic_dim)).astype(np.float32)
        label = float(rng.uniform(-3.0, 3.0))
        out.append(((words, visual, acoustic), label, f"synthetic_{i}"))
    return out


def make_dataset(path, visual_dim, acoustic_dim, n_train, n_dev, n_test, seed=0, max_words=40):
    """README.md:129-159: {"train": [((words, visual, acoustic), label_id, segment), ...], "dev": [...], "test": [...]}"""
    import numpy as np
    rng = np.random.default_rng(seed)
    data = {"train": _synthetic_split(n_train, visual_dim, acoustic_dim, rng, max_words),
            "dev": _synthetic_split(n_dev, visual_dim, acoustic_dim, rng, max_words),
            "test": _synthetic_split(n_test, visual_dim, acoustic_dim, rng, max_words)}
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(data, f)
    return data


def make_bert_dir(path, full_geometry, seed=0):
    """config.json + vocab.txt + model.safetensors (random init, keys `bert.*` like a bert-base-uncased checkpoint)."""
    import torch
    from safetensors.torch import save_file
    from transformers import BertConfig, BertModel
    os.makedirs(path, exist_ok=True)
    specials = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    vocab = specials + sorted(set(WORDS)) + ["##s", "##ing", "##ed"]
    if full_geometry:
        vocab += [f"[filler{i}]" for i in range(30522 - len(vocab))]
        cfg = BertConfig()                                   # bert-base-uncased geometry, 12 layers, vocab 30522
    else:
        cfg = BertConfig(vocab_size=len(vocab), num_hidden_layers=2)
    with open(os.path.join(path, "vocab.txt"), "w") as f:
        f.write("\n".join(vocab) + "\n")
    cfg.save_pretrained(path)
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"do_lower_case": True, "model_max_length": 512, "tokenizer_class": "BertTokenizer"}, f)
    torch.manual_seed(seed)
    sd = {"bert." + k: v.contiguous() for k, v in BertModel(cfg).state_dict().items() if v.dtype.is_floating_point}
    save_file(sd, os.path.join(path, "model.safetensors"))


def make_xlnet_dir(path, full_geometry, seed=0):
    """config.json + spiece.model (trained on the synthetic corpus with XLNet's special-token ids) + weights."""
    import sentencepiece as spm
    import torch
    from safetensors.torch import save_file
    from transformers import XLNetConfig, XLNetModel
    os.makedirs(path, exist_ok=True)
    corpus = os.path.join(path, "_corpus.txt")
    with open(corpus, "w") as f:
        for i in range(400):
            f.write(" ".join(WORDS[(i * 7 + j * 3) % len(WORDS)] for j in range(12)) + "\n")
    # xlnet-base-cased ids: <unk>=0 <s>=1 </s>=2 <cls>=3 <sep>=4 <pad>=5 <mask>=6 <eod>=7 <eop>=8
    spm.SentencePieceTrainer.train(input=corpus, model_prefix=os.path.join(path, "spiece"), vocab_size=300,
                                   model_type="unigram", unk_id=0, bos_id=1, eos_id=2, pad_id=-1,
                                   user_defined_symbols=["<cls>", "<sep>", "<pad>", "<mask>", "<eod>", "<eop>"],
                                   hard_vocab_limit=False, minloglevel=2)
    os.remove(corpus)
    vocab_size = spm.SentencePieceProcessor(model_file=os.path.join(path, "spiece.model")).get_piece_size()
    base = dict(d_model=768, n_head=12, d_inner=3072, mem_len=None)      # xlnet-base-cased geometry
    if full_geometry:
        cfg = XLNetConfig(vocab_size=32000, n_layer=12, **base)
    else:
        cfg = XLNetConfig(vocab_size=vocab_size, n_layer=2, **base)
    cfg.save_pretrained(path)
    torch.manual_seed(seed)
    sd = {"transformer." + k: v.contiguous() for k, v in XLNetModel(cfg).state_dict().items() if v.dtype.is_floating_point}
    save_file(sd, os.path.join(path, "model.safetensors"))


def make_fixtures(workdir, dataset="mosi", model="bert-base-uncased", full_geometry=False, n_train=96, n_dev=32, n_test=32):
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import global_configs
    ds = os.path.join(workdir, "datasets", f"{dataset}.pkl")
    if not os.path.exists(ds):
        make_dataset(ds, global_configs.VISUAL_DIM, global_configs.ACOUSTIC_DIM, n_train, n_dev, n_test)
    mdir = os.path.join(workdir, model)
    if not os.path.exists(os.path.join(mdir, "model.safetensors")):
        (make_bert_dir if model == "bert-base-uncased" else make_xlnet_dir)(mdir, full_geometry)


def run(driver_path, driver_args, workdir):
    driver_path = os.path.abspath(driver_path)
    if not os.path.isfile(driver_path):
        raise SystemExit(f"driver not found: {driver_path}")
    install_shims()
    # the reference directory holds its own bert.py / xlnet.py / modeling.py: it must not shadow the drop-ins
    ref_dir = os.path.dirname(driver_path)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ref_dir]
    os.chdir(workdir)
    sys.argv = [driver_path] + list(driver_args)
    return runpy.run_path(driver_path, run_name="__main__")


def _arg(driver_args, name, default):
    for i, a in enumerate(driver_args):
        if a == name and i + 1 < len(driver_args):
            return driver_args[i + 1]
        if a.startswith(name + "="):
            return a.split("=", 1)[1]
    return default


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    driver_args = []
    if "--" in argv:
        i = argv.index("--")
        argv, driver_args = argv[:i], argv[i + 1:]
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("driver", help="path to the reference's multimodal_driver.py")
    ap.add_argument("--workdir", default=".", help="directory the driver runs in (datasets/, <model name>/ live here)")
    ap.add_argument("--make-fixtures", action="store_true", help="write the synthetic dataset and local model directory if absent")
    ap.add_argument("--full-geometry", action="store_true", help="fixtures with the real 12-layer / full-vocabulary geometry")
    ap.add_argument("--train", type=int, default=96)
    ap.add_argument("--dev", type=int, default=32)
    ap.add_argument("--test", type=int, default=32)
    a = ap.parse_args(argv)
    workdir = os.path.abspath(a.workdir)
    os.makedirs(workdir, exist_ok=True)
    dataset = _arg(driver_args, "--dataset", "mosi")
    model = _arg(driver_args, "--model", "bert-base-uncased")
    if dataset == "mosei":
        os.environ.setdefault("MAG_VISUAL_DIM", "35")     # global_configs.py:14-17 (the reference edits the file instead)
    if a.make_fixtures:
        make_fixtures(workdir, dataset, model, a.full_geometry, a.train, a.dev, a.test)
    run(a.driver, driver_args, workdir)


if __name__ == "__main__":
    main()
