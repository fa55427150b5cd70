"""CPU (no GPU needed): the C-ABI library builds, loads and exports every symbol include/magb200.h
declares; host-side logic (flat layout, weight-decay flags, LR schedule, drop-in module surface)."""
import argparse
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bert_multimodal_transformer_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from bert_multimodal_transformer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "magb200.h")).read()
    declared = set(re.findall(r"\b(magb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/magb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.magb200_version() == 100


def test_header_is_plain_c():
    """include/magb200.h is the C-ABI contract: it must parse as C (no C++ constructs, every declaration at file scope)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for h in sorted(os.listdir(os.path.join(root, "include"))):
        if h.endswith(".h"):
            r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(root, "include", h)],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr


def test_dims_struct_matches_header():
    from bert_multimodal_transformer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "magb200.h")).read()
    body = hdr[hdr.index("typedef struct magb200_bert_dims {"):hdr.index("} magb200_bert_dims;")]
    fields = []
    for decl in body.split(";"):
        decl = re.sub(r"/\*.*?\*/", "", decl, flags=re.S).strip()
        m = re.match(r"(?:typedef struct magb200_bert_dims \{)?\s*(int32_t|float|uint64_t|const uint64_t\*)\s+(.*)", decl, re.S)
        if m:
            fields += [f.strip() for f in m.group(2).split(",")]
    assert fields == [n for n, _ in _lib.BertDims._fields_]
    assert ctypes.sizeof(_lib.BertDims) == 13 * 4 + 8 * 4 + 4 + 8 + 8   # 4 bytes of padding before the uint64


def test_no_cpu_fallback_and_no_oracle_import_in_product():
    import bert_multimodal_transformer_b200 as pkg
    pkg_dir = os.path.dirname(pkg.__file__)
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    from bert_multimodal_transformer_b200.modeling import MAG
    from bert_multimodal_transformer_b200 import _lib
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = MAG(768, 1.0, 0.5, visual_dim=47, acoustic_dim=74)
    with pytest.raises(_lib.MagB200Error, match="no CPU fallback"):
        m(torch.zeros(2, 768), torch.zeros(2, 47), torch.zeros(2, 74))


def test_flat_layout_adjacency_alignment_and_decay_flags():
    from bert_multimodal_transformer_b200 import _lib as L
    from bert_multimodal_transformer_b200.flat import FlatLayout, bert_offsets_table, bert_param_order
    from oracle import weights
    layers = 2
    sd = weights.bert_state_dict(num_layers=layers)
    order = bert_param_order(layers)
    assert sorted(order) == sorted(sd)
    lay = FlatLayout([(n, tuple(sd[n].shape)) for n in order])
    assert all(off % 256 == 0 for off, _ in lay.entries.values()) and lay.total % 256 == 0
    t = bert_offsets_table(lay, layers, 768, 896)
    assert t[L.OFF_TOTAL] == lay.total and t[L.OFF_MAG_PACKED16] == lay.total
    b = L.OFF_LAYER0
    assert t[b + L.LOFF_QKV_W] + 3 * 768 * 768 <= t[b + L.LOFF_QKV_B]
    flat = torch.zeros(lay.total)
    for n in order:
        lay.view(flat, n).copy_(sd[n])
    q = lay.view(flat, "bert.encoder.layer.1.attention.self.query.weight")
    assert torch.equal(q, sd["bert.encoder.layer.1.attention.self.query.weight"])
    flags = lay.decay_chunks()
    def flag(name):
        return int(flags[lay.offset(name) // 256])
    # multimodal_driver.py:328-343: no decay for names containing bias / LayerNorm.bias / LayerNorm.weight
    assert flag("bert.encoder.layer.0.output.dense.weight") == 1
    assert flag("bert.encoder.layer.0.output.dense.bias") == 0
    assert flag("bert.embeddings.LayerNorm.weight") == 0 and flag("bert.MAG.LayerNorm.weight") == 0
    assert flag("bert.embeddings.word_embeddings.weight") == 1 and flag("classifier.bias") == 0


def test_flat_order_is_embeddings_layers_head():
    """magb200_bert_backward_adamw and the all-reduce buckets of trainer.py slice the flat buffers per encoder layer:
    they rely on the order [embeddings | layer 0 .. L-1 | pooler, MAG, classifier] with 256-aligned layer starts
    (model_bert.cu checks the same at run time)."""
    from bert_multimodal_transformer_b200 import _lib as L
    from bert_multimodal_transformer_b200.flat import FlatLayout, bert_offsets_table, bert_param_order
    from oracle import weights
    layers = 3
    sd = weights.bert_state_dict(num_layers=layers)
    lay = FlatLayout([(n, tuple(sd[n].shape)) for n in bert_param_order(layers)])
    t = bert_offsets_table(lay, layers, 768, 896)
    starts = [int(t[L.OFF_LAYER0 + L.OFF_PER_LAYER * l + L.LOFF_QKV_W]) for l in range(layers)] + [int(t[L.OFF_POOL_W])]
    assert all(a < b for a, b in zip(starts, starts[1:])) and all(x % 256 == 0 for x in starts)
    assert int(t[L.OFF_EMB_LN_B]) < starts[0]                       # every embedding tensor sits in front of layer 0
    assert starts[-1] < int(t[L.OFF_TOTAL]) == lay.total
    per_layer = {n for n in lay.entries if ".encoder.layer." in n}
    for n, (off, _) in lay.entries.items():
        if n in per_layer:
            l = int(n.split(".encoder.layer.")[1].split(".")[0])
            assert starts[l] <= off < starts[l + 1], n              # a layer's tensors are one contiguous slice
        elif n.startswith("bert.embeddings."):
            assert off < starts[0], n
        else:
            assert off >= starts[-1], n                              # pooler, MAG, classifier


def test_linear_schedule_matches_transformers():
    from transformers import get_linear_schedule_with_warmup
    from bert_multimodal_transformer_b200.trainer import linear_schedule
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = get_linear_schedule_with_warmup(opt, num_warmup_steps=7, num_training_steps=50)
    for step in range(55):
        assert abs(opt.param_groups[0]["lr"] - linear_schedule(step, 7, 50)) < 1e-12
        opt.step()
        sch.step()


def test_drop_in_module_surface():
    """The names multimodal_driver.py:29-33 imports exist at the repo root with the reference's signatures."""
    import inspect
    import argparse_utils
    import bert
    import global_configs
    import modeling
    for n in ("ACOUSTIC_DIM", "VISUAL_DIM", "TEXT_DIM", "DEVICE", "XLNET_INJECTION_INDEX"):
        assert hasattr(global_configs, n)
    assert (global_configs.VISUAL_DIM, global_configs.ACOUSTIC_DIM, global_configs.TEXT_DIM) == (47, 74, 768)
    assert os.environ.get("CUDA_VISIBLE_DEVICES") != "0" or "CUDA_VISIBLE_DEVICES" in os.environ  # never pinned by us
    sig = inspect.signature(bert.MAG_BertForSequenceClassification.forward)
    assert list(sig.parameters)[:6] == ["self", "input_ids", "visual", "acoustic", "attention_mask", "token_type_ids"]
    assert list(inspect.signature(modeling.MAG.__init__).parameters)[:4] == ["self", "hidden_size", "beta_shift", "dropout_prob"]
    assert argparse_utils.str2bool("Yes") is True and argparse_utils.str2bool("0") is False
    with pytest.raises(argparse.ArgumentTypeError):
        argparse_utils.str2bool("maybe")
    assert 0 <= argparse_utils.seed("random") <= 9999 and argparse_utils.seed(17) == 17
    with pytest.raises(argparse.ArgumentTypeError):
        argparse_utils.seed(10000)
    with pytest.raises(argparse.ArgumentTypeError):
        argparse_utils.seed("42")       # the reference rejects numeric strings too (argparse_utils.py:27-31)


def test_state_dict_keys_equal_the_reference_models():
    import contextlib, io
    from transformers import BertConfig
    from bert_multimodal_transformer_b200.bert import MAG_BertForSequenceClassification
    from oracle import weights

    class MC:
        beta_shift, dropout_prob = 1.0, 0.5
    with contextlib.redirect_stdout(io.StringIO()):
        m = MAG_BertForSequenceClassification(BertConfig(num_labels=1, num_hidden_layers=2), MC())
    assert set(m.state_dict()) == set(weights.bert_state_dict(num_layers=2))
    # reference init (bert.py:90,249): N(0, 0.02) weights, zero biases, unit LayerNorm -- MAG included
    assert abs(m.bert.MAG.W_hv.weight.std().item() - 0.02) < 2e-3 and m.bert.MAG.W_hv.bias.abs().max().item() == 0


def test_default_valued_extras_are_recognised():
    """bert.py:147-156 / 208-209: position_ids=None means arange, head_mask=None means 'keep every head'.  The native
    path accepts exactly the values that mean the same thing and refuses the rest (no kernel for them)."""
    from bert_multimodal_transformer_b200.bert import _is_default_position_ids, _is_identity_head_mask
    assert _is_default_position_ids(None, 50)
    assert _is_default_position_ids(torch.arange(50), 50)
    assert _is_default_position_ids(torch.arange(50).unsqueeze(0).expand(4, 50), 50)
    assert not _is_default_position_ids(torch.arange(50).flip(0), 50)
    assert not _is_default_position_ids(torch.arange(49), 50)
    assert not _is_default_position_ids(torch.zeros(2, 3, 50, dtype=torch.long), 50)
    assert _is_identity_head_mask(None)
    assert _is_identity_head_mask([None] * 12)
    assert _is_identity_head_mask(torch.ones(12, 12))
    assert _is_identity_head_mask([torch.ones(12)] * 12)
    m = torch.ones(12, 12)
    m[3, 5] = 0
    assert not _is_identity_head_mask(m)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle arm the driver runs next to ours): one JSON line with the arm's keys,
    same metric / unit / workload string as the GPU arm, zero transfers and launches."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    assert d["metric"] == bench.WORKLOADS["mosi48"]["metric"] and d["config"]["workload"] == bench.WORKLOADS["mosi48"]["desc"]


def test_documents_cite_files_that_exist():
    """DESIGN.md / INTEGRATION.md / README.md / profiles/README.md / tools/README.md name evidence and source files in back-ticks;
    a citation of a file that is not in the tree is a claim without evidence."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md"), os.path.join("tools", "README.md")):
        with open(os.path.join(root, doc)) as fh:
            text = fh.read()
        for tok in set(re.findall(r"`([^`\n]+)`", text)):
            t = tok.strip().split("::")[0]
            if re.match(r"^(profiles/)?r0\d_[A-Za-z0-9_.*-]+$", t):
                path = t if t.startswith("profiles/") else "profiles/" + t
            elif re.match(r"^(tools|tests|oracle|include|profiles|bert_multimodal_transformer_b200)/[A-Za-z0-9_./*-]+$", t):
                path = t
            else:
                continue
            if path.endswith(".*"):
                path = path[:-2] + "*"
            if not glob.glob(os.path.join(root, path)) and not glob.glob(os.path.join(root, path + "*")):
                missing.append((doc, tok))
    assert not missing, missing
