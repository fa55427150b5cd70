"""CPU: the f1 row of SURVEY.md section 8 as far as a GPU-less box allows.  tools/run_driver.py runs the reference's
UNMODIFIED multimodal_driver.py against the drop-in modules: argument parsing, synthetic datasets/mosi.pkl in the
documented format, tokenisation against a local `bert-base-uncased/` directory, `from_pretrained`, the named-parameter
weight-decay grouping, `transformers.optimization.AdamW` (shimmed) and the LR scheduler all execute
(multimodal_driver.py:249-351); the run stops in the first forward with MagB200Error because there is no CPU path."""
import math
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = "/root/reference/multimodal_driver.py"

sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_driver  # noqa: E402


def test_synthetic_dataset_follows_the_documented_format(tmp_path):
    path = str(tmp_path / "datasets" / "mosi.pkl")
    run_driver.make_dataset(path, 47, 74, 5, 3, 2, seed=1)
    with open(path, "rb") as f:
        data = pickle.load(f)
    assert set(data) == {"train", "dev", "test"} and [len(data[k]) for k in ("train", "dev", "test")] == [5, 3, 2]
    (words, visual, acoustic), label, segment = data["train"][0]
    assert isinstance(words, list) and all(isinstance(w, str) for w in words)
    assert visual.shape == (len(words), 47) and acoustic.shape == (len(words), 74)
    assert isinstance(label, float) and -3.0 <= label <= 3.0 and isinstance(segment, str)


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="the reference driver only exists in the build container")
@pytest.mark.parametrize("model", ["bert-base-uncased", "xlnet-base-cased"])
def test_unmodified_driver_runs_up_to_the_first_forward(tmp_path, model):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH="")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_driver.py"), DRIVER, "--workdir", str(tmp_path), "--make-fixtures",
           "--train", "16", "--dev", "8", "--test", "8", "--", "--n_epochs", "1", "--train_batch_size", "8", "--model", model]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert "Initializing MAG with beta_shift:1.0 hidden_prob:0.5" in out                 # modeling.py:9-13 via prep_for_training
    assert "Iteration" in out                                                            # train_epoch's tqdm bar started
    assert "MagB200Error" in out and "no CPU fallback" in out                            # and the first forward refused
    assert "newly initialised" in out and "MAG.W_hv.weight" in out                       # from_pretrained reports what it did
    assert os.path.exists(tmp_path / "datasets" / "mosi.pkl") and os.path.exists(tmp_path / model / "model.safetensors")


def test_from_pretrained_refuses_a_directory_without_weights(tmp_path):
    from transformers import BertConfig
    import bert as drop_in

    class MC:
        beta_shift, dropout_prob = 1.0, 0.5
    d = tmp_path / "empty-bert"
    d.mkdir()
    BertConfig(num_hidden_layers=1, vocab_size=64).save_pretrained(str(d))
    with pytest.raises(OSError, match="no weights file"):
        drop_in.MAG_BertForSequenceClassification.from_pretrained(str(d), multimodal_config=MC(), num_labels=1)


def test_adamw_per_tensor_path_is_transformers_3_0_2_arithmetic():
    """transformers 3.0.2 optimization.AdamW.step: bias-corrected Adam update, then decoupled decay `p -= lr*wd*p`
    (applied after the update, with the un-corrected lr), eps = 1e-6 by default."""
    from bert_multimodal_transformer_b200.optim import AdamW
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(7, 5))
    q = torch.nn.Parameter(torch.randn(5))
    opt = AdamW([{"params": [p], "weight_decay": 0.01}, {"params": [q], "weight_decay": 0.0}], lr=1e-2)
    ref = {id(t): (t.detach().clone().double(), torch.zeros_like(t).double(), torch.zeros_like(t).double()) for t in (p, q)}
    b1, b2, eps, lr = 0.9, 0.999, 1e-6, 1e-2
    for step in range(1, 4):
        for t in (p, q):
            t.grad = torch.randn_like(t)
        for t, wd in ((p, 0.01), (q, 0.0)):
            w, m, v = ref[id(t)]
            g = t.grad.double()
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            w.addcdiv_(m, v.sqrt() + eps, value=-lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step))
            w.add_(w, alpha=-lr * wd)
        opt.step()
    for t in (p, q):
        assert np.allclose(t.detach().double().numpy(), ref[id(t)][0].numpy(), rtol=0, atol=2e-6)
    assert opt.fused_steps_taken == 0          # CPU tensors: the flat-buffer kernel path is never taken
    with pytest.raises(ValueError):
        AdamW([p], lr=-1.0)
