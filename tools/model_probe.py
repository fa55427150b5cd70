"""GPU probe: MAG-BERT forward/backward through the module API vs the CPU oracle, stage by stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from transformers import BertConfig
import bert as mag_bert
from bert_multimodal_transformer_b200 import _lib as L
from oracle import inputs as gen, restatement as R, weights

B = int(os.environ.get("PROBE_B", "8"))
class MC: beta_shift = 1.0; dropout_prob = 0.5
cfg = BertConfig(num_labels=1)
model = mag_bert.MAG_BertForSequenceClassification(cfg, MC())
sd = weights.bert_state_dict(seed=0)
model.load_state_dict(sd, strict=True)
model = model.to("cuda").eval()
b = gen.bert_batch(B, 50, seed=1234)
cb = {k: v.cuda() for k, v in b.items()}

with torch.no_grad():
    logits = model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"],
                   token_type_ids=cb["token_type_ids"])[0]
torch.cuda.synchronize()
want = {}
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
ref = R.mag_bert_forward(sdg, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"], want=want)
eng = model._engine
dims, ws = model._last
def cmp(name, got, ref):
    got = got.detach().float().cpu(); ref = ref.detach().float()
    print(f"{name:16s} max|d|={float((got-ref).abs().max()):.3e}  ref_absmean={float(ref.abs().mean()):.3e}")
cmp("emb", eng.peek(dims, ws, L.PEEK_EMB), want["emb"])
cmp("mag", eng.peek(dims, ws, L.PEEK_MAG), want["mag"])
for l in (0, 1, 5, 11):
    cmp(f"layer{l}", eng.peek(dims, ws, L.PEEK_LAYER_OUT, l), want[f"L{l}.out"])
cmp("pooled", eng.peek(dims, ws, L.PEEK_POOLED), want["pooled"])
cmp("logits", logits, ref)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bert_mosi_b8.npz"))
if B == 8:
    print("logits vs golden(reference) max|d| =", float(np.abs(logits.cpu().numpy() - g["logits"]).max()))

# backward (eval mode: dropout off) vs oracle autograd
model.zero_grad()
out = model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"],
            token_type_ids=cb["token_type_ids"])[0]
loss = torch.nn.functional.mse_loss(out.view(-1), cb["labels"].view(-1))
loss.backward()
torch.cuda.synchronize()
rl = R.mse_loss(ref, b["labels"]); rl.backward()
print("loss", loss.item(), rl.item())
worst = []
for n, p in model.named_parameters():
    gg = p.grad.detach().float().cpu(); rg = sdg[n].grad
    den = rg.norm().item() + 1e-12
    rel = (gg - rg).norm().item() / den
    worst.append((rel, n, den))
worst.sort(reverse=True)
for rel, n, den in worst[:12]:
    print(f"grad relerr {rel:.3e}  |ref|={den:.3e}  {n}")
print("median relerr", sorted(w[0] for w in worst)[len(worst)//2])

# timing (eager, no graphs)
for mode in ("eval", "train"):
    model.train(mode == "train")
    def step():
        if mode == "eval":
            with torch.no_grad():
                model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"], token_type_ids=cb["token_type_ids"])
        else:
            model.zero_grad(set_to_none=True)
            o = model(cb["input_ids"], cb["visual"], cb["acoustic"], attention_mask=cb["attention_mask"], token_type_ids=cb["token_type_ids"])[0]
            torch.nn.functional.mse_loss(o.view(-1), cb["labels"].view(-1)).backward()
    for _ in range(3): step()
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10): step()
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print(f"{mode}: {dt*1e3:.3f} ms/iter  {B/dt:.1f} samples/s (eager, wall clock)")
