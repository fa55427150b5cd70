"""Where does the host time of MagBertTrainer.train_step go?"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers import BertConfig
import bert as mag_bert
from bert_multimodal_transformer_b200.trainer import MagBertTrainer
from oracle import inputs as gen
class MC: beta_shift, dropout_prob = 1.0, 0.5
with contextlib.redirect_stdout(io.StringIO()):
    model = mag_bert.MAG_BertForSequenceClassification(BertConfig(num_labels=1), MC())
model = model.cuda().train()
tr = MagBertTrainer(model, 48, 50, total_steps=1000, warmup_steps=10)
b = gen.bert_batch(48, 50, seed=1)
for _ in range(5):
    tr.train_step(b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"], b["labels"])
acc = dict(stage=0, ctl=0, replay=0, sync=0)
N = 50
for _ in range(N):
    t0 = time.perf_counter(); tr._stage(b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"], b["labels"])
    t1 = time.perf_counter(); tr.set_step_controls()
    t2 = time.perf_counter()
    with torch.cuda.stream(tr.stream):
        tr._train_graph[tr.slot].replay()
    tr.step_count += 1
    t3 = time.perf_counter(); tr.stream.synchronize(); float(tr.host_loss[tr.slot][0])
    t4 = time.perf_counter()
    acc["stage"] += t1 - t0; acc["ctl"] += t2 - t1; acc["replay"] += t3 - t2; acc["sync"] += t4 - t3
print({k: round(v / N * 1e3, 3) for k, v in acc.items()}, "ms per step")
# device time of the e2e graph vs the resident graph
for name, g in (("e2e_graph", tr._train_graph[0]),):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(tr.stream):
        e0.record()
        for _ in range(20): g.replay()
        e1.record()
    tr.stream.synchronize()
    print(name, e0.elapsed_time(e1) / 20, "ms device time per replay")
t0 = time.perf_counter(); prev = None
for _ in range(N):
    tk = tr.submit_train_step(b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"], b["labels"])
    if prev is not None: tr.result(prev)
    prev = tk
tr.result(prev)
print("pipelined submit/result:", (time.perf_counter() - t0) / N * 1e3, "ms per step")
tr.upload_staged()
tr.launch_train_step(transfers=False); tr.stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(tr.stream):
    e0.record()
    for _ in range(20): tr._train_graph_resident.replay()
    e1.record()
tr.stream.synchronize()
print("resident_graph", e0.elapsed_time(e1) / 20, "ms device time per replay")
