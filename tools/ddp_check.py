"""torchrun entry (one rank per GPU): BASELINE.json configs[2] in miniature -- MOSEI-shaped batch sharded over
the ranks, gradient all-reduce inside the captured training step; after N optimizer steps the parameters must
equal a single-process run on the whole batch (dropout disabled so both runs are deterministic)."""
import contextlib, io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from transformers import BertConfig

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import global_configs
global_configs.VISUAL_DIM, global_configs.ACOUSTIC_DIM = 35, 74
from bert_multimodal_transformer_b200.bert import MAG_BertForSequenceClassification
from bert_multimodal_transformer_b200.parallel import shard_batch
from bert_multimodal_transformer_b200.trainer import MagBertTrainer
from oracle import inputs as gen, weights

LAYERS, GLOBAL_B, STEPS = 4, 16 * world, 3

def make():
    class MC: beta_shift, dropout_prob = 1.0, 0.0
    cfg = BertConfig(num_labels=1, num_hidden_layers=LAYERS, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = MAG_BertForSequenceClassification(cfg, MC())
    m.load_state_dict(weights.bert_state_dict(num_layers=LAYERS, visual_dim=35, acoustic_dim=74, seed=1))
    return m.cuda().train()

full = gen.bert_batch(GLOBAL_B, 50, 35, 74, seed=99)
mine = shard_batch(full, rank, world)
args = lambda b: (b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"], b["labels"])

model = make()
tr = MagBertTrainer(model, GLOBAL_B // world, 50, lr=1e-3, weight_decay=0.01, total_steps=0, use_cuda_graph=True)
losses = [tr.train_step(*args(mine)) for _ in range(STEPS)]
torch.cuda.synchronize()
ok = True
if rank == 0:
    ref_model = make()
    ref = MagBertTrainer(ref_model, GLOBAL_B, 50, lr=1e-3, weight_decay=0.01, total_steps=0, use_cuda_graph=False,
                         world_size=1)
    ref_losses = [ref.train_step(*args(full)) for _ in range(STEPS)]
    torch.cuda.synchronize()
    a, b = model._engine.flat, ref_model._engine.flat
    num = (a - b).norm().item(); den = (b - torch.zeros_like(b)).norm().item()
    upd = (b - MagBertTrainer.__mro__[0].__name__ and b).norm().item()
    init = make()._engine if False else None
    drift = (a - b).abs().max().item()
    print(json.dumps({"world": world, "max_abs_param_diff": drift, "rel_param_diff": num / den,
                      "rank0_local_losses": losses, "single_process_losses": ref_losses}))
    ok = drift < 2e-4      # lr 1e-3 Adam steps are ~1e-3 each: a wrong reduction shows up as >1e-3
# every rank must hold identical parameters after the all-reduced steps
chk = model._engine.flat.double().sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print("replica checksum spread", float(hi - lo))
    ok = ok and float(hi - lo) == 0.0
    print("DDP_CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
