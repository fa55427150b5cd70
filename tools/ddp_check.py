"""torchrun entry (one rank per GPU): BASELINE.json configs[2] in miniature -- MOSEI-shaped batch sharded over
the ranks, gradient all-reduce inside the captured training step; after N optimizer steps the parameters must
equal a single-process run on the whole batch (dropout disabled so both runs are deterministic)."""
import contextlib, faulthandler, io, json, os, sys
faulthandler.enable()
faulthandler.dump_traceback_later(90, exit=False)      # a hang leaves a Python traceback in the log
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from transformers import BertConfig

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host_group = dist.new_group(backend="gloo")     # rank 0 works alone for a while: the others wait on the host, not in a spinning kernel
import global_configs
global_configs.VISUAL_DIM, global_configs.ACOUSTIC_DIM = 35, 74
from bert_multimodal_transformer_b200.bert import MAG_BertForSequenceClassification
from bert_multimodal_transformer_b200.parallel import shard_batch
from bert_multimodal_transformer_b200.trainer import MagBertTrainer
from oracle import inputs as gen, weights

LAYERS, GLOBAL_B, STEPS = 4, 16 * world, int(os.environ.get("DDP_STEPS", "1"))

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

def note(msg):
    print(f"[ddp_check rank {rank}] {msg}", flush=True)


model = make()
note("model built")
tr = MagBertTrainer(model, GLOBAL_B // world, 50, lr=1e-4, weight_decay=0.01, total_steps=0, use_cuda_graph=True)
note(f"trainer built (peer-memory optimizer: {getattr(tr, 'peer', None) is not None}, "
     f"multicast: {getattr(tr, 'peer_mc_grad', None) is not None})")
losses = [tr.train_step(*args(mine)) for _ in range(STEPS)]
torch.cuda.synchronize()
note(f"steps done, losses {losses}")
if getattr(tr, "peer", None) is not None:
    # peer-memory optimizer: the Adam moments are sharded (each rank updates the slice it owns, the rest stays zero)
    dist.all_reduce(tr.exp_avg)
    torch.cuda.synchronize()
    note(f"moments gathered, owned slice {tr.peer_slice}")
ok = True
if rank == 0:
    ref_model = make()
    ref = MagBertTrainer(ref_model, GLOBAL_B, 50, lr=1e-4, weight_decay=0.01, total_steps=0, use_cuda_graph=False,
                         world_size=1)
    ref_losses = [ref.train_step(*args(full)) for _ in range(STEPS)]
    torch.cuda.synchronize()
    note("single-process reference done")
    # Adam's update is invariant to the gradient scale, so compare the first moment after one step instead:
    # exp_avg = (1 - beta1) * (mean gradient over the GLOBAL batch); a sum-instead-of-mean or a slice missing from
    # the all-reduce shows up as a factor `world` / a zero block.
    a, b = tr.exp_avg, ref.exp_avg
    rel = ((a - b).norm() / b.norm()).item()
    lay = model._engine.layout
    worst = max((((lay.view(a, n) - lay.view(b, n)).norm() / (lay.view(b, n).norm() + 1e-20)).item(), n)
                for n in model._engine.order if lay.view(b, n).norm().item() > 1e-4 * b.norm().item())  # key biases have zero true gradient
    pa, pb = model._engine.flat, ref_model._engine.flat
    print(json.dumps({"world": world, "rel_first_moment_diff": rel, "worst_tensor": worst,
                      "rel_param_diff": ((pa - pb).norm() / pb.norm()).item(),
                      "rank0_local_losses": losses, "single_process_losses": ref_losses}))
    ok = rel < 1e-3 and worst[0] < 2e-2
# every rank must hold identical parameters after the all-reduced steps
dist.barrier(group=host_group)
note("checksum exchange")
chk = model._engine.flat.double().sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print("replica checksum spread", float(hi - lo))
    ok = ok and float(hi - lo) == 0.0
    print("DDP_CHECK", "PASS" if ok else "FAIL")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
note("releasing")
tr.release()
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if int(flag.item()) == 1 else 1)
