"""torchrun entry: times the pieces of the peer-memory optimizer step (csrc/peer_optim.cu) on N ranks with CUDA events:
forward+loss, backward, barrier, reduce+AdamW+broadcast kernel, barrier, clean-up -- eager launches, max over ranks.

    MAGB200_PEER_ADAMW=1 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MAGB200_PEER_ADAMW", "1")
os.environ["MAGB200_PEER_LAZY_CLEAR"] = "0"      # the clean-up below clears the gradients (timed separately from the recast)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from bert_multimodal_transformer_b200 import _lib as L
from bert_multimodal_transformer_b200.trainer import MagBertTrainer
from oracle import inputs as gen

wl = bench.WORKLOADS["mosi48"]
dev = torch.device("cuda", local)
model = bench.make_model(torch, wl, dev)
B, S = wl["per_gpu"](world), wl["seq"]
tr = MagBertTrainer(model, B, S, lr=1e-5, weight_decay=0.01, warmup_steps=10, total_steps=1000, use_cuda_graph=False, seed=rank)
assert tr.peer is not None, "peer-memory optimizer not active"
b = gen.bert_batch(B, S, wl["V"], wl["A"], seed=77 + rank)
tr._stage(b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"], b["labels"])
tr.upload_staged()
eng, lib, i = tr.eng, L.load(), tr.dev_in
total = eng.layout.total
lo0 = eng.layer_grad_range(0, 1)[0]
st = L.current_stream()
own = sum(e - b for b, e in (tr._peer_piece(0, lo0), tr._peer_piece(lo0, total)))


def pieces():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    ev[0].record()
    eng.forward(tr.train_dims, tr.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], i["visual"], i["acoustic"], tr.logits)
    L.check(lib.magb200_mse_loss(L.ptr(tr.logits), L.ptr(i["labels"]), B * eng.num_labels, eng.grad_scale, L.ptr(tr.loss),
                                 L.ptr(tr.dlogits), st))
    ev[1].record()
    eng.backward(tr.train_dims, tr.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], tr.dlogits, tr.grads)
    ev[2].record()
    tr._peer_barrier(st)
    ev[3].record()
    tr._peer_adamw(tr._peer_piece(lo0, total), (0, 0), st)
    tr._peer_adamw(tr._peer_piece(0, lo0), (0, 0), st)
    ev[4].record()
    tr._peer_barrier(st)
    ev[5].record()
    tr._peer_adamw((0, 0), (lo0, total), st, keep=tr._peer_piece(lo0, total))
    tr._peer_adamw((0, 0), (0, lo0), st, keep=tr._peer_piece(0, lo0))
    ev[6].record()
    eng.after_optimizer_step()
    ev[7].record()
    torch.cuda.synchronize()
    return [ev[k].elapsed_time(ev[k + 1]) for k in range(7)]


names = ["forward+loss", "backward", "barrier 1 (skew between ranks)", "reduce + AdamW + broadcast", "barrier 2", "clean-up (recast + clear)",
         "MAG repack"]
for _ in range(3):
    pieces()
acc = torch.zeros(7, device="cuda")
R = 10
for _ in range(R):
    dist.barrier()
    acc += torch.tensor(pieces(), device="cuda")
acc /= R
mx = acc.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    own = own * 4 / 1e6
    print(f"world {world}: owned slice {own:.1f} MB of {total * 4 / 1e6:.1f} MB; "
          f"NVSwitch multicast {'on' if tr.peer_mc_grad is not None else 'off'}; "
          f"eager launches, ms (max over ranks / rank 0):")
    for n, a, m in zip(names, acc.tolist(), mx.tolist()):
        print(f"  {n:34s} {m:8.3f} {a:8.3f}")
    k = mx[3].item()
    # each direction of a GPU's NVLink carries the gradients it pulls (in) / serves (out) AND the parameters it receives
    # (in) / stores (out): 2 (W-1)/W of the model per direction
    print(f"  kernel: {2 * (world - 1) * own / k:.1f} GB/s per NVLink direction per GPU "
          f"(gradient pulls + parameter stores), {(own * 7 + own * 0.5) / k:.1f} GB/s local HBM")
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
