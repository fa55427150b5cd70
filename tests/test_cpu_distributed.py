"""CPU, world_size 2 over gloo: the data-parallel recipe of trainer.py / parallel.py (contiguous batch
shards, local MEAN loss, SUM all-reduce of the flat gradient, 1/world unscale) reproduces the gradient of
one process on the whole batch.  Local gradients come from the CPU oracle -- the product kernels need a
GPU; what is under test is the host-side sharding / reduction logic."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from bert_multimodal_transformer_b200 import parallel
    from bert_multimodal_transformer_b200.flat import FlatLayout, bert_param_order
    from oracle import inputs as gen, restatement as R, weights
    layers = 1
    sd = {k: v.requires_grad_(True) for k, v in weights.bert_state_dict(num_layers=layers, seed=1).items()}
    full = gen.bert_batch(4, 16, seed=3)
    mine = parallel.shard_batch(full, rank, world)
    assert mine["input_ids"].shape[0] == 2
    logits = R.mag_bert_forward(sd, mine["input_ids"], mine["visual"], mine["acoustic"], mine["attention_mask"],
                                mine["token_type_ids"], n_layer=layers)
    R.mse_loss(logits, mine["labels"]).backward()
    order = bert_param_order(layers)
    lay = FlatLayout([(n, tuple(sd[n].shape)) for n in order])
    flat = torch.zeros(lay.total)
    for n in order:
        lay.view(flat, n).copy_(sd[n].grad)
    parallel.allreduce_sum_(flat)
    flat.mul_(parallel.grad_unscale(world))
    if rank == 0:
        torch.save({n: lay.view(flat, n).clone() for n in order}, os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    from oracle import inputs as gen, restatement as R, weights
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    sd = {k: v.requires_grad_(True) for k, v in weights.bert_state_dict(num_layers=1, seed=1).items()}
    full = gen.bert_batch(4, 16, seed=3)
    logits = R.mag_bert_forward(sd, full["input_ids"], full["visual"], full["acoustic"], full["attention_mask"],
                                full["token_type_ids"], n_layer=1)
    R.mse_loss(logits, full["labels"]).backward()
    for k, g in got.items():
        ref = sd[k].grad
        assert (g - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), k


def test_shard_slice_contract():
    from bert_multimodal_transformer_b200.parallel import shard_slice
    assert shard_slice(128, 3, 8) == slice(48, 64)
    with pytest.raises(ValueError):
        shard_slice(50, 0, 8)


def _sparse_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bert_multimodal_transformer_b200 import parallel
    g = torch.Generator().manual_seed(100 + rank)
    vocab, H = 97, 8
    ids = torch.randint(0, vocab, (3, 11), generator=g)
    ids[0, :4] = 5                                              # duplicates inside a rank, and id 5 on both ranks
    contrib = torch.randn(ids.numel(), H, generator=g)
    dense = torch.zeros(vocab, H).index_add_(0, ids.reshape(-1), contrib)     # what the embedding backward leaves
    ref = dense.clone()
    dist.all_reduce(ref)                                        # dense all-reduce = the definition
    got = parallel.sparse_rows_allreduce_(dense.clone(), ids)
    ok = torch.allclose(got, ref, atol=1e-5)
    torch.save({"ok": ok, "err": (got - ref).abs().max().item(), "got": got}, os.path.join(out_dir, f"sparse{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sparse_embedding_gradient_exchange_equals_dense_allreduce(tmp_path):
    port = 31500 + os.getpid() % 2000
    mp.spawn(_sparse_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"sparse{r}.pt")) for r in range(2)]
    for r in res:
        assert r["ok"], r["err"]
    # replicas must not drift: every rank holds the same BITS (additions happen in rank order everywhere)
    assert torch.equal(res[0]["got"], res[1]["got"])


def test_peer_optimizer_ownership_tiles_every_bucket():
    """parallel.owned_piece (who updates what in the peer-memory optimizer step): for any bucket and world size the pieces
    are 256-aligned, disjoint, in rank order, and cover the bucket exactly -- an element with two owners would be updated
    twice, one with none never."""
    from bert_multimodal_transformer_b200.parallel import owned_piece
    for world in (1, 2, 3, 4, 7, 8):
        for begin, end in ((0, 256), (0, 23_835_648), (23_835_648, 109_483_776 // 256 * 256), (512, 512), (256, 256 * 6)):
            pieces = [owned_piece(begin, end, world, r) for r in range(world)]
            cursor = begin
            for b, e in pieces:
                assert b % 256 == 0 and e % 256 == 0 and b <= e
                assert b == min(cursor, end) and e <= end
                cursor = max(cursor, e)
            assert cursor == end
            sizes = [e - b for b, e in pieces]
            assert sizes == sorted(sizes, reverse=True)


def test_peer_step_schedule_equals_allreduce_then_adamw():
    """oracle/peer_step.py restates the schedule of the peer-memory optimizer step (ownership by bucket piece, early pass over
    the embedding rows nobody touched, late pass over the rest, parameters copied from the owner).  Over several steps with
    changing batches it must equal the textbook data-parallel step -- mean gradient on every rank, then transformers-3.0.2
    AdamW on the full vector -- and keep the replicas bit-identical.  Untouched rows really have zero gradient here, as in
    the embedding backward."""
    from oracle import peer_step as PS
    g = torch.Generator().manual_seed(5)
    rows, rl, extra = 37, 64, 256 * 9 + 0
    table_elems = rows * rl
    n = (table_elems + 255) // 256 * 256 + extra
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.999, 1e-6, 0.01
    decay = (torch.rand(n, generator=g) > 0.3).float()
    for world in (1, 2, 3, 8):
        p0 = torch.randn(n, generator=g)
        params = [p0.clone() for _ in range(world)]
        moments = [(torch.zeros(n), torch.zeros(n)) for _ in range(world)]
        ref_p, ref_m, ref_v = p0.clone(), torch.zeros(n), torch.zeros(n)
        for step in (1, 2, 3, 4):
            ids = [torch.randint(0, rows, (2, 5), generator=g) for _ in range(world)]
            grads = []
            for r in range(world):
                gr = torch.randn(n, generator=g) * 0.01
                hit = torch.zeros(rows, dtype=torch.bool)
                hit[ids[r].reshape(-1)] = True
                gr[:table_elems].view(rows, rl)[~hit] = 0.0            # nn.Embedding backward: only the batch's rows
                gr[table_elems:(table_elems + 255) // 256 * 256] = 0.0  # alignment padding carries nothing
                grads.append(gr)
            mean = sum(grads) / world
            PS.adamw_(ref_p, mean, ref_m, ref_v, decay, lr, b1, b2, eps, wd, step)
            PS.peer_step(params, grads, moments, decay, (0, rows, rl), ids, lr, b1, b2, eps, wd, step)
            for q in range(1, world):
                assert torch.equal(params[q], params[0])
            assert (params[0] - ref_p).abs().max().item() < 2e-6, (world, step)
            assert all(float(gr.abs().max()) == 0.0 for gr in grads)
        gathered_m = sum(m for m, _ in moments)                        # every element has exactly one owner
        assert (gathered_m - ref_m).abs().max().item() < 1e-7
