"""Host-side data-parallel plumbing: one process per GPU, batch sharded by
contiguous slices, ONE collective per step -- a sum all-reduce of the flat gradient buffer; the mean
over ranks is folded into the AdamW kernel's grad_unscale.  torch.distributed is the transport (NCCL over
NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous B/world slice of the global batch for `rank` (config 3: 128 -> 16 per GPU on 8 GPUs)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    n = next(iter(batch.values())).shape[0]
    s = shard_slice(n, rank, world)
    return {k: v[s] for k, v in batch.items()}


def allreduce_sum_(flat_grads: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of the flat gradient buffer (a no-op for a single process)."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return flat_grads


def grad_unscale(world: int) -> float:
    """Each rank's loss is a MEAN over its local batch (multimodal_driver.py:371-373), so the global-batch
    gradient is the mean of the per-rank gradients: sum-all-reduce, then multiply by 1/world."""
    return 1.0 / float(world)


def sparse_rows_allreduce_(table_grad: torch.Tensor, ids: torch.Tensor, group=None) -> torch.Tensor:
    """Sum over ranks of an embedding-table gradient that is non-zero only in the rows `ids` touched on each rank
    (word embeddings: <= B*S of 30522 rows, 7 MB per rank instead of a 94 MB dense all-reduce).

    Every rank contributes (ids, rows): its local row sum at the FIRST occurrence of each id in its list, zeros at the
    repeats.  Both are all-gathered; each rank clears the rows it touched and adds the contributions rank by rank.
    The result is bit-identical on every rank -- replicas must not drift: the order of additions into a row is the
    rank order everywhere, and the only duplicates inside one index_add_ call add exact zeros, so the (unordered)
    atomics of index_add_ cannot change a bit.  Fixed shapes, no host synchronisation (no unique / nonzero): the whole
    exchange can be captured in a CUDA graph.  In place; returns `table_grad`.

    table_grad: [vocab, H] fp32 (a view of the flat gradient buffer); ids: this rank's token ids, any shape."""
    rank, world = world_info(group)
    if world == 1:
        return table_grad
    flat_ids = ids.reshape(-1).to(torch.int64)
    n = flat_ids.numel()
    pos = torch.arange(n, device=flat_ids.device)
    same = flat_ids[:, None] == flat_ids[None, :]
    repeat = (same & (pos[None, :] < pos[:, None])).any(dim=1)         # an earlier position holds the same id
    rows = table_grad.index_select(0, flat_ids)
    rows = torch.where(repeat[:, None], torch.zeros((), dtype=rows.dtype, device=rows.device), rows)
    all_ids = [torch.empty_like(flat_ids) for _ in range(world)]
    all_rows = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(all_ids, flat_ids, group=group)
    dist.all_gather(all_rows, rows, group=group)
    table_grad.index_fill_(0, flat_ids, 0.0)
    for r in range(world):
        table_grad.index_add_(0, all_ids[r], all_rows[r])
    return table_grad


def owned_piece(begin: int, end: int, world: int, rank: int):
    """The 256-aligned 1 / world of bucket [begin, end) that `rank` updates in the peer-memory optimizer step
    (csrc/peer_optim.cu): equal chunks rounded up to 256 elements, the last ranks get what is left (possibly nothing).
    The pieces of all ranks tile the bucket exactly, so every element has one owner."""
    assert begin % 256 == 0 and end % 256 == 0 and begin <= end and 0 <= rank < world
    chunk = -(-(end - begin) // (world * 256)) * 256
    b = min(end, begin + rank * chunk)
    return b, min(end, b + chunk)


class PeerMemory:
    """Symmetric (peer-mapped) device buffers of the ranks of ONE node, for the NVLink-native optimizer step
    (csrc/peer_optim.cu): torch's symmetric-memory allocator + rendezvous give every rank the addresses of every other
    rank's copy of a buffer; the kernels of this package then read / write them directly over NVLink (no collective
    library on that path).  PyTorch is plumbing here: allocation and the one-time exchange of the IPC handles."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = world_info(group)
        if self.world > 8:
            raise ValueError("peer-memory optimizer: one NVLink domain of at most 8 ranks")
        self._keep = []
        flags = self.empty(64, torch.int32)
        flags.zero_()
        self.flag_ptrs = self.pointers(flags)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)          # every rank has zeroed its flags before anyone signals

    def empty(self, numel, dtype):
        t = self.symm.empty(numel, dtype=dtype, device=self.device)
        h = self.symm.rendezvous(t, self.group)
        self._keep.append((t, h))
        return t

    def pointers(self, tensor):
        """ctypes array of the device pointers of `tensor` (allocated by self.empty) on every rank, in rank order."""
        import ctypes
        for t, h in self._keep:
            if t.data_ptr() == tensor.data_ptr():
                ptrs = [int(p) for p in h.buffer_ptrs]
                assert len(ptrs) == self.world and ptrs[self.rank] == tensor.data_ptr()
                arr = (ctypes.c_void_p * 8)(*([ctypes.c_void_p(p) for p in ptrs] + [None] * (8 - len(ptrs))))
                return arr
        raise KeyError("tensor was not allocated through this PeerMemory")

    def multicast_pointer(self, tensor):
        """NVSwitch multicast address of `tensor` (allocated by self.empty), or 0 when the node / allocator has none."""
        for t, h in self._keep:
            if t.data_ptr() == tensor.data_ptr():
                try:
                    return int(h.multicast_ptr or 0)
                except Exception:        # pragma: no cover  (older allocator without the attribute)
                    return 0
        raise KeyError("tensor was not allocated through this PeerMemory")

    def owned_slice(self, total):
        """256-aligned [begin, end) of a flat buffer of `total` elements (a multiple of 256) that this rank updates."""
        chunks = total // 256
        b = (chunks * self.rank) // self.world * 256
        e = (chunks * (self.rank + 1)) // self.world * 256
        return b, e
