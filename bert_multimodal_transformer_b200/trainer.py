"""Training / inference steps of MAG-BERT as single CUDA-graph replays.

The reference's hot loop (multimodal_driver.py:354-388) is, per batch: H2D copy of six small tensors,
forward, MSE loss, backward, AdamW step, LR-schedule step, zero_grad, `loss.item()`.  At bs=48, S=50
the whole step is ~1.3 TFLOP spread over ~300 kernels, so launch latency dominates unless the step
is replayed as one graph.  `MagBertTrainer.train_step` does exactly that:

    pinned host batch --H2D--> [ forward | mse | backward | gradient exchange | AdamW+zero_grad+fp16 refresh ] --D2H--> loss

Everything that changes from step to step (dropout seed, learning rate, Adam bias correction) is read
by the kernels from a small device control block that is part of the H2D copy, so one captured graph
serves every step.  Data parallelism: one process per GPU, each rank runs this on
its own slice of the batch.  Gradient exchange and update at world > 1 are, by default, ONE kernel per bucket over
NVLink peer memory (reduce-scatter + AdamW on the owned 1/world + all-gather of the parameters, csrc/peer_optim.cu,
_peer_step below); the NCCL all-reduce schedule (_backward_overlapped) is the fallback.  DESIGN.md section 6.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
from typing import Optional

import torch

from . import _lib as L
from .engine import BertEngine  # noqa: F401  (re-exported: the trainer drives a BertEngine or an XlnetEngine)
from .parallel import allreduce_sum_, grad_unscale, sparse_rows_allreduce_


def linear_schedule(step: int, warmup: int, total: int) -> float:
    """transformers.get_linear_schedule_with_warmup lambda (multimodal_driver.py:346-350)."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    return max(0.0, float(total - step) / float(max(1, total - warmup)))


class MagBertTrainer:
    def __init__(self, model, batch_size: int, seq_len: int, lr: float = 1e-5, weight_decay: float = 0.01,
                 betas=(0.9, 0.999), eps: float = 1e-6, warmup_steps: int = 0, total_steps: int = 0,
                 use_cuda_graph: bool = True, process_group=None, seed: int = 0, world_size: Optional[int] = None,
                 overlap_allreduce: bool = True):
        model._ensure_bound()
        self.model = model
        self.eng = model._engine      # BertEngine or XlnetEngine
        eng = self.eng
        self.B, self.S = batch_size, seq_len
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        self.pg = process_group
        self.world = 1
        if world_size is not None:
            self.world = int(world_size)          # explicit override (1 = no gradient all-reduce)
        elif process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        dev = eng.flat.device
        self.dev = dev
        n = eng.layout.total
        # Data parallel on one NVLink node: parameters and gradients move to symmetric (peer-mapped) memory and the
        # optimizer step becomes reduce-scatter + AdamW + all-gather in ONE kernel over peer pointers
        # (csrc/peer_optim.cu) -- measured 85.2 k vs 74.2 k samples/s at 8 GPUs, 24.0 k vs 22.2 k at 2 (profiles/r02_n8_*,
        # r02_n2_*); MAGB200_PEER_ADAMW=0 (or an unavailable symmetric-memory allocator) keeps the NCCL all-reduce schedule.
        self.peer = None
        if self.world > 1 and os.environ.get("MAGB200_PEER_ADAMW", "1") == "1" and hasattr(eng, "allocator") and dev.type == "cuda":
            try:
                from .parallel import PeerMemory
                pm = PeerMemory(dev, process_group)
                eng.allocator = pm.empty
                eng.flat = None
                model._ensure_bound()                      # re-binds the parameters as views of a symmetric flat buffer
                self.peer = pm
            except Exception as e:                         # pragma: no cover  (depends on the node)
                import warnings
                warnings.warn(f"peer-memory optimizer unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce schedule")
                eng.allocator = None
                if eng.flat is None:
                    model._ensure_bound()
                self.peer = None
        if self.peer is not None:
            self.grads = self.peer.empty(n, torch.float32)
            self.grads.zero_()
            self.peer_grad_ptrs = self.peer.pointers(self.grads)
            self.peer_param_ptrs = self.peer.pointers(eng.flat)
            self.peer_slice = self.peer.owned_slice(n)
            # NVSwitch multicast addresses of the two buffers (0 when the allocator has no multicast object for them)
            # NVLink bytes per direction and GPU: peer loads / stores 2 (W-1)/W of the model, multicast (W+1)/W (own data
            # loops through the switch): worse at 2 GPUs (kernel 1.17 vs 0.69 ms), better at 8 (0.99 vs 1.19 ms, step 4.33
            # vs 4.51 ms; profiles/r02_peer_probe_n8*.txt).  "auto" (default) = from 6 GPUs up, where it has been run;
            # "1" / "0" force it on / off.  Without a multicast object for the buffers the peer-pointer form is used.
            mc_env = os.environ.get("MAGB200_PEER_MULTICAST", "auto")
            use_mc = mc_env == "1" or (mc_env == "auto" and self.world >= 6)
            mg, mp = (self.peer.multicast_pointer(self.grads), self.peer.multicast_pointer(eng.flat)) if use_mc else (0, 0)
            self.peer_mc_grad = ctypes.c_void_p(mg) if mg and mp else None
            self.peer_mc_param = ctypes.c_void_p(mp) if mg and mp else None
            self.peer_blocks = int(os.environ.get("MAGB200_PEER_ADAMW_BLOCKS", "0"))
            self.peer_overlap = os.environ.get("MAGB200_PEER_OVERLAP", "0") == "1"
            self.peer_lazy_clear = os.environ.get("MAGB200_PEER_LAZY_CLEAR", "0") == "1"   # measured slower (4.20 vs 4.16 ms at 2 GPUs)
            self.peer_early_blocks = int(os.environ.get("MAGB200_PEER_EARLY_BLOCKS", "32"))
            # word-embedding matrix by touched rows (csrc/peer_optim.cu RowTable): marks in symmetric memory
            self.peer_sparse_table = os.environ.get("MAGB200_PEER_SPARSE_TABLE", "1") == "1" and hasattr(eng, "vocab")
            if self.peer_sparse_table:
                self.peer_marks = self.peer.empty(eng.vocab, torch.int32)
                self.peer_marks.zero_()
                self.peer_mark_ptrs = self.peer.pointers(self.peer_marks)
                self.peer_mark_counter = torch.zeros(1, dtype=torch.int32, device=dev)
                self.peer_row_any = torch.zeros(eng.vocab, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            torch.distributed.barrier(group=process_group)
        else:
            self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.decay = eng.layout.decay_chunks().to(dev)
        self.step_count = 0
        self._gen = torch.Generator().manual_seed(seed)
        B, S, V, A, Lb = batch_size, seq_len, eng.visual_dim, eng.acoustic_dim, eng.num_labels
        # static device inputs + pinned host staging
        spec = dict(input_ids=((B, S), torch.int64), token_type_ids=((B, S), torch.int64),
                    attention_mask=((B, S), torch.int64), visual=((B, S, V), torch.float32),
                    acoustic=((B, S, A), torch.float32), labels=((B * Lb,), torch.float32))
        self.dev_in = {k: torch.zeros(s, dtype=t, device=dev) for k, (s, t) in spec.items()}
        # two pinned staging slots: while the GPU runs step k out of slot k%2, the host fills slot (k+1)%2
        self.NSLOT = 2
        self.slots = [{k: torch.zeros(s, dtype=t).pin_memory() for k, (s, t) in spec.items()} for _ in range(self.NSLOT)]
        self.slots_np = [{k: v.numpy() for k, v in sl.items()} for sl in self.slots]
        self.slot = 0
        self.host_in = self.slots[0]
        # per-step control block (device words the kernels read): dropout seed, lr, Adam step size.  It is copied
        # eagerly (outside the captured graphs) from a ring of pinned host blocks, each guarded by an event recorded
        # right after its copy: steps can be queued back to back without a later step's controls overwriting a block
        # an earlier, still queued copy has not read yet
        self.NCTRL = 8
        self.host_seed = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(self.NCTRL)]
        self.dev_seed = torch.zeros(1, dtype=torch.int64, device=dev)
        self.host_hyper = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(self.NCTRL)]
        self.dev_hyper = torch.zeros(2, dtype=torch.float32, device=dev)
        self.ctrl_ev = [None] * self.NCTRL
        self.ctrl_i = 0
        self.logits = torch.zeros(B, Lb, dtype=torch.float32, device=dev)
        self.dlogits = torch.zeros(B, Lb, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.host_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(self.NSLOT)]
        self.host_logits = torch.zeros(B, Lb, dtype=torch.float32).pin_memory()
        self.done = [torch.cuda.Event() for _ in range(self.NSLOT)]
        self.train_dims = eng.dims(B, S, True, 0)
        self.train_dims.seed_dev = self.dev_seed.data_ptr()
        self.eval_dims = eng.dims(B, S, False, 0)
        lib = L.load()
        self.ws = torch.empty(eng.workspace_bytes(self.train_dims), dtype=torch.uint8, device=dev)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.slots[0].values()) + 8 + 8
        self.d2h_bytes = 4
        self.use_graph = use_cuda_graph
        self._train_graph = [None] * self.NSLOT     # one captured graph per staging slot (fixed memcpy sources)
        self._eval_graph = [None] * self.NSLOT
        self._train_graph_resident = None
        self._eval_graph_resident = None
        self.stream = torch.cuda.Stream(device=dev)
        self.comm_stream = torch.cuda.Stream(device=dev, priority=-1)   # its kernels go first whenever an SM has room
        self.overlap = overlap_allreduce
        self.sparse_embedding_allreduce = os.environ.get("MAGB200_SPARSE_EMB_ALLREDUCE", "1") == "1" and \
            hasattr(eng, "vocab") and hasattr(eng, "backward_part")
        self.bucket_adamw = os.environ.get("MAGB200_BUCKET_ADAMW", "1") == "1"
        self.fused_optimizer = os.environ.get("MAGB200_FUSED_OPTIMIZER", "0") == "1"   # measured: no gain (L2-bound GEMMs and the HBM-bound update slow each other), see DESIGN.md
        eng.refresh_operands()
        self.kernel_launches_per_train_step = None

    # ------------------------------------------------------------------ device-side step bodies
    def _h2d(self, slot, with_labels=True):
        for k, d in self.dev_in.items():
            if k == "labels" and not with_labels:
                continue
            d.copy_(self.slots[slot][k], non_blocking=True)

    def _train_body(self, transfers=True, slot=0):
        eng, lib = self.eng, L.load()
        st = L.current_stream()
        if transfers:
            self._h2d(slot)
        i = self.dev_in
        if self.peer is not None and self.peer_lazy_clear:
            self._peer_clear_gradients_early()
        eng.forward(self.train_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], i["visual"],
                    i["acoustic"], self.logits)
        L.check(lib.magb200_mse_loss(L.ptr(self.logits), L.ptr(i["labels"]), self.B * eng.num_labels, eng.grad_scale,
                                     L.ptr(self.loss), L.ptr(self.dlogits), st))
        if self.world == 1 and self.fused_optimizer and hasattr(eng, "backward_adamw"):
            # backward + AdamW as one schedule: per-layer updates overlap the backward of the layers below
            eng.backward_adamw(self.train_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"],
                               self.dlogits, self.grads, self.exp_avg, self.exp_avg_sq, self.decay, self.lr, self.betas,
                               self.eps, self.wd, 1, grad_unscale(self.world), self.dev_hyper)
            eng.after_optimizer_step()
            if transfers:
                self.host_loss[slot].copy_(self.loss, non_blocking=True)
            return
        if self.peer is not None:
            if self.overlap and self.peer_overlap and hasattr(eng, "backward_part"):
                self._peer_step_overlapped(i)
            else:
                self._peer_step(i, st)
            if transfers:
                self.host_loss[slot].copy_(self.loss, non_blocking=True)
            return
        if self.world > 1 and self.overlap and hasattr(eng, "backward_part"):
            self._backward_overlapped(i)
            if self.bucket_adamw:            # every slice was updated behind its all-reduce
                eng.after_optimizer_step()
                if transfers:
                    self.host_loss[slot].copy_(self.loss, non_blocking=True)
                return
        else:
            eng.backward(self.train_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], self.dlogits,
                         self.grads)
            if self.world > 1:
                allreduce_sum_(self.grads, self.pg)
        L.check(lib.magb200_adamw_step(L.ptr(eng.flat), L.ptr(self.grads), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                       L.ptr(eng.flat16), L.ptr(self.decay), eng.layout.total, self.lr, self.betas[0],
                                       self.betas[1], self.eps, self.wd, 1, grad_unscale(self.world), 1, L.ptr(self.dev_hyper),
                                       st))
        eng.after_optimizer_step()
        if transfers:
            self.host_loss[slot].copy_(self.loss, non_blocking=True)

    def _peer_adamw(self, owned, clean, stream_handle, row_mode=0, blocks=None, keep=(0, 0)):
        """csrc/peer_optim.cu on one bucket: reduce + AdamW + broadcast of `owned`, clean-up (recast the fp16 mirror, and
        clear the gradients unless they are cleared at the start of the next step) of `clean`; both (begin, end) in
        elements of the flat buffers, 256-aligned, possibly empty.  row_mode 1 / 2: the word-embedding matrix by
        untouched / touched rows (magb200.h)."""
        eng, lib, pm = self.eng, L.load(), self.peer
        table = (L.ptr(self.peer_row_any), row_mode, int(eng.offsets[L.OFF_WORD]), eng.vocab, eng.hidden) if row_mode else \
            (None, 0, 0, 0, 0)
        L.check(lib.magb200_adamw_peer_step(self.peer_grad_ptrs, self.peer_param_ptrs, pm.world, pm.rank, L.ptr(self.exp_avg),
                                            L.ptr(self.exp_avg_sq), L.ptr(eng.flat16), L.ptr(self.decay), owned[0], owned[1],
                                            clean[0], clean[1], self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 1,
                                            grad_unscale(self.world), L.ptr(self.dev_hyper),
                                            self.peer_blocks if blocks is None else blocks, *table,
                                            1 if self.peer_lazy_clear else 3, keep[0], keep[1], self.peer_mc_grad,
                                            self.peer_mc_param, stream_handle))

    def _peer_barrier(self, stream_handle):
        pm = self.peer
        L.check(L.load().magb200_peer_barrier(pm.flag_ptrs, pm.world, pm.rank, L.ptr(pm.epoch), stream_handle))

    def _peer_piece(self, s, e):
        """The 256-aligned 1 / world of bucket [s, e) this rank owns (every rank owns a piece of every bucket)."""
        from .parallel import owned_piece
        return owned_piece(s, e, self.peer.world, self.peer.rank)

    def _peer_clear_gradients_early(self):
        """MAGB200_PEER_LAZY_CLEAR: the gradients of the previous step are cleared on the communication stream while the
        forward pass runs (it does not touch them) instead of in the clean-up after the optimizer step, where every byte is
        exposed.  Safe: the previous step ended with a barrier after which no rank reads this rank's gradients."""
        main, side = torch.cuda.current_stream(self.dev), self.comm_stream
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            self.grads.zero_()
            self._peer_clear_done = torch.cuda.Event()
            self._peer_clear_done.record(side)

    def _peer_wait_cleared(self):
        """backward may write gradients only after the early clearing has finished"""
        if self.peer_lazy_clear:
            torch.cuda.current_stream(self.dev).wait_event(self._peer_clear_done)

    def _peer_early_table_pass(self, i, lo0):
        """Word-embedding rows no rank touches in this step: their update needs no gradient (it is exactly zero), so it is
        issued before backward on the communication stream and hides under it.  All ranks have finished forward once the
        barrier is passed, and neither forward nor backward reads an untouched row."""
        eng, lib, pm = self.eng, L.load(), self.peer
        main, side = torch.cuda.current_stream(self.dev), self.comm_stream
        side_h = ctypes.c_void_p(side.cuda_stream)
        L.check(lib.magb200_peer_mark_rows(L.ptr(i["input_ids"]), self.B * self.S, eng.vocab, L.ptr(self.peer_marks),
                                           L.ptr(self.peer_mark_counter), L.current_stream()))
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            self._peer_barrier(side_h)
            L.check(lib.magb200_peer_combine_marks(self.peer_mark_ptrs, pm.world, L.ptr(self.peer_mark_counter), eng.vocab,
                                                   L.ptr(self.peer_row_any), side_h))
            self._peer_adamw(self._peer_piece(0, lo0), (0, 0), side_h, row_mode=1, blocks=self.peer_early_blocks)

    def _peer_step(self, i, st):
        """backward -> [barrier: all gradients written] -> reduce-scatter + AdamW + all-gather over NVLink peer memory for
        the pieces this rank owns -> [barrier: all parameters landed, all gradients read] -> clean-up (fp16 operand mirror;
        gradients).  No collective library call.  The untouched rows of the word-embedding matrix go first, under backward."""
        eng = self.eng
        total = eng.layout.total
        lo0 = eng.layer_grad_range(0, 1)[0] if hasattr(eng, "layer_grad_range") else 0
        sparse = self.peer_sparse_table
        main, side = torch.cuda.current_stream(self.dev), self.comm_stream
        if sparse:
            self._peer_early_table_pass(i, lo0)
        self._peer_wait_cleared()
        eng.backward(self.train_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], self.dlogits, self.grads)
        main.wait_stream(side)             # early table pass / early clearing of the gradients
        mode = 2 if sparse else 0
        self._peer_barrier(st)
        self._peer_adamw(self._peer_piece(lo0, total), (0, 0), st)
        self._peer_adamw(self._peer_piece(0, lo0), (0, 0), st, row_mode=mode)
        self._peer_barrier(st)
        self._peer_adamw((0, 0), (lo0, total), st, keep=self._peer_piece(lo0, total))
        self._peer_adamw((0, 0), (0, lo0), st, row_mode=mode, keep=self._peer_piece(0, lo0))
        eng.after_optimizer_step()

    def _peer_step_overlapped(self, i):
        """The peer-memory optimizer step bucket by bucket behind the backward pass (same pieces as _backward_overlapped;
        MAGB200_PEER_OVERLAP=1).  Measured SLOWER than _peer_step at 2 GPUs (4.24 vs 4.16 ms / step): the reduce kernels
        take SM slots, L2 and HBM bandwidth from the backward GEMMs, which are sensitive to all three.  Kept for A/B.

        communication stream, per bucket (a layer group; at the end embeddings and pooler / MAG / classifier):
            wait for the local backward of the bucket -> barrier (every rank has written the bucket's gradients, and has
            finished the previous bucket's kernel) -> one kernel: reduce + AdamW + broadcast of the 1/world of the bucket
            this rank owns, plus the clean-up of the PREVIOUS bucket (clear its gradients, recast the parameters the
            peers stored into it to fp16)
        then a last barrier and the clean-up of the last buckets.  The main stream only runs backward; what remains after
        it is the bottom bucket."""
        eng = self.eng
        main = torch.cuda.current_stream(self.dev)
        side = self.comm_stream
        side_h = ctypes.c_void_p(side.cuda_stream)
        Ln = eng.layers
        ng = max(1, min(Ln, int(os.environ.get("MAGB200_ALLREDUCE_GROUPS", "12"))))
        cuts = sorted({(k * Ln) // ng for k in range(ng + 1)}, reverse=True)
        groups = [(cuts[k + 1], cuts[k]) for k in range(len(cuts) - 1)]
        args = (self.train_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], self.dlogits,
                self.grads)
        total = eng.layout.total
        tail0 = int(eng.offsets[L.OFF_POOL_W])
        piece = self._peer_piece
        lo0 = eng.layer_grad_range(groups[-1][0], groups[-1][1])[0]
        sparse = self.peer_sparse_table
        if sparse:
            self._peer_early_table_pass(i, lo0)
        self._peer_wait_cleared()
        prev = (0, 0)
        for gi, (lo, hi) in enumerate(groups):
            last = gi == len(groups) - 1
            eng.backward_part(*args, hi - 1, lo, (1 if gi == 0 else 0) | (2 if last else 0))
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            s, e = eng.layer_grad_range(lo, hi)
            with torch.cuda.stream(side):
                self._peer_barrier(side_h)
                self._peer_adamw(piece(s, e), prev, side_h, keep=piece(*prev) if prev[1] > prev[0] else (0, 0))
            prev = (s, e)
        mode = 2 if sparse else 0
        with torch.cuda.stream(side):
            self._peer_barrier(side_h)
            self._peer_adamw(piece(tail0, total), prev, side_h, keep=piece(*prev) if prev[1] > prev[0] else (0, 0))
            self._peer_adamw(piece(0, lo0), (0, 0), side_h, row_mode=mode)
            self._peer_barrier(side_h)
            self._peer_adamw((0, 0), (tail0, total), side_h, keep=piece(tail0, total))
            self._peer_adamw((0, 0), (0, lo0), side_h, row_mode=mode, keep=piece(0, lo0))
        main.wait_stream(side)
        eng.after_optimizer_step()

    def _adamw_range(self, s0, s1, stream_handle):
        """AdamW over elements [s0, s1) of the flat buffers (256-aligned slice) on the given stream."""
        if s1 <= s0:
            return
        eng, lib = self.eng, L.load()
        assert s0 % 256 == 0 and s1 % 256 == 0, (s0, s1)
        f = lambda t, es: ctypes.c_void_p(t.data_ptr() + es * s0)   # noqa: E731
        L.check(lib.magb200_adamw_step(f(eng.flat, 4), f(self.grads, 4), f(self.exp_avg, 4), f(self.exp_avg_sq, 4),
                                       f(eng.flat16, 2), ctypes.c_void_p(self.decay.data_ptr() + s0 // 256), s1 - s0, self.lr,
                                       self.betas[0], self.betas[1], self.eps, self.wd, 1, grad_unscale(self.world), 1,
                                       L.ptr(self.dev_hyper), stream_handle))

    def _backward_overlapped(self, i):
        """Data-parallel backward.  The order of the flat buffers is
        [embeddings | layer 0 .. L-1 | pooler, MAG, classifier], so a group of layers is one contiguous slice.

        Backward runs in pieces (MAGB200_ALLREDUCE_GROUPS layer groups, default 12 = one per layer); as soon as a piece
        has finished, its gradient slice is all-reduced on the communication stream AND, right behind it on the same
        stream, updated by AdamW (MAGB200_BUCKET_ADAMW=0 turns that off) while the main stream continues with the next
        piece: the NVLink transfer and the HBM-bound update both hide under the tensor-core-bound backward of the layers
        below.  What is left when backward ends is only the bottom piece: the last layer group, the word-embedding
        gradient -- exchanged as (token ids, rows) instead of a dense 94 MB all-reduce, it is non-zero in at most B*S
        of its 30522 rows (MAGB200_SPARSE_EMB_ALLREDUCE=0 for the dense form) -- the small position / type / LayerNorm
        tensors, and pooler / MAG / classifier."""
        eng = self.eng
        main = torch.cuda.current_stream(self.dev)
        side = self.comm_stream
        side_h = ctypes.c_void_p(side.cuda_stream)
        Ln = eng.layers
        ng = max(1, min(Ln, int(os.environ.get("MAGB200_ALLREDUCE_GROUPS", "12"))))
        cuts = sorted({(k * Ln) // ng for k in range(ng + 1)}, reverse=True)  # e.g. 12, 10, 8, 6, 4, 2, 0
        groups = [(cuts[k + 1], cuts[k]) for k in range(len(cuts) - 1)]       # [(8,12), (4,8), (0,4)]
        args = (self.train_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], self.dlogits,
                self.grads)
        total = eng.layout.total
        tail0 = int(eng.offsets[L.OFF_POOL_W])
        for gi, (lo, hi) in enumerate(groups):
            last = gi == len(groups) - 1
            eng.backward_part(*args, hi - 1, lo, (1 if gi == 0 else 0) | (2 if last else 0))
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            s, e = eng.layer_grad_range(lo, hi)
            with torch.cuda.stream(side):
                allreduce_sum_(self.grads[s:e], self.pg)
                if self.bucket_adamw:
                    self._adamw_range(s, e, side_h)
        # the bottom of the network (all produced by the last piece): embeddings in front, pooler / MAG / classifier behind
        lo0 = eng.layer_grad_range(groups[-1][0], groups[-1][1])[0]
        with torch.cuda.stream(side):
            if self.sparse_embedding_allreduce:
                w0, w1 = int(eng.offsets[L.OFF_WORD]), int(eng.offsets[L.OFF_POS])
                table = self.grads[w0:w0 + eng.vocab * eng.hidden].view(eng.vocab, eng.hidden)
                sparse_rows_allreduce_(table, i["input_ids"], self.pg)
                allreduce_sum_(self.grads[w1:lo0], self.pg)
            else:
                allreduce_sum_(self.grads[:lo0], self.pg)
            allreduce_sum_(self.grads[tail0:], self.pg)
            if self.bucket_adamw:
                self._adamw_range(0, lo0, side_h)
                self._adamw_range(tail0, total, side_h)
        main.wait_stream(side)

    def release(self):
        """Drop the captured graphs (they hold NCCL work when world > 1) -- call before destroying the process group."""
        self._train_graph = [None] * self.NSLOT
        self._eval_graph = [None] * self.NSLOT
        self._train_graph_resident = None
        self._eval_graph_resident = None
        torch.cuda.synchronize(self.dev)

    def _eval_body(self, transfers=True, slot=0):
        eng = self.eng
        if transfers:
            self._h2d(slot, with_labels=False)
        i = self.dev_in
        eng.forward(self.eval_dims, self.ws, i["input_ids"], i["token_type_ids"], i["attention_mask"], i["visual"],
                    i["acoustic"], self.logits)
        if transfers:
            self.host_logits.copy_(self.logits, non_blocking=True)

    def _capture(self, body):
        # warm up on the side stream (descriptor caches, func attributes), then capture
        s = self.stream
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            body()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body()
        return g

    # ------------------------------------------------------------------ host API
    def _stage(self, input_ids, visual, acoustic, attention_mask, token_type_ids, labels=None):
        """Host batch -> the current pinned staging slot.  Plain single-threaded memcpy through numpy
        (torch.Tensor.copy_ spins up the intra-op thread pool, which costs milliseconds on a busy host)."""
        self.done[self.slot].synchronize()          # the previous step that read this slot has finished
        h = self.slots_np[self.slot]

        def put(name, t, dtype):
            if torch.is_tensor(t):
                t = t.detach()
                if t.is_cuda:
                    t = t.cpu()
                t = t.to(dtype).contiguous().numpy()
            np.copyto(h[name], np.asarray(t).reshape(h[name].shape), casting="unsafe")

        put("input_ids", input_ids, torch.int64)
        put("visual", visual, torch.float32)
        put("acoustic", acoustic, torch.float32)
        if attention_mask is None:
            h["attention_mask"].fill(1)
        else:
            put("attention_mask", attention_mask, torch.int64)
        if token_type_ids is None:
            h["token_type_ids"].fill(0)
        else:
            put("token_type_ids", token_type_ids, torch.int64)
        if labels is not None:
            put("labels", labels, torch.float32)

    def upload_staged(self):
        """Copy the staged (pinned) batch to its device buffers once; later steps launched with
        transfers=False then run on that HBM-resident batch."""
        with torch.cuda.stream(self.stream):
            self._h2d(self.slot)
        self.stream.synchronize()

    def set_step_controls(self):
        """Per-step control block: dropout seed, scheduled LR, Adam bias correction -> device words, enqueued on the
        trainer's stream ahead of the step that reads them."""
        t = self.step_count
        lam = linear_schedule(t, self.warmup_steps, self.total_steps) if self.total_steps > 0 else 1.0
        lr = self.lr * lam
        b1, b2 = self.betas
        k = t + 1
        step_size = lr * math.sqrt(1.0 - b2 ** k) / (1.0 - b1 ** k)
        r = self.ctrl_i % self.NCTRL
        self.ctrl_i += 1
        if self.ctrl_ev[r] is not None:
            self.ctrl_ev[r].synchronize()          # the copy that last read this pinned block has executed
        self.host_hyper[r][0] = lr
        self.host_hyper[r][1] = step_size
        self.host_seed[r][0] = int(torch.randint(0, 2 ** 62, (1,), generator=self._gen).item())
        with torch.cuda.stream(self.stream):
            self.dev_seed.copy_(self.host_seed[r], non_blocking=True)
            self.dev_hyper.copy_(self.host_hyper[r], non_blocking=True)
            if self.ctrl_ev[r] is None:
                self.ctrl_ev[r] = torch.cuda.Event()
            self.ctrl_ev[r].record()

    def launch_train_step(self, transfers=True):
        """Enqueue one training step on the inputs already staged in pinned memory (no host sync).
        transfers=False replays the variant without the batch H2D / loss D2H copies (inputs resident in
        HBM): the kernel-only number of bench.py."""
        self.set_step_controls()
        slot = self.slot
        if self.use_graph and not transfers:
            if self._train_graph_resident is None:
                torch.cuda.synchronize(self.dev)   # earlier steps may still be running on self.stream: clone AFTER them
                saved = (self.eng.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
                self._train_graph_resident = self._capture(lambda: self._train_body(False, slot))
                self.eng.flat.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2])
                self.grads.zero_()
                self.eng.refresh_operands()
                torch.cuda.synchronize(self.dev)
            with torch.cuda.stream(self.stream):
                self._train_graph_resident.replay()
        elif self.use_graph:
            if self._train_graph[slot] is None:
                # the capture's warm-up run applies an optimizer step that is undone below: the saved copies must be
                # taken after every step already queued on self.stream has finished (pipelined submit_train_step
                # reaches this point for slot 1 while step 0 is still running)
                torch.cuda.synchronize(self.dev)
                saved = (self.eng.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
                self._train_graph[slot] = self._capture(lambda: self._train_body(True, slot))
                # the warm-up run inside _capture applied one optimizer step: undo it
                self.eng.flat.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2])
                self.grads.zero_()
                self.eng.refresh_operands()
                torch.cuda.synchronize(self.dev)
            with torch.cuda.stream(self.stream):
                self._train_graph[slot].replay()
                self.done[slot].record()
        else:
            with torch.cuda.stream(self.stream):
                self._train_body(transfers, slot)
                self.done[slot].record()
        self.step_count += 1
        self.eng.operands_dirty = False
        if transfers:
            self.slot = (slot + 1) % self.NSLOT      # next batch is staged into the other pinned slot
        return slot

    def submit_train_step(self, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None, labels=None) -> int:
        """Stage a host batch and enqueue its optimisation step; returns a ticket for `result()`.  Does not
        wait for the GPU: the next batch can be staged while this step runs."""
        self._stage(input_ids, visual, acoustic, attention_mask, token_type_ids, labels)
        return self.launch_train_step()

    def result(self, ticket: int) -> float:
        """Loss of a submitted step (blocks until that step's D2H copy has landed)."""
        self.done[ticket].synchronize()
        return float(self.host_loss[ticket][0])

    def train_step(self, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None, labels=None) -> float:
        """One optimisation step on a host batch; returns the loss (host float, like `loss.item()`)."""
        return self.result(self.submit_train_step(input_ids, visual, acoustic, attention_mask, token_type_ids, labels))

    def launch_eval_step(self, transfers=True):
        slot = self.slot
        if self.use_graph and not transfers:
            if self._eval_graph_resident is None:
                self._eval_graph_resident = self._capture(lambda: self._eval_body(False, slot))
            with torch.cuda.stream(self.stream):
                self._eval_graph_resident.replay()
        elif self.use_graph:
            if self._eval_graph[slot] is None:
                self._eval_graph[slot] = self._capture(lambda: self._eval_body(True, slot))
            with torch.cuda.stream(self.stream):
                self._eval_graph[slot].replay()
                self.done[slot].record()
        else:
            with torch.cuda.stream(self.stream):
                self._eval_body(transfers, slot)
                self.done[slot].record()
        if transfers:
            self.slot = (slot + 1) % self.NSLOT

    def eval_step(self, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None) -> torch.Tensor:
        """Eval-mode logits for a host batch (host tensor [B, num_labels])."""
        self._stage(input_ids, visual, acoustic, attention_mask, token_type_ids)
        if self.eng.operands_dirty:
            self.eng.refresh_operands()
        self.launch_eval_step()
        self.stream.synchronize()
        return self.host_logits.clone()
