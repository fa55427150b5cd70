"""Host-side driver of the native MAG-BERT schedules (magb200_bert_forward / _backward).

Owns the flat fp32 parameter buffer (the nn.Parameters are views into it), its fp16 operand mirror,
the offsets table and the workspaces.  PyTorch is used for device memory and streams only; every
FLOP runs in libmagb200.
"""
from __future__ import annotations

import weakref
from ctypes import byref, c_void_p
from typing import Dict, Optional

import torch

from . import _lib as L
from .flat import FlatLayout, bert_offsets_table, bert_param_order, packed16_numel


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise L.MagB200Error(
            f"{what} is on {t.device}: bert_multimodal_transformer_b200 only runs on a CUDA (sm_100a) device; "
            "there is no CPU fallback")


class _GradBuffers:
    """Two persistent flat gradient buffers for the autograd entry points (loss.backward() through the class API).
    The native backward accumulates into a zeroed buffer and returns per-parameter VIEWS of it; autograd adopts those
    views as `p.grad` (no copy).  A buffer is reused once no parameter's .grad points into it any more (after
    optimizer.zero_grad(), whether it sets the grads to None or zeroes them in place the other buffer is taken), so a
    training loop allocates nothing per step."""

    def __init__(self):
        self.bufs = [None, None]
        self.last = None

    def take(self, flat: torch.Tensor, probe: torch.nn.Parameter) -> torch.Tensor:
        n, dev = flat.numel(), flat.device
        busy = None
        g = probe.grad
        if g is not None and g.is_cuda:
            for i, b in enumerate(self.bufs):
                if b is not None and b.data_ptr() <= g.data_ptr() < b.data_ptr() + 4 * b.numel():
                    busy = i
        i = 1 if busy == 0 else 0
        b = self.bufs[i]
        if b is None or b.numel() != n or b.device != dev:
            b = torch.empty(n, dtype=torch.float32, device=dev)
            self.bufs[i] = b
        b.zero_()
        self.last = b
        return b


def _tag_parameters(engine, named_params, order):
    ref = weakref.ref(engine)
    for n in order:
        named_params[n]._magb200_engine = ref


class BertEngine:
    def __init__(self, *, hidden, heads, inter, layers, vocab, max_pos, type_vocab, visual_dim, acoustic_dim,
                 num_labels, ln_eps, beta_shift, p_hidden, p_attn, p_mag, grad_scale=1024.0):
        if hidden != 768 or heads != 12:
            raise NotImplementedError("libmagb200 kernels are specialised for hidden_size=768, 12 heads (bert-base)")
        self.hidden, self.heads, self.inter, self.layers = hidden, heads, inter, layers
        self.vocab, self.max_pos, self.type_vocab = vocab, max_pos, type_vocab
        self.visual_dim, self.acoustic_dim, self.num_labels = visual_dim, acoustic_dim, num_labels
        self.ln_eps, self.beta_shift = ln_eps, beta_shift
        self.p_hidden, self.p_attn, self.p_mag = p_hidden, p_attn, p_mag
        self.grad_scale = float(grad_scale)
        self.ka = hidden + 128
        self.order = bert_param_order(layers)
        self.layout: Optional[FlatLayout] = None
        self.flat: Optional[torch.Tensor] = None       # fp32 master parameters
        self.flat16: Optional[torch.Tensor] = None     # fp16 mirror + packed MAG gate matrix
        self.offsets = None
        self.operands_dirty = True
        self.grad_buffers = _GradBuffers()
        self.allocator = None          # (numel, dtype) -> tensor; the data-parallel trainer installs a peer-memory one
        self._eval_ws: Dict[tuple, torch.Tensor] = {}
        self._expected_ptrs = None

    @property
    def last_grad_buffer(self):
        return self.grad_buffers.last

    # ------------------------------------------------------------------ parameters
    def bind(self, named_params: Dict[str, torch.nn.Parameter]):
        """(Re)build the flat buffer on the parameters' device and make every parameter a view of it."""
        missing = [n for n in self.order if n not in named_params]
        if missing:
            raise KeyError(f"parameters missing from the model: {missing[:4]}...")
        dev = named_params[self.order[0]].device
        _require_cuda(named_params[self.order[0]], "model parameters")
        if self.layout is None:
            self.layout = FlatLayout([(n, tuple(named_params[n].shape)) for n in self.order])
            self.offsets = bert_offsets_table(self.layout, self.layers, self.hidden, self.ka)
        if self.allocator is not None:
            flat = self.allocator(self.layout.total, torch.float32)
            flat.zero_()
        else:
            flat = torch.zeros(self.layout.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for n in self.order:
                p = named_params[n]
                v = self.layout.view(flat, n)
                v.copy_(p.detach().to(device=dev, dtype=torch.float32))
                p.data = v
        self.flat = flat
        self.flat16 = torch.zeros(self.layout.total + packed16_numel(self.hidden, self.ka), dtype=torch.float16,
                                  device=dev)
        self._expected_ptrs = [(named_params[n], self.layout.view(flat, n).data_ptr()) for n in self.order]
        _tag_parameters(self, named_params, self.order)
        self.operands_dirty = True
        self._eval_ws.clear()

    def is_bound(self) -> bool:
        if self.flat is None:
            return False
        for p, ptr in self._expected_ptrs:
            if p.data_ptr() != ptr:
                return False
        return True

    def dims(self, batch, seq, training, seed=0) -> L.BertDims:
        d = L.BertDims()
        d.batch, d.seq, d.hidden, d.heads, d.inter, d.layers = batch, seq, self.hidden, self.heads, self.inter, self.layers
        d.vocab, d.max_pos, d.type_vocab = self.vocab, self.max_pos, self.type_vocab
        d.visual_dim, d.acoustic_dim, d.num_labels, d.ka = self.visual_dim, self.acoustic_dim, self.num_labels, self.ka
        d.ln_eps, d.mag_ln_eps, d.beta_shift = self.ln_eps, 1e-5, self.beta_shift
        if training:
            d.p_hidden, d.p_attn, d.p_mag, d.p_cls = self.p_hidden, self.p_attn, self.p_mag, self.p_hidden
        else:
            d.p_hidden = d.p_attn = d.p_mag = d.p_cls = 0.0
        d.grad_scale = self.grad_scale
        d.seed = seed
        return d

    def refresh_operands(self, dims=None):
        lib = L.load()
        d = dims or self.dims(1, 1, False)
        L.check(lib.magb200_bert_refresh_operands(byref(d), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                                  L.current_stream()))
        self.operands_dirty = False

    def workspace_bytes(self, dims):
        return L.load().magb200_bert_workspace_bytes(byref(dims))

    def after_optimizer_step(self):
        """fp32 masters changed in place and the fused AdamW already rewrote the fp16 mirror: only the packed
        MAG gate matrix has to be re-derived."""
        lib, o, f = L.load(), self.offsets, self.flat
        L.check(lib.magb200_mag_pack_weights(
            c_void_p(f.data_ptr() + 4 * o[L.OFF_MAG_WHV]), c_void_p(f.data_ptr() + 4 * o[L.OFF_MAG_WHA]),
            c_void_p(f.data_ptr() + 4 * o[L.OFF_MAG_WV]), c_void_p(f.data_ptr() + 4 * o[L.OFF_MAG_WA]),
            self.visual_dim, self.acoustic_dim, self.ka, c_void_p(self.flat16.data_ptr() + 2 * o[L.OFF_MAG_PACKED16]),
            L.current_stream()))

    def workspace(self, dims, keep: bool) -> torch.Tensor:
        lib = L.load()
        nbytes = lib.magb200_bert_workspace_bytes(byref(dims))
        if keep:   # saved for backward: a fresh buffer per autograd graph
            return torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
        key = (dims.batch, dims.seq)
        ws = self._eval_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self._eval_ws.clear()
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
            self._eval_ws[key] = ws
        return ws

    # ------------------------------------------------------------------ schedules
    def forward(self, dims, ws, input_ids, token_type_ids, attention_mask, visual, acoustic, logits):
        lib = L.load()
        L.check(lib.magb200_bert_forward(byref(dims), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                         L.ptr(input_ids), L.ptr(token_type_ids), L.ptr(attention_mask),
                                         L.ptr(visual), L.ptr(acoustic), L.ptr(ws), ws.numel(), L.ptr(logits),
                                         L.current_stream()))

    def backward(self, dims, ws, input_ids, token_type_ids, attention_mask, dlogits_scaled, grads):
        lib = L.load()
        L.check(lib.magb200_bert_backward(byref(dims), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                          L.ptr(input_ids), L.ptr(token_type_ids), L.ptr(attention_mask), L.ptr(ws),
                                          ws.numel(), L.ptr(dlogits_scaled), L.ptr(grads), L.current_stream()))

    def backward_part(self, dims, ws, input_ids, token_type_ids, attention_mask, dlogits_scaled, grads, layer_hi,
                      layer_lo, flags):
        L.check(L.load().magb200_bert_backward_part(byref(dims), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                                    L.ptr(input_ids), L.ptr(token_type_ids), L.ptr(attention_mask),
                                                    L.ptr(ws), ws.numel(), L.ptr(dlogits_scaled), L.ptr(grads), layer_hi,
                                                    layer_lo, flags, L.current_stream()))

    def backward_adamw(self, dims, ws, input_ids, token_type_ids, attention_mask, dlogits_scaled, grads, exp_avg,
                       exp_avg_sq, decay_chunks, lr, betas, eps, weight_decay, step, grad_unscale, hyper_dev):
        """Backward and the AdamW update as one schedule (magb200_bert_backward_adamw): same result as backward()
        followed by magb200_adamw_step over the flat buffer, with the per-layer updates overlapped."""
        L.check(L.load().magb200_bert_backward_adamw(
            byref(dims), self.offsets, L.ptr(self.flat), L.ptr(self.flat16), L.ptr(input_ids), L.ptr(token_type_ids),
            L.ptr(attention_mask), L.ptr(ws), ws.numel(), L.ptr(dlogits_scaled), L.ptr(grads), L.ptr(exp_avg),
            L.ptr(exp_avg_sq), L.ptr(decay_chunks), lr, betas[0], betas[1], eps, weight_decay, step, grad_unscale, 1,
            L.ptr(hyper_dev) if hyper_dev is not None else None, L.current_stream()))

    def layer_grad_range(self, lo, hi):
        """[start, end) element range of the flat buffers covering encoder layers lo .. hi-1 (contiguous)."""
        start = self.offsets[L.OFF_LAYER0 + L.OFF_PER_LAYER * lo + L.LOFF_QKV_W]
        end = self.offsets[L.OFF_LAYER0 + L.OFF_PER_LAYER * hi + L.LOFF_QKV_W] if hi < self.layers \
            else self.offsets[L.OFF_POOL_W]
        return int(start), int(end)

    def attention_probs(self, dims, ws, layer, mask):
        out = torch.empty(dims.batch, self.heads, dims.seq, dims.seq, dtype=torch.float32, device=self.flat.device)
        L.check(L.load().magb200_bert_attention_probs(byref(dims), L.ptr(ws), layer, L.ptr(mask), L.ptr(out), out.numel() * 4,
                                                      L.current_stream()))
        return out

    def peek(self, dims, ws, what, layer=0):
        lib = L.load()
        T = dims.batch * dims.seq
        shape = {L.PEEK_EMB: (dims.batch, dims.seq, self.hidden), L.PEEK_MAG: (dims.batch, dims.seq, self.hidden),
                 L.PEEK_LAYER_OUT: (dims.batch, dims.seq, self.hidden), L.PEEK_POOLED: (dims.batch, self.hidden),
                 L.PEEK_GATES: (T, 4 * self.hidden)}[what]
        out = torch.empty(shape, dtype=torch.float32, device=self.flat.device)
        L.check(lib.magb200_bert_peek(byref(dims), L.ptr(ws), what, layer, L.ptr(out), out.numel() * 4,
                                      L.current_stream()))
        return out


def prep_inputs(device, input_ids, visual, acoustic, attention_mask, token_type_ids):
    """Contiguous device tensors of the dtypes the C ABI expects."""
    ids = input_ids.to(device=device, dtype=torch.int64).contiguous()
    vis = visual.to(device=device, dtype=torch.float32).contiguous()
    aco = acoustic.to(device=device, dtype=torch.float32).contiguous()
    mask = None if attention_mask is None else attention_mask.to(device=device, dtype=torch.int64).contiguous()
    tt = None if token_type_ids is None else token_type_ids.to(device=device, dtype=torch.int64).contiguous()
    return ids, vis, aco, mask, tt


class XlnetEngine:
    """Host-side driver of magb200_xlnet_forward / _backward (same role as BertEngine)."""

    def __init__(self, *, hidden, heads, inter, layers, vocab, visual_dim, acoustic_dim, num_labels, ln_eps, beta_shift,
                 p_drop, p_mag, p_summary, injection_index=1, grad_scale=1024.0):
        if hidden != 768 or heads != 12:
            raise NotImplementedError("libmagb200 kernels are specialised for d_model=768, 12 heads (xlnet-base)")
        from .flat import xlnet_param_order
        self.hidden, self.heads, self.inter, self.layers, self.vocab = hidden, heads, inter, layers, vocab
        self.visual_dim, self.acoustic_dim, self.num_labels = visual_dim, acoustic_dim, num_labels
        self.ln_eps, self.beta_shift = ln_eps, beta_shift
        self.p_drop, self.p_mag, self.p_summary = p_drop, p_mag, p_summary
        self.injection_index = injection_index
        self.grad_scale = float(grad_scale)
        self.ka = hidden + 128
        self.order = xlnet_param_order(layers)
        self.layout = None
        self.flat = None
        self.flat16 = None
        self.offsets = None
        self.operands_dirty = True
        self.grad_buffers = _GradBuffers()
        self._eval_ws = {}
        self._expected_ptrs = None

    @property
    def last_grad_buffer(self):
        return self.grad_buffers.last

    def bind(self, named_params):
        from .flat import xlnet_fp16_numel, xlnet_offsets_table
        missing = [n for n in self.order if n not in named_params]
        if missing:
            raise KeyError(f"parameters missing from the model: {missing[:4]}...")
        dev = named_params[self.order[0]].device
        _require_cuda(named_params[self.order[0]], "model parameters")
        if self.layout is None:
            self.layout = FlatLayout([(n, tuple(named_params[n].shape)) for n in self.order])
            self.offsets = xlnet_offsets_table(self.layout, self.layers, self.hidden, self.ka)
        flat = torch.zeros(self.layout.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for n in self.order:
                p = named_params[n]
                v = self.layout.view(flat, n)
                v.copy_(p.detach().to(device=dev, dtype=torch.float32))
                p.data = v
        self.flat = flat
        self.flat16 = torch.zeros(xlnet_fp16_numel(self.layout, self.layers, self.hidden, self.ka), dtype=torch.float16,
                                  device=dev)
        self._expected_ptrs = [(named_params[n], self.layout.view(flat, n).data_ptr()) for n in self.order]
        _tag_parameters(self, named_params, self.order)
        self.operands_dirty = True
        self._eval_ws.clear()

    def is_bound(self):
        if self.flat is None:
            return False
        return all(p.data_ptr() == ptr for p, ptr in self._expected_ptrs)

    def dims(self, batch, seq, training, seed=0) -> L.XlnetDims:
        d = L.XlnetDims()
        d.batch, d.seq, d.hidden, d.heads, d.inter, d.layers, d.vocab = batch, seq, self.hidden, self.heads, self.inter, \
            self.layers, self.vocab
        d.visual_dim, d.acoustic_dim, d.num_labels, d.ka = self.visual_dim, self.acoustic_dim, self.num_labels, self.ka
        d.injection_index = self.injection_index
        d.ln_eps, d.mag_ln_eps, d.beta_shift = self.ln_eps, 1e-5, self.beta_shift
        if training:
            d.p_drop, d.p_mag, d.p_summary = self.p_drop, self.p_mag, self.p_summary
        else:
            d.p_drop = d.p_mag = d.p_summary = 0.0
        d.grad_scale = self.grad_scale
        d.seed = seed
        return d

    def refresh_operands(self, dims=None):
        lib = L.load()
        d = dims or self.dims(1, 1, False)
        L.check(lib.magb200_xlnet_refresh_operands(byref(d), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                                   L.current_stream()))
        self.operands_dirty = False

    def workspace_bytes(self, dims):
        return L.load().magb200_xlnet_workspace_bytes(byref(dims))

    def workspace(self, dims, keep):
        nbytes = self.workspace_bytes(dims)
        if keep:
            return torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
        key = (dims.batch, dims.seq, dims.p_drop > 0)
        ws = self._eval_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self._eval_ws.clear()
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
            self._eval_ws[key] = ws
        return ws

    def forward(self, dims, ws, input_ids, token_type_ids, attention_mask, visual, acoustic, logits):
        L.check(L.load().magb200_xlnet_forward(byref(dims), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                               L.ptr(input_ids), L.ptr(token_type_ids), L.ptr(attention_mask),
                                               L.ptr(visual), L.ptr(acoustic), L.ptr(ws), ws.numel(), L.ptr(logits),
                                               L.current_stream()))

    def backward(self, dims, ws, input_ids, token_type_ids, attention_mask, dlogits_scaled, grads):
        L.check(L.load().magb200_xlnet_backward(byref(dims), self.offsets, L.ptr(self.flat), L.ptr(self.flat16),
                                                L.ptr(input_ids), L.ptr(token_type_ids), L.ptr(attention_mask),
                                                L.ptr(ws), ws.numel(), L.ptr(dlogits_scaled), L.ptr(grads),
                                                L.current_stream()))

    def peek(self, dims, ws, what, layer=0):
        shape = {L.PEEK_MAG: (dims.batch, dims.seq, self.hidden), L.PEEK_LAYER_OUT: (dims.batch, dims.seq, self.hidden),
                 L.PEEK_POOLED: (dims.batch, self.hidden)}[what]
        out = torch.empty(shape, dtype=torch.float32, device=self.flat.device)
        L.check(L.load().magb200_xlnet_peek(byref(dims), L.ptr(ws), what, layer, L.ptr(out), out.numel() * 4,
                                            L.current_stream()))
        return out

    def after_optimizer_step(self):
        """fp32 masters changed in place (fused AdamW already rewrote the fp16 mirror): re-pack the derived
        operands (MAG gate matrix, per-layer q|k|v).  The cast inside refresh is redundant but cheap."""
        self.refresh_operands()
