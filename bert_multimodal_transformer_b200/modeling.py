"""MAG -- Multimodal Adaptation Gate with the reference's API (modeling.py:6-51), on libmagb200.

`MAG(hidden_size, beta_shift, dropout_prob)`; `forward(text_embedding, visual, acoustic)` with any
leading dims.  Children are named W_hv, W_ha, W_v, W_a, LayerNorm, dropout so state dicts are
interchangeable with the reference.  Inside MAG_BertForSequenceClassification the gate runs as part
of the whole-model schedule; this module's own forward serves standalone use (BASELINE.json
configs[0]) through magb200_mag_fwd / magb200_mag_bwd.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _lib as L

TEXT_DIM = 768
KA = TEXT_DIM + 128


def _dims_from_global_configs():
    # modeling.py:4,15-19 reads the modality widths from global_configs at construction time
    try:
        import global_configs as g
        return int(g.VISUAL_DIM), int(g.ACOUSTIC_DIM)
    except Exception:
        return 47, 74   # MOSI (global_configs.py:9-12)


class _MagFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, training, need_grad, *tensors):
        with torch.cuda.device(tensors[0].device):     # enqueue on the current stream of the tensors' device
            return _MagFunction._forward(ctx, mod, training, need_grad, *tensors)

    @staticmethod
    def backward(ctx, dout):
        with torch.cuda.device(dout.device):
            return _MagFunction._backward(ctx, dout)

    @staticmethod
    def _forward(ctx, mod, training, need_grad, text, visual, acoustic, W_hv, b_hv, W_ha, b_ha, W_v, b_v, W_a, b_a,
                 gamma, beta):
        lib = L.load()
        T = text.shape[0]
        dev = text.device
        st = L.current_stream()
        packed = torch.empty(4 * TEXT_DIM * KA, dtype=torch.float16, device=dev)
        L.check(lib.magb200_mag_pack_weights(L.ptr(W_hv), L.ptr(W_ha), L.ptr(W_v), L.ptr(W_a), mod.visual_dim,
                                             mod.acoustic_dim, KA, L.ptr(packed), st))
        bias4 = torch.cat([b_hv, b_ha, b_v, b_a]).contiguous()
        amag = torch.empty(T * KA, dtype=torch.float16, device=dev)
        gates = torch.empty(T * 4 * TEXT_DIM, dtype=torch.float32, device=dev)
        out = torch.empty(T, TEXT_DIM, dtype=torch.float32, device=dev)
        p = mod.dropout.p if training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        if not mod.exact_forward or need_grad:
            # fp16-operand gate GEMM (the training path; also produces the packed operands backward reads)
            L.check(lib.magb200_mag_fwd(L.ptr(text), L.ptr(visual), L.ptr(acoustic), L.ptr(packed), L.ptr(bias4),
                                        L.ptr(gamma), L.ptr(beta), float(mod.beta_shift), float(mod.LayerNorm.eps), p,
                                        seed, 2, mod.visual_dim, mod.acoustic_dim, KA, T, L.ptr(amag), L.ptr(gates),
                                        L.ptr(out), None, st))
        if mod.exact_forward:
            # split-TF32 gate GEMMs: fp32-accurate gates / output (BASELINE.json configs[0] tolerance 1e-5)
            scratch = torch.empty(2 * (T + 4 * TEXT_DIM) * KA, dtype=torch.float32, device=dev)
            L.check(lib.magb200_mag_fwd_exact(L.ptr(text), L.ptr(visual), L.ptr(acoustic), L.ptr(W_hv), L.ptr(W_ha),
                                              L.ptr(W_v), L.ptr(W_a), L.ptr(bias4), L.ptr(gamma), L.ptr(beta),
                                              float(mod.beta_shift), float(mod.LayerNorm.eps), p, seed, 2,
                                              mod.visual_dim, mod.acoustic_dim, KA, T, L.ptr(scratch), L.ptr(gates),
                                              L.ptr(out), None, st))
        if need_grad:
            ctx.mod, ctx.p, ctx.seed = mod, p, seed
            ctx.save_for_backward(text, amag, gates, packed, gamma)
            ctx.shapes = (W_hv.shape, W_ha.shape, W_v.shape, W_a.shape)
        return out

    @staticmethod
    def _backward(ctx, dout):
        lib = L.load()
        mod = ctx.mod
        text, amag, gates, packed, gamma = ctx.saved_tensors
        T = text.shape[0]
        dev = text.device
        st = L.current_stream()
        scale = mod.grad_scale
        dscaled = (dout.to(torch.float32) * scale).contiguous()
        dgates = torch.empty(T * 4 * TEXT_DIM, dtype=torch.float16, device=dev)
        dpacked = torch.empty(4 * TEXT_DIM * KA, dtype=torch.float32, device=dev)
        dtext = torch.empty(T, TEXT_DIM, dtype=torch.float32, device=dev)
        dW = [torch.zeros(s, dtype=torch.float32, device=dev) for s in ctx.shapes]
        dbias4 = torch.zeros(4 * TEXT_DIM, dtype=torch.float32, device=dev)
        dgamma = torch.zeros(TEXT_DIM, dtype=torch.float32, device=dev)
        dbeta = torch.zeros(TEXT_DIM, dtype=torch.float32, device=dev)
        L.check(lib.magb200_mag_bwd(L.ptr(dscaled), L.ptr(text), L.ptr(amag), L.ptr(gates), L.ptr(packed), L.ptr(gamma),
                                    float(mod.beta_shift), float(mod.LayerNorm.eps), ctx.p, ctx.seed, 2, mod.visual_dim,
                                    mod.acoustic_dim, KA, T, 1.0 / scale, L.ptr(dgates), L.ptr(dpacked), L.ptr(dtext),
                                    L.ptr(dW[0]), L.ptr(dW[1]), L.ptr(dW[2]), L.ptr(dW[3]), L.ptr(dbias4),
                                    L.ptr(dgamma), L.ptr(dbeta), st))
        dtext = dtext * (1.0 / scale)
        db = dbias4.view(4, TEXT_DIM)
        # visual / acoustic are data, not parameters: the reference never asks for their gradient
        return (None, None, None, dtext, None, None, dW[0], db[0], dW[1], db[1], dW[2], db[2], dW[3], db[3], dgamma,
                dbeta)


class MAG(nn.Module):
    def __init__(self, hidden_size, beta_shift, dropout_prob, visual_dim=None, acoustic_dim=None):
        super().__init__()
        print("Initializing MAG with beta_shift:{} hidden_prob:{}".format(beta_shift, dropout_prob))  # modeling.py:9-13
        if hidden_size != TEXT_DIM:
            raise NotImplementedError("libmagb200's MAG kernels are specialised for hidden_size=768")
        gv, ga = _dims_from_global_configs()
        self.visual_dim = int(visual_dim) if visual_dim is not None else gv
        self.acoustic_dim = int(acoustic_dim) if acoustic_dim is not None else ga
        if self.visual_dim + self.acoustic_dim > KA - TEXT_DIM:
            raise NotImplementedError("visual_dim + acoustic_dim must be <= 128")
        self.W_hv = nn.Linear(self.visual_dim + TEXT_DIM, TEXT_DIM)
        self.W_ha = nn.Linear(self.acoustic_dim + TEXT_DIM, TEXT_DIM)
        self.W_v = nn.Linear(self.visual_dim, TEXT_DIM)
        self.W_a = nn.Linear(self.acoustic_dim, TEXT_DIM)
        self.beta_shift = beta_shift
        self.LayerNorm = nn.LayerNorm(hidden_size)
        self.dropout = nn.Dropout(dropout_prob)
        self.grad_scale = 1024.0
        # exact_forward: evaluate the four Linears to fp32 accuracy (three TF32 GEMMs on split operands) instead of
        # one fp16-operand GEMM; backward is unchanged.  MAGB200_MAG_EXACT=1 turns it on for every MAG instance.
        self.exact_forward = os.environ.get("MAGB200_MAG_EXACT", "0") == "1"

    def forward(self, text_embedding, visual, acoustic):
        if not text_embedding.is_cuda:
            raise L.MagB200Error("MAG.forward needs CUDA tensors: there is no CPU fallback in this package")
        lead = text_embedding.shape[:-1]
        text = text_embedding.reshape(-1, TEXT_DIM).to(torch.float32).contiguous()
        vis = visual.reshape(-1, self.visual_dim).to(device=text.device, dtype=torch.float32).contiguous()
        aco = acoustic.reshape(-1, self.acoustic_dim).to(device=text.device, dtype=torch.float32).contiguous()
        if vis.shape[0] != text.shape[0] or aco.shape[0] != text.shape[0]:
            raise ValueError("text_embedding, visual and acoustic must share their leading dimensions")
        need_grad = torch.is_grad_enabled() and (text_embedding.requires_grad or
                                                 any(p.requires_grad for p in self.parameters()))
        out = _MagFunction.apply(self, self.training, need_grad, text, vis, aco, self.W_hv.weight, self.W_hv.bias,
                                 self.W_ha.weight, self.W_ha.bias, self.W_v.weight, self.W_v.bias, self.W_a.weight,
                                 self.W_a.bias, self.LayerNorm.weight, self.LayerNorm.bias)
        return out.view(*lead, TEXT_DIM)
