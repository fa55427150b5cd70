"""AdamW with the semantics of `transformers.optimization.AdamW` 3.0.2 -- the optimizer the reference driver
constructs at multimodal_driver.py:345 (`AdamW(optimizer_grouped_parameters, lr=args.learning_rate)`; that class was
removed from transformers 5.x, so the launcher `tools/run_driver.py` installs this one under the old name).

    exp_avg    = b1 * exp_avg    + (1 - b1) * g
    exp_avg_sq = b2 * exp_avg_sq + (1 - b2) * g * g
    p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * exp_avg / (sqrt(exp_avg_sq) + eps)        (correct_bias=True)
    p -= lr * weight_decay * p                       (decoupled decay applied AFTER the Adam update; eps default 1e-6)

When every parameter handed to the optimizer is a view of one libmagb200 flat buffer (a MAG_Bert / MAG_XLNet model
bound to its engine) and the gradients are the views the native backward returned, `step()` is ONE launch of
`magb200_adamw_step` over the flat buffers: it also rewrites the fp16 operand mirror, so the next training forward
does not have to re-cast 110 M parameters.  Any other parameter set (foreign tensors, per-group learning rates that
differ, sparse gradients ...) takes the per-tensor path below, written with torch ops -- that path is what makes the
class a general drop-in; it is not used by the models of this package on a GPU.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch.optim import Optimizer

from . import _lib as L


class AdamW(Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0,
                 correct_bias: bool = True, **_ignored):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self._fused = None          # (engine, exp_avg, exp_avg_sq, decay_chunks) once the flat layout is recognised
        self._extras = []
        self._fused_checked = False
        self._fused_step = 0
        self.fused_steps_taken = 0

    # ------------------------------------------------------------------ flat-buffer fast path
    @staticmethod
    def _engine_of(p) -> Optional[object]:
        eng = getattr(p, "_magb200_engine", None)
        return eng() if callable(eng) else eng

    def _try_bind_fused(self):
        """All parameters views of one engine's flat fp32 buffer, covering it exactly once, one lr / betas / eps."""
        self._fused_checked = True
        groups = self.param_groups
        if not groups:
            return
        g0 = groups[0]
        for g in groups:
            if (g["betas"] != g0["betas"] or g["eps"] != g0["eps"] or not g["correct_bias"]):
                return
        wds = {float(g["weight_decay"]) for g in groups if g["weight_decay"] != 0.0}
        if len(wds) > 1:
            return
        eng = None
        extras = []       # parameters outside the flat buffer (XLNet's transformer.mask_emb: never used on this path,
        for g in groups:  # never receives a gradient): they keep the per-tensor path and do not block the fused one
            for p in g["params"]:
                e = self._engine_of(p)
                if e is None:
                    extras.append(p)
                    continue
                if eng is not None and e is not eng:
                    return
                eng = e
        if eng is None or eng.flat is None or not eng.flat.is_cuda:
            return
        base, n = eng.flat.data_ptr(), eng.layout.total
        seen = 0
        decay = torch.zeros(n // 256, dtype=torch.uint8)
        self._extras = extras
        for g in groups:
            for p in g["params"]:
                if self._engine_of(p) is None:
                    continue
                off = (p.data_ptr() - base) // 4
                if off < 0 or off >= n or off % 256 != 0:
                    return
                seen += p.numel()
                if g["weight_decay"] != 0.0:
                    decay[off // 256:(off + p.numel() + 255) // 256] = 1
        if seen != sum(eng.layout.numel(name) for name in eng.layout.entries):
            return
        dev = eng.flat.device
        self._fused = (eng, torch.zeros(n, dtype=torch.float32, device=dev), torch.zeros(n, dtype=torch.float32, device=dev),
                       decay.to(dev), float(next(iter(wds))) if wds else 0.0)

    def _fused_grad_buffer(self, eng):
        """The flat gradient buffer the native backward filled, if every .grad is the view it returned."""
        gbuf = getattr(eng, "last_grad_buffer", None)
        if gbuf is None:
            return None
        gbase, pbase = gbuf.data_ptr(), eng.flat.data_ptr()
        for g in self.param_groups:
            for p in g["params"]:
                if self._engine_of(p) is None:
                    continue
                if p.grad is None or p.grad.data_ptr() - gbase != p.data_ptr() - pbase or p.grad.is_sparse:
                    return None
        return gbuf

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self._fused_checked:
            self._try_bind_fused()
        if self._fused is not None:
            eng, m, v, decay, wd = self._fused
            lrs = {float(g["lr"]) for g in self.param_groups}
            gbuf = self._fused_grad_buffer(eng) if (len(lrs) == 1 and eng.flat is not None and eng.flat.data_ptr() == self._flat_ptr(eng)) else None
            if gbuf is not None:
                self._fused_step += 1
                g0 = self.param_groups[0]
                with torch.cuda.device(eng.flat.device):
                    L.check(L.load().magb200_adamw_step(
                        L.ptr(eng.flat), L.ptr(gbuf), L.ptr(m), L.ptr(v), L.ptr(eng.flat16), L.ptr(decay), eng.layout.total,
                        lrs.pop(), g0["betas"][0], g0["betas"][1], g0["eps"], wd, self._fused_step, 1.0, 0, None,
                        L.current_stream()))
                    eng.after_optimizer_step()
                eng.operands_dirty = False       # the kernel rewrote the fp16 operand copies
                self.fused_steps_taken += 1
                if any(p.grad is not None for p in getattr(self, "_extras", ())):
                    self._step_per_tensor(only=self._extras)
                return loss
            if self._fused_step > 0:
                raise RuntimeError("AdamW: the flat-buffer layout changed after fused steps were taken (model moved or "
                                   "gradients produced outside the native backward); re-create the optimizer")
        self._step_per_tensor()
        return loss

    def _flat_ptr(self, eng):
        if not hasattr(self, "_bound_flat_ptr"):
            self._bound_flat_ptr = eng.flat.data_ptr()
        return self._bound_flat_ptr

    # ------------------------------------------------------------------ general path (transformers 3.0.2, line by line)
    def _step_per_tensor(self, only=None):
        chosen = None if only is None else {id(p) for p in only}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None or (chosen is not None and id(p) not in chosen):
                    continue
                grad = p.grad
                if grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
                beta1, beta2 = group["betas"]
                state["step"] += 1
                exp_avg.mul_(beta1).add_(grad, alpha=1.0 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
                denom = exp_avg_sq.sqrt().add_(group["eps"])
                step_size = group["lr"]
                if group["correct_bias"]:
                    bias_correction1 = 1.0 - beta1 ** state["step"]
                    bias_correction2 = 1.0 - beta2 ** state["step"]
                    step_size = step_size * math.sqrt(bias_correction2) / bias_correction1
                p.addcdiv_(exp_avg, denom, value=-step_size)
                if group["weight_decay"] > 0.0:
                    p.add_(p, alpha=-group["lr"] * group["weight_decay"])
                eng = self._engine_of(p)
                if eng is not None:
                    eng.operands_dirty = True
