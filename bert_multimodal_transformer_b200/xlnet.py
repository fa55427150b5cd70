"""MAG-XLNet model classes with the reference's API (xlnet.py:15-527), running on libmagb200.

`MAG_XLNetForSequenceClassification(config, multimodal_config)`: same constructor, `forward`
signature / positional order (xlnet.py:443-460), tuple return (xlnet.py:526-527) and state-dict keys
as the reference.  Only the path the driver exercises is native (content stream: mems / perm_mask /
target_mapping None, mem_len None); anything else raises NotImplementedError rather than taking a
different route.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _lib as L
from .engine import XlnetEngine, prep_inputs
from .modeling import MAG


def _xlnet_injection_index():
    try:
        import global_configs as g           # xlnet.py:6 star-imports XLNET_INJECTION_INDEX
        return int(g.XLNET_INJECTION_INDEX)
    except Exception:
        return 1


class _RelAttn(nn.Module):
    """Parameter holder with XLNetRelativeAttention's names and shapes (transformers 3.0.2)."""

    def __init__(self, config):
        super().__init__()
        d, n, dh = config.d_model, config.n_head, config.d_head
        for name in ("q", "k", "v", "o", "r"):
            setattr(self, name, nn.Parameter(torch.empty(d, n, dh)))
        for name in ("r_r_bias", "r_s_bias", "r_w_bias"):
            setattr(self, name, nn.Parameter(torch.empty(n, dh)))
        self.seg_embed = nn.Parameter(torch.empty(2, n, dh))
        self.layer_norm = nn.LayerNorm(d, eps=config.layer_norm_eps)


class _FeedForward(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer_norm = nn.LayerNorm(config.d_model, eps=config.layer_norm_eps)
        self.layer_1 = nn.Linear(config.d_model, config.d_inner)
        self.layer_2 = nn.Linear(config.d_inner, config.d_model)


class _Layer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.rel_attn = _RelAttn(config)
        self.ff = _FeedForward(config)


class _Summary(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.summary = nn.Linear(config.d_model, config.d_model)


def _init_weights(module, std):
    """transformers 3.0.2 XLNetPreTrainedModel._init_weights (applied through xlnet.py:40,441)."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        elif isinstance(m, _RelAttn):
            for p in (m.q, m.k, m.v, m.o, m.r, m.r_r_bias, m.r_s_bias, m.r_w_bias, m.seg_embed):
                p.data.normal_(mean=0.0, std=std)


def _check_supported(config):
    ff = getattr(config, "ff_activation", "gelu")
    if ff != "gelu":
        raise NotImplementedError("libmagb200 implements ff_activation='gelu' only")
    if getattr(config, "attn_type", "bi") != "bi" or getattr(config, "bi_data", False) or \
            getattr(config, "clamp_len", -1) > 0 or getattr(config, "same_length", False):
        raise NotImplementedError("only attn_type='bi', bi_data=False, clamp_len=-1, same_length=False (xlnet-base) is native")
    if getattr(config, "mem_len", None):
        raise NotImplementedError("mem_len must be None (xlnet-base-cased; xlnet.py:363-365 is dead on this path)")
    if getattr(config, "summary_type", "last") != "last" or not getattr(config, "summary_use_proj", True) or \
            getattr(config, "summary_activation", "tanh") != "tanh":
        raise NotImplementedError("SequenceSummary: only summary_type='last' with tanh projection is native")


class MAG_XLNetModel(nn.Module):
    """xlnet.py:15-429 -- parameter holder; the arithmetic runs in the owning classification model."""

    def __init__(self, config, multimodal_config):
        super().__init__()
        _check_supported(config)
        self.config = config
        self.n_layer = config.n_layer
        self.d_model = config.d_model
        self.word_embedding = nn.Embedding(config.vocab_size, config.d_model)
        # never read on the content-stream path; the reference leaves it uninitialised
        self.mask_emb = nn.Parameter(torch.zeros(1, 1, config.d_model), requires_grad=False)
        self.layer = nn.ModuleList([_Layer(config) for _ in range(config.n_layer)])
        self.dropout = nn.Dropout(config.dropout)
        self.MAG = MAG(config.d_model, multimodal_config.beta_shift, multimodal_config.dropout_prob)
        _init_weights(self, config.initializer_range)

    def get_input_embeddings(self):
        return self.word_embedding

    def set_input_embeddings(self, new_embeddings):
        self.word_embedding = new_embeddings

    def _prune_heads(self, heads_to_prune):
        raise NotImplementedError            # same as the reference (xlnet.py:48-49)

    def forward(self, input_ids, visual, acoustic, attention_mask=None, mems=None, perm_mask=None, target_mapping=None,
                token_type_ids=None, input_mask=None, head_mask=None, inputs_embeds=None, use_cache=True,
                output_attentions=None, output_hidden_states=None):
        """Returns (last_hidden_state,) like xlnet.py:399-429 (mem_len None).  Inference-only when called on
        its own; training goes through MAG_XLNetForSequenceClassification."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("MAG_XLNetModel.forward is inference-only (wrap in torch.no_grad())")
        owner = self.__dict__.get("_owner")
        if owner is None:
            raise NotImplementedError("construct MAG_XLNetForSequenceClassification and call model.transformer(...)")
        owner._run(input_ids, visual, acoustic, attention_mask, mems, perm_mask, target_mapping, token_type_ids, input_mask,
                   head_mask, inputs_embeds, output_attentions, output_hidden_states)
        dims, ws = owner._last
        # output dropout (xlnet.py:396) is the identity in eval mode
        return (owner._engine.peek(dims, ws, L.PEEK_LAYER_OUT, owner._engine.layers - 1),)


class _MagXlnetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ids, tt, mask, visual, acoustic, training, need_grad, *params):
        eng = model._engine
        B, S = ids.shape
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0    # host draw, no device sync
        dims = eng.dims(B, S, training, seed)
        with torch.cuda.device(eng.flat.device):
            ws = eng.workspace(dims, keep=need_grad)
            logits = torch.empty(B, eng.num_labels, dtype=torch.float32, device=ids.device)
            eng.forward(dims, ws, ids, tt, mask, visual, acoustic, logits)
        ctx.model, ctx.dims, ctx.ws = model, dims, ws
        ctx.inputs = (ids, tt, mask)
        model._last = (dims, ws)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        eng = model._engine
        ids, tt, mask = ctx.inputs
        with torch.cuda.device(eng.flat.device):
            grads = eng.grad_buffers.take(eng.flat, model._param_list[0])    # persistent, zeroed (engine._GradBuffers)
            dl = (dlogits.to(torch.float32) * eng.grad_scale).contiguous()
            eng.backward(ctx.dims, ctx.ws, ids, tt, mask, dl, grads)
        ctx.ws = None
        return (None,) * 8 + tuple(eng.layout.view(grads, n) for n in eng.order)


class MAG_XLNetForSequenceClassification(nn.Module):
    """xlnet.py:432-527."""

    def __init__(self, config, multimodal_config):
        super().__init__()
        _check_supported(config)
        self.config = config
        self.num_labels = config.num_labels
        self.transformer = MAG_XLNetModel(config, multimodal_config)
        self.sequence_summary = _Summary(config)
        self.logits_proj = nn.Linear(config.d_model, config.num_labels)
        _init_weights(self.sequence_summary, config.initializer_range)
        _init_weights(self.logits_proj, config.initializer_range)
        self.logits_proj.weight.data.normal_(mean=0.0, std=config.initializer_range)
        self.logits_proj.bias.data.zero_()
        self.transformer.__dict__["_owner"] = self
        mag = self.transformer.MAG
        self._engine = XlnetEngine(
            hidden=config.d_model, heads=config.n_head, inter=config.d_inner, layers=config.n_layer,
            vocab=config.vocab_size, visual_dim=mag.visual_dim, acoustic_dim=mag.acoustic_dim,
            num_labels=config.num_labels, ln_eps=config.layer_norm_eps, beta_shift=mag.beta_shift,
            p_drop=config.dropout, p_mag=mag.dropout.p, p_summary=getattr(config, "summary_last_dropout", 0.1),
            injection_index=_xlnet_injection_index(), grad_scale=float(os.environ.get("MAGB200_GRAD_SCALE", "1024")))
        self._last = None
        self._param_list = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, multimodal_config=None, num_labels=None,
                        config=None, state_dict=None, **kwargs):
        """multimodal_driver.py:321-323.  Offline resolution as in bert.MAG_BertForSequenceClassification."""
        from transformers import XLNetConfig
        from .bert import _load_checkpoint_state_dict, _load_pretrained_into
        if multimodal_config is None and model_args:
            multimodal_config = model_args[0]
        if config is None:
            cfg_kwargs = {} if num_labels is None else {"num_labels": num_labels}
            config = XLNetConfig.from_pretrained(pretrained_model_name_or_path, **cfg_kwargs)
        model = cls(config, multimodal_config)
        if state_dict is None:
            state_dict = _load_checkpoint_state_dict(pretrained_model_name_or_path)
        _load_pretrained_into(model, state_dict, pretrained_model_name_or_path)
        return model

    def _ensure_bound(self):
        eng = self._engine
        if not eng.is_bound():
            named = dict(self.named_parameters())
            eng.bind(named)
            self._param_list = [named[n] for n in eng.order]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine.flat = None
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._engine.operands_dirty = True
        return out

    def refresh_operands(self):
        self._engine.operands_dirty = True

    def _run(self, input_ids, visual, acoustic, attention_mask, mems, perm_mask, target_mapping, token_type_ids,
             input_mask, head_mask, inputs_embeds, output_attentions, output_hidden_states):
        if input_ids is not None and inputs_embeds is not None:                        # xlnet.py:200-203
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:                                # xlnet.py:211-213
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        assert input_mask is None or attention_mask is None, \
            "You can only use one of input_mask (uses 1 for padding) "                 # xlnet.py:259-262
        if inputs_embeds is not None or mems is not None or perm_mask is not None or target_mapping is not None \
                or head_mask is not None:
            raise NotImplementedError("inputs_embeds / mems / perm_mask / target_mapping / head_mask are not on the native "
                                      "content-stream path")
        if output_attentions or output_hidden_states or getattr(self.config, "output_attentions", False) \
                or getattr(self.config, "output_hidden_states", False):
            raise NotImplementedError("output_attentions / output_hidden_states are not materialised by the fused kernels")
        if input_mask is not None:                                                     # xlnet.py:263-264 (inverse)
            attention_mask = 1.0 - input_mask.to(torch.float32)
        self._ensure_bound()
        eng = self._engine
        dev = eng.flat.device
        ids, vis, aco, mask, tt = prep_inputs(dev, input_ids, visual, acoustic, attention_mask, token_type_ids)
        if ids.dim() != 2 or vis.shape[:2] != ids.shape or aco.shape[:2] != ids.shape:
            raise ValueError("input_ids [B,S], visual [B,S,V], acoustic [B,S,A] expected")
        if vis.shape[-1] != eng.visual_dim or aco.shape[-1] != eng.acoustic_dim:
            raise ValueError(f"visual/acoustic feature dims {vis.shape[-1]}/{aco.shape[-1]} do not match the model "
                             f"({eng.visual_dim}/{eng.acoustic_dim})")
        training = self.training
        if eng.operands_dirty:                                             # see bert.MAG_BertForSequenceClassification._run
            with torch.cuda.device(dev):
                eng.refresh_operands()
        eng.operands_dirty = training
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        params = self._param_list if need_grad else []
        return _MagXlnetFunction.apply(self, ids, tt, mask, vis, aco, training, need_grad, *params)

    def forward(self, input_ids, visual, acoustic, attention_mask=None, mems=None, perm_mask=None, target_mapping=None,
                token_type_ids=None, input_mask=None, head_mask=None, inputs_embeds=None, use_cache=True, labels=None,
                output_attentions=None, output_hidden_states=None):
        logits = self._run(input_ids, visual, acoustic, attention_mask, mems, perm_mask, target_mapping, token_type_ids,
                           input_mask, head_mask, inputs_embeds, output_attentions, output_hidden_states)
        outputs = (logits,)                       # mem_len None: no `mems` entry in the tuple (xlnet.py:401-402)
        if labels is not None:                    # xlnet.py:513-524
            if self.num_labels == 1:
                loss = nn.functional.mse_loss(logits.view(-1), labels.view(-1).to(logits.dtype))
            else:
                loss = nn.functional.cross_entropy(logits.view(-1, self.num_labels), labels.view(-1))
            outputs = (loss,) + outputs
        return outputs
