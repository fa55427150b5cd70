"""MAG-BERT model classes with the reference's API (bert.py:76-324), running on libmagb200.

`MAG_BertForSequenceClassification(config, multimodal_config)` has the same constructor, the same
`forward` signature / return tuple and the same state-dict keys as the reference class
(bert.py:240-324), so multimodal_driver.py:311-388 (named-parameter grouping, `.to(DEVICE)`,
`model(...)[0]`, `loss.backward()`, external optimizer) works unchanged.  The nn.Modules below only
*hold* parameters; the arithmetic is one native enqueue per forward and one per backward.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _lib as L
from .engine import BertEngine, prep_inputs
from .modeling import MAG


class _Embeddings(nn.Module):
    """Parameter holder with BertEmbeddings' key names (transformers 3.0.2)."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class _SelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(config.hidden_size, config.hidden_size)
        self.value = nn.Linear(config.hidden_size, config.hidden_size)


class _DenseLN(nn.Module):
    def __init__(self, n_in, n_out, eps):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)
        self.LayerNorm = nn.LayerNorm(n_out, eps=eps)


class _Dense(nn.Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)


class _Attention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = _SelfAttention(config)
        self.output = _DenseLN(config.hidden_size, config.hidden_size, config.layer_norm_eps)


class _Layer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = _Attention(config)
        self.intermediate = _Dense(config.hidden_size, config.intermediate_size)
        self.output = _DenseLN(config.intermediate_size, config.hidden_size, config.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(config) for _ in range(config.num_hidden_layers)])


def _init_weights(module, std):
    """transformers 3.0.2 BertPreTrainedModel._init_weights, applied to every submodule incl. MAG
    through bert.py:90,249."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()


def _check_supported(config):
    if getattr(config, "hidden_act", "gelu") != "gelu":
        raise NotImplementedError("libmagb200 implements hidden_act='gelu' (erf) only")
    if getattr(config, "is_decoder", False):
        raise NotImplementedError("decoder / cross-attention path (bert.py:186-201) is not on the MAG hot path")


def _is_default_position_ids(position_ids, seq_len):
    """True when position_ids is None or spells out the default 0 .. seq_len-1 for every row ([S], [1, S] or [B, S])."""
    if position_ids is None:
        return True
    p = torch.as_tensor(position_ids)
    if p.dim() == 0 or p.shape[-1] != seq_len or p.dim() > 2:
        return False
    want = torch.arange(seq_len, device=p.device, dtype=p.dtype)
    return bool((p == want).all())


def _is_identity_head_mask(head_mask):
    """True when head_mask keeps every head: None, a list of None, or a tensor / list of tensors of ones."""
    if head_mask is None:
        return True
    if isinstance(head_mask, (list, tuple)):
        return all(_is_identity_head_mask(h) for h in head_mask)
    h = torch.as_tensor(head_mask)
    return bool((h == 1).all())


class MAG_BertModel(nn.Module):
    """bert.py:76-237.  Holds embeddings / encoder / pooler / MAG parameters."""

    def __init__(self, config, multimodal_config):
        super().__init__()
        _check_supported(config)
        self.config = config
        self.embeddings = _Embeddings(config)
        self.encoder = _Encoder(config)
        self.pooler = _Dense(config.hidden_size, config.hidden_size)
        self.MAG = MAG(config.hidden_size, multimodal_config.beta_shift, multimodal_config.dropout_prob)
        _init_weights(self, config.initializer_range)
        self._standalone = None   # lazily built classification wrapper for standalone use

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    def _prune_heads(self, heads_to_prune):
        raise NotImplementedError("head pruning is not supported by the fused attention kernels")

    def forward(self, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None, position_ids=None,
                head_mask=None, inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None,
                output_attentions=None, output_hidden_states=None):
        """Returns (sequence_output, pooled_output) like bert.py:233-237.  Standalone use is
        inference-only; training goes through MAG_BertForSequenceClassification."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "MAG_BertModel.forward is inference-only (wrap in torch.no_grad()); the differentiable entry point is "
                "MAG_BertForSequenceClassification")
        owner = self.__dict__.get("_owner")
        if owner is None:
            raise NotImplementedError("construct MAG_BertForSequenceClassification and call model.bert(...)")
        return owner._encode(input_ids, visual, acoustic, attention_mask, token_type_ids, position_ids, head_mask,
                             inputs_embeds, output_attentions, output_hidden_states)


class _MagBertFunction(torch.autograd.Function):
    """One native forward enqueue, one native backward enqueue (multimodal_driver.py:363-378)."""

    @staticmethod
    def forward(ctx, model, ids, tt, mask, visual, acoustic, training, need_grad, *params):
        eng = model._engine
        B, S = ids.shape
        # dropout seed of this call: drawn on the HOST from torch's default CPU generator (no device sync; follows
        # torch.manual_seed like the reference's nn.Dropout layers do)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        dims = eng.dims(B, S, training, seed)
        with torch.cuda.device(eng.flat.device):      # kernels are enqueued on THIS device's current stream
            ws = eng.workspace(dims, keep=need_grad)
            logits = torch.empty(B, eng.num_labels, dtype=torch.float32, device=ids.device)
            eng.forward(dims, ws, ids, tt, mask, visual, acoustic, logits)
        ctx.model, ctx.dims, ctx.ws = model, dims, ws
        ctx.inputs = (ids, tt, mask)
        ctx.nparams = len(params)
        model._last = (dims, ws)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        eng = model._engine
        ids, tt, mask = ctx.inputs
        with torch.cuda.device(eng.flat.device):
            # persistent flat gradient buffer (zeroed, no allocation per step); the per-parameter views returned below
            # become the .grad tensors, which is what lets optim.AdamW run as one fused launch
            grads = eng.grad_buffers.take(eng.flat, model._param_list[0])
            dl = (dlogits.to(torch.float32) * eng.grad_scale).contiguous()
            eng.backward(ctx.dims, ctx.ws, ids, tt, mask, dl, grads)
        ctx.ws = None
        out = tuple(eng.layout.view(grads, n) for n in eng.order)
        return (None,) * 8 + out


class MAG_BertForSequenceClassification(nn.Module):
    """bert.py:240-324."""

    def __init__(self, config, multimodal_config):
        super().__init__()
        _check_supported(config)
        self.config = config
        self.num_labels = config.num_labels
        self.bert = MAG_BertModel(config, multimodal_config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.classifier = nn.Linear(config.hidden_size, config.num_labels)
        _init_weights(self, config.initializer_range)
        self.bert.__dict__["_owner"] = self
        mag = self.bert.MAG
        self._engine = BertEngine(
            hidden=config.hidden_size, heads=config.num_attention_heads, inter=config.intermediate_size,
            layers=config.num_hidden_layers, vocab=config.vocab_size, max_pos=config.max_position_embeddings,
            type_vocab=config.type_vocab_size, visual_dim=mag.visual_dim, acoustic_dim=mag.acoustic_dim,
            num_labels=config.num_labels, ln_eps=config.layer_norm_eps, beta_shift=mag.beta_shift,
            p_hidden=config.hidden_dropout_prob, p_attn=config.attention_probs_dropout_prob, p_mag=mag.dropout.p,
            grad_scale=float(os.environ.get("MAGB200_GRAD_SCALE", "1024")))
        self._last = None

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, multimodal_config=None, num_labels=None,
                        config=None, state_dict=None, **kwargs):
        """multimodal_driver.py:317-319 calls `from_pretrained(name, multimodal_config=..., num_labels=1)`.
        Resolves a local directory / HF cache entry (no network is assumed); weights present in the
        checkpoint are loaded by key, everything else (MAG, classifier) keeps its fresh init."""
        from transformers import BertConfig
        if multimodal_config is None and model_args:
            multimodal_config = model_args[0]
        if config is None:
            cfg_kwargs = {} if num_labels is None else {"num_labels": num_labels}
            config = BertConfig.from_pretrained(pretrained_model_name_or_path, **cfg_kwargs)
        model = cls(config, multimodal_config)
        if state_dict is None:
            state_dict = _load_checkpoint_state_dict(pretrained_model_name_or_path)
        _load_pretrained_into(model, state_dict, pretrained_model_name_or_path,
                              rename=lambda k: k.replace("gamma", "weight").replace("beta", "bias") if "LayerNorm" in k else k)
        return model

    # ------------------------------------------------------------------ plumbing
    def _ensure_bound(self):
        eng = self._engine
        if not eng.is_bound():
            named = dict(self.named_parameters())
            eng.bind(named)
            self._param_list = [named[n] for n in eng.order]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine.flat = None   # .to()/.cuda()/.float() re-created the parameter storages
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._engine.operands_dirty = True
        return out

    def refresh_operands(self):
        """Re-derive the fp16 GEMM operand copies after editing weights by hand in eval mode."""
        self._engine.operands_dirty = True

    def _run(self, input_ids, visual, acoustic, attention_mask, token_type_ids, position_ids, head_mask,
             inputs_embeds, output_attentions, output_hidden_states):
        # argument checks with the reference's messages (bert.py:158-168)
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        if inputs_embeds is not None:
            raise NotImplementedError("inputs_embeds is not supported by the native path (the embedding kernels gather by id)")
        self._ensure_bound()
        eng = self._engine
        dev = eng.flat.device
        ids, vis, aco, mask, tt = prep_inputs(dev, input_ids, visual, acoustic, attention_mask, token_type_ids)
        if ids.dim() != 2 or vis.shape[:2] != ids.shape or aco.shape[:2] != ids.shape:
            raise ValueError("input_ids [B,S], visual [B,S,V], acoustic [B,S,A] expected")
        # Values that mean "the default" are accepted like the reference accepts them (bert.py:147-156 builds arange
        # positions, bert.py:208-209 turns head_mask=None into [None] * layers); anything else has no kernel.
        if not _is_default_position_ids(position_ids, ids.shape[-1]):
            raise NotImplementedError("position_ids other than arange(seq_len) are not supported by the native path")
        if not _is_identity_head_mask(head_mask):
            raise NotImplementedError("a head_mask other than None / all ones is not supported by the fused attention kernels")
        if vis.shape[-1] != eng.visual_dim or aco.shape[-1] != eng.acoustic_dim:
            raise ValueError(f"visual/acoustic feature dims {vis.shape[-1]}/{aco.shape[-1]} do not match the model "
                             f"({eng.visual_dim}/{eng.acoustic_dim})")
        training = self.training
        # fp16 operand mirror: re-derived when flagged dirty.  A training forward flags it for the next call (an external
        # optimizer may change the fp32 masters in place -- `p.data` updates do not bump version counters, so this
        # cannot be detected); this package's fused AdamW rewrites the mirror in its kernel and clears the flag
        if eng.operands_dirty:
            with torch.cuda.device(dev):
                eng.refresh_operands()
        eng.operands_dirty = training    # an optimizer step may follow; optim.AdamW's fused step clears the flag again
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        params = self._param_list if need_grad else []
        return _MagBertFunction.apply(self, ids, tt, mask, vis, aco, training, need_grad, *params)

    def _extras(self, attention_mask, output_attentions, output_hidden_states):
        """The optional tail of the reference's output tuple (bert.py:122-124, 233-237; transformers 3.0.2 BertEncoder):
        hidden_states = (MAG output, layer 1 .. L outputs), attentions = per-layer probabilities after dropout.  They
        are detached copies read back from the workspace of the forward that just ran (the fused kernels keep the
        residual stream there and never materialise the probabilities, which are recomputed on request)."""
        cfg = self.config
        if output_attentions is None:
            output_attentions = getattr(cfg, "output_attentions", False)
        if output_hidden_states is None:
            output_hidden_states = getattr(cfg, "output_hidden_states", False)
        extras = ()
        if not (output_attentions or output_hidden_states):
            return extras
        dims, ws = self._last
        eng = self._engine
        with torch.cuda.device(eng.flat.device):
            if output_hidden_states:
                extras += ((eng.peek(dims, ws, L.PEEK_MAG),) +
                           tuple(eng.peek(dims, ws, L.PEEK_LAYER_OUT, l) for l in range(eng.layers)),)
            if output_attentions:
                mask = None if attention_mask is None else attention_mask.to(device=eng.flat.device,
                                                                             dtype=torch.int64).contiguous()
                extras += (tuple(eng.attention_probs(dims, ws, l, mask) for l in range(eng.layers)),)
        return extras

    def _encode(self, input_ids, visual, acoustic, attention_mask, token_type_ids, position_ids, head_mask, inputs_embeds,
                output_attentions, output_hidden_states):
        self._run(input_ids, visual, acoustic, attention_mask, token_type_ids, position_ids, head_mask, inputs_embeds,
                  output_attentions, output_hidden_states)
        dims, ws = self._last
        eng = self._engine
        seq = eng.peek(dims, ws, L.PEEK_LAYER_OUT, eng.layers - 1)
        pooled = eng.peek(dims, ws, L.PEEK_POOLED)
        return (seq, pooled) + self._extras(attention_mask, output_attentions, output_hidden_states)

    def forward(self, input_ids, visual, acoustic, attention_mask=None, token_type_ids=None, position_ids=None,
                head_mask=None, inputs_embeds=None, labels=None, output_attentions=None, output_hidden_states=None):
        logits = self._run(input_ids, visual, acoustic, attention_mask, token_type_ids, position_ids, head_mask,
                           inputs_embeds, output_attentions, output_hidden_states)
        outputs = (logits,) + self._extras(attention_mask, output_attentions, output_hidden_states)   # bert.py:309
        if labels is not None:                                                  # bert.py:313-322
            if self.num_labels == 1:
                loss = nn.functional.mse_loss(logits.view(-1), labels.view(-1).to(logits.dtype))
            else:
                loss = nn.functional.cross_entropy(logits.view(-1, self.num_labels), labels.view(-1))
            outputs = (loss,) + outputs
        return outputs


def _load_pretrained_into(model, state_dict, source, rename=lambda k: k):
    """Copy a checkpoint into `model` the way transformers' from_pretrained does, and say what happened: no weights
    file is an error (the reference would fail to resolve the checkpoint too -- fine-tuning silently from a random
    BERT is the worst outcome); keys missing from the checkpoint (MAG, classifier: expected), unexpected keys and
    shape mismatches are reported."""
    import warnings
    if state_dict is None:
        raise OSError(f"no weights file (model.safetensors / pytorch_model.bin) found for '{source}': refusing to return "
                      "a randomly initialised model from from_pretrained(); construct the class from a config instead")
    own = model.state_dict()
    loadable, mismatched, unexpected = {}, [], []
    for k, v in state_dict.items():
        k2 = rename(k)
        if k2 not in own:
            unexpected.append(k)
        elif tuple(own[k2].shape) != tuple(v.shape):
            mismatched.append((k2, tuple(v.shape), tuple(own[k2].shape)))
        else:
            loadable[k2] = v
    model.load_state_dict(loadable, strict=False)
    missing = [k for k in own if k not in loadable]
    model.pretrained_load_report = {"source": str(source), "loaded": len(loadable), "missing": missing,
                                    "unexpected": unexpected, "mismatched": mismatched}
    if mismatched:
        warnings.warn(f"from_pretrained('{source}'): {len(mismatched)} tensors have a different shape in the checkpoint "
                      f"and keep their fresh initialisation, e.g. {mismatched[:3]}")
    if missing:
        warnings.warn(f"from_pretrained('{source}'): {len(missing)} tensors are not in the checkpoint and are newly "
                      f"initialised: {missing[:6]}{' ...' if len(missing) > 6 else ''}")
    if unexpected:
        warnings.warn(f"from_pretrained('{source}'): {len(unexpected)} checkpoint tensors are not used: {unexpected[:6]}"
                      f"{' ...' if len(unexpected) > 6 else ''}")


def _load_checkpoint_state_dict(name_or_path):
    """Local checkpoint lookup (safetensors or torch pickle, directory or HF cache entry); None when absent."""
    cands = []
    if os.path.isdir(name_or_path):
        cands = [os.path.join(name_or_path, f) for f in ("model.safetensors", "pytorch_model.bin")]
    else:
        try:
            from transformers.utils import cached_file
            for f in ("model.safetensors", "pytorch_model.bin"):
                try:
                    pth = cached_file(name_or_path, f, local_files_only=True)
                    if pth:
                        cands.append(pth)
                except Exception:
                    pass
        except Exception:
            pass
    for c in cands:
        if os.path.exists(c):
            if c.endswith(".safetensors"):
                from safetensors.torch import load_file
                return load_file(c)
            return torch.load(c, map_location="cpu")
    return None
