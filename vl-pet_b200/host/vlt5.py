"""Host model for BASELINE config 3: a frozen T5-base encoder/decoder in stock PyTorch whose module tree, parameter names
and op order follow the reference's VLT5 (src/modeling_t5.py:177-401 JointEncoder, 404-798 VLT5;
src/my_transformers/modeling_t5.py:235-252 T5LayerNorm, 255-285 DenseReluDense, 288-409 T5LayerFF, 412-676 T5Attention,
679-826 T5LayerSelfAttention, 829-893 T5LayerCrossAttention, 896-1001 T5Block), so a reference ``state_dict`` loads key
for key.  The PET op groups are ours (CUDA, include/vlpet.h): two encoder PET sites per block (K1 with s = 0.3 and no
LayerNorm behind it: pre-norm architecture), the decoder cross-attention value parallel adapter (K2,
modeling_t5.py:588-613) and the visual projection with RMS norms (K3).  Everything else is plumbing in stock PyTorch.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import encoder as E
from .. import functional as F_
from ..adapters import AdapterController
from ..visual import T5LayerNorm, VisualEmbedding
from .config import VLPetConfig
from .vlbart import Downsample, _sdpa_policy


def relative_position_bucket(relative_position, bidirectional: bool, num_buckets: int = 32, max_distance: int = 128):
    """Bucket index of memory_position - query_position (my_transformers/modeling_t5.py:465-507): half of the buckets for
    exact small offsets, the other half logarithmic up to max_distance; bidirectional stacks the sign on top."""
    ret = torch.zeros_like(relative_position)
    n = num_buckets
    if bidirectional:
        n //= 2
        ret = ret + (relative_position > 0).long() * n
        rp = relative_position.abs()
    else:
        rp = -torch.clamp(relative_position, max=0)
    max_exact = n // 2
    large = max_exact + (torch.log(rp.float().clamp(min=1) / max_exact) / math.log(max_distance / max_exact) * (n - max_exact)).long()
    large = torch.clamp(large, max=n - 1)
    return ret + torch.where(rp < max_exact, rp, large)


class T5Attention(nn.Module):
    """q/k/v/o without biases and WITHOUT the 1/sqrt(d) scale; the first block of a stack owns the relative-position table.
    Cross-attention instances carry ``attn_value_parallel_adapter`` (modeling_t5.py:439-442, 588-613)."""

    def __init__(self, config: VLPetConfig, is_decoder: bool, has_relative_attention_bias: bool, value_adapter: bool):
        super().__init__()
        self.is_decoder = is_decoder
        self.has_relative_attention_bias = has_relative_attention_bias
        self.num_buckets = config.relative_attention_num_buckets
        self.d_model, self.d_kv, self.n_heads = config.d_model, config.d_kv, config.num_heads
        self.inner_dim = self.n_heads * self.d_kv
        self.dropout = config.dropout_rate
        self.q = nn.Linear(self.d_model, self.inner_dim, bias=False)
        self.k = nn.Linear(self.d_model, self.inner_dim, bias=False)
        self.v = nn.Linear(self.d_model, self.inner_dim, bias=False)
        self.o = nn.Linear(self.inner_dim, self.d_model, bias=False)
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(self.num_buckets, self.n_heads)
        self.attn_value_parallel_adapter = AdapterController(config.vpa_adapter_config()) if value_adapter else None
        self._qkv = None      # (key, [3 * inner, d] weight): fused projection of a frozen self-attention, see _fused_qkv

    def compute_bias(self, q_len: int, k_len: int) -> torch.Tensor:
        dev = self.relative_attention_bias.weight.device
        ctx = torch.arange(q_len, dtype=torch.long, device=dev)[:, None]
        mem = torch.arange(k_len, dtype=torch.long, device=dev)[None, :]
        bucket = relative_position_bucket(mem - ctx, bidirectional=not self.is_decoder, num_buckets=self.num_buckets)
        return self.relative_attention_bias(bucket).permute(2, 0, 1).unsqueeze(0)      # [1, H, q, k]

    def _fused_qkv(self, dtype):
        """Frozen self-attention: q, k, v out of ONE GEMM over the row-concatenated weights (a cached copy; the state-dict
        keeps the reference's three Linears) -- and one dgrad GEMM + one stack instead of three + two accumulation passes."""
        ws = (self.q.weight, self.k.weight, self.v.weight)
        if any(t.requires_grad for t in ws):
            return None
        key = tuple((t.data_ptr(), t._version, t.dtype) for t in ws)
        if self._qkv is None or self._qkv[0] != key:
            self._qkv = (key, torch.cat([t.detach() for t in ws], 0).contiguous())
        w = self._qkv[1]
        return w if (w.dtype == dtype or torch.is_autocast_enabled()) else None

    def forward(self, hidden_states, key_value_states=None, bias=None, task=None, k_override=None):
        B, L, _ = hidden_states.shape
        src = hidden_states if key_value_states is None else key_value_states
        wqkv = self._fused_qkv(hidden_states.dtype) if key_value_states is None else None
        if wqkv is not None:
            q, k, v = F.linear(hidden_states, wqkv).view(B, L, 3, self.inner_dim).unbind(2)
        else:
            q, k, v = self.q(hidden_states), (k_override if k_override is not None else self.k(src)), self.v(src)
        if key_value_states is not None and self.attn_value_parallel_adapter is not None:
            v = self.attn_value_parallel_adapter(key_value_states, task, y=v)                  # K2
        sh = lambda t: t.view(B, -1, self.n_heads, self.d_kv).transpose(1, 2)  # noqa: E731
        m = bias
        if m is not None and m.dtype != q.dtype:
            m = m.to(q.dtype)
        with _sdpa_policy(q):
            o = F.scaled_dot_product_attention(sh(q), sh(k), sh(v), attn_mask=m, dropout_p=self.dropout if self.training else 0.0,
                                               scale=1.0)
        return self.o(o.transpose(1, 2).reshape(B, L, self.inner_dim))


    # ---- KV-cached decode (my_transformers/modeling_t5.py:535-613) --------------------------------------
    def _sh(self, t: torch.Tensor) -> torch.Tensor:
        return t.view(t.shape[0], -1, self.n_heads, self.d_kv).transpose(1, 2)

    def cross_kv(self, encoder_hidden_states, task=None):
        """Cross-attention heads for a whole generation, the values through the value parallel adapter (K2, forward only):
        the pair the reference keeps in ``past_key_value`` after the first step."""
        k, v = self.k(encoder_hidden_states), self.v(encoder_hidden_states)
        if self.attn_value_parallel_adapter is not None:
            v = self.attn_value_parallel_adapter(encoder_hidden_states, task, y=v)                  # K2
        return self._sh(k), self._sh(v)

    def self_kv(self, hidden_states, past=None):
        kh, vh = self._sh(self.k(hidden_states)), self._sh(self.v(hidden_states))
        if past is not None:
            kh, vh = torch.cat([past[0], kh], dim=2), torch.cat([past[1], vh], dim=2)
        return kh, vh

    def attend_cached(self, hidden_states, kh, vh, bias=None):
        B, T, _ = hidden_states.shape
        m = bias.to(hidden_states.dtype) if bias is not None else None
        o = F.scaled_dot_product_attention(self._sh(self.q(hidden_states)), kh, vh, attn_mask=m, scale=1.0)
        return self.o(o.transpose(1, 2).reshape(B, T, self.inner_dim))


def _rms(ln: T5LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """T5LayerNorm (my_transformers/modeling_t5.py:235-252: fp32 statistics, no mean subtraction, weight only).  Frozen-side
    plumbing: torch's fused rms_norm on the GPU (the weight is an fp32 master under --unfreeze_encoder_layer_norms)."""
    if x.is_cuda and x.dtype == torch.bfloat16:
        return F.rms_norm(x, (x.shape[-1],), ln.weight.to(x.dtype), ln.variance_epsilon)
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)      # "layer norm should always be calculated in float32"
    return ln.weight.to(x.dtype) * (x * torch.rsqrt(var + ln.variance_epsilon)).to(x.dtype)


class _PetSiteMixin:
    def _make_pet(self, config: VLPetConfig, site: str):
        d = config.d_model
        if not config.use_encoder_adapter_down_multihead:
            raise NotImplementedError("vlpet host model: the encoder PET path needs use_encoder_adapter_down_multihead")
        h = config.encoder_adapter_multihead_num_head
        hr = int(config.adapter_down_dim / h)
        setattr(self, f"{site}_adapter_multihead_down", nn.ModuleList([nn.Linear(d, hr) for _ in range(h)]))
        setattr(self, f"{site}_adapter_multihead_up", nn.Linear(config.adapter_down_dim, d))
        stem = f"encoder_{site}_adapter_gating"
        if config.use_encoder_adapter_gating_large_x_lowrank:
            setattr(self, stem + "_large_x_down", nn.Linear(d, config.adapter_gating_down_dim))
            setattr(self, stem + "_large_x_up", nn.Linear(config.adapter_gating_down_dim, d))
        if config.use_encoder_adapter_gating_small_xy_cat:
            setattr(self, stem + "_small_xy_cat", nn.Linear(2 * d, 1))
        if config.use_encoder_adapter_gating_middle_xy_add:
            setattr(self, stem + "_middle_xy_add", nn.Linear(d, 1))
        if config.use_encoder_adapter_gating_middle_ia3_add:
            setattr(self, stem + "_middle_ia3_add", nn.Parameter(torch.zeros(d).normal_(std=0.02)))
        self._vlpet_site_cfg = E.site_config(config, is_t5=True, impl=config.pet_impl)

    def _pet(self, site: str, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        return E.encoder_pet(self, site, x1, x2)


class T5LayerSelfAttention(nn.Module, _PetSiteMixin):
    def __init__(self, config: VLPetConfig, is_decoder: bool, has_relative_attention_bias: bool):
        super().__init__()
        self.config, self.is_decoder = config, is_decoder
        self.SelfAttention = T5Attention(config, is_decoder, has_relative_attention_bias, value_adapter=False)
        self.layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = config.dropout_rate
        if not is_decoder:
            self._make_pet(config, "attn")

    def forward(self, hidden_states, bias=None):
        a = self.SelfAttention(_rms(self.layer_norm, hidden_states), bias=bias)
        if self.is_decoder:
            return hidden_states + F.dropout(a, p=self.dropout, training=self.training)
        return self._pet("attn", hidden_states, a)                                   # K1 (modeling_t5.py:777-824)


class T5LayerCrossAttention(nn.Module):
    def __init__(self, config: VLPetConfig):
        super().__init__()
        self.EncDecAttention = T5Attention(config, True, False,
                                           value_adapter=config.use_decoder_enc_attn_value_parallel_adapter_down_dim)
        self.layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = config.dropout_rate

    def forward(self, hidden_states, encoder_hidden_states, bias=None, task=None, cross_k=None):
        a = self.EncDecAttention(_rms(self.layer_norm, hidden_states), key_value_states=encoder_hidden_states, bias=bias, task=task,
                                 k_override=cross_k)
        return hidden_states + F.dropout(a, p=self.dropout, training=self.training)


class T5DenseReluDense(nn.Module):
    def __init__(self, config: VLPetConfig):
        super().__init__()
        self.wi = nn.Linear(config.d_model, config.d_ff, bias=False)
        self.wo = nn.Linear(config.d_ff, config.d_model, bias=False)
        self.dropout = config.dropout_rate

    def forward(self, x):
        return self.wo(F.dropout(F.relu(self.wi(x)), p=self.dropout, training=self.training))


class T5LayerFF(nn.Module, _PetSiteMixin):
    def __init__(self, config: VLPetConfig, is_decoder: bool):
        super().__init__()
        self.config, self.is_decoder = config, is_decoder
        self.DenseReluDense = T5DenseReluDense(config)
        self.layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = config.dropout_rate
        if not is_decoder:
            self._make_pet(config, "ff")

    def forward(self, hidden_states):
        f = self.DenseReluDense(_rms(self.layer_norm, hidden_states))
        if self.is_decoder:
            return hidden_states + F.dropout(f, p=self.dropout, training=self.training)
        return self._pet("ff", hidden_states, f)                                     # K1 (modeling_t5.py:359-409)


class T5Block(nn.Module):
    def __init__(self, config: VLPetConfig, is_decoder: bool, has_relative_attention_bias: bool):
        super().__init__()
        self.is_decoder = is_decoder
        layers = [T5LayerSelfAttention(config, is_decoder, has_relative_attention_bias)]
        if is_decoder:
            layers.append(T5LayerCrossAttention(config))
        layers.append(T5LayerFF(config, is_decoder))
        self.layer = nn.ModuleList(layers)

    def forward(self, hidden_states, bias=None, encoder_hidden_states=None, cross_bias=None, task=None, cross_k=None):
        hidden_states = self.layer[0](hidden_states, bias=bias)
        if self.is_decoder:
            hidden_states = self.layer[1](hidden_states, encoder_hidden_states, bias=cross_bias, task=task, cross_k=cross_k)
        return self.layer[-1](hidden_states)


    @torch.no_grad()
    def step(self, hidden_states, cross, self_past=None, bias=None, cross_bias=None):
        """Incremental decoder block (eval mode): T new positions against the cached self-attention heads and the
        generation-long cross-attention heads.  Returns (hidden_states, new self-attention cache)."""
        sa, ca = self.layer[0], self.layer[1]
        kv = sa.SelfAttention.self_kv(_rms(sa.layer_norm, hidden_states), self_past)
        hidden_states = hidden_states + sa.SelfAttention.attend_cached(_rms(sa.layer_norm, hidden_states), kv[0], kv[1], bias)
        hidden_states = hidden_states + ca.EncDecAttention.attend_cached(_rms(ca.layer_norm, hidden_states), cross[0], cross[1], cross_bias)
        return self.layer[-1](hidden_states), kv


def _additive(mask_2d: Optional[torch.Tensor], dtype):
    """[B, S] 1/0 -> additive [B, 1, 1, S] (0 / -1e4-style large negative; the reference uses -10000.0)."""
    if mask_2d is None:
        return None
    return (1.0 - mask_2d[:, None, None, :].to(dtype)) * -10000.0


class JointEncoder(nn.Module):
    def __init__(self, config: VLPetConfig, embed_tokens: nn.Embedding):
        super().__init__()
        self.config = config
        self.embed_tokens = embed_tokens
        self.block = nn.ModuleList([T5Block(config, False, i == 0) for i in range(config.num_layers)])
        self.final_layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = config.dropout_rate
        self.visual_embedding = VisualEmbedding(config, embed_tokens, rms=True)
        self.downsample = None
        if config.downsample:
            s = int(config.n_boxes ** 0.5)
            self.downsample = Downsample((s, s))

    def forward(self, input_ids, vis_inputs, attention_mask=None, vis_attention_mask=None, task=None):
        B, L = input_ids.shape
        x = self.embed_tokens(input_ids)
        if self.downsample is not None:
            vis_inputs = self.downsample(vis_inputs)
        feats, boxes = vis_inputs[0], vis_inputs[1]
        img_ids = vis_inputs[2] if len(vis_inputs) >= 3 else None
        obj_ids = vis_inputs[3] if len(vis_inputs) == 4 else None
        vis = self.visual_embedding(feats.to(x.dtype), boxes, img_ids, obj_ids)                 # K3
        x = torch.cat([x, vis], dim=1)
        S = x.shape[1]
        mask = None
        if not self.config.assume_no_padding or attention_mask is not None or vis_attention_mask is not None:
            if attention_mask is None:
                attention_mask = input_ids.ne(self.config.pad_token_id)
            if vis_attention_mask is None:
                vis_attention_mask = attention_mask.new_ones(B, vis.shape[1])
            mask = torch.cat([attention_mask.to(vis_attention_mask.dtype), vis_attention_mask], dim=1)
        # relative position bias between text tokens only (src/modeling_t5.py:311-327)
        text_bias = self.block[0].layer[0].SelfAttention.compute_bias(L, L)
        bias = text_bias.new_zeros(1, text_bias.shape[1], S, S)
        bias[:, :, :L, :L] = text_bias
        if mask is not None:
            bias = bias + _additive(mask, bias.dtype)
        x = F.dropout(x, p=self.dropout, training=self.training)
        bias = bias.to(x.dtype)
        for blk in self.block:
            x = blk(x, bias=bias, task=task)
        x = _rms(self.final_layer_norm, x)
        return F.dropout(x, p=self.dropout, training=self.training), mask


class T5DecoderStack(nn.Module):
    def __init__(self, config: VLPetConfig, embed_tokens: nn.Embedding):
        super().__init__()
        self.config = config
        self.embed_tokens = embed_tokens
        self.block = nn.ModuleList([T5Block(config, True, i == 0) for i in range(config.num_decoder_layers)])
        self.final_layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = config.dropout_rate
        self._kcat = None          # (key, [layers * inner, d] weight): see _cross_keys

    def _cross_keys(self, encoder_hidden_states):
        """Cross-attention keys of ALL decoder blocks in one GEMM over the row-concatenated frozen k weights (every block
        projects the same encoder output): 12 GEMMs forward, 12 dgrad GEMMs + 11 accumulation passes over [tokens, d] backward
        become one GEMM each way + one stack (T5 twin of host.vlbart.BartDecoder._cross_keys)."""
        ks = [blk.layer[1].EncDecAttention.k for blk in self.block]
        if any(k.weight.requires_grad for k in ks):
            return None
        key = tuple((k.weight.data_ptr(), k.weight._version, k.weight.dtype) for k in ks)
        if self._kcat is None or self._kcat[0] != key:
            self._kcat = (key, torch.cat([k.weight.detach() for k in ks], 0).contiguous())
        w = self._kcat[1]
        if w.dtype != encoder_hidden_states.dtype and not torch.is_autocast_enabled():
            return None
        B, S, _ = encoder_hidden_states.shape
        return F.linear(encoder_hidden_states, w).view(B, S, len(ks), -1).unbind(2)

    def forward(self, input_ids, encoder_hidden_states, encoder_mask=None, task=None):
        B, T = input_ids.shape
        x = F.dropout(self.embed_tokens(input_ids), p=self.dropout, training=self.training)
        bias = self.block[0].layer[0].SelfAttention.compute_bias(T, T)
        causal = torch.ones(T, T, dtype=torch.bool, device=x.device).tril()
        bias = (bias.float() + torch.where(causal, 0.0, -10000.0)[None, None]).to(x.dtype)
        cross = None
        if encoder_mask is not None:
            cross = ((1.0 - encoder_mask[:, None, None, :].float()) * -1e9).to(x.dtype).expand(B, 1, T, -1)
        keys = self._cross_keys(encoder_hidden_states)
        for i, blk in enumerate(self.block):
            x = blk(x, bias=bias, encoder_hidden_states=encoder_hidden_states, cross_bias=cross, task=task,
                    cross_k=keys[i] if keys is not None else None)
        x = _rms(self.final_layer_norm, x)
        return F.dropout(x, p=self.dropout, training=self.training)


    @torch.no_grad()
    def init_cache(self, encoder_hidden_states, encoder_mask=None, task=None):
        cross = [blk.layer[1].EncDecAttention.cross_kv(encoder_hidden_states, task) for blk in self.block]
        return {"cross": cross, "self": [None] * len(self.block), "len": 0, "mask": encoder_mask}

    @torch.no_grad()
    def step(self, input_ids, cache):
        """``input_ids`` [B, T]: the positions after the ``cache['len']`` decoded ones -> normalised hidden states [B, T, d]."""
        B, T = input_ids.shape
        S = cache["len"] + T
        x = self.embed_tokens(input_ids)
        bias = self.block[0].layer[0].SelfAttention.compute_bias(S, S)[:, :, S - T:, :]
        causal = torch.ones(T, S, dtype=torch.bool, device=x.device).tril(diagonal=S - T)
        bias = (bias.float() + torch.where(causal, 0.0, -10000.0)[None, None]).to(x.dtype)
        cross = None
        if cache["mask"] is not None:
            cross = ((1.0 - cache["mask"][:, None, None, :].float()) * -1e9).to(x.dtype).expand(B, 1, T, -1)
        for i, blk in enumerate(self.block):
            x, cache["self"][i] = blk.step(x, cache["cross"][i], cache["self"][i], bias, cross)
        cache["len"] = S
        return _rms(self.final_layer_norm, x)


class VLT5(nn.Module):
    """LM head (tied, output scaled by d^-0.5) + token-level cross-entropy with reduction='none'
    (src/modeling_t5.py:655-690) and the per-task loss shaping of the reference task models."""

    def __init__(self, config: VLPetConfig):
        super().__init__()
        self.config = config
        self.model_dim = config.d_model
        self.shared = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = JointEncoder(config, self.shared)
        self.decoder = T5DecoderStack(config, self.shared)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.apply(self._init_weights)
        self.lm_head.weight = self.shared.weight
        self._nlvr_ids = {}

    def _init_weights(self, m):
        """my_transformers/modeling_t5.py:1026-1066 (factor 1.0).  PET Linears keep the default nn.Linear init: the reference's
        _init_weights does not touch them."""
        cfg = self.config
        if isinstance(m, T5LayerNorm):
            m.weight.data.fill_(1.0)
        elif isinstance(m, T5Attention):
            d, kv, h = cfg.d_model, cfg.d_kv, cfg.num_heads
            m.q.weight.data.normal_(mean=0.0, std=(d * kv) ** -0.5)
            m.k.weight.data.normal_(mean=0.0, std=d ** -0.5)
            m.v.weight.data.normal_(mean=0.0, std=d ** -0.5)
            m.o.weight.data.normal_(mean=0.0, std=(h * kv) ** -0.5)
            if m.has_relative_attention_bias:
                m.relative_attention_bias.weight.data.normal_(mean=0.0, std=d ** -0.5)
        elif isinstance(m, T5DenseReluDense):
            m.wi.weight.data.normal_(mean=0.0, std=cfg.d_model ** -0.5)
            m.wo.weight.data.normal_(mean=0.0, std=cfg.d_ff ** -0.5)
        elif isinstance(m, VLT5):
            m.shared.weight.data.normal_(mean=0.0, std=1.0)

    def _shift_right(self, labels):
        cfg = self.config
        s = labels.new_zeros(labels.shape)
        s[:, 1:] = labels[:, :-1]
        s[:, 0] = cfg.decoder_start_token_id
        return s.masked_fill(s == -100, cfg.pad_token_id)

    def forward(self, input_ids, vis_inputs, labels, attention_mask=None, vis_attention_mask=None, task=None):
        enc, mask = self.encoder(input_ids, vis_inputs, attention_mask, vis_attention_mask, task=task)
        h = self.decoder(self._shift_right(labels), enc, encoder_mask=mask, task=task)
        h = h * (self.model_dim ** -0.5)
        logits = self.lm_head(h)
        lg = logits.view(-1, logits.shape[-1])
        if F_.cross_entropy_supported(lg):
            loss = F_.cross_entropy_bf16(lg, labels.reshape(-1), -100)
        else:
            if lg.dtype in (torch.bfloat16, torch.float16):
                lg = lg.float()
            loss = F.cross_entropy(lg, labels.reshape(-1), ignore_index=-100, reduction="none")
        return loss, logits

    def _vis_inputs(self, batch: dict, dev):
        """(feats, boxes[, img_ids, obj_ids]) of a task batch on the device; NLVR pairs are flattened (nlvr_model.py:28-62)."""
        feats = batch["vis_feats"].to(dev, non_blocking=True)
        boxes = batch["boxes"].to(dev, non_blocking=True)
        if batch["task"] != "nlvr":
            return (feats, boxes)
        B, V_L = feats.shape[0], feats.shape[2]
        feats = feats.reshape(B, 2 * V_L, -1)
        boxes = boxes.reshape(B, 2 * V_L, 4)
        key = (V_L, str(dev))
        if key not in self._nlvr_ids:
            self._nlvr_ids[key] = (torch.tensor([0] * V_L + [1] * V_L, dtype=torch.long, device=dev).view(1, -1),
                                   torch.arange(V_L, dtype=torch.long, device=dev).repeat(2).view(1, -1))
        return (feats, boxes, self._nlvr_ids[key][0].expand(B, -1), self._nlvr_ids[key][1].expand(B, -1))

    @torch.no_grad()
    def generate(self, input_ids, vis_inputs, task=None, max_length: int = 20, min_length: int = 0, num_beams: int = 1,
                 logits_processor=None, attention_mask=None, vis_attention_mask=None, return_step_logits: bool = False,
                 length_penalty: float = 1.0, early_stopping: bool = False, return_scores: bool = False):
        """Greedy decoding through the KV cache (src/modeling_t5.py:560-690 with ``past_key_values``; T5 twin of
        host.VLBart.generate): encoder once, cross-attention heads -- values through the value parallel adapter (K2) -- once,
        one new position per step.  ``num_beams > 1``: host/generation.py's beam search over the same cached step."""
        cfg = self.config
        eos = getattr(cfg, "eos_token_id", 1)
        was_training = self.training
        self.eval()
        try:
            enc, mask = self.encoder(input_ids, vis_inputs, attention_mask, vis_attention_mask, task=task)
            B = input_ids.shape[0]
            if num_beams > 1:
                # beam search (caption: --num_beams 5): every sample's encoder output is expanded to num_beams rows, as the
                # reference's _expand_inputs_for_generation does; the self-attention caches follow the beams
                from .generation import beam_search
                rep = torch.arange(B, device=input_ids.device).repeat_interleave(num_beams)
                cache = self.decoder.init_cache(enc.index_select(0, rep), mask.index_select(0, rep) if mask is not None else None, task=task)

                def _step(new_tokens):
                    h = self.decoder.step(new_tokens, cache)
                    return self.lm_head(h[:, -1] * (self.model_dim ** -0.5))

                def _reorder(beam_idx):
                    cache["self"] = [(k.index_select(0, beam_idx), v.index_select(0, beam_idx)) for k, v in cache["self"]]

                return beam_search(_step, _reorder, B, num_beams, input_ids.device, cfg.decoder_start_token_id, cfg.pad_token_id,
                                   eos, max_length, min_length, length_penalty, early_stopping, logits_processor, return_scores)
            cache = self.decoder.init_cache(enc, mask, task=task)
            tokens = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long, device=input_ids.device)
            done = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
            steps = []
            while tokens.shape[1] < max_length:
                h = self.decoder.step(tokens if cache["len"] == 0 else tokens[:, -1:], cache)
                scores = self.lm_head(h[:, -1] * (self.model_dim ** -0.5))
                if scores.dtype in (torch.bfloat16, torch.float16):
                    scores = scores.float()
                if return_step_logits:
                    steps.append(scores.clone())
                if logits_processor is not None:
                    scores = logits_processor(tokens.shape[1] - 1, tokens, scores)
                if tokens.shape[1] < min_length:
                    scores[:, eos] = -float("inf")
                nxt = scores.argmax(-1)
                nxt = torch.where(done, torch.full_like(nxt, cfg.pad_token_id), nxt)
                done |= nxt == eos
                tokens = torch.cat([tokens, nxt[:, None]], dim=1)
                if bool(done.all()):
                    break
        finally:
            self.train(was_training)
        return (tokens, torch.stack(steps, 1)) if return_step_logits else tokens

    @torch.no_grad()
    def test_step(self, batch: dict, **gen_kwargs) -> dict:
        """vqa_model.py:112-164 (the T5 task model's generative test_step): batch -> {'token_ids'}."""
        dev = self.shared.weight.device
        out = self.generate(batch["input_ids"].to(dev, non_blocking=True), self._vis_inputs(batch, dev), task=batch["task"], **gen_kwargs)
        result = {"token_ids": out}
        tok = getattr(self, "tokenizer", None)
        if tok is not None:
            result["pred_ans"] = tok.batch_decode(out, skip_special_tokens=True)
        return result

    def train_step(self, batch: dict) -> dict:
        """Same batch schema and loss shaping as host.VLBart.train_step (vqa_model.py:44-110 etc. are the T5 twins)."""
        dev = self.shared.weight.device
        task = batch["task"]
        input_ids = batch["input_ids"].to(dev, non_blocking=True)
        labels = batch["target_ids"].to(dev, non_blocking=True)
        B = input_ids.shape[0]
        vis_inputs = self._vis_inputs(batch, dev)
        loss, _ = self(input_ids, vis_inputs, labels, task=task)
        T = labels.shape[1]
        mask = (labels != -100).float()
        loss = loss.view(B, T) * mask
        if task in ("caption", "tvc", "yc2c"):
            loss = loss.sum() / mask.sum().clamp(min=1)
        else:
            loss = loss.sum(dim=1) / mask.sum(dim=1).clamp(min=1)
            if "scores" in batch and batch["scores"] is not None:
                loss = loss * batch["scores"].to(dev, non_blocking=True)
            loss = loss.mean()
        return {"loss": loss}
