"""Host model that CALLS the PET hot path: a frozen BART-base encoder/decoder in stock PyTorch whose module tree,
parameter names and op order follow the reference's VLBart (src/modeling_bart.py:696-898 JointEncoder, 1296-1455
VLBartModel, 1458-1602 VLBart; src/my_transformers/modeling_bart.py:143-280 attention, 882-1388 encoder layer,
1391-1788 decoder layer, 1953-2358 encoder / decoder stacks), so a reference ``state_dict`` loads key-for-key and
the name-substring unfreezing rules of trainer_base.py:308-542 select the same parameters.

Only the three PET op groups are ours (CUDA, include/vlpet.h): the two encoder PET sites per layer (K1), the
decoder cross-attention value parallel adapter (K2) and the visual projection (K3).  Everything else here is
plumbing in stock PyTorch (nn.Linear -> cuBLAS, F.scaled_dot_product_attention, nn.LayerNorm): it exists so that
BASELINE.json's metric (multitask samples/s on BART-base) can be measured and so that the drop-in claim is tested
end to end; it is not a re-implementation of the reference's trainer, data pipeline or generation code.
"""
from __future__ import annotations

import contextlib
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import encoder as E
from .. import functional as F_
from ..adapters import AdapterController
from ..visual import VisualEmbedding
from .config import VLPetConfig

try:
    from torch.nn.attention import SDPBackend as _SDPBackend, sdpa_kernel as _sdpa_kernel
    _SDPA_ORDER = [_SDPBackend.EFFICIENT_ATTENTION, _SDPBackend.CUDNN_ATTENTION, _SDPBackend.FLASH_ATTENTION, _SDPBackend.MATH]
except Exception:                                                    # older torch: leave the default selection
    _SDPA_ORDER = None


SHORT_ATTENTION = True     # frozen attention through vlpet_attn_fwd/_bwd when both sequence lengths are <= 64 (register-resident
                           # kernels: 187 vs 314 us forward + backward per encoder call at bs = 300); longer sequences (NLVR: 92)
                           # and masked attention stay on torch SDPA


def _ln(ln: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """LayerNorm whose (trainable, fp32-master) affine parameters may be wider than the activation dtype.  On the GPU in
    bf16 it runs the one-launch row kernels of libvlpet.so; anywhere else stock F.layer_norm."""
    if x.is_cuda and x.dtype == torch.bfloat16 and F_.layer_norm_supported(x):
        return F_.layer_norm(x, ln.weight, ln.bias, ln.eps)
    w, b = ln.weight, ln.bias
    if w.dtype != x.dtype:
        w, b = w.to(x.dtype), b.to(x.dtype)
    return F.layer_norm(x, ln.normalized_shape, w, b, ln.eps)


def _drop_add_ln(ln: nn.LayerNorm, res: torch.Tensor, h: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """LayerNorm(res + dropout(h)): the post-LN residual step (my_transformers/modeling_bart.py:1663-1665 etc.).  On the GPU in
    bf16 one kernel each way (vlpet_dropout_add_layernorm_fwd / _bwd) instead of dropout, add and LayerNorm."""
    if h.is_cuda and h.dtype == torch.bfloat16 and res.dtype == torch.bfloat16 and F_.layer_norm_supported(h):
        return F_.dropout_add_layer_norm(h, res, ln.weight, ln.bias, ln.eps, p, training)
    return _ln(ln, res + F.dropout(h, p=p, training=training))


def shift_tokens_right(input_ids: torch.Tensor, pad_token_id: int, decoder_start_token_id: int) -> torch.Tensor:
    """Decoder inputs from labels (my_transformers/modeling_bart.py:63-77): shift right, start token first,
    -100 -> pad."""
    shifted = input_ids.new_empty(input_ids.shape)
    shifted[:, 1:] = input_ids[:, :-1]
    shifted[:, 0] = decoder_start_token_id
    return shifted.masked_fill(shifted == -100, pad_token_id)


class BartLearnedPositionalEmbedding(nn.Embedding):
    """Learned positions with BART's +2 offset (my_transformers/modeling_bart.py:122-140)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: int):
        self.offset = 2
        super().__init__(num_embeddings + self.offset, embedding_dim, padding_idx=padding_idx)

    def forward(self, seq_len: int, past: int = 0):
        pos = torch.arange(past, past + seq_len, dtype=torch.long, device=self.weight.device)
        return super().forward(pos + self.offset)


class BartAttention(nn.Module):
    """Multi-head attention with the reference's parameter names (q/k/v/out_proj).  Cross-attention instances carry
    ``attn_value_parallel_adapter`` (AdapterController, K2) exactly where BartAttentionWithValueAdapter does
    (my_transformers/modeling_bart.py:329-340, 427-430)."""

    def __init__(self, config: VLPetConfig, num_heads: int, is_decoder: bool, value_adapter: bool):
        super().__init__()
        d = config.d_model
        self.embed_dim, self.num_heads, self.head_dim = d, num_heads, d // num_heads
        self.dropout = config.attention_dropout
        self.is_decoder = is_decoder
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)
        self.attn_value_parallel_adapter = AdapterController(config.vpa_adapter_config()) if value_adapter else None
        self._qkv = None      # (key, [3d, d] weight, [3d] bias): fused projection of a FROZEN self-attention, see _fused_qkv

    def _fused_qkv(self):
        """Self-attention with frozen projections on the GPU: q, k, v come out of ONE GEMM over the row-concatenated
        weights (a cached copy; the state-dict keeps the reference's three Linears), and the backward is one dgrad GEMM
        instead of three plus two gradient-accumulation passes over [tokens, d]."""
        ws = (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight)
        bs = (self.q_proj.bias, self.k_proj.bias, self.v_proj.bias)
        if any(t.requires_grad for t in ws + bs) or not ws[0].is_cuda:
            return None
        key = tuple((t.data_ptr(), t._version, t.dtype) for t in ws + bs)
        if self._qkv is None or self._qkv[0] != key:
            self._qkv = (key, torch.cat([t.detach() for t in ws], 0).contiguous(), torch.cat([t.detach() for t in bs], 0).contiguous())
        return self._qkv[1], self._qkv[2]

    def _heads(self, t: torch.Tensor) -> torch.Tensor:
        B, L, _ = t.shape
        return t.view(B, L, self.num_heads, self.head_dim).transpose(1, 2)

    def forward(self, hidden_states, key_value_states=None, attn_mask=None, is_causal=False, task=None, k_override=None):
        src = hidden_states if key_value_states is None else key_value_states
        fused = self._fused_qkv() if key_value_states is None else None
        B, L, _ = hidden_states.shape
        if fused is not None:
            qkv = F.linear(hidden_states, fused[0], fused[1]).view(B, L, 3, self.embed_dim)
            if attn_mask is None and self.head_dim == 64 and SHORT_ATTENTION and L <= 64 and qkv.dtype == torch.bfloat16:
                # fused projection -> attention -> out_proj: the backward writes dq | dk | dv into one buffer
                return self.out_proj(F_.short_self_attention(qkv, self.num_heads, is_causal, self.dropout, self.training))
            q, k, v = qkv.unbind(2)
        else:
            q = self.q_proj(hidden_states)
            k = k_override if k_override is not None else self.k_proj(src)      # (cross-attention keys of all layers: BartDecoder._cross_keys)
            v = self.v_proj(src)
            if key_value_states is not None and self.attn_value_parallel_adapter is not None:
                v = self.attn_value_parallel_adapter(key_value_states, task, y=v)          # K2
        if attn_mask is None and self.head_dim == 64 and SHORT_ATTENTION and F_.short_attention_supported(q, k, v, self.num_heads) \
                and F_.short_attention_profitable(q, k):
            # one CTA per (batch, head), whole score tile in shared memory (include/vlpet.h vlpet_attn_fwd); output is
            # already [B, L, d]: no head transposes either way
            o = F_.short_attention(q, k, v, self.num_heads, is_causal, self.dropout, self.training)
            return self.out_proj(o)
        qh, kh, vh = self._heads(q), self._heads(k), self._heads(v)
        with _sdpa_policy(qh):
            o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=attn_mask,
                                               dropout_p=self.dropout if self.training else 0.0, is_causal=is_causal)
        return self.out_proj(o.transpose(1, 2).reshape(B, L, self.embed_dim))


    # ---- KV-cached decode (my_transformers/modeling_bart.py:419-430, 486-508) -------------------------
    def cross_kv(self, encoder_hidden_states, task=None):
        """Keys / values of a cross-attention for a whole generation: computed once, the values already through the value
        parallel adapter (K2, forward only) -- the reference caches exactly this pair in ``past_key_value``."""
        k = self.k_proj(encoder_hidden_states)
        v = self.v_proj(encoder_hidden_states)
        if self.attn_value_parallel_adapter is not None:
            v = self.attn_value_parallel_adapter(encoder_hidden_states, task, y=v)                  # K2
        return self._heads(k), self._heads(v)

    def attend_cached(self, hidden_states, kh, vh, attn_mask=None, is_causal=False):
        """``hidden_states`` [B, T, d] (T new positions) against cached heads kh / vh [B, H, S, hd]."""
        B, T, _ = hidden_states.shape
        qh = self._heads(self.q_proj(hidden_states))
        o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=attn_mask, is_causal=is_causal and T > 1 and attn_mask is None)
        return self.out_proj(o.transpose(1, 2).reshape(B, T, self.embed_dim))

    def self_kv(self, hidden_states, past=None):
        """Self-attention keys / values of the new positions appended to the cache (past = (kh, vh) or None)."""
        kh, vh = self._heads(self.k_proj(hidden_states)), self._heads(self.v_proj(hidden_states))
        if past is not None:
            kh, vh = torch.cat([past[0], kh], dim=2), torch.cat([past[1], vh], dim=2)
        return kh, vh


def _sdpa_policy(q: torch.Tensor):
    """Backend order for the frozen attention (stock PyTorch SDPA, host code).  At the sequence lengths of this workload
    (<= 92 encoder tokens, <= 40 decoder tokens) the memory-efficient kernel beats the cuDNN / flash kernels, whose
    128-wide tiles are mostly padding (tools/sdpa_probe.py: 0.57 vs 0.60 ms encoder, 0.27 vs 0.38 ms decoder self,
    forward + backward)."""
    if q.is_cuda and _SDPA_ORDER is not None:
        try:
            return _sdpa_kernel(_SDPA_ORDER, set_priority=True)
        except TypeError:
            pass
    return contextlib.nullcontext()


def _ffn_act(layer, h: torch.Tensor) -> torch.Tensor:
    """activation_fn + activation dropout of the frozen FFN (my_transformers/modeling_bart.py:1264-1266): one fused CUDA
    kernel each way for bf16 gelu, stock PyTorch otherwise."""
    if layer.activation_fn is F.gelu and F_.gelu_dropout_supported(h):
        return F_.gelu_dropout(h, layer.activation_dropout, layer.training)
    return F.dropout(layer.activation_fn(h), p=layer.activation_dropout, training=layer.training)


def _act(name: str):
    if name == "gelu":
        return F.gelu
    if name == "relu":
        return F.relu
    raise ValueError(name)


class BartEncoderLayer(nn.Module):
    """Post-LN encoder block with the two VL-PET sites.  PET parameters live on the layer under the reference's
    attribute names (my_transformers/modeling_bart.py:976-1056); the forward replaces the inline
    op sequence of modeling_bart.py:1145-1261 / 1268-1377 by ``encoder_pet`` (K1)."""

    def __init__(self, config: VLPetConfig):
        super().__init__()
        self.config = config
        d = config.d_model
        self.embed_dim = d
        self.self_attn = BartAttention(config, config.encoder_attention_heads, is_decoder=False, value_adapter=False)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.dropout = config.dropout
        self.activation_fn = _act(config.activation_function)
        self.activation_dropout = config.activation_dropout
        self.fc1 = nn.Linear(d, config.encoder_ffn_dim)
        self.fc2 = nn.Linear(config.encoder_ffn_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)
        if not config.use_encoder_adapter_down_multihead:
            raise NotImplementedError("vlpet host model: the encoder PET path needs use_encoder_adapter_down_multihead")
        h = config.encoder_adapter_multihead_num_head
        hr = int(config.adapter_down_dim / h)
        for site in ("attn", "ff"):
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
        self._vlpet_site_cfg = E.site_config(config, is_t5=False, impl=config.pet_impl)
        if config.pet_bwd_impl != "auto":
            import dataclasses
            self._vlpet_site_cfg = dataclasses.replace(self._vlpet_site_cfg, bwd_impl=config.pet_bwd_impl)

    def _pet(self, site: str, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        return E.encoder_pet(self, site, x1, x2)

    def forward(self, hidden_states, attn_mask=None, task=None):
        x1 = hidden_states
        x2 = self.self_attn(hidden_states, attn_mask=attn_mask)
        hidden_states = _ln(self.self_attn_layer_norm, self._pet("attn", x1, x2))
        x1 = hidden_states
        x2 = self.fc2(_ffn_act(self, self.fc1(hidden_states)))
        return _ln(self.final_layer_norm, self._pet("ff", x1, x2))


class BartDecoderLayer(nn.Module):
    """Frozen post-LN decoder block; the only PET op is the value parallel adapter inside ``encoder_attn``
    (my_transformers/modeling_bart.py:1452-1464 wiring, 1611-1788 forward)."""

    def __init__(self, config: VLPetConfig):
        super().__init__()
        d = config.d_model
        self.self_attn = BartAttention(config, config.decoder_attention_heads, is_decoder=True, value_adapter=False)
        self.dropout = config.dropout
        self.activation_fn = _act(config.activation_function)
        self.activation_dropout = config.activation_dropout
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = BartAttention(config, config.decoder_attention_heads, is_decoder=True,
                                          value_adapter=config.use_decoder_enc_attn_value_parallel_adapter_down_dim)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, config.decoder_ffn_dim)
        self.fc2 = nn.Linear(config.decoder_ffn_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)

    def forward(self, hidden_states, encoder_hidden_states, self_mask=None, cross_mask=None, task=None, cross_k=None):
        h = self.self_attn(hidden_states, attn_mask=self_mask, is_causal=self_mask is None)
        hidden_states = _drop_add_ln(self.self_attn_layer_norm, hidden_states, h, self.dropout, self.training)
        h = self.encoder_attn(hidden_states, key_value_states=encoder_hidden_states, attn_mask=cross_mask, task=task, k_override=cross_k)
        hidden_states = _drop_add_ln(self.encoder_attn_layer_norm, hidden_states, h, self.dropout, self.training)
        h = self.fc2(_ffn_act(self, self.fc1(hidden_states)))
        return _drop_add_ln(self.final_layer_norm, hidden_states, h, self.dropout, self.training)

    @torch.no_grad()
    def step(self, hidden_states, cross, self_past=None, cross_mask=None):
        """Incremental form of ``forward`` (eval mode): T new positions, cached self-attention heads, cross = (kh, vh) from
        ``encoder_attn.cross_kv``.  Returns (hidden_states, new self-attention cache)."""
        T = hidden_states.shape[1]
        kv = self.self_attn.self_kv(hidden_states, self_past)
        mask = None
        if self_past is not None and T > 1:                       # new positions see the whole past and a causal triangle
            S = kv[0].shape[2]
            mask = torch.ones(T, S, dtype=torch.bool, device=hidden_states.device).tril(diagonal=S - T)
        h = self.self_attn.attend_cached(hidden_states, kv[0], kv[1], attn_mask=mask, is_causal=self_past is None)
        hidden_states = _ln(self.self_attn_layer_norm, hidden_states + h)
        h = self.encoder_attn.attend_cached(hidden_states, cross[0], cross[1], attn_mask=cross_mask)
        hidden_states = _ln(self.encoder_attn_layer_norm, hidden_states + h)
        h = self.fc2(self.activation_fn(self.fc1(hidden_states)))
        return _ln(self.final_layer_norm, hidden_states + h), kv



class Downsample(nn.Module):
    """CLIP 7x7 grid -> sqrt(n_boxes)^2 by adaptive max-pool (src/modeling_bart.py:556-613); the NLVR 4-tuple pools
    each of the two images separately."""

    def __init__(self, output_size: Tuple[int, int]):
        super().__init__()
        self.output_size = output_size
        self.out_dtype = None        # set by the trainer: pool and cast to the compute dtype in one CUDA kernel

    def _pool(self, x):
        B, L, dim = x.shape
        if x.is_cuda and not x.requires_grad and dim % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16) and \
                self.output_size[0] == self.output_size[1]:
            from .. import functional as F_
            return F_.grid_maxpool(x, self.output_size[0], self.out_dtype or x.dtype)
        s = int(L ** 0.5)
        x = F.adaptive_max_pool2d(x.permute(0, 2, 1).reshape(B, dim, s, s), self.output_size)
        return x.reshape(B, dim, -1).permute(0, 2, 1)

    def forward(self, vis_inputs):
        if len(vis_inputs) == 4:
            feats, boxes, img_ids, obj_ids = vis_inputs
            B, L2, dim = feats.shape
            pooled = self._pool(feats.reshape(B * 2, L2 // 2, dim))
            n = pooled.shape[1]
            cut = lambda t: t.reshape(B, 2, L2 // 2, *t.shape[2:])[:, :, :n].reshape(B, 2 * n, *t.shape[2:])  # noqa: E731
            return pooled.reshape(B, 2 * n, dim), cut(boxes), cut(img_ids), cut(obj_ids)
        feats, boxes = vis_inputs
        feats = self._pool(feats)
        return feats, boxes[:, :feats.shape[1]]


def _pad_mask(mask_2d: Optional[torch.Tensor], dtype, tgt_len: int):
    """[B, S] 1/0 mask -> boolean [B, 1, T, S] for SDPA (True = attend); None stays None."""
    if mask_2d is None:
        return None
    B, S = mask_2d.shape
    return mask_2d.bool()[:, None, None, :].expand(B, 1, tgt_len, S)


class JointEncoder(nn.Module):
    """Text tokens (+ visual tokens appended) through the encoder stack (src/modeling_bart.py:696-898)."""

    def __init__(self, config: VLPetConfig, embed_tokens: nn.Embedding):
        super().__init__()
        self.config = config
        d = config.d_model
        self.dropout = config.dropout
        self.embed_scale = math.sqrt(d) if config.scale_embedding else 1.0
        self.embed_tokens = embed_tokens
        self.embed_positions = BartLearnedPositionalEmbedding(config.max_position_embeddings, d, config.pad_token_id)
        self.layers = nn.ModuleList([BartEncoderLayer(config) for _ in range(config.encoder_layers)])
        self.layernorm_embedding = nn.LayerNorm(d)
        self.visual_embedding = VisualEmbedding(config, self.embed_tokens)
        self.downsample = None
        if config.downsample:
            s = int(config.n_boxes ** 0.5)
            self.downsample = Downsample((s, s))

    def forward(self, input_ids, vis_inputs, attention_mask=None, vis_attention_mask=None, task=None):
        B, L = input_ids.shape
        x = self.embed_tokens(input_ids) * self.embed_scale + self.embed_positions(L)
        if self.downsample is not None:
            vis_inputs = self.downsample(vis_inputs)
        feats, boxes = vis_inputs[0], vis_inputs[1]
        img_ids = vis_inputs[2] if len(vis_inputs) >= 3 else None
        obj_ids = vis_inputs[3] if len(vis_inputs) == 4 else None
        vis = self.visual_embedding(feats.to(x.dtype), boxes, img_ids, obj_ids)                  # K3
        if self.config.share_vis_lang_layer_norm:
            x = _ln(self.layernorm_embedding, torch.cat([x, vis], dim=1))
        else:
            x = torch.cat([_ln(self.layernorm_embedding, x), vis], dim=1)
        mask = None
        if not self.config.assume_no_padding or attention_mask is not None or vis_attention_mask is not None:
            if attention_mask is None:
                attention_mask = input_ids.ne(self.config.pad_token_id)
            if vis_attention_mask is None:
                vis_attention_mask = attention_mask.new_ones(B, vis.shape[1])
            mask = torch.cat([attention_mask.to(vis_attention_mask.dtype), vis_attention_mask], dim=1)
        x = F.dropout(x, p=self.dropout, training=self.training)
        sdpa_mask = _pad_mask(mask, x.dtype, x.shape[1])
        for layer in self.layers:
            x = layer(x, attn_mask=sdpa_mask, task=task)
        return x, mask


class BartDecoder(nn.Module):
    def __init__(self, config: VLPetConfig, embed_tokens: nn.Embedding):
        super().__init__()
        self.config = config
        d = config.d_model
        self.dropout = config.dropout
        self.embed_scale = math.sqrt(d) if config.scale_embedding else 1.0
        self.embed_tokens = embed_tokens
        self.embed_positions = BartLearnedPositionalEmbedding(config.max_position_embeddings, d, config.pad_token_id)
        self.layers = nn.ModuleList([BartDecoderLayer(config) for _ in range(config.decoder_layers)])
        self.layernorm_embedding = nn.LayerNorm(d)
        self._kcat = None          # (key, [layers * d, d] weight, [layers * d] bias): see _cross_keys

    def _cross_keys(self, encoder_hidden_states):
        """Cross-attention keys of ALL decoder layers in one GEMM over the row-concatenated (frozen) k_proj weights: every layer
        projects the same encoder output, so 6 GEMMs forward and 6 dgrad GEMMs + 5 gradient-accumulation passes over
        [tokens, d] backward become one GEMM each way (+ one stack of the per-layer key gradients).  The state-dict keeps the
        reference's per-layer Linears; the concatenation is a cached copy."""
        ks = [layer.encoder_attn.k_proj for layer in self.layers]
        if any(k.weight.requires_grad or k.bias.requires_grad for k in ks):
            return None
        key = tuple((k.weight.data_ptr(), k.weight._version, k.weight.dtype, k.bias._version) for k in ks)
        if self._kcat is None or self._kcat[0] != key:
            self._kcat = (key, torch.cat([k.weight.detach() for k in ks], 0).contiguous(),
                          torch.cat([k.bias.detach() for k in ks], 0).contiguous())
        w, b = self._kcat[1], self._kcat[2]
        if w.dtype != encoder_hidden_states.dtype and not torch.is_autocast_enabled():
            return None
        B, S, d = encoder_hidden_states.shape
        return F.linear(encoder_hidden_states, w, b).view(B, S, len(ks), d).unbind(2)

    def forward(self, input_ids, encoder_hidden_states, encoder_mask=None, task=None):
        B, T = input_ids.shape
        x = self.embed_tokens(input_ids) * self.embed_scale + self.embed_positions(T)
        x = _ln(self.layernorm_embedding, x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        cross = _pad_mask(encoder_mask, x.dtype, T)
        keys = self._cross_keys(encoder_hidden_states)
        for i, layer in enumerate(self.layers):
            x = layer(x, encoder_hidden_states, self_mask=None, cross_mask=cross, task=task, cross_k=keys[i] if keys is not None else None)
        return x


    @torch.no_grad()
    def init_cache(self, encoder_hidden_states, encoder_mask=None, task=None):
        """Per-layer cross-attention heads for a whole generation (the value parallel adapter runs here, once) + empty
        self-attention caches."""
        cross = [layer.encoder_attn.cross_kv(encoder_hidden_states, task) for layer in self.layers]
        return {"cross": cross, "self": [None] * len(self.layers), "len": 0, "mask": encoder_mask}

    @torch.no_grad()
    def step(self, input_ids, cache):
        """``input_ids`` [B, T]: the positions after the ``cache['len']`` already decoded ones -> hidden states [B, T, d]."""
        B, T = input_ids.shape
        x = self.embed_tokens(input_ids) * self.embed_scale + self.embed_positions(T, past=cache["len"])
        x = _ln(self.layernorm_embedding, x)
        cross_mask = _pad_mask(cache["mask"], x.dtype, T)
        for i, layer in enumerate(self.layers):
            x, cache["self"][i] = layer.step(x, cache["cross"][i], cache["self"][i], cross_mask)
        cache["len"] += T
        return x


class VLBartModel(nn.Module):
    def __init__(self, config: VLPetConfig):
        super().__init__()
        self.config = config
        self.shared = nn.Embedding(config.vocab_size, config.d_model, config.pad_token_id)
        self.encoder = JointEncoder(config, self.shared)
        self.decoder = BartDecoder(config, self.shared)

    def forward(self, input_ids, vis_inputs, decoder_input_ids, attention_mask=None, vis_attention_mask=None, task=None):
        enc, mask = self.encoder(input_ids, vis_inputs, attention_mask, vis_attention_mask, task=task)
        return self.decoder(decoder_input_ids, enc, encoder_mask=mask, task=task)


class VLBart(nn.Module):
    """LM head + token-level cross-entropy with reduction='none' (src/modeling_bart.py:1522-1602) and the task
    ``train_step`` loss shaping of vqa_model.py:167-233 / nlvr_model.py:140-262 / caption (plain mean over tokens)."""

    def __init__(self, config: VLPetConfig):
        super().__init__()
        self.config = config
        self.model = VLBartModel(config)
        self.register_buffer("final_logits_bias", torch.zeros(1, config.vocab_size))
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.apply(self._init_weights)
        self.lm_head.weight = self.model.shared.weight           # tied, as BartForConditionalGeneration
        self._lm_pad = None                                      # (key, padded weight, padded bias) cache, see _lm_operands
        self._nlvr_ids = {}

    def _init_weights(self, m):
        std = self.config.init_std                               # my_transformers/modeling_bart.py:1819-1828
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()

    def _lm_operands(self, dtype):
        """LM-head weight / bias for the logits GEMM.  The reference vocabulary (50 265 + 200 = 50 465 rows) is odd, which
        sends cuBLAS to an unaligned legacy kernel; while the table is frozen on the GPU (always, under the VL-PET flags)
        the GEMM runs on a copy padded to a multiple of 64 rows whose extra logits get a -1e4 bias, i.e. exp() == 0: the
        cross-entropy over the padded width equals the reference's over `vocab_size` classes."""
        w, V = self.lm_head.weight, self.config.vocab_size
        if w.requires_grad or not w.is_cuda or V % 64 == 0:
            return w, self.final_logits_bias.to(dtype), V
        key = (w.data_ptr(), w.dtype, w._version, self.final_logits_bias._version)
        if self._lm_pad is None or self._lm_pad[0] != key:
            Vp = (V + 63) // 64 * 64
            wp = w.new_zeros(Vp, w.shape[1])
            wp[:V] = w.detach()
            bp = torch.full((1, Vp), -1e4, dtype=w.dtype, device=w.device)
            bp[:, :V] = self.final_logits_bias.to(w.dtype)
            self._lm_pad = (key, wp, bp)
        return self._lm_pad[1], self._lm_pad[2].to(dtype), self._lm_pad[1].shape[0]

    def forward(self, input_ids, vis_inputs, labels, attention_mask=None, vis_attention_mask=None, task=None):
        """-> per-token loss [B*T] (reduction='none', ignore_index=-100) and logits."""
        cfg = self.config
        dec_in = shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)
        h = self.model(input_ids, vis_inputs, dec_in, attention_mask, vis_attention_mask, task=task)
        w, b, width = self._lm_operands(h.dtype)
        logits = F.linear(h, w, b.reshape(-1))          # bias folded into the GEMM epilogue: no extra pass over the logits
        lg = logits.view(-1, width)
        if F_.cross_entropy_supported(lg):     # bf16 on the GPU: fused kernel, no fp32 copy of the [tokens, vocab] logits
            loss = F_.cross_entropy_bf16(lg, labels.reshape(-1), -100)
        else:
            if lg.dtype in (torch.bfloat16, torch.float16):
                lg = lg.float()
            loss = F.cross_entropy(lg, labels.reshape(-1), ignore_index=-100, reduction="none")
        return loss, logits[..., :cfg.vocab_size]

    def _vis_inputs(self, batch: dict, dev):
        """(feats, boxes[, img_ids, obj_ids]) of a task batch on the device; NLVR pairs are flattened (nlvr_model.py:156-176)."""
        feats = batch["vis_feats"].to(dev, non_blocking=True)
        boxes = batch["boxes"].to(dev, non_blocking=True)
        if batch["task"] != "nlvr":
            return (feats, boxes)
        B, V_L = feats.shape[0], feats.shape[2]
        feats = feats.reshape(B, 2 * V_L, -1)
        boxes = boxes.reshape(B, 2 * V_L, 4)
        key = (V_L, str(dev))
        if key not in self._nlvr_ids:                        # built once: no host->device copy inside a step
            self._nlvr_ids[key] = (torch.tensor([0] * V_L + [1] * V_L, dtype=torch.long, device=dev).view(1, -1),
                                   torch.arange(V_L, dtype=torch.long, device=dev).repeat(2).view(1, -1))
        return (feats, boxes, self._nlvr_ids[key][0].expand(B, -1), self._nlvr_ids[key][1].expand(B, -1))

    def _logits(self, h):
        w, b, _ = self._lm_operands(h.dtype)
        return F.linear(h, w, b.reshape(-1))[..., :self.config.vocab_size]

    @torch.no_grad()
    def generate(self, input_ids, vis_inputs, task=None, max_length: int = 20, min_length: int = 0, num_beams: int = 1,
                 logits_processor=None, attention_mask=None, vis_attention_mask=None, return_step_logits: bool = False,
                 length_penalty: float = 1.0, early_stopping: bool = False, return_scores: bool = False):
        """Greedy decoding with the KV cache of the reference's decode path (src/modeling_bart.py:1522-1602 with
        ``past_key_values``; what ``test_step`` -> ``generate(num_beams=1)`` runs for VQA / GQA / NLVR, multitask.py:480, 516):
        the encoder runs once, every decoder layer's cross-attention keys / values -- the values through the value parallel
        adapter (K2) -- are formed once, each step feeds ONE new position.  ``logits_processor(step, tokens, scores)`` is the
        caller's hook (HF's logits processors); ``min_length`` masks EOS as HF's MinLengthLogitsProcessor does.
        ``num_beams > 1`` (caption: --num_beams 5, multitask.py:587) runs host/generation.py's beam search over the same
        cached step (the processor then sees log-probabilities, as in HF)."""
        cfg = self.config
        was_training = self.training
        self.eval()
        try:
            enc, mask = self.model.encoder(input_ids, vis_inputs, attention_mask, vis_attention_mask, task=task)
            B = input_ids.shape[0]
            if num_beams > 1:
                # beam search (caption: --num_beams 5): every sample's encoder output is expanded to num_beams rows, as the
                # reference's _expand_inputs_for_generation does; the self-attention caches follow the beams
                from .generation import beam_search
                rep = torch.arange(B, device=input_ids.device).repeat_interleave(num_beams)
                cache = self.model.decoder.init_cache(enc.index_select(0, rep), mask.index_select(0, rep) if mask is not None else None, task=task)

                def _step(new_tokens):
                    h = self.model.decoder.step(new_tokens, cache)
                    return self._logits(h[:, -1])

                def _reorder(beam_idx):
                    cache["self"] = [(k.index_select(0, beam_idx), v.index_select(0, beam_idx)) for k, v in cache["self"]]

                return beam_search(_step, _reorder, B, num_beams, input_ids.device, cfg.decoder_start_token_id, cfg.pad_token_id,
                                   cfg.eos_token_id, max_length, min_length, length_penalty, early_stopping, logits_processor, return_scores)
            cache = self.model.decoder.init_cache(enc, mask, task=task)
            tokens = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long, device=input_ids.device)
            done = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
            steps = []
            while tokens.shape[1] < max_length:
                h = self.model.decoder.step(tokens if cache["len"] == 0 else tokens[:, -1:], cache)
                scores = self._logits(h[:, -1])
                if scores.dtype in (torch.bfloat16, torch.float16):
                    scores = scores.float()
                if return_step_logits:
                    steps.append(scores.clone())
                if logits_processor is not None:
                    scores = logits_processor(tokens.shape[1] - 1, tokens, scores)     # (step index, tokens so far, scores)
                if tokens.shape[1] < min_length:
                    scores[:, cfg.eos_token_id] = -float("inf")
                nxt = scores.argmax(-1)
                nxt = torch.where(done, torch.full_like(nxt, cfg.pad_token_id), nxt)
                done |= nxt == cfg.eos_token_id
                tokens = torch.cat([tokens, nxt[:, None]], dim=1)
                if bool(done.all()):
                    break
        finally:
            self.train(was_training)
        return (tokens, torch.stack(steps, 1)) if return_step_logits else tokens

    @torch.no_grad()
    def test_step(self, batch: dict, **gen_kwargs) -> dict:
        """vqa_model.py:235-288 (generative branch): batch -> {'token_ids'} (and 'pred_ans' when a tokenizer was attached as
        ``self.tokenizer``; tokenizers need files this offline image does not have)."""
        dev = self.model.shared.weight.device
        out = self.generate(batch["input_ids"].to(dev, non_blocking=True), self._vis_inputs(batch, dev), task=batch["task"],
                            **gen_kwargs)
        result = {"token_ids": out}
        tok = getattr(self, "tokenizer", None)
        if tok is not None:
            result["pred_ans"] = tok.batch_decode(out, skip_special_tokens=True)
        return result

    def train_step(self, batch: dict) -> dict:
        """One task batch -> {'loss': scalar}.  Batch schema = the reference collate (vqa_clip_data.py:365-390):
        input_ids [B,Lt] int64, vis_feats [B,N,F] ([B,2,N,F] for nlvr), boxes [B,N,4], target_ids [B,T] int64 with
        -100 padding, optional scores [B], task str."""
        dev = self.model.shared.weight.device
        task = batch["task"]
        input_ids = batch["input_ids"].to(dev, non_blocking=True)
        feats = batch["vis_feats"].to(dev, non_blocking=True)
        boxes = batch["boxes"].to(dev, non_blocking=True)
        labels = batch["target_ids"].to(dev, non_blocking=True)
        B = input_ids.shape[0]
        if task == "nlvr":                                       # nlvr_model.py:156-176
            V_L = feats.shape[2]
            feats = feats.reshape(B, 2 * V_L, -1)
            boxes = boxes.reshape(B, 2 * V_L, 4)
            key = (V_L, str(dev))
            if key not in self._nlvr_ids:                        # built once: no host->device copy inside a step
                self._nlvr_ids[key] = (torch.tensor([0] * V_L + [1] * V_L, dtype=torch.long, device=dev).view(1, -1),
                                       torch.arange(V_L, dtype=torch.long, device=dev).repeat(2).view(1, -1))
            img_ids = self._nlvr_ids[key][0].expand(B, -1)
            obj_ids = self._nlvr_ids[key][1].expand(B, -1)
            vis_inputs = (feats, boxes, img_ids, obj_ids)
        else:
            vis_inputs = (feats, boxes)
        loss, _ = self(input_ids, vis_inputs, labels, task=task)
        T = labels.shape[1]
        mask = (labels != -100).float()
        loss = loss.view(B, T) * mask
        if task in ("caption", "tvc", "yc2c"):                   # caption_model.py: sum / count over the whole batch
            loss = loss.sum() / mask.sum().clamp(min=1)
        else:
            loss = loss.sum(dim=1) / mask.sum(dim=1).clamp(min=1)
            if "scores" in batch and batch["scores"] is not None:
                loss = loss * batch["scores"].to(dev, non_blocking=True)
            loss = loss.mean()
        return {"loss": loss}
