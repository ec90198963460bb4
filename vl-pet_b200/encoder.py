"""Drop-in boundary for the ENCODER PET sites.

In the reference the encoder PET math is not a module: it is written inline in ``BartEncoderLayer.forward``
(my_transformers/modeling_bart.py:1145-1261 attention site, 1268-1377 FFN site), ``T5LayerSelfAttention.forward``
(my_transformers/modeling_t5.py:777-824) and ``T5LayerFF.forward`` (359-409), with the parameters living as plain
attributes of the layer.  The boundary is therefore *the layer's parameter names + flags*:
``encoder_pet(layer, site, x1, x2)`` reads them off any layer object that follows the reference naming and runs
the fused kernel; ``patch_*`` swap the forward of reference-style layers for one that calls it, leaving
``named_parameters()`` / ``state_dict()`` untouched so name-based unfreezing (trainer_base.py:308-542) and
released checkpoints keep working.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import functional as F_

# attribute-name stems per site (Appendix B)
_ADAPTER = {"attn": "attn_adapter_multihead", "ff": "ff_adapter_multihead"}
_GATE = {"attn": "encoder_attn_adapter_gating", "ff": "encoder_ff_adapter_gating"}

_UNSUPPORTED_FLAGS = (
    "use_encoder_adapter_up_multihead", "use_encoder_adapter_down_up_multihead",
    "use_encoder_adapter_down_up_pair_multihead", "use_encoder_adapter_gating_large_x",
    "use_encoder_gating_large_x_lowrank", "use_encoder_adapter_gating_layernorm", "use_encoder_adapter_gating_l2norm",
    "use_hyperformer", "use_store_gate_large",
)


def gate_kind(config) -> str:
    """Which granularity-control matrix the flags select (same precedence as modeling_bart.py:1195-1231)."""
    g = lambda n: bool(getattr(config, n, False))  # noqa: E731
    if g("use_encoder_adapter_gating_large_x_lowrank"):
        return "large"
    if g("use_encoder_adapter_gating_small_xy_cat"):
        return "small"
    if g("use_encoder_adapter_gating_middle_xy_add"):
        return "middle_x"
    if g("use_encoder_adapter_gating_middle_ia3_add"):
        return "middle_y"
    return "none"


def site_config(config, is_t5: bool = False, impl: str = "auto") -> F_.PetSiteConfig:
    """PetSiteConfig from the reference flag names (param.py:262-376)."""
    g = lambda n, dflt=False: getattr(config, n, dflt)  # noqa: E731
    for flag in _UNSUPPORTED_FLAGS:
        if g(flag):
            raise NotImplementedError(f"vlpet: --{flag} is an ablation branch outside the VL-PET hot path")
    if not g("use_encoder_adapter_down_multihead"):
        raise NotImplementedError("vlpet: the encoder PET path needs --use_encoder_adapter_down_multihead")
    s = float(g("encoder_gating_scaling_factor", 1.0)) if g("use_encoder_gating_scaling") else 1.0
    # alpha / kappa exist only in the T5 layers (modeling_t5.py:789-793); the BART layer ignores the flags
    alpha = float(g("encoder_adapter_scaling_factor", 1.0)) if (is_t5 and g("use_encoder_adapter_scaling")) else 1.0
    kappa = float(g("encoder_x2_scaling_factor", 1.0)) if (is_t5 and g("use_encoder_x2_scaling")) else 1.0
    p = float(g("dropout_rate", 0.0) if is_t5 else g("dropout", 0.0))
    return F_.PetSiteConfig(gate=gate_kind(config), add_gate=bool(g("use_encoder_adapter_gating_add")) and not is_t5,
                            s=s, alpha=alpha, kappa=kappa, p_drop=p, impl=impl)


def site_params(layer: nn.Module, site: str, gate: str):
    """-> (down_ws, down_bs, up_w, up_b, gate_params) read off the layer by the reference attribute names."""
    down = getattr(layer, _ADAPTER[site] + "_down")
    up = getattr(layer, _ADAPTER[site] + "_up")
    if not isinstance(down, nn.ModuleList) or not isinstance(up, nn.Linear):
        raise NotImplementedError("vlpet: expected multi-head down (ModuleList) + single up Linear")
    down_ws = [h.weight for h in down]
    down_bs = [h.bias for h in down]
    stem = _GATE[site]
    if gate == "large":
        gd, gu = getattr(layer, stem + "_large_x_down"), getattr(layer, stem + "_large_x_up")
        gp = (gd.weight, gd.bias, gu.weight, gu.bias)
    elif gate == "middle_x":
        lin = getattr(layer, stem + "_middle_xy_add")
        gp = (lin.weight.view(-1), lin.bias)
    elif gate == "small":
        lin = getattr(layer, stem + "_small_xy_cat")
        gp = (lin.weight.view(-1), lin.bias)
    elif gate == "middle_y":
        gp = (getattr(layer, stem + "_middle_ia3_add"),)
    else:
        gp = ()
    return down_ws, down_bs, up.weight, up.bias, gp


def encoder_pet(layer: nn.Module, site: str, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """x1 + dropout(s * gate(x1, x2 + adapter(x2))) for ``site`` in {"attn", "ff"} of a reference-style layer.
    The BART LayerNorm that follows stays with the caller."""
    cfg = layer._vlpet_site_cfg
    down_ws, down_bs, up_w, up_b, gp = site_params(layer, site, cfg.gate)
    if x1.is_cuda and (torch.is_autocast_enabled() or x1.dtype != x2.dtype):
        # torch.autocast (the reference's only reduced-precision mode, multitask.py:229-234): the residual stream stays fp32,
        # the sub-layer output arrives in the autocast dtype.  The kernel runs in that dtype; the sum goes back to the stream's.
        ct = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x2.dtype
        out = F_.gated_pet(x1.to(ct), x2.to(ct), down_ws, down_bs, up_w, up_b, gp, cfg, training=layer.training)
        return out.to(x1.dtype)
    return F_.gated_pet(x1, x2, down_ws, down_bs, up_w, up_b, gp, cfg, training=layer.training)


# ---- patched forwards (same signatures / return tuples as the reference layers) -----------------------------------
def _bart_encoder_layer_forward(self, hidden_states, attention_mask, past_key_value=None, block_adapters=None,
                                task=None, output_attentions=False):
    x1 = hidden_states
    x2, attn_weights, _ = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask,
                                         past_key_value=past_key_value, output_attentions=output_attentions)
    hidden_states = self.self_attn_layer_norm(encoder_pet(self, "attn", x1, x2))
    x1 = hidden_states
    h = self.activation_fn(self.fc1(hidden_states))
    h = nn.functional.dropout(h, p=self.activation_dropout, training=self.training)
    x2 = self.fc2(h)
    hidden_states = self.final_layer_norm(encoder_pet(self, "ff", x1, x2))
    # the reference clamps when it sees inf/nan (modeling_bart.py:1379-1381) behind a device->host sync; a
    # LayerNorm output is finite unless its input was not, so the clamp is kept sync-free: nan_to_num is skipped
    outputs = (hidden_states,)
    if output_attentions:
        outputs += (attn_weights,)
    return outputs


def _t5_self_attention_forward(self, hidden_states, attention_mask=None, position_bias=None, head_mask=None,
                               past_key_value=None, use_cache=False, output_attentions=False, block_adapters=None,
                               task=None):
    normed = self.layer_norm(hidden_states)
    attention_output = self.SelfAttention(normed, mask=attention_mask, position_bias=position_bias, head_mask=head_mask,
                                          past_key_value=past_key_value, use_cache=use_cache,
                                          output_attentions=output_attentions)
    hidden_states = encoder_pet(self, "attn", hidden_states, attention_output[0])
    return (hidden_states,) + attention_output[1:]


def _t5_ff_forward(self, hidden_states, block_adapters=None, task=None):
    forwarded = self.DenseReluDense(self.layer_norm(hidden_states))
    return encoder_pet(self, "ff", hidden_states, forwarded)


def _is_encoder_pet_layer(m: nn.Module, site: str) -> bool:
    return isinstance(getattr(m, _ADAPTER[site] + "_down", None), nn.ModuleList)


def patch_layer(layer: nn.Module, impl: str = "auto") -> str:
    """Swap the forward of ONE reference-style layer for the fused one.  Returns the kind patched
    ('bart_encoder_layer' | 't5_self_attention' | 't5_ff') or '' if the layer carries no encoder PET site."""
    cls = type(layer).__name__
    cfg = getattr(layer, "config", None)
    if cfg is None:
        return ""
    if cls == "BartEncoderLayer" and _is_encoder_pet_layer(layer, "attn"):
        if getattr(layer, "attn_adapter", None) is not None or getattr(cfg, "use_lora", False):
            raise NotImplementedError("vlpet: layer mixes VL-PET with another PET method")
        layer._vlpet_site_cfg = site_config(cfg, is_t5=False, impl=impl)
        layer.forward = types.MethodType(_bart_encoder_layer_forward, layer)
        return "bart_encoder_layer"
    if cls == "T5LayerSelfAttention" and not getattr(layer, "is_decoder", False) and _is_encoder_pet_layer(layer, "attn"):
        layer._vlpet_site_cfg = site_config(cfg, is_t5=True, impl=impl)
        layer.forward = types.MethodType(_t5_self_attention_forward, layer)
        return "t5_self_attention"
    if cls == "T5LayerFF" and not getattr(layer, "is_decoder", False) and _is_encoder_pet_layer(layer, "ff"):
        layer._vlpet_site_cfg = site_config(cfg, is_t5=True, impl=impl)
        layer.forward = types.MethodType(_t5_ff_forward, layer)
        return "t5_ff"
    return ""


def patch_reference_model(model: nn.Module, impl: str = "auto") -> dict:
    """Walk a reference model (VLBart / VLT5 built by the reference's own code) and route every VL-PET site
    through the fused kernels: encoder layers via ``patch_layer``; decoder ``attn_value_parallel_adapter`` and
    ``visual_embedding`` modules get their forward replaced by ours (parameters stay where they are)."""
    from . import adapters as A, visual as V
    counts = {"bart_encoder_layer": 0, "t5_self_attention": 0, "t5_ff": 0, "vpa": 0, "visual_embedding": 0}
    for name, m in model.named_modules():
        kind = patch_layer(m, impl)
        if kind:
            counts[kind] += 1
            continue
        cls = type(m).__name__
        if cls == "AdapterController" and name.endswith("attn_value_parallel_adapter") and not isinstance(m, A.AdapterController):
            m.forward = types.MethodType(A.AdapterController.forward, m)
            m.get_adapter = types.MethodType(A.AdapterController.get_adapter, m)
            m.get_task = types.MethodType(A.AdapterController.get_task, m)
            for ad in set(m.adapters.values()):
                ad.track_z = getattr(ad, "track_z", False)
            counts["vpa"] += 1
        elif cls == "VisualEmbedding" and not isinstance(m, V.VisualEmbedding):
            V.adopt_reference_visual_embedding(m)
            counts["visual_embedding"] += 1
    return counts
