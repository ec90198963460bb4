"""Host-side mirror of the reference's ``VisualEmbedding`` (src/modeling_bart.py:77-192; T5 flavour
src/modeling_t5.py:44-174): same constructor signature ``(config, obj_order_embedding)``, same parameter names
(``feat_embedding.0/.1``, ``absolute_vis_pos_embedding.0/.1``, ``img_order_embedding``; ``obj_order_embedding``
aliases the token table), same ``forward(feats, pos, img_order_ids=None, obj_order_ids=None)``; the forward runs
the CUDA kernels of include/vlpet.h K3."""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import functional as F_


class T5LayerNorm(nn.Module):
    """Parameter container for the RMS norm of my_transformers/modeling_t5.py:235-252 (weight only)."""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):  # host-side use outside the fused visual projection (text stream of the backbone)
        var = x.float().pow(2).mean(-1, keepdim=True)
        return self.weight * (x * torch.rsqrt(var + self.variance_epsilon)).to(self.weight.dtype)


def _check_flags(config):
    g = lambda n, d: getattr(config, n, d)  # noqa: E731
    if g("vis_use_transformer", False) or not g("use_vis_layer_norm", True) or not g("individual_vis_layer_norm", True) \
            or not g("use_vis_order_embedding", True) or g("additional_visual_embedding_layers", 0):
        raise NotImplementedError("vlpet: VisualEmbedding supports the VL-PET defaults only (individual LayerNorms, "
                                  "order embeddings, no BERT block, no extra layers)")


class VisualEmbedding(nn.Module):
    def __init__(self, config, obj_order_embedding: nn.Embedding, rms: bool = False):
        super().__init__()
        _check_flags(config)
        self.config = config
        d = config.d_model
        self.rms = rms
        if rms:
            eps = getattr(config, "layer_norm_epsilon", 1e-6)
            norm = lambda: T5LayerNorm(d, eps=eps)  # noqa: E731
        else:
            norm = lambda: nn.LayerNorm(d)  # noqa: E731
        self.feat_embedding = nn.Sequential(nn.Linear(config.feat_dim, d), norm())
        self.absolute_vis_pos_embedding = nn.Sequential(nn.Linear(config.pos_dim + 1, d), norm())
        self.obj_order_embedding = obj_order_embedding
        self.img_order_embedding = nn.Embedding(config.n_images, d)
        self.default_obj_order_ids = getattr(config, "default_obj_order_ids", None)

    def forward(self, feats, pos, img_order_ids=None, obj_order_ids=None):
        return _forward(self, feats, pos, img_order_ids, obj_order_ids)


def _forward(self, feats, pos, img_order_ids=None, obj_order_ids=None):
    B, N, _ = feats.size()
    assert pos.size() == (B, N, 4)
    lin_f, ln_f = self.feat_embedding[0], self.feat_embedding[1]
    lin_p, ln_p = self.absolute_vis_pos_embedding[0], self.absolute_vis_pos_embedding[1]
    rms = not hasattr(ln_f, "bias") or getattr(ln_f, "bias", None) is None
    eps = getattr(ln_f, "variance_epsilon", None) if rms else ln_f.eps
    if feats.is_cuda and torch.is_autocast_enabled():      # torch.autocast: the projection GEMM runs in the autocast dtype
        feats = feats.to(torch.get_autocast_dtype("cuda"))
    return F_.visual_projection(feats, pos, img_order_ids, obj_order_ids, lin_f.weight, lin_f.bias, ln_f.weight,
                                None if rms else ln_f.bias, lin_p.weight, lin_p.bias, ln_p.weight,
                                None if rms else ln_p.bias, self.img_order_embedding.weight,
                                self.obj_order_embedding.weight, rms=rms, eps=float(eps))


def adopt_reference_visual_embedding(module: nn.Module) -> nn.Module:
    """Route an already-constructed reference ``VisualEmbedding`` through the CUDA kernels (parameters untouched)."""
    _check_flags(module.config)
    if len(module.feat_embedding) != 2 or len(module.absolute_vis_pos_embedding) != 2:
        raise NotImplementedError("vlpet: unexpected VisualEmbedding layout")
    module.forward = types.MethodType(_forward, module)
    return module


class LowRankVisualEmbedding(nn.Module):
    """Mirror of the reference's PET-shaped visual projector (src/modeling_bart.py:195-334, --use_lowrank_visual_projector):
    same constructor ``(config, obj_order_embedding)``, same parameter names (``visual_projector_multihead_down.{h}``,
    ``visual_projector_multihead_up``, ``visual_projector_gating_large_x_{down,up}``, ``visual_projector_layer_norm``,
    ``absolute_vis_pos_embedding.0/.1``, ``img_order_embedding``), forward through include/vlpet.h K3-LR."""

    def __init__(self, config, obj_order_embedding: nn.Embedding):
        super().__init__()
        _check_flags(config)
        self.config = config
        d, F = config.d_model, config.feat_dim
        h = config.visual_projector_multihead_num_head
        r = config.visual_projector_down_dim
        self.embed_dim = d
        self.visual_projector_multihead_dim = int(r / h)
        self.visual_projector_multihead_down = nn.ModuleList([nn.Linear(F, self.visual_projector_multihead_dim) for _ in range(h)])
        self.visual_projector_multihead_up = nn.Linear(r, d)
        self.gated = bool(getattr(config, "use_visual_projector_gating_large_x_lowrank", False))
        if self.gated:
            rg = config.visual_projector_gating_down_dim
            self.visual_projector_gating_large_x_down = nn.Linear(F, rg)
            self.visual_projector_gating_large_x_up = nn.Linear(rg, d)
        self.visual_projector_layer_norm = nn.LayerNorm(d)
        self.absolute_vis_pos_embedding = nn.Sequential(nn.Linear(config.pos_dim + 1, d), nn.LayerNorm(d))
        self.obj_order_embedding = obj_order_embedding
        self.img_order_embedding = nn.Embedding(config.n_images, d)
        self.default_obj_order_ids = getattr(config, "default_obj_order_ids", None)

    def forward(self, feats, pos, img_order_ids=None, obj_order_ids=None):
        B, N, _ = feats.size()
        assert pos.size() == (B, N, 4)
        down = self.visual_projector_multihead_down
        # the multi-head down projection is ONE Linear whose weight is the row-concatenation of the heads;
        # the differentiable cat hands autograd the per-head gradient slices
        Wd = down[0].weight if len(down) == 1 else torch.cat([h.weight for h in down], dim=0)
        bd = down[0].bias if len(down) == 1 else torch.cat([h.bias for h in down], dim=0)
        gate = (self.visual_projector_gating_large_x_down.weight, self.visual_projector_gating_large_x_down.bias,
                self.visual_projector_gating_large_x_up.weight, self.visual_projector_gating_large_x_up.bias) \
            if self.gated else (None, None, None, None)
        ln, pe = self.visual_projector_layer_norm, self.absolute_vis_pos_embedding
        params = (Wd, bd, self.visual_projector_multihead_up.weight, self.visual_projector_multihead_up.bias, *gate,
                  ln.weight, ln.bias, pe[0].weight, pe[0].bias, pe[1].weight, pe[1].bias, self.img_order_embedding.weight,
                  self.obj_order_embedding.weight)
        return F_.lowrank_visual_projection(feats, pos, img_order_ids, obj_order_ids, params, self.gated,
                                            bool(getattr(self.config, "use_visual_projector_residual_connection", False)),
                                            eps=ln.eps)
