"""ORACLE — test infrastructure, NOT product code.

Stock-PyTorch *eager* restatement of the reference's PET op sequences, written op by op the way the reference
issues them (separate per-head Linears + ``torch.cat``, the un-fused Python ``gelu_new``, ``torch.sigmoid``,
``F.dropout``, ``nn.functional.linear`` ...), runnable on CPU.  It serves two purposes:

  * ``use_eager_pet(model)`` swaps the PET sites of a ``vlpet_b200.host.VLBart`` for these eager versions, which
    gives (a) the CPU checker of the host model against golden vectors made by the reference's own ``VLBart``
    (tests/golden/make_golden_vlbart.py) and (b) the "reference's stock PyTorch CPU path" that
    ``bench.py --impl reference`` / ``cpu_baseline`` time on the GPU box's host cores, where /root/reference does
    not exist (kind = "port").
  * the isolated PET op sequence for the kernel-level CPU baseline.

Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this module.

Reference lines restated (paths relative to /root/reference/src):
  my_transformers/modeling_bart.py:1145-1155 (adapter), 1195-1231 (gates), 1256-1260 (scale, dropout, residual);
  adapters/adapter_controller.py:131-162 + adapters/adapter_modeling.py:55-61 (VPA);
  modeling_bart.py:143-192 (VisualEmbedding.forward); transformers.activations.NewGELUActivation (gelu_new).
"""
from __future__ import annotations

import math
import types

import torch
import torch.nn.functional as F


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def eager_encoder_pet(layer, site: str, x1, x2):
    """One encoder PET site exactly as BartEncoderLayer.forward runs it inline (returns the pre-LayerNorm sum)."""
    cfg = layer.config
    g = lambda n, d=False: getattr(cfg, n, d)  # noqa: E731
    down = getattr(layer, f"{site}_adapter_multihead_down")
    up = getattr(layer, f"{site}_adapter_multihead_up")
    stem = f"encoder_{site}_adapter_gating"
    z = torch.cat([down[i](x2) for i in range(len(down))], dim=-1)
    z = gelu_new(z)
    h = x2 + up(z)
    add = g("use_encoder_adapter_gating_add")
    if g("use_encoder_adapter_gating_large_x_lowrank"):
        gate = getattr(layer, stem + "_large_x_down")(x1)
        gate = gelu_new(gate)
        gate = getattr(layer, stem + "_large_x_up")(gate)
        gate = torch.sigmoid(gate)
        h = h + gate if add else h * gate
    elif g("use_encoder_adapter_gating_small_xy_cat"):
        gate = getattr(layer, stem + "_small_xy_cat")(torch.cat([x1, h], dim=-1))
        gate = torch.sigmoid(gate)
        gate = torch.mean(gate, dim=1).unsqueeze(-1)
        h = h + gate if add else h * gate
    elif g("use_encoder_adapter_gating_middle_xy_add"):
        gate = getattr(layer, stem + "_middle_xy_add")(x1 + h)
        gate = torch.sigmoid(gate)
        h = h + gate if add else h * gate
    elif g("use_encoder_adapter_gating_middle_ia3_add"):
        gate = getattr(layer, stem + "_middle_ia3_add")
        h = (h + 1 + gate) if add else (h + h * gate)
    if g("use_encoder_gating_scaling"):
        h = h * g("encoder_gating_scaling_factor", 1.0)
    h = F.dropout(h, p=layer.dropout, training=layer.training)
    return x1 + h


def eager_encoder_pet_t5(layer, site: str, x1, x2):
    """One encoder PET site as T5LayerSelfAttention.forward / T5LayerFF.forward run it inline
    (my_transformers/modeling_t5.py:777-824, 359-409): optional adapter / x2 scales, no add-gate, no LayerNorm behind."""
    cfg = layer.config
    g = lambda n, d=False: getattr(cfg, n, d)  # noqa: E731
    down = getattr(layer, f"{site}_adapter_multihead_down")
    up = getattr(layer, f"{site}_adapter_multihead_up")
    stem = f"encoder_{site}_adapter_gating"
    u = up(gelu_new(torch.cat([down[i](x2) for i in range(len(down))], dim=-1)))
    if g("use_encoder_adapter_scaling"):
        u = u * g("encoder_adapter_scaling_factor", 1.0)
    y = x2 * g("encoder_x2_scaling_factor", 1.0) if g("use_encoder_x2_scaling") else x2
    y = y + u
    if g("use_encoder_adapter_gating_large_x_lowrank"):
        gate = torch.sigmoid(getattr(layer, stem + "_large_x_up")(gelu_new(getattr(layer, stem + "_large_x_down")(x1))))
        y = y * gate
    elif g("use_encoder_adapter_gating_small_xy_cat"):
        gate = torch.sigmoid(getattr(layer, stem + "_small_xy_cat")(torch.cat([x1, y], dim=2)))
        y = y * torch.mean(gate, dim=1).unsqueeze(-1)
    elif g("use_encoder_adapter_gating_middle_xy_add"):
        y = y * torch.sigmoid(getattr(layer, stem + "_middle_xy_add")(x1 + y))
    elif g("use_encoder_adapter_gating_middle_ia3_add"):
        y = y + y * getattr(layer, stem + "_middle_ia3_add")
    if g("use_encoder_gating_scaling"):
        y = y * g("encoder_gating_scaling_factor", 1.0)
    return x1 + F.dropout(y, p=layer.dropout, training=layer.training)


def eager_adapter_controller_forward(self, inputs, task, y=None):
    adapter = self.adapters[task]
    z = self.pre_layer_norm(inputs) if self.add_layer_norm_before_adapter else inputs
    out = adapter.up_sampler(gelu_new(adapter.down_sampler(z)))
    if self.config.use_scaling_factor:
        out = self.config.scaling_factor * out
    if self.add_layer_norm_after_adapter:
        out = self.post_layer_norm(out)
    return out + (y if self.config.use_parallel_adapter else inputs)


def eager_visual_embedding_forward(self, feats, pos, img_order_ids=None, obj_order_ids=None):
    B, N, _ = feats.size()
    feat_embedding = self.feat_embedding(feats)
    height = pos[:, :, 3] - pos[:, :, 2]
    width = pos[:, :, 1] - pos[:, :, 0]
    area = (height * width).unsqueeze(2)
    pos5 = torch.cat([pos, area], dim=2).to(feats.dtype)
    absolute_vis_pos_embedding = self.absolute_vis_pos_embedding(pos5)
    device = feats.device
    if img_order_ids is None:
        img_order_ids = torch.zeros(N, dtype=torch.long, device=device).unsqueeze(0)
    img_order_embedding = self.img_order_embedding(img_order_ids)
    if obj_order_ids is None:
        obj_order_ids = torch.arange(N, dtype=torch.long, device=device).unsqueeze(0)
    obj_order_ids = self.obj_order_embedding.num_embeddings - obj_order_ids - 1
    obj_order_embedding = self.obj_order_embedding(obj_order_ids)
    return feat_embedding + absolute_vis_pos_embedding + img_order_embedding + obj_order_embedding


def use_eager_pet(model):
    """Route every PET site of a host model through the eager restatement (instance-level patch, parameters and
    names untouched).  Returns the model."""
    n = 0
    for m in model.modules():
        cls = type(m).__name__
        if cls == "BartEncoderLayer" and hasattr(m, "attn_adapter_multihead_down"):
            m._pet = types.MethodType(lambda self, site, x1, x2: eager_encoder_pet(self, site, x1, x2), m)
            n += 1
        elif cls in ("T5LayerSelfAttention", "T5LayerFF") and hasattr(m, "_vlpet_site_cfg"):
            m._pet = types.MethodType(lambda self, site, x1, x2: eager_encoder_pet_t5(self, site, x1, x2), m)
            n += 1
        elif cls == "AdapterController":
            m.forward = types.MethodType(eager_adapter_controller_forward, m)
            n += 1
        elif cls == "VisualEmbedding":
            m.forward = types.MethodType(eager_visual_embedding_forward, m)
            n += 1
    assert n > 0, "no PET site found"
    return model


def isolated_pet_step(x1, x2, dout, params, nheads: int = 4):
    """The reference's encoder PET op sequence (large gate) forward + backward on raw tensors: the kernel-level
    CPU baseline.  params: Wd [r,d], bd, Wu, bu, Gd, gbd, Gu, gbu (requires_grad tensors)."""
    r = params["Wd"].shape[0]
    hr = r // nheads
    heads = [F.linear(x2, params["Wd"][i * hr:(i + 1) * hr], params["bd"][i * hr:(i + 1) * hr]) for i in range(nheads)]
    z = gelu_new(torch.cat(heads, dim=-1))
    h = x2 + F.linear(z, params["Wu"], params["bu"])
    gate = torch.sigmoid(F.linear(gelu_new(F.linear(x1, params["Gd"], params["gbd"])), params["Gu"], params["gbu"]))
    out = x1 + h * gate
    out.backward(dout)
    return out
