"""Host-side mirror of the reference's adapter API for the decoder value-parallel-adapter (VPA) path:
``AdapterConfig`` (src/adapters/config.py:4-55), ``Activations`` (adapter_utils.py:7-13), ``Adapter``
(adapter_modeling.py:36-61) and ``AdapterController`` (adapter_controller.py:11-162) -- same constructor
arguments, attribute / parameter names (``adapters.<task>.down_sampler.weight`` ...), aliasing of the single
shared adapter across tasks, and error behaviour; the forward runs the fused CUDA kernel (include/vlpet.h K2).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from . import functional as F_


@dataclass
class AdapterConfig:
    """Defaults of src/adapters/config.py:4-55 plus the fields trainer_base.py:141-178 sets."""
    add_layer_norm_before_adapter: bool = False
    add_layer_norm_after_adapter: bool = False
    non_linearity: str = "gelu_new"
    reduction_factor: int = 16
    weight_init_range: float = 1e-2
    hidden_dim: int = 128
    task_embedding_dim: int = 64
    tasks: Optional[List[str]] = None
    d_model: int = 768
    input_dim: int = 768
    use_single_adapter: bool = False
    use_adapter_down_dim: bool = False
    adapter_down_dim: int = 96
    use_parallel_adapter: bool = False
    use_scaling_factor: bool = False
    scaling_factor: float = 1.0
    track_z: bool = False
    share_up_sampler: bool = False
    share_down_sampler: bool = False
    hypercomplex_adapters: bool = False
    low_rank_adapters: bool = False
    shared_phm_rule: bool = True
    shared_phm_rule_over_tasks: bool = False
    learn_phm: bool = True


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """transformers.activations.NewGELUActivation, for host-side (non-hot-path) uses such as ``track_z``."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class Activations(nn.Module):
    """adapter_utils.py:7-13.  Only ``gelu_new`` exists in the fused kernels; anything else is an error."""

    def __init__(self, activation_type: str):
        super().__init__()
        if activation_type.lower() != "gelu_new":
            raise ValueError(f"vlpet: the PET kernels implement non_linearity='gelu_new' only, got {activation_type!r}")
        self.f = gelu_new

    def forward(self, x):
        return self.f(x)


class Adapter(nn.Module):
    """Bottleneck adapter: parameters ``down_sampler`` Linear(d, r), ``up_sampler`` Linear(r, d)
    (adapter_modeling.py:36-61); r = adapter_down_dim if use_adapter_down_dim else d // reduction_factor."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.input_dim = config.d_model
        if getattr(config, "use_adapter_down_dim", False):
            self.down_sample_size = config.adapter_down_dim
        else:
            self.down_sample_size = self.input_dim // config.reduction_factor
        self.activation = Activations(config.non_linearity.lower())
        self.down_sampler = nn.Linear(self.input_dim, self.down_sample_size)
        self.up_sampler = nn.Linear(self.down_sample_size, self.input_dim)
        self.track_z = getattr(config, "track_z", False)

    def forward(self, x):
        """Up(gelu_new(Down(x))) with no residual (the controller adds it)."""
        if self.track_z:  # debugging aid read by multitask.py:243-257; off in every shipped script
            with torch.no_grad():
                self.z = gelu_new(torch.nn.functional.linear(x, self.down_sampler.weight, self.down_sampler.bias))
        return F_.vpa(x, None, self.down_sampler.weight, self.down_sampler.bias, self.up_sampler.weight,
                      self.up_sampler.bias, 1.0)


class AdapterController(nn.Module):
    """Per-task dictionary of adapters + residual / parallel / scaling logic (adapter_controller.py:11-162)."""

    def __init__(self, config):
        super().__init__()
        if getattr(config, "hypercomplex_adapters", False) or getattr(config, "low_rank_adapters", False):
            raise NotImplementedError("vlpet: Compacter / low-rank adapters are other PET baselines, not the VL-PET path")
        self.config = config
        self.low_rank_adapters = False
        self.hypercomplex_adapters = False
        self.tasks = list(config.tasks)
        self.shared_phm_rule = getattr(config, "shared_phm_rule", True)
        self.use_single_adapter = config.use_single_adapter
        self.share_up_sampler = getattr(config, "share_up_sampler", False)
        self.share_down_sampler = getattr(config, "share_down_sampler", False)
        self.shared_phm_rule_over_tasks = getattr(config, "shared_phm_rule_over_tasks", False)
        self.adapters = self.construct_adapters(self.tasks)
        self.add_layer_norm_before_adapter = config.add_layer_norm_before_adapter
        self.add_layer_norm_after_adapter = config.add_layer_norm_after_adapter
        if self.add_layer_norm_before_adapter:
            self.pre_layer_norm = nn.LayerNorm(config.input_dim)
        if self.add_layer_norm_after_adapter:
            self.post_layer_norm = nn.LayerNorm(config.input_dim)

    def get_task(self, task):
        return task

    def construct_adapters(self, tasks):
        adapters = nn.ModuleDict()
        if self.use_single_adapter:
            shared = Adapter(self.config)       # ONE module registered under every task key (aliasing contract)
            for t in tasks:
                adapters[t] = shared
        else:
            for t in tasks:
                adapters[t] = Adapter(self.config)
            if self.share_up_sampler:
                up = adapters[tasks[0]].up_sampler
                for t in tasks:
                    adapters[t].up_sampler = up
            if self.share_down_sampler:
                down = adapters[tasks[0]].down_sampler
                for t in tasks:
                    adapters[t].down_sampler = down
        return adapters

    @staticmethod
    def convert_to_list(tasks):
        return tasks if isinstance(tasks, list) else [tasks]

    def get_adapter(self, task):
        return self.adapters[task]      # KeyError for an unknown task, as the reference

    def disable_adapters(self, tasks):
        for t in self.convert_to_list(tasks):
            for p in self.get_adapter(t).parameters():
                p.requires_grad = False

    def enable_adapters(self, tasks):
        for t in self.convert_to_list(tasks):
            for p in self.get_adapter(t).parameters():
                p.requires_grad = True

    def forward(self, inputs, task, y=None):
        """outputs = [post_LN]( sf * adapter([pre_LN](inputs)) ) + (y if use_parallel_adapter else inputs)."""
        adapter = self.get_adapter(self.get_task(task))
        cfg = self.config
        z = self.pre_layer_norm(inputs) if self.add_layer_norm_before_adapter else inputs
        sf = float(cfg.scaling_factor) if cfg.use_scaling_factor else 1.0
        if cfg.use_parallel_adapter:
            if y is None:
                raise TypeError("AdapterController.forward: use_parallel_adapter=True needs y")
            residual = y
        else:
            residual = inputs
        if adapter.track_z:
            with torch.no_grad():
                adapter.z = gelu_new(torch.nn.functional.linear(z, adapter.down_sampler.weight, adapter.down_sampler.bias))
        if z.is_cuda and (torch.is_autocast_enabled() or (residual is not None and residual.dtype != z.dtype)):
            # torch.autocast: `inputs` comes out of a LayerNorm in fp32, y = v_proj(inputs) in the autocast dtype
            ct = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else residual.dtype
            z = z.to(ct)
            residual = residual.to(ct) if residual is not None else None
        if self.add_layer_norm_after_adapter:   # LayerNorm sits between the scaled adapter output and the residual
            out = F_.vpa(z, None, adapter.down_sampler.weight, adapter.down_sampler.bias, adapter.up_sampler.weight,
                         adapter.up_sampler.bias, sf)
            return self.post_layer_norm(out) + residual
        return F_.vpa(z, residual, adapter.down_sampler.weight, adapter.down_sampler.bias, adapter.up_sampler.weight,
                      adapter.up_sampler.bias, sf)
