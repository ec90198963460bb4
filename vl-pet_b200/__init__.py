"""vl-pet_b200: B200-native (sm_100a) implementation of VL-PET's PET hot path behind the reference's module API.

Import as ``vlpet_b200`` (the repo-root ``vlpet_b200.py`` maps that name onto this directory, whose name is not
a valid Python identifier).  Importing loads ``libvlpet.so`` (built in-tree by ``build.py``); there is no CPU
fallback -- if the CUDA library is missing the import fails.
"""
from . import _lib
from ._lib import LIB_PATH, VlpetError, launch_count
from .functional import PetSiteConfig, gated_pet, vpa, visual_projection, fwd_is_fused, grid_maxpool, layer_norm
from .adapters import AdapterConfig, Activations, Adapter, AdapterController
from .encoder import encoder_pet, gate_kind, patch_layer, patch_reference_model, site_config, site_params
from .visual import VisualEmbedding, LowRankVisualEmbedding, T5LayerNorm, adopt_reference_visual_embedding
from . import host

__all__ = ["LIB_PATH", "VlpetError", "launch_count", "PetSiteConfig", "gated_pet", "vpa", "visual_projection",
           "fwd_is_fused", "grid_maxpool", "layer_norm", "AdapterConfig", "Activations", "Adapter", "AdapterController", "encoder_pet", "gate_kind",
           "patch_layer", "patch_reference_model", "site_config", "site_params", "VisualEmbedding", "LowRankVisualEmbedding", "T5LayerNorm",
           "adopt_reference_visual_embedding"]
