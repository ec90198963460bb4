"""Configuration of the host model: BART-base hyper-parameters plus the VL-PET flags, under the attribute names
the reference copies onto its HF config (param.py:59-419 via trainer_base.py:86-87, 141-178, 210-213)."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List

from ..adapters import AdapterConfig

TASKS = ["vqa", "gqa", "nlvr", "caption"]


@dataclass
class VLPetConfig:
    # ---- BART-base (facebook/bart-base config.json; vocab +200 added tokens, multitask.py:59-74)
    vocab_size: int = 50465
    d_model: int = 768
    encoder_layers: int = 6
    decoder_layers: int = 6
    encoder_attention_heads: int = 12
    decoder_attention_heads: int = 12
    encoder_ffn_dim: int = 3072
    decoder_ffn_dim: int = 3072
    max_position_embeddings: int = 1024
    activation_function: str = "gelu"
    scale_embedding: bool = False
    init_std: float = 0.02
    pad_token_id: int = 1
    bos_token_id: int = 0
    eos_token_id: int = 2
    decoder_start_token_id: int = 2
    dropout: float = 0.1              # trainer_base.py:210-213 sets all three from --dropout
    attention_dropout: float = 0.1
    activation_dropout: float = 0.1
    # ---- visual side (param.py:95-110, 217)
    feat_dim: int = 2048
    pos_dim: int = 4
    n_images: int = 2
    n_boxes: int = 36
    downsample: bool = True
    use_vis_order_embedding: bool = True
    use_vis_layer_norm: bool = True
    individual_vis_layer_norm: bool = True
    share_vis_lang_layer_norm: bool = False
    # ---- VL-PET flags (scripts/image-text/VL-PET-large.sh:54-66)
    tasks: List[str] = field(default_factory=lambda: list(TASKS))
    use_encoder_adapter_down_multihead: bool = True
    adapter_down_dim: int = 96
    encoder_adapter_multihead_num_head: int = 4
    use_encoder_adapter_gating_large_x_lowrank: bool = True
    use_encoder_adapter_gating_middle_xy_add: bool = False
    use_encoder_adapter_gating_middle_ia3_add: bool = False
    use_encoder_adapter_gating_small_xy_cat: bool = False
    adapter_gating_down_dim: int = 96
    use_encoder_adapter_gating_add: bool = False
    use_encoder_gating_scaling: bool = False
    encoder_gating_scaling_factor: float = 1.0
    unfreeze_encoder_layer_norms: bool = True
    use_decoder_enc_attn_value_parallel_adapter_down_dim: bool = True
    decoder_enc_attn_value_parallel_adapter_down_dim: int = 96
    use_single_adapter: bool = True
    freeze_vis_emb: bool = False
    # zero-init flags of the T5 scripts (trainer_base.py:544-599 weight_initialization; scripts/image-text/T5-VL-PET-*.sh)
    use_encoder_multihead_up_zero_init: bool = False
    use_encoder_gating_large_x_lowrank_up_zero_init: bool = False
    use_decoder_enc_vpa_up_zero_init: bool = False
    use_encoder_gating_small_up_zero_init: bool = False
    use_encoder_gating_middle_up_zero_init: bool = False
    # ---- T5-base (t5-base config.json; vocab 32 100 + 100 <vis_extra_id> tokens, tokenization.py:58-60); used when arch == "t5"
    arch: str = "bart"
    d_kv: int = 64
    d_ff: int = 3072
    num_layers: int = 12
    num_decoder_layers: int = 12
    num_heads: int = 12
    relative_attention_num_buckets: int = 32
    layer_norm_epsilon: float = 1e-6
    dropout_rate: float = 0.1
    # ---- host-side execution policy (ours)
    assume_no_padding: bool = False   # synthetic batches carry no pad tokens: skip building attention masks
    pet_impl: str = "auto"            # forward kernel selection passed to the C ABI
    pet_bwd_impl: str = "auto"

    def vpa_adapter_config(self) -> AdapterConfig:
        """AdapterConfig of the decoder value-parallel-adapter as my_transformers/modeling_bart.py:329-340 derives
        it from config.adapter_config (deepcopy + use_adapter_down_dim / adapter_down_dim / use_parallel_adapter)."""
        return AdapterConfig(tasks=list(self.tasks), d_model=self.d_model, input_dim=self.d_model,
                             use_single_adapter=self.use_single_adapter, use_adapter_down_dim=True,
                             adapter_down_dim=self.decoder_enc_attn_value_parallel_adapter_down_dim,
                             use_parallel_adapter=True, reduction_factor=8)

    def clone(self, **kw) -> "VLPetConfig":
        c = copy.deepcopy(self)
        for k, v in kw.items():
            if not hasattr(c, k):
                raise AttributeError(k)
            setattr(c, k, v)
        return c


def bart_base_vlpet_large(r: int = 96, heads: int = 4, rg: int = 96, dec_r: int = 96, **kw) -> VLPetConfig:
    """BASELINE config 2: BART-base + VL-PET-large (script arguments `96 4 96 96`)."""
    return VLPetConfig(adapter_down_dim=r, encoder_adapter_multihead_num_head=heads, adapter_gating_down_dim=rg,
                       decoder_enc_attn_value_parallel_adapter_down_dim=dec_r).clone(**kw)


def t5_base_vlpet_large(r: int = 96, heads: int = 4, rg: int = 96, dec_r: int = 96, s: float = 0.3, **kw) -> VLPetConfig:
    """BASELINE config 3: T5-base + VL-PET-large at r = rg = 96 (scripts/image-text/T5-VL-PET-large.sh:41-58 with the
    ranks BASELINE.json names; the script's own `192 4 192 96` is a legal value of the same flags), gate scale 0.3."""
    return VLPetConfig(arch="t5", vocab_size=32200, pad_token_id=0, decoder_start_token_id=0, eos_token_id=1,
                       adapter_down_dim=r, encoder_adapter_multihead_num_head=heads, adapter_gating_down_dim=rg,
                       decoder_enc_attn_value_parallel_adapter_down_dim=dec_r, use_encoder_gating_scaling=True,
                       encoder_gating_scaling_factor=s, use_encoder_multihead_up_zero_init=True,
                       use_encoder_gating_large_x_lowrank_up_zero_init=True, use_decoder_enc_vpa_up_zero_init=True).clone(**kw)


def bart_base_vlpet_small(r: int = 4, heads: int = 4, dec_r: int = 4, **kw) -> VLPetConfig:
    """BASELINE config 4: BART-base + VL-PET-small (scripts/image-text/VL-PET-small.sh flag set) at rank 4."""
    return VLPetConfig(adapter_down_dim=r, encoder_adapter_multihead_num_head=heads, use_encoder_adapter_gating_large_x_lowrank=False,
                       use_encoder_adapter_gating_small_xy_cat=True, decoder_enc_attn_value_parallel_adapter_down_dim=dec_r).clone(**kw)


def bart_base_vlpet_large_video(r: int = 96, heads: int = 4, rg: int = 96, dec_r: int = 96, **kw) -> VLPetConfig:
    """BASELINE config 5: BART-base + VL-PET-large on the video-text multitask (scripts/video-text/VL-PET-large.sh:
    --feat_dim 512 --n_boxes 64 --downsample (an 8x8 -> 8x8 no-op), tasks tvqa / how2qa / tvc / yc2c)."""
    return VLPetConfig(adapter_down_dim=r, encoder_adapter_multihead_num_head=heads, adapter_gating_down_dim=rg,
                       decoder_enc_attn_value_parallel_adapter_down_dim=dec_r, feat_dim=512, n_boxes=64,
                       tasks=["tvqa", "how2qa", "tvc", "yc2c"]).clone(**kw)


def tiny_t5_test_config(**kw) -> VLPetConfig:
    """2+2-layer d=64 T5 used by the parity test against the reference's VLT5 (tests/golden/vlt5_tiny_large.npz)."""
    return VLPetConfig(arch="t5", vocab_size=300, pad_token_id=0, decoder_start_token_id=0, eos_token_id=1, d_model=64, d_kv=16,
                       d_ff=128, num_layers=2, num_decoder_layers=2, num_heads=4, feat_dim=128, adapter_down_dim=16,
                       encoder_adapter_multihead_num_head=4, adapter_gating_down_dim=16,
                       decoder_enc_attn_value_parallel_adapter_down_dim=16, use_encoder_gating_scaling=True,
                       encoder_gating_scaling_factor=0.3).clone(**kw)


def tiny_test_config(**kw) -> VLPetConfig:
    """2+2-layer d=64 model used by the parity tests against the reference's VLBart (tests/golden/vlbart_*.npz)."""
    return VLPetConfig(vocab_size=300, d_model=64, encoder_layers=2, decoder_layers=2, encoder_attention_heads=4,
                       decoder_attention_heads=4, encoder_ffn_dim=128, decoder_ffn_dim=128, max_position_embeddings=128,
                       feat_dim=128, adapter_down_dim=16, encoder_adapter_multihead_num_head=4,
                       adapter_gating_down_dim=16, decoder_enc_attn_value_parallel_adapter_down_dim=16).clone(**kw)
