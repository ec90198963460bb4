"""The reference arm of bench.py: the UNMODIFIED reference (HenryHZY/VL-PET ``src/``) driven through its own public API
(``VLBartMultiTask.train_step`` / ``VLT5MultiTask.train_step``, multitask_model.py:54-89) on the synthetic multitask
cycle -- none of this repository's models, kernels or engine is on that path (``vlpet_b200`` is never imported here).

``stage()`` copies ``/root/reference/src`` to ``baseline/_ref/src`` (git-ignored, travels to the GPU box with the gpurun
snapshot: the box has no /root/reference).  The reference pins transformers 4.2.1; the image has 5.5, so the import shims
of SURVEY Appendix C are applied (``tests/golden/ref_import.py`` holds the same list) -- they patch *transformers*, not
the reference.  The reference's trainer (trainer_base.py) cannot be imported offline (vis_encoder -> clip -> ftfy / timm):
its freeze / unfreeze-by-substring rule (trainer_base.py:268-270, 308-542) and AdamW groups (627-732) are restated in
``prepare_training`` below, the model and the task steps are the reference's own code.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DST = os.path.join(HERE, "_ref", "src")
REF_ORIGIN = "/root/reference/src"


def stage(force: bool = False) -> str | None:
    """Copy the reference's Python sources next to the bench (no edits).  Returns the staged path or None."""
    if os.path.isdir(REF_ORIGIN) and (force or not os.path.isdir(REF_DST)):
        tmp = REF_DST + ".tmp"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(REF_ORIGIN, tmp, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        shutil.rmtree(REF_DST, ignore_errors=True)
        os.makedirs(os.path.dirname(REF_DST), exist_ok=True)
        os.replace(tmp, REF_DST)
    return REF_DST if os.path.isdir(REF_DST) else None


def available() -> str | None:
    if os.path.isdir(REF_DST):
        return REF_DST
    return REF_ORIGIN if os.path.isdir(REF_ORIGIN) else None


def _synthetic():
    """host/synthetic.py by file path: the batch generator is plain torch, and importing the package would load libvlpet.so."""
    path = os.path.join(ROOT, "vl-pet_b200", "host", "synthetic.py")
    spec = importlib.util.spec_from_file_location("_vlpet_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_vlpet_synthetic"] = mod   # dataclasses / typing look the module up by name
    spec.loader.exec_module(mod)
    return mod


_PATCHED_METHODS = ("init_weights", "get_extended_attention_mask", "invert_attention_mask", "get_head_mask")
_STOCK: dict = {}


def remove_shims():
    """Restore the ``PreTrainedModel`` methods install_shims replaced (for a process that goes on to use stock transformers
    models, e.g. a test session); names that were only added stay."""
    import transformers.modeling_utils as mu
    for name, fn in _STOCK.items():
        if fn is not None:
            setattr(mu.PreTrainedModel, name, fn)


def install_shims(src: str):
    import torch
    import transformers
    import transformers.file_utils as fu
    import transformers.modeling_utils as mu

    fu.add_code_sample_docstrings = lambda *a, **k: (lambda f: f)

    def _stub(*a, **k):
        raise NotImplementedError("head pruning is not used")

    mu.find_pruneable_heads_and_indices = _stub
    mu.prune_linear_layer = _stub
    if not hasattr(mu, "apply_chunking_to_forward"):
        from transformers.pytorch_utils import apply_chunking_to_forward
        mu.apply_chunking_to_forward = apply_chunking_to_forward
    mp = types.ModuleType("transformers.utils.model_parallel_utils")
    mp.assert_device_map = lambda *a, **k: None
    mp.get_device_map = lambda *a, **k: None
    sys.modules["transformers.utils.model_parallel_utils"] = mp
    for name in _PATCHED_METHODS:                       # remembered once, so remove_shims() can give stock models theirs back
        _STOCK.setdefault(name, getattr(mu.PreTrainedModel, name, None))
    mu.PreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)

    # transformers 4.2.1 semantics of the mask helpers the T5 wrapper calls (src/modeling_t5.py:282-292)
    def get_extended_attention_mask(self, attention_mask, input_shape, device=None, *a, **k):
        if attention_mask.dim() == 3:
            ext = attention_mask[:, None, :, :]
        elif getattr(self.config, "is_decoder", False):
            B, S = input_shape
            ids = torch.arange(S, device=attention_mask.device)
            causal = (ids[None, None, :].repeat(B, S, 1) <= ids[None, :, None]).to(attention_mask.dtype)
            ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
        else:
            ext = attention_mask[:, None, None, :]
        ext = ext.to(dtype=self.dtype)
        return (1.0 - ext) * -10000.0

    def invert_attention_mask(self, m):
        ext = m[:, None, None, :] if m.dim() == 2 else m[:, None, :, :]
        return (1.0 - ext.to(dtype=self.dtype)) * -1e9

    mu.PreTrainedModel.get_extended_attention_mask = get_extended_attention_mask
    mu.PreTrainedModel.invert_attention_mask = invert_attention_mask
    mu.PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    if src not in sys.path:
        sys.path.insert(0, src)


def _import_reference(kind: str):
    src = available()
    if src is None:
        raise RuntimeError("reference sources not staged (baseline/_ref/src) and /root/reference absent")
    install_shims(src)
    import transformers
    import my_transformers.modeling_bart as mb
    import transformers.models.bart.modeling_bart as hfb
    hfb._make_causal_mask = mb._make_causal_mask
    hfb._expand_mask = mb._expand_mask

    class _Dummy:   # only generate() uses the beam scorers
        pass

    for n in ("BeamScorer", "BeamSearchScorer"):
        setattr(sys.modules["transformers"], n, _Dummy)
    import multitask_model
    return multitask_model, transformers


# flag sets of scripts/image-text/{VL-PET-large,T5-VL-PET-large}.sh (read, never executed: SURVEY F10)
BASE_FLAGS = ("--use_adapter --use_single_adapter --no_encoder_adapter --use_adapter_down_dim "
              "--use_encoder_adapter_down_multihead --unfreeze_encoder_layer_norms --no_decoder_adapter "
              "--use_decoder_enc_attn_value_parallel_adapter_down_dim --tasks vqa,gqa,nlvr,caption "
              "--feature RN101 --n_boxes 36 --downsample --image_size (224,224) --optim adamw --warmup_ratio 0.1 "
              "--clip_grad_norm 5 --num_beams 5 --max_text_length 20")
GATE_FLAG = {"large": "--use_encoder_adapter_gating_large_x_lowrank", "middle_x": "--use_encoder_adapter_gating_middle_xy_add",
             "middle_y": "--use_encoder_adapter_gating_middle_ia3_add", "small": "--use_encoder_adapter_gating_small_xy_cat"}


def build_model(kind: str = "bart", r: int = 96, rg: int | None = None, dec_r: int | None = None, gate: str = "large",
                heads: int = 4, feat_dim: int = 2048, n_boxes: int = 36, dropout: float = 0.1, seed: int = 0,
                layers: int | None = None):
    """The reference's own VLBartMultiTask / VLT5MultiTask, random-init from a hand-written base config (no network)."""
    import torch
    mm, transformers = _import_reference(kind)
    import param
    rg = rg or r
    dec_r = dec_r or r
    flags = BASE_FLAGS.split() + [GATE_FLAG[gate], "--adapter_down_dim", str(r), "--encoder_adapter_multihead_num_head",
                                  str(heads), "--adapter_gating_down_dim", str(rg),
                                  "--decoder_enc_attn_value_parallel_adapter_down_dim", str(dec_r), "--dropout", str(dropout)]
    if kind == "t5":   # T5-VL-PET-large.sh:41-58
        flags += ["--use_encoder_gating_scaling", "--encoder_gating_scaling_factor", "0.3"]
    old = sys.argv
    sys.argv = ["x"] + flags
    try:
        args = param.parse_args()
    finally:
        sys.argv = old
    if kind == "bart":
        config = transformers.BartConfig(vocab_size=50465, d_model=768, encoder_layers=layers or 6, decoder_layers=layers or 6,
                                         encoder_attention_heads=12, decoder_attention_heads=12, encoder_ffn_dim=3072,
                                         decoder_ffn_dim=3072, max_position_embeddings=1024, activation_function="gelu",
                                         init_std=0.02, pad_token_id=1, bos_token_id=0, eos_token_id=2,
                                         decoder_start_token_id=2)
        d = 768
    else:
        config = transformers.T5Config(vocab_size=32200, d_model=768, d_kv=64, d_ff=3072, num_layers=layers or 12,
                                       num_decoder_layers=layers or 12, num_heads=12, relative_attention_num_buckets=32,
                                       feed_forward_proj="relu", layer_norm_epsilon=1e-6, pad_token_id=0,
                                       decoder_start_token_id=0, tie_word_embeddings=True)
        d = 768
    for k, v in vars(args).items():          # trainer_base.py:86-87
        setattr(config, k, v)
    config.dropout = config.dropout_rate = dropout
    config.attention_dropout = config.activation_dropout = dropout
    from adapters import AdapterConfig       # trainer_base.py:141-178
    ac = AdapterConfig()
    ac.tasks = args.tasks.split(",") if isinstance(args.tasks, str) else args.tasks
    ac.input_dim = ac.d_model = d
    ac.use_single_adapter = args.use_single_adapter
    ac.reduction_factor = args.reduction_factor
    ac.add_layer_norm_before_adapter = args.add_layer_norm_before_adapter
    ac.add_layer_norm_after_adapter = args.add_layer_norm_after_adapter
    ac.track_z = args.track_z
    ac.use_adapter_down_dim = bool(args.use_adapter_down_dim)
    ac.adapter_down_dim = args.adapter_down_dim
    ac.use_parallel_adapter = False
    ac.use_scaling_factor = False
    ac.scaling_factor = 1.0
    for k in ("unique_hyper_net", "efficient_unique_hyper_net", "hypercomplex_division", "phm_rank", "shared_phm_rule",
              "factorized_phm", "low_rank_rank", "phm_init_range", "share_down_sampler", "share_up_sampler",
              "shared_phm_rule_over_tasks"):
        if hasattr(args, k):
            setattr(ac, k, getattr(args, k))
    config.adapter_config = ac
    config.encoder_prompt_config = config.decoder_prompt_config = None
    config.feat_dim, config.pos_dim, config.n_images = feat_dim, 4, 2
    config.n_boxes = n_boxes
    config.use_vis_order_embedding, config.use_vis_layer_norm, config.individual_vis_layer_norm = True, True, True
    config.share_vis_lang_layer_norm = False
    config.default_obj_order_ids = None
    config.losses = "lm"
    config.classifier = False
    torch.manual_seed(seed)
    model = (mm.VLBartMultiTask if kind == "bart" else mm.VLT5MultiTask)(config)
    if kind == "bart":
        model.lm_head.weight = model.model.shared.weight   # 4.2.1's init_weights() ties them; the shimmed one does not
    # multitask.py:78-79 sets these from the tokenizer ('true' / 'false'); no tokenizer offline: two fixed vocabulary ids
    model.true_id, model.false_id = 1528, 3950
    return model, config


def prepare_training(model, lr: float = 1e-3):
    """freeze_whole_model + unfreeze by name substring (trainer_base.py:268-270, 308-542 for the VL-PET flag set) and the
    two AdamW groups of trainer_base.py:627-732 (torch.optim.AdamW: transformers.optimization.AdamW is gone in 5.x)."""
    import torch
    names = []
    for n, p in model.named_parameters():
        on = any(t in n for t in ("adapter", "gating", "visual_embedding")) or \
            ("encoder." in n and ("layer_norm" in n or "layernorm" in n))
        p.requires_grad_(on)
        if on:
            names.append(n)
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    no_decay = ("bias", "LayerNorm.weight")
    opt = torch.optim.AdamW([{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
                             {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}],
                            lr=lr, eps=1e-6)
    return [p for _, p in named], opt


def train_steps(model, params, opt, cycle, steps: int, warmup: int, device, autocast_dtype=None):
    """fwd + bwd + clip 5 + AdamW (multitask.py:217-300) over the task cycle.  -> (samples, seconds, last loss)."""
    import torch
    cuda = torch.device(device).type == "cuda"
    n, t_total, loss_v = 0, 0.0, float("nan")
    for i in range(warmup + steps):
        b = cycle[i % len(cycle)]
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        if autocast_dtype is not None:
            with torch.autocast(device_type="cuda", dtype=autocast_dtype):
                loss = model.train_step(b)["loss"]
        else:
            loss = model.train_step(b)["loss"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        loss_v = float(loss.detach())
        if i >= warmup:
            n += b["input_ids"].shape[0]
            t_total += dt
    return n, t_total, loss_v


def reference_batches(batch_size: int, tasks, feat_dim: int = 2048, seed: int = 0, vocab_hi: int = 50000):
    """The same synthetic cycle as the GPU arm, in the reference's collate schema (the NLVR batch needs the [B, 2, ...]
    image pair layout of nlvr_model.py:150-170, which is what the generator produces)."""
    S = _synthetic()
    return S.multitask_cycle(batch_size, list(tasks), feat_dim=feat_dim, seed=seed, vocab_hi=vocab_hi), S.task_batch_sizes(batch_size)


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
