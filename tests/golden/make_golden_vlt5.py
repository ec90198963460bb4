"""Golden fixture for the T5 caller of the hot path: the reference's own ``VLT5`` (src/modeling_t5.py) on a tiny
configuration with the flag set of scripts/image-text/T5-VL-PET-large.sh (large gate, gate scale 0.3), forward + backward of
one VQA-shaped and one NLVR-shaped batch (test infrastructure; twin of make_golden_vlbart.py).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden_vlt5.py

Writes ``vlt5_tiny_large.npz``: the reference state_dict (every key), the batches, the per-task loss after the reference's
loss shaping and the gradient of every trainable parameter.  ``tests/test_host_model.py`` loads the state_dict into
``vlpet_b200.host.VLT5`` key for key and must reproduce loss and gradients."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from baseline import ref_arm as RA  # noqa: E402  (the T5 import shims of SURVEY Appendix C, items 7-9, live there)

D, R_, HEADS, FEAT = 64, 16, 4, 128


def build(gate: str = "large"):
    mm, transformers = RA._import_reference("t5")
    import param
    flags = RA.BASE_FLAGS.split() + [RA.GATE_FLAG[gate], "--adapter_down_dim", str(R_), "--encoder_adapter_multihead_num_head",
                                     str(HEADS), "--adapter_gating_down_dim", str(R_),
                                     "--decoder_enc_attn_value_parallel_adapter_down_dim", str(R_), "--dropout", "0.0",
                                     "--use_encoder_gating_scaling", "--encoder_gating_scaling_factor", "0.3"]
    old = sys.argv
    sys.argv = ["x"] + flags
    try:
        args = param.parse_args()
    finally:
        sys.argv = old
    config = transformers.T5Config(vocab_size=300, d_model=D, d_kv=16, d_ff=128, num_layers=2, num_decoder_layers=2, num_heads=4,
                                   relative_attention_num_buckets=32, feed_forward_proj="relu", layer_norm_epsilon=1e-6,
                                   pad_token_id=0, decoder_start_token_id=0, tie_word_embeddings=True, dropout_rate=0.0)
    for k, v in vars(args).items():
        setattr(config, k, v)
    config.dropout = config.dropout_rate = 0.0
    from adapters import AdapterConfig
    ac = AdapterConfig()
    ac.tasks = args.tasks.split(",") if isinstance(args.tasks, str) else args.tasks
    ac.input_dim = ac.d_model = D
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
    config.feat_dim, config.pos_dim, config.n_images = FEAT, 4, 2
    config.use_vis_order_embedding, config.use_vis_layer_norm, config.individual_vis_layer_norm = True, True, True
    config.share_vis_lang_layer_norm = False
    config.default_obj_order_ids = None
    config.losses = "lm"
    config.classifier = False
    torch.manual_seed(0)
    model = mm.VLT5MultiTask(config).double().eval()
    model.lm_head.weight = model.shared.weight        # tie_word_embeddings (the shimmed init_weights skips tie_weights)
    model.true_id, model.false_id = 11, 12            # multitask.py:78-79 (tokenizer ids of 'true' / 'false'; logging only)
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():                              # trained-like PET weights (the scripts' zero-init verifies nothing)
        for n, p in model.named_parameters():
            if any(t in n for t in ("adapter", "gating", "visual_embedding")) or ("encoder." in n and "layer_norm" in n):
                if "norm" in n or (n.endswith(".1.weight") and "visual_embedding" in n):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen, dtype=torch.float64))
                else:
                    p.copy_(torch.randn(p.shape, generator=gen, dtype=torch.float64) * (0.02 if p.dim() == 1 else 0.05))
            elif p.dim() == 2 and "shared" not in n and "embed_tokens" not in n:
                p.mul_(0.5)                            # keep the un-scaled T5 attention logits moderate in the tiny model
    return model, config


def main():
    model, config = build()
    g = torch.Generator().manual_seed(2)
    B, Lt, T = 3, 7, 4
    out = {"meta_gate": np.array("large")}
    sd = model.state_dict()
    out["meta_state_keys"] = np.array(list(sd.keys()))
    for k, v in sd.items():
        out["sd/" + k] = v.detach().cpu().numpy()
    trainable = [n for n, _ in model.named_parameters()
                 if any(t in n for t in ("adapter", "gating", "visual_embedding")) or
                 ("encoder." in n and ("layer_norm" in n or "layernorm" in n))]
    out["meta_trainable"] = np.array(trainable)
    params = dict(model.named_parameters())
    for task in ("vqa", "nlvr"):
        ids = torch.randint(3, 300, (B, Lt), generator=g)
        tgt = torch.randint(3, 300, (B, T), generator=g)
        tgt[0, 3] = -100
        tgt[2, 2:] = -100
        if task == "nlvr":
            feats = torch.randn(B, 2, 49, FEAT, generator=g, dtype=torch.float64)
            boxes = torch.zeros(B, 2, 49, 4, dtype=torch.float64)
        else:
            feats = torch.randn(B, 49, FEAT, generator=g, dtype=torch.float64)
            boxes = torch.zeros(B, 49, 4, dtype=torch.float64)
        batch = {"task": task, "input_ids": ids, "target_ids": tgt, "vis_feats": feats, "boxes": boxes}
        if task == "vqa":
            batch["scores"] = torch.rand(B, generator=g, dtype=torch.float64)
        model.zero_grad()
        loss = model.train_step(batch)["loss"]           # the reference's own task step (vqa_model.py / nlvr_model.py)
        loss.backward()
        for k in ("input_ids", "target_ids", "vis_feats", "boxes", "scores"):
            if k in batch:
                out[f"{task}/{k}"] = batch[k].numpy()
        out[f"{task}/loss"] = np.array(loss.item())
        for n in trainable:
            out[f"{task}/grad/{n}"] = params[n].grad.detach().numpy().copy()
    path = os.path.join(HERE, "vlt5_tiny_large.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", "loss", {t: float(out[t + '/loss']) for t in ("vqa", "nlvr")},
          "trainable", sum(params[n].numel() for n in trainable))


if __name__ == "__main__":
    main()
