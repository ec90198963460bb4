"""Golden fixture for the CALLER of the hot path: the reference's own ``VLBart`` (src/modeling_bart.py) on a tiny
configuration, forward + backward of one VQA-shaped and one NLVR-shaped batch (test infrastructure).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden_vlbart.py

Writes ``vlbart_tiny_<gate>.npz`` holding the reference state_dict (every key), the batch, the per-task loss after
the reference's loss shaping (vqa_model.py:211-227 / nlvr_model.py:184-193), and the gradient of every trainable
parameter.  ``tests/test_host_model.py`` loads the state_dict into ``vlpet_b200.host.VLBart`` key for key and
must reproduce loss and gradients -- on CPU with the eager PET restatement, on the GPU with the CUDA kernels.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import as R  # noqa: E402
import make_golden as G  # noqa: E402

D, R_, HEADS, RG, FEAT = 64, 16, 4, 16, 128


def build(gate):
    vlb = R.import_vl_bart()
    import transformers
    flags = G.BASE_FLAGS.split() + [G.GATE_FLAG[gate], "--adapter_down_dim", str(R_),
                                    "--encoder_adapter_multihead_num_head", str(HEADS),
                                    "--adapter_gating_down_dim", str(RG),
                                    "--decoder_enc_attn_value_parallel_adapter_down_dim", str(R_),
                                    "--n_boxes", "36", "--downsample"]
    args = R.parse_args(flags)
    config = transformers.BartConfig(vocab_size=300, d_model=D, encoder_layers=2, decoder_layers=2,
                                     encoder_attention_heads=4, decoder_attention_heads=4, encoder_ffn_dim=128,
                                     decoder_ffn_dim=128, max_position_embeddings=128, dropout=0.0,
                                     attention_dropout=0.0, activation_dropout=0.0, pad_token_id=1, bos_token_id=0,
                                     eos_token_id=2, decoder_start_token_id=2, activation_function="gelu")
    for k, v in vars(args).items():
        setattr(config, k, v)
    config.dropout = config.attention_dropout = config.activation_dropout = 0.0
    ref_cfg = G.make_config("bart", D, R_, HEADS, RG, gate)
    config.adapter_config = ref_cfg.adapter_config
    config.encoder_prompt_config = config.decoder_prompt_config = None
    config.feat_dim, config.pos_dim, config.n_images = FEAT, 4, 2
    config.use_vis_order_embedding, config.use_vis_layer_norm, config.individual_vis_layer_norm = True, True, True
    config.share_vis_lang_layer_norm = False
    config.default_obj_order_ids = None
    config.losses = "lm"
    config.classifier = False
    torch.manual_seed(0)
    model = vlb.VLBart(config).double().eval()
    # transformers 4.2.1's init_weights() ends with tie_weights(): lm_head shares model.shared.  The init_weights shim
    # needed under transformers 5.x (ref_import.py) skips that step, so the tie is restored here.
    model.lm_head.weight = model.model.shared.weight
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():                                   # trained-like PET weights, so gates are not pinned at 0.5
        for n, p in model.named_parameters():
            if any(t in n for t in ("adapter", "gating", "visual_embedding")) or ("encoder." in n and "layer" in n and "norm" in n):
                if "norm" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen, dtype=torch.float64))
                else:
                    std = 0.02 if p.dim() == 1 else 0.05
                    p.copy_(torch.randn(p.shape, generator=gen, dtype=torch.float64) * std)
    return model, config


def main():
    for gate in ("large", "small"):
        model, config = build(gate)
        g = torch.Generator().manual_seed(2)
        B, Lt, T = 3, 7, 4
        out = {"meta_gate": np.array(gate)}
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
                V_L = 49
                f2, b2 = feats.view(B, 2 * V_L, -1), boxes.view(B, 2 * V_L, 4)
                img = torch.tensor([0] * V_L + [1] * V_L).view(1, -1).expand(B, -1)
                obj = torch.arange(V_L).view(1, 1, V_L).expand(B, 2, -1).contiguous().view(B, 2 * V_L)
                vis_inputs = (f2, b2, img, obj)
            else:
                feats = torch.randn(B, 49, FEAT, generator=g, dtype=torch.float64)
                boxes = torch.zeros(B, 49, 4, dtype=torch.float64)
                vis_inputs = (feats, boxes)
            scores = torch.rand(B, generator=g, dtype=torch.float64) if task == "vqa" else None
            model.zero_grad()
            o = model(input_ids=ids, vis_inputs=vis_inputs, labels=tgt, return_dict=True, task=task)
            mask = (tgt != -100).double()
            loss = (o["loss"].view(B, T) * mask).sum(dim=1) / mask.sum(dim=1).clamp(min=1)
            if scores is not None:
                loss = loss * scores
            loss = loss.mean()
            loss.backward()
            out[f"{task}/input_ids"], out[f"{task}/target_ids"] = ids.numpy(), tgt.numpy()
            out[f"{task}/vis_feats"], out[f"{task}/boxes"] = feats.numpy(), boxes.numpy()
            if scores is not None:
                out[f"{task}/scores"] = scores.numpy()
            out[f"{task}/loss"] = np.array(loss.item())
            out[f"{task}/logits"] = o["logits"].detach().numpy()
            for n in trainable:
                out[f"{task}/grad/{n}"] = params[n].grad.detach().numpy().copy()
        path = os.path.join(HERE, f"vlbart_tiny_{gate}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KB", "loss", {t: float(out[t + '/loss']) for t in ("vqa", "nlvr")},
              "trainable", sum(params[n].numel() for n in trainable))


if __name__ == "__main__":
    main()
