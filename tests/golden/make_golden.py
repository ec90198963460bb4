"""Generate the golden fixtures in this directory from the UNMODIFIED reference (test infrastructure).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

For every PET site it instantiates the reference's own layer class (BartEncoderLayer, T5LayerSelfAttention,
T5LayerFF, AdapterController, VisualEmbedding) with the flag set of the shipped VL-PET scripts, replaces the
*frozen* sub-modules around the PET block (self-attention, fc2, LayerNorm) by stubs that return a fixed leaf
tensor -- so x1 and x2 are independent inputs -- and runs the reference forward + autograd backward in
float64.  Inputs, "trained-like" weights (N(0,0.05) / biases N(0,0.02): the zero-init of the T5 scripts would
verify nothing), outputs and every gradient go into one ``.npz`` per case.

The GPU box has no /root/reference: tests read only the committed ``.npz`` files.
"""
import copy
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import as R  # noqa: E402

BASE_FLAGS = ("--use_adapter --use_single_adapter --no_encoder_adapter --use_adapter_down_dim "
              "--use_encoder_adapter_down_multihead --unfreeze_encoder_layer_norms --no_decoder_adapter "
              "--use_decoder_enc_attn_value_parallel_adapter_down_dim --tasks vqa,gqa,nlvr,caption --dropout 0.0")
GATE_FLAG = {
    "large": "--use_encoder_adapter_gating_large_x_lowrank",
    "middle_x": "--use_encoder_adapter_gating_middle_xy_add",
    "middle_y": "--use_encoder_adapter_gating_middle_ia3_add",
    "small": "--use_encoder_adapter_gating_small_xy_cat",
}


def make_config(kind, d, r, heads, rg, gate, add_gate=False, s=None, alpha=None, vpa_r=None, vpa_sf=None):
    flags = BASE_FLAGS.split() + [GATE_FLAG[gate], "--adapter_down_dim", str(r),
                                  "--encoder_adapter_multihead_num_head", str(heads),
                                  "--adapter_gating_down_dim", str(rg),
                                  "--decoder_enc_attn_value_parallel_adapter_down_dim", str(vpa_r or r)]
    if add_gate:
        flags.append("--use_encoder_adapter_gating_add")
    if s is not None:
        flags += ["--use_encoder_gating_scaling", "--encoder_gating_scaling_factor", str(s)]
    if alpha is not None:
        flags += ["--use_encoder_adapter_scaling", "--encoder_adapter_scaling_factor", str(alpha)]
    if vpa_sf is not None:
        flags += ["--use_decoder_enc_attn_value_parallel_adapter_scaling",
                  "--decoder_enc_attn_value_parallel_adapter_scaling_factor", str(vpa_sf)]
    args = R.parse_args(flags)
    import transformers
    if kind == "bart":
        config = transformers.BartConfig(vocab_size=200, d_model=d, encoder_layers=1, decoder_layers=1,
                                         encoder_attention_heads=4, decoder_attention_heads=4,
                                         encoder_ffn_dim=2 * d, decoder_ffn_dim=2 * d, dropout=0.0,
                                         attention_dropout=0.0, activation_dropout=0.0)
    else:
        config = transformers.T5Config(vocab_size=200, d_model=d, d_kv=d // 4, d_ff=2 * d, num_layers=1,
                                       num_decoder_layers=1, num_heads=4, dropout_rate=0.0,
                                       feed_forward_proj="relu")
    for k, v in vars(args).items():          # trainer_base.py:86-87
        setattr(config, k, v)
    config.dropout = 0.0
    config.dropout_rate = 0.0
    # trainer_base.py:141-178
    from adapters import AdapterConfig
    ac = AdapterConfig()
    ac.tasks = args.tasks.split(",") if isinstance(args.tasks, str) else args.tasks
    ac.input_dim = d
    ac.d_model = d
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
    for k in ("unique_hyper_net", "efficient_unique_hyper_net", "hypercomplex_division", "phm_rank",
              "shared_phm_rule", "factorized_phm", "low_rank_rank", "phm_init_range", "share_down_sampler",
              "share_up_sampler", "shared_phm_rule_over_tasks"):
        if hasattr(args, k):
            setattr(ac, k, getattr(args, k))
    config.adapter_config = ac
    config.encoder_prompt_config = None
    config.decoder_prompt_config = None
    return config


class Fixed(nn.Module):
    """Stub for a frozen sub-module: ignores its input, returns a fixed leaf tensor (optionally as tuple)."""

    def __init__(self, value, as_tuple=0):
        super().__init__()
        self.value, self.as_tuple = value, as_tuple

    def forward(self, *a, **k):
        return (self.value,) + (None,) * (self.as_tuple - 1) if self.as_tuple else self.value


def trained_like_(module, gen):
    with torch.no_grad():
        for n, p in module.named_parameters():
            std = 0.02 if (n.endswith("bias") or p.dim() == 1) else 0.05
            p.copy_(torch.randn(p.shape, generator=gen, dtype=torch.float64) * std)


def t2n(t):
    return t.detach().cpu().numpy().copy()


def site_params(layer, prefix_adapter, prefix_gate, gate):
    """Collect one site's parameters in oracle naming from the reference layer's own attribute names
    (SURVEY Appendix B)."""
    down = getattr(layer, prefix_adapter + "_adapter_multihead_down")
    up = getattr(layer, prefix_adapter + "_adapter_multihead_up")
    out = {"Wd_heads": [h.weight for h in down], "bd_heads": [h.bias for h in down], "Wu": up.weight, "bu": up.bias}
    if gate == "large":
        gd = getattr(layer, f"encoder_{prefix_gate}_adapter_gating_large_x_down")
        gu = getattr(layer, f"encoder_{prefix_gate}_adapter_gating_large_x_up")
        out.update(Gd=gd.weight, gbd=gd.bias, Gu=gu.weight, gbu=gu.bias)
    elif gate == "middle_x":
        g = getattr(layer, f"encoder_{prefix_gate}_adapter_gating_middle_xy_add")
        out.update(gw=g.weight, gb=g.bias)
    elif gate == "middle_y":
        out.update(gz=getattr(layer, f"encoder_{prefix_gate}_adapter_gating_middle_ia3_add"))
    elif gate == "small":
        g = getattr(layer, f"encoder_{prefix_gate}_adapter_gating_small_xy_cat")
        out.update(gw=g.weight, gb=g.bias)
    return out


def dump_site(name, meta, x1, x2, dout, out, params):
    rec = {"x1": t2n(x1), "x2": t2n(x2), "dout": t2n(dout), "out": t2n(out),
           "dx1": t2n(x1.grad), "dx2": t2n(x2.grad)}
    for k, v in params.items():
        if isinstance(v, list):
            rec[k[:-6]] = np.concatenate([t2n(h) for h in v], 0)               # Wd / bd (stacked heads)
            rec["d" + k[:-6]] = np.concatenate([t2n(h.grad) for h in v], 0)
        else:
            rec[k] = t2n(v).reshape(-1) if k in ("gw",) else t2n(v)
            rec["d" + k] = t2n(v.grad).reshape(-1) if k in ("gw",) else t2n(v.grad)
    for k, v in meta.items():
        rec["meta_" + k] = np.array(v)
    np.savez(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, {k: v.shape for k, v in rec.items() if not k.startswith("meta_")})


def bart_case(name, B, L, d, r, heads, rg, gate, seed, add_gate=False, s=None, ff=True):
    mb = R.import_bart_backbone()
    torch.manual_seed(seed)
    cfg = make_config("bart", d, r, heads, rg, gate, add_gate=add_gate, s=s)
    layer = mb.BartEncoderLayer(cfg).double().train()        # train(): dropout p=0 is still the identity
    gen = torch.Generator().manual_seed(seed)
    trained_like_(layer, gen)
    rnd = lambda *sh: torch.randn(*sh, generator=gen, dtype=torch.float64)  # noqa: E731
    meta = dict(B=B, L=L, d=d, r=r, heads=heads, rg=rg, gate=gate, add_gate=int(add_gate),
                s=1.0 if s is None else s, alpha=1.0, kappa=1.0, arch="bart")
    # ---- attention site: x1 = layer input, x2 = self_attn output (modeling_bart.py:1132-1141)
    x1 = rnd(B, L, d).requires_grad_()
    x2 = (0.5 * rnd(B, L, d)).requires_grad_()
    dout = rnd(B, L, d)
    layer.self_attn = Fixed(x2, as_tuple=3)
    grab = {}
    h = layer.self_attn_layer_norm.register_forward_pre_hook(lambda m, inp: grab.__setitem__("o", inp[0]))
    layer(x1, None, task="vqa")
    h.remove()
    (grab["o"] * dout).sum().backward()
    dump_site(name + "_attn", dict(meta, site="attn"), x1, x2, dout, grab["o"], site_params(layer, "attn", "attn", gate))
    if not ff:
        return
    # ---- FFN site: x1' = post-attn-LN stream, x2' = fc2 output (modeling_bart.py:1263-1377)
    layer.zero_grad()
    x1f = rnd(B, L, d).requires_grad_()
    x2f = (0.5 * rnd(B, L, d)).requires_grad_()
    doutf = rnd(B, L, d)
    layer.self_attn_layer_norm = Fixed(x1f)
    layer.fc2 = Fixed(x2f)
    h = layer.final_layer_norm.register_forward_pre_hook(lambda m, inp: grab.__setitem__("f", inp[0]))
    layer(x1.detach(), None, task="vqa")
    h.remove()
    (grab["f"] * doutf).sum().backward()
    dump_site(name + "_ff", dict(meta, site="ff"), x1f, x2f, doutf, grab["f"], site_params(layer, "ff", "ff", gate))


def t5_case(name, B, L, d, r, heads, rg, gate, seed, s=0.3, alpha=None):
    mt = R.import_t5_backbone()
    torch.manual_seed(seed)
    cfg = make_config("t5", d, r, heads, rg, gate, s=s, alpha=alpha)
    cfg.use_encoder_x2_scaling = getattr(cfg, "use_encoder_x2_scaling", False)
    gen = torch.Generator().manual_seed(seed)
    rnd = lambda *sh: torch.randn(*sh, generator=gen, dtype=torch.float64)  # noqa: E731
    meta = dict(B=B, L=L, d=d, r=r, heads=heads, rg=rg, gate=gate, add_gate=0, s=s,
                alpha=1.0 if alpha is None else alpha, kappa=1.0, arch="t5")
    # ---- T5LayerSelfAttention (modeling_t5.py:766-824): x1 = un-normalised residual, x2 = SelfAttention out
    lay = mt.T5LayerSelfAttention(cfg, has_relative_attention_bias=False).double().train()
    trained_like_(lay, gen)
    x1 = rnd(B, L, d).requires_grad_()
    x2 = (0.5 * rnd(B, L, d)).requires_grad_()
    dout = rnd(B, L, d)
    lay.layer_norm = Fixed(x1.detach())
    lay.SelfAttention = Fixed(x2, as_tuple=2)
    out = lay(x1, task="vqa")[0]
    (out * dout).sum().backward()
    dump_site(name + "_attn", dict(meta, site="attn"), x1, x2, dout, out, site_params(lay, "attn", "attn", gate))
    # ---- T5LayerFF (modeling_t5.py:359-409)
    lay = mt.T5LayerFF(cfg).double().train()
    trained_like_(lay, gen)
    x1 = rnd(B, L, d).requires_grad_()
    x2 = (0.5 * rnd(B, L, d)).requires_grad_()
    dout = rnd(B, L, d)
    lay.layer_norm = Fixed(x1.detach())
    lay.DenseReluDense = Fixed(x2)
    out = lay(x1, None, "vqa")
    (out * dout).sum().backward()
    dump_site(name + "_ff", dict(meta, site="ff"), x1, x2, dout, out, site_params(lay, "ff", "ff", gate))


def vpa_case(name, B, L, d, r, seed, sf=None):
    """Decoder VPA through the reference's BartAttentionWithValueAdapter constructor (so the adapter config is
    wired exactly as modeling_bart.py:329-340) and AdapterController.forward (adapter_controller.py:131-162)."""
    mb = R.import_bart_backbone()
    torch.manual_seed(seed)
    cfg = make_config("bart", d, r, 4, r, "large", vpa_r=r, vpa_sf=sf)
    pac = copy.deepcopy(cfg.adapter_config)
    pac.use_adapter_down_dim = True                       # modeling_bart.py:1452-1464
    pac.adapter_down_dim = cfg.decoder_enc_attn_value_parallel_adapter_down_dim
    attn = mb.BartAttentionWithValueAdapter(d, 4, dropout=0.0, is_decoder=True, adapter_config=pac, config=cfg).double()
    ctrl = attn.attn_value_parallel_adapter
    gen = torch.Generator().manual_seed(seed)
    trained_like_(ctrl, gen)
    rnd = lambda *sh: torch.randn(*sh, generator=gen, dtype=torch.float64)  # noqa: E731
    kv = rnd(B, L, d).requires_grad_()
    y = rnd(B, L, d).requires_grad_()
    dout = rnd(B, L, d)
    out = ctrl(kv, "gqa", y=y)
    (out * dout).sum().backward()
    ad = ctrl.adapters["vqa"]
    assert ad is ctrl.adapters["gqa"]                     # use_single_adapter aliasing (adapter_controller.py:49-58)
    rec = dict(kv=t2n(kv), y=t2n(y), dout=t2n(dout), out=t2n(out), dkv=t2n(kv.grad), dy=t2n(y.grad),
               Wd=t2n(ad.down_sampler.weight), bd=t2n(ad.down_sampler.bias), Wu=t2n(ad.up_sampler.weight),
               bu=t2n(ad.up_sampler.bias), dWd=t2n(ad.down_sampler.weight.grad), dbd=t2n(ad.down_sampler.bias.grad),
               dWu=t2n(ad.up_sampler.weight.grad), dbu=t2n(ad.up_sampler.bias.grad),
               meta_sf=np.array(1.0 if sf is None else sf), meta_B=np.array(B), meta_L=np.array(L),
               meta_d=np.array(d), meta_r=np.array(r),
               meta_state_keys=np.array(sorted(ctrl.state_dict().keys())),
               meta_param_names=np.array([n for n, _ in ctrl.named_parameters()]))
    np.savez(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name)


def visproj_case(name, kind, B, N, F, d, seed, zero_boxes, explicit_ids):
    torch.manual_seed(seed)
    if kind == "bart":
        vl = R.import_vl_bart()
        cfg = make_config("bart", d, 16, 4, 16, "large")
    else:
        R.import_vl_bart()
        import modeling_t5 as vl
        cfg = make_config("t5", d, 16, 4, 16, "large")
    cfg.feat_dim, cfg.pos_dim, cfg.n_images = F, 4, 2
    cfg.default_obj_order_ids = None
    V = 150
    emb = nn.Embedding(V, d).double()
    ve = vl.VisualEmbedding(cfg, emb).double()
    gen = torch.Generator().manual_seed(seed)
    trained_like_(ve, gen)
    with torch.no_grad():   # LN weights ~ 1
        for n, p in ve.named_parameters():
            if (".1.weight" in n) or n.endswith("layer_norm.weight"):
                p.add_(1.0)
    rnd = lambda *sh: torch.randn(*sh, generator=gen, dtype=torch.float64)  # noqa: E731
    feats = rnd(B, N, F).requires_grad_()
    pos = torch.zeros(B, N, 4, dtype=torch.float64) if zero_boxes else torch.rand(B, N, 4, generator=gen, dtype=torch.float64)
    img_ids = obj_ids = None
    if explicit_ids:   # NLVR-style two images (nlvr_model.py)
        img_ids = torch.cat([torch.zeros(N // 2, dtype=torch.long), torch.ones(N - N // 2, dtype=torch.long)])[None].expand(B, -1)
        obj_ids = torch.cat([torch.arange(N // 2), torch.arange(N - N // 2)])[None].expand(B, -1)
    dout = rnd(B, N, d)
    out = ve(feats, pos, img_ids, obj_ids)
    (out * dout).sum().backward()
    fe, pe = ve.feat_embedding, ve.absolute_vis_pos_embedding
    rec = dict(feats=t2n(feats), pos=t2n(pos), dout=t2n(dout), out=t2n(out), dfeats=t2n(feats.grad),
               Wf=t2n(fe[0].weight), bf=t2n(fe[0].bias), ln_f_w=t2n(fe[1].weight),
               Wp=t2n(pe[0].weight), bp=t2n(pe[0].bias), ln_p_w=t2n(pe[1].weight),
               E_img=t2n(ve.img_order_embedding.weight), E_obj=t2n(emb.weight),
               dWf=t2n(fe[0].weight.grad), dbf=t2n(fe[0].bias.grad), dln_f_w=t2n(fe[1].weight.grad),
               dWp=t2n(pe[0].weight.grad), dbp=t2n(pe[0].bias.grad), dln_p_w=t2n(pe[1].weight.grad),
               dE_img=t2n(ve.img_order_embedding.weight.grad),
               meta_kind=np.array(kind), meta_eps=np.array(1e-5 if kind == "bart" else cfg.layer_norm_epsilon),
               meta_param_names=np.array([n for n, _ in ve.named_parameters()]))
    if kind == "bart":
        rec.update(ln_f_b=t2n(fe[1].bias), ln_p_b=t2n(pe[1].bias), dln_f_b=t2n(fe[1].bias.grad), dln_p_b=t2n(pe[1].bias.grad))
    if explicit_ids:
        rec.update(img_ids=t2n(img_ids), obj_ids=t2n(obj_ids))
    np.savez(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name)


def main():
    assert R.available(), "reference not found; goldens can only be regenerated in the build container"
    # small-d cases for every granularity (fast, tiny files)
    bart_case("k1_bart_large_d64", 2, 7, 64, 16, 4, 16, "large", 1)
    bart_case("k1_bart_large_add_s_d64", 2, 5, 64, 16, 2, 8, "large", 2, add_gate=True, s=0.7)
    bart_case("k1_bart_middlex_d64", 3, 6, 64, 16, 4, 16, "middle_x", 3)
    bart_case("k1_bart_middley_d64", 2, 7, 64, 16, 4, 16, "middle_y", 4)
    bart_case("k1_bart_small_d64", 3, 5, 64, 16, 4, 16, "small", 5)
    bart_case("k1_bart_middlex_add_d64", 2, 4, 64, 8, 1, 8, "middle_x", 6, add_gate=True)
    bart_case("k1_bart_small_add_d64", 2, 4, 64, 8, 1, 8, "small", 7, add_gate=True, s=0.5)
    bart_case("k1_bart_middley_add_d64", 2, 4, 64, 8, 1, 8, "middle_y", 8, add_gate=True)
    t5_case("k1_t5_large_d64", 2, 6, 64, 16, 4, 16, "large", 9, s=0.3)
    t5_case("k1_t5_large_alpha_d64", 2, 6, 64, 16, 4, 8, "large", 10, s=0.3, alpha=0.6)
    t5_case("k1_t5_small_d64", 2, 6, 64, 16, 4, 16, "small", 11, s=0.3)
    # the headline shape d=768, r=rg=96, 4 heads (few tokens to keep the file small)
    bart_case("k1_bart_large_d768", 1, 9, 768, 96, 4, 96, "large", 12, ff=False)   # 5 MB: one site only
    vpa_case("k2_vpa_d64", 2, 7, 64, 16, 13)
    vpa_case("k2_vpa_sf_d64", 2, 5, 64, 24, 14, sf=0.7)
    visproj_case("k3_bart_f128_d64", "bart", 2, 6, 128, 64, 16, zero_boxes=False, explicit_ids=False)
    visproj_case("k3_bart_nlvr_d64", "bart", 2, 8, 128, 64, 17, zero_boxes=True, explicit_ids=True)
    visproj_case("k3_t5_f128_d64", "t5", 2, 6, 128, 64, 18, zero_boxes=False, explicit_ids=False)


if __name__ == "__main__":
    main()
