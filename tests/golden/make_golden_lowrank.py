"""Golden vectors for SURVEY §8 row a7: the reference's LowRankVisualEmbedding (src/modeling_bart.py:195-334), forward +
autograd backward in float64 (test infrastructure; run in the build container: python tests/golden/make_golden_lowrank.py)."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import as R  # noqa: E402
import make_golden as G  # noqa: E402


def case(name, B, N, F, d, r, heads, rg, seed, gated, residual, explicit_ids):
    vl = R.import_vl_bart()
    cfg = G.make_config("bart", d, 16, 4, 16, "large")
    cfg.feat_dim, cfg.pos_dim, cfg.n_images, cfg.default_obj_order_ids = F, 4, 2, None
    cfg.use_lowrank_visual_projector = True
    cfg.visual_projector_down_dim, cfg.visual_projector_multihead_num_head = r, heads
    cfg.use_visual_projector_gating_large_x_lowrank, cfg.visual_projector_gating_down_dim = gated, rg
    cfg.use_visual_projector_residual_connection = residual
    torch.manual_seed(seed)
    emb = nn.Embedding(150, d).double()
    ve = vl.LowRankVisualEmbedding(cfg, emb).double()
    gen = torch.Generator().manual_seed(seed)
    G.trained_like_(ve, gen)
    with torch.no_grad():
        for n, p in ve.named_parameters():
            if n.endswith("layer_norm.weight") or ".1.weight" in n:
                p.add_(1.0)
    rnd = lambda *sh: torch.randn(*sh, generator=gen, dtype=torch.float64)  # noqa: E731
    feats = rnd(B, N, F).requires_grad_()
    pos = torch.rand(B, N, 4, generator=gen, dtype=torch.float64)
    img_ids = obj_ids = None
    if explicit_ids:
        img_ids = torch.cat([torch.zeros(N // 2, dtype=torch.long), torch.ones(N - N // 2, dtype=torch.long)])[None].expand(B, -1)
        obj_ids = torch.cat([torch.arange(N // 2), torch.arange(N - N // 2)])[None].expand(B, -1)
    dout = rnd(B, N, d)
    out = ve(feats, pos, img_ids, obj_ids)
    (out * dout).sum().backward()
    t = G.t2n
    down = ve.visual_projector_multihead_down
    pe = ve.absolute_vis_pos_embedding
    rec = dict(feats=t(feats), pos=t(pos), dout=t(dout), out=t(out), dfeats=t(feats.grad),
               Wd=np.concatenate([t(h.weight) for h in down]), bd=np.concatenate([t(h.bias) for h in down]),
               Wu=t(ve.visual_projector_multihead_up.weight), bu=t(ve.visual_projector_multihead_up.bias),
               ln_f_w=t(ve.visual_projector_layer_norm.weight), ln_f_b=t(ve.visual_projector_layer_norm.bias),
               Wp=t(pe[0].weight), bp=t(pe[0].bias), ln_p_w=t(pe[1].weight), ln_p_b=t(pe[1].bias),
               E_img=t(ve.img_order_embedding.weight), E_obj=t(emb.weight),
               dWd=np.concatenate([t(h.weight.grad) for h in down]), dbd=np.concatenate([t(h.bias.grad) for h in down]),
               dWu=t(ve.visual_projector_multihead_up.weight.grad), dbu=t(ve.visual_projector_multihead_up.bias.grad),
               dln_f_w=t(ve.visual_projector_layer_norm.weight.grad), dln_f_b=t(ve.visual_projector_layer_norm.bias.grad),
               dWp=t(pe[0].weight.grad), dbp=t(pe[0].bias.grad), dln_p_w=t(pe[1].weight.grad), dln_p_b=t(pe[1].bias.grad),
               dE_img=t(ve.img_order_embedding.weight.grad),
               meta_gated=np.array(int(gated)), meta_residual=np.array(int(residual)), meta_heads=np.array(heads),
               meta_param_names=np.array([n for n, _ in ve.named_parameters()]))
    if gated:
        gd, gu = ve.visual_projector_gating_large_x_down, ve.visual_projector_gating_large_x_up
        rec.update(Gd=t(gd.weight), gbd=t(gd.bias), Gu=t(gu.weight), gbu=t(gu.bias), dGd=t(gd.weight.grad),
                   dgbd=t(gd.bias.grad), dGu=t(gu.weight.grad), dgbu=t(gu.bias.grad))
    if explicit_ids:
        rec.update(img_ids=t(img_ids), obj_ids=t(obj_ids))
    np.savez(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name)


if __name__ == "__main__":
    assert R.available()
    case("k3lr_gated_f128_d64", 2, 6, 128, 64, 16, 4, 8, 31, True, False, False)
    case("k3lr_gated_res_nlvr_f128_d64", 2, 8, 128, 64, 16, 2, 16, 32, True, True, True)
    case("k3lr_plain_f128_d64", 2, 6, 128, 64, 16, 1, 16, 33, False, False, False)
